"""small frames through every rasteriser path, for compute-sanitizer (memcheck / racecheck):
compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rsr_b200 as R
from rsr_b200 import scenes

g = R.GPU(0)
size = (320, 192)
for sc, kw in ((scenes.WavyGridScene(n=24), {}), (scenes.CubesScene(instances=40), {}), (scenes.SoupScene(n=150, seed=3), {}),
               (scenes.SoupScene(n=120, seed=7, blend=True, cull=R.GL_BACK), {}), (scenes.GeometryStressScene(spheres=2, divs=4, size=size, radius_px=40.0), {}),
               (scenes.BundledLikeScene(cubes=60, field=6), {"t": 0.4}), (scenes.ColortestScene(), {})):
    out = np.zeros((size[1], size[0]), np.uint32)
    sc.record(g, size, out, **kw)
    g.Run()
    print(type(sc).__name__, g.stats()["fragments_shaded"], hex(int(out.sum()) & 0xffffffff))
g.set_overlap(True)
outs = [np.zeros((size[1], size[0]), np.uint32) for _ in range(4)]
for i in range(4):
    scenes.CubesScene(instances=30).record(g, size, outs[i], t=0.3 * i)
    g.Run(sync=False)
g.Sync()
h = g.Retain()
g.Replay(h, sync=True)
g.Release(h)
g.close()
print("done")
