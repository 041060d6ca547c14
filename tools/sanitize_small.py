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
# SURVEY 8(f) kernels: device canvases, the depth-only program, Kawase / glow, span bars, marching cubes (small sizes)
g.set_overlap(False)
w, h = size
cq, ch, tc = g.Canvas("quads", w, h), g.Canvas("fp", w // 2, h // 2), g.Canvas("tc", w, h)
sc = scenes.SoupScene(n=100, seed=5)
g.Reset(size, (8, 8))
g.ClearColor((0.1, 0.2, 0.3)); g.ClearDepth(1.0); g.Clear(R.GL_COLOR_BUFFER_BIT | R.GL_DEPTH_BUFFER_BIT)
sc.draw(g, size)
g.UseProgram(R.PROGRAM_DEFAULT_POST)
g.StoreToCanvas(cq); g.StoreToCanvas(ch, half=True)
g.Run(sync=False)
blurred = g.Kawase(ch, 3)
g.Glow(cq, blurred, tc, True)
g.DrawSpans(tc, 5, 5, 2.0, [(0.01 * i, 0.01 * i + 0.02, 77 * i, i % 4) for i in range(12)])
print("post chain", hex(int(tc.read().sum()) & 0xffffffff))
depth = g.Canvas("depth", 128, 128)
g.Reset((128, 128), (8, 8))
g.RenderbufferType(R.GL_COLOR_ATTACHMENT0, R.RB_RGBF32); g.RenderbufferType(R.GL_DEPTH_ATTACHMENT, R.RB_F32)
g.ColorWriteMask(False); g.Enable(R.GL_CULL_FACE); g.CullFace(R.GL_FRONT); g.ClearDepth(1.0); g.Clear(R.GL_DEPTH_BUFFER_BIT)
g.UseProgram(0)
g.ViewMatrix(scenes.translate(0.3, 0.2, -2.0)); g.ProjectionMatrix(scenes.perspective(60.0, 1.0, 0.5, 60.0))
g.UseBuffer(0, sc.pos); g.DrawElements(len(sc.idx), sc.idx, 0)
g.StoreToCanvas(depth)
g.Run()
print("shadow map", float(depth.read().min()))
for args in ((0.7, 32, 2, 5.0), (1.9, 64, 1, 5.0), (0.2, 8, 0, 5.0)):
    soa, blocks, total = g.MarchSurface(*args)
    print("march", args, len(blocks), total, float(g.read_device(soa[0], total).sum()))
g.close()
print("done")
