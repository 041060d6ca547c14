#!/usr/bin/env python3
"""stage timings of a few probe frames (device events): python tools/stage_probe.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rsr_b200 as R
from rsr_b200 import scenes

g = R.GPU(0)
g.set_profiling(2)
size = (1920, 1080)

class Empty:
    triangles = 0
    def record(self, gl, size, out, **kw):
        scenes.begin(gl, size); scenes.finish(gl, out)

def probe(name, sc, **kw):
    sc.record(g, size, None, **kw)
    rec = g.Finish()
    acc = {}
    host = []
    for i in range(25):
        t0 = time.perf_counter(); g.Submit(rec, sync=False); host.append(time.perf_counter() - t0); g.Sync()
        if i >= 5:
            for k, v in g.stage_ms().items():
                acc[k] = acc.get(k, 0) + v / 20
    print(f"{name:28s}", {k: round(v, 3) for k, v in acc.items()}, "host_submit_ms", round(1e3 * float(np.median(host)), 3), "record_us", g.stats()["host_record_ns"] // 1000, "end_frame_us", g.stats()["host_submit_ns"] // 1000, g.stats()["bin_entries"])

probe("empty", Empty())
c2 = scenes.BundledLikeScene()
probe("c2", c2, static=True)
probe("c2 cubes only", scenes.BundledLikeScene(field=0), static=True)
probe("c2 no cubes", scenes.BundledLikeScene(cubes=1, groups=1), static=True)
probe("grid 20k tris", scenes.WavyGridScene(n=100))
