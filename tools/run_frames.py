#!/usr/bin/env python3
"""renders a few frames of a bench workload (for ncu captures): python tools/run_frames.py c2 4"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import rsr_b200  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
scene, size, workload = bench.make_scene(name)
gpu = rsr_b200.GPU(0)
scene.record(gpu, size, None, t=0.0, static=True)
rec = gpu.Finish()
for i in range(n):
    gpu.Submit(rec)
print(workload, gpu.stats())
