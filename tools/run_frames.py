#!/usr/bin/env python3
"""renders a few frames of a bench workload (for ncu captures), the L2 flushed before every frame like bench.py's
timed region: python tools/run_frames.py c2 4"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
import rsr_b200  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
wl = bench.Workload(name)
gpu = rsr_b200.GPU(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
stream = torch.cuda.ExternalStream(gpu.stream(), device=0)
recs = []
for sf in wl.subframes:
    wl.record(gpu, sf, None, t=0.0, static=True)
    recs.append(gpu.Finish())
for i in range(n):
    for rec in recs:
        with torch.cuda.stream(stream):
            flush.fill_(i & 0xff)
        gpu.Submit(rec)
print(wl.name, gpu.stats())
