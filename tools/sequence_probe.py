#!/usr/bin/env python3
"""frame-sequence throughput: a ring of distinct retained frames (own tables, instance data and output buffer each)
replayed back to back, with and without frame overlap: python tools/sequence_probe.py c2 [ring] [frames]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
import rsr_b200  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
ring = int(sys.argv[2]) if len(sys.argv) > 2 else 24
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 480
wl = bench.Workload(name)
gpu = rsr_b200.GPU(0)
stream = torch.cuda.ExternalStream(gpu.stream(), device=0)
retained = []
for i in range(ring):
    wl.record(gpu, wl.subframes[0], None, t=i / 60.0, static=True)
    gpu.Submit(gpu.Finish())
    retained.append(gpu.Retain())
for overlap in (False, True):
    gpu.set_overlap(overlap)
    for i in range(ring):
        gpu.Replay(retained[i])
    gpu.Sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        e0.record(stream)
    for i in range(frames):
        gpu.Replay(retained[i % ring])
    gpu.Join()
    with torch.cuda.stream(stream):
        e1.record(stream)
    t1 = time.perf_counter()
    gpu.Sync()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / frames
    print(f"{wl.name} ring {ring} overlap {overlap}: {ms * 1e3:.1f} us/frame = {1e3 / ms:.0f} frames/s (host enqueue {1e6 * (t1 - t0) / frames:.1f} us/frame)")
gpu.set_overlap(False)
