#!/usr/bin/env python3
"""c4 with every quad bound to ONE texture and no L2 flush: the tile kernel with its texels L2-resident -- an upper
bound on what hiding DRAM latency (prefetching) could give the fill-stress frame"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
import rsr_b200  # noqa: E402

wl = bench.Workload("c4")
gpu = rsr_b200.GPU(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
stream = torch.cuda.ExternalStream(gpu.stream(), device=0)
for label in ("distinct textures, L2 flushed", "distinct textures, no flush", "one shared texture, no flush"):
    if label.startswith("one"):
        t0 = wl.scene.items[0][2]
        wl.scene.items = [(p, uv, t0) for (p, uv, _) in wl.scene.items]
    wl.record(gpu, wl.subframes[0], None, t=0.0, static=True)
    gpu.Submit(gpu.Finish())
    fr = gpu.Retain()
    gpu.set_profiling(1)
    ts = []
    for i in range(8):
        if "flushed" in label:
            with torch.cuda.stream(stream):
                flush.fill_(i)
        gpu.Replay(fr, sync=True)
        ts.append(gpu.stage_ms()["tile"])
    gpu.set_profiling(0)
    gpu.Release(fr)
    print(f"{label}: tile kernel {1e3 * min(ts[2:]):.1f} us (median {1e3 * sorted(ts[2:])[len(ts[2:]) // 2]:.1f})")
