#!/usr/bin/env python3
"""e2e pipeline probe: frames/s of Submit+SyncFrame loops with and without the D2H readback"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, rsr_b200

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
wl = bench.Workload(name)
scene, size, workload = wl.scene, wl.sub_size, wl.name
W, H = size
gpu = rsr_b200.GPU(0)
host_out = [torch.empty((H, W), dtype=torch.int32).pin_memory().numpy().view(np.uint32) for _ in range(3)]
for label, outs in (("readback", host_out), ("device-only", [None, None, None])):
    frames = []
    for i in range(62):
        scene.record(gpu, size, outs[i % 3], t=i / 60.0, static=(label == "static"))
        frames.append(gpu.Finish())
    for rec in frames[:2]:
        gpu.Submit(rec)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sub = []
    for i, rec in enumerate(frames[2:]):
        a = time.perf_counter()
        gpu.Submit(rec, sync=False)
        b = time.perf_counter()
        if i > 1:
            gpu.SyncFrame(2)
        sub.append((b - a, time.perf_counter() - b))
    gpu.Sync()
    dt = time.perf_counter() - t0
    print(workload, label, "fps %.0f" % (60 / dt), "us/frame %.0f" % (dt / 60 * 1e6), "submit_us %.0f" % (1e6 * np.median([s[0] for s in sub])),
          "syncframe_us %.0f" % (1e6 * np.median([s[1] for s in sub])), gpu.stats()["host_record_ns"] // 1000, gpu.stats()["host_submit_ns"] // 1000)
