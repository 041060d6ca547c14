#!/usr/bin/env python3
"""end-to-end frames/s with K submission threads, each with its own context on the same GPU (frames dealt round robin):
python tools/e2e_threads.py c2 [threads] [frames]"""
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
import rsr_b200  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
nthreads = int(sys.argv[2]) if len(sys.argv) > 2 else 2
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 240
wl = bench.Workload(name)
W, H = wl.sub_size
per = frames // nthreads
ctxs = []
for k in range(nthreads):
    gpu = rsr_b200.GPU(0)
    gpu.set_overlap(True)
    outs = [torch.empty((H, W), dtype=torch.int32).pin_memory().numpy().view(np.uint32) for _ in range(3)]
    recs = []
    for i in range(min(per + 2, 33)):
        wl.record(gpu, wl.subframes[0], outs[i % 3], t=(i * nthreads + k) / 60.0)
        recs.append(gpu.Finish())
    order = [recs[i % (len(recs) // 3 * 3)] for i in range(per + 2)]
    for rec in order[:2]:
        gpu.Submit(rec)
    ctxs.append((gpu, order, outs))
start = threading.Barrier(nthreads + 1)


def worker(k):
    gpu, order, _ = ctxs[k]
    start.wait()
    for i, rec in enumerate(order[2:]):
        gpu.Submit(rec, sync=False)
        if i > 1:
            gpu.SyncFrame(2)
    gpu.Sync()


ths = [threading.Thread(target=worker, args=(k,)) for k in range(nthreads)]
for t in ths:
    t.start()
torch.cuda.synchronize()
start.wait()
t0 = time.perf_counter()
for t in ths:
    t.join()
dt = time.perf_counter() - t0
print(f"{wl.name}: {nthreads} submission threads, {per * nthreads} frames: {per * nthreads / dt:.0f} frames/s ({1e6 * dt / (per * nthreads):.0f} us/frame)")
