"""e2e frame timeline (RSRCU_TRACE): python tools/trace_run.py c2 [overlap]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["RSRCU_TRACE"] = "1"
import numpy as np, torch, bench, rsr_b200
scene, size, workload = bench.make_scene(sys.argv[1] if len(sys.argv) > 1 else "c2")
W, H = size
gpu = rsr_b200.GPU(0)
gpu.set_overlap(len(sys.argv) > 2 and sys.argv[2] == "overlap")
host_out = [torch.empty((H, W), dtype=torch.int32).pin_memory().numpy().view(np.uint32) for _ in range(3)]
frames = []
for i in range(24):
    scene.record(gpu, size, host_out[i % 3], t=i / 60.0)
    frames.append(gpu.Finish())
for rec in frames[:2]:
    gpu.Submit(rec)
host = []
t00 = time.perf_counter()
for i, rec in enumerate(frames[2:]):
    t0 = time.perf_counter()
    gpu.Submit(rec, sync=False)
    t1 = time.perf_counter()
    if i > 1: gpu.SyncFrame(2)
    host.append((t1 - t0, time.perf_counter() - t1))
gpu.Sync()
tot = time.perf_counter() - t00
print("frames/s", 22 / tot, "host submit us (median)", 1e6 * float(np.median([h[0] for h in host])), "sync wait us (median)", 1e6 * float(np.median([h[1] for h in host])))
st = gpu.stats()
print({k: st[k] for k in ("host_record_ns", "host_submit_ns", "h2d_bytes", "d2h_bytes")})
gpu.close()
