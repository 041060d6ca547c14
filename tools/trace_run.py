import os, sys
sys.path.insert(0, "/root/repo")
os.environ["RSRCU_TRACE"] = "1"
import numpy as np, torch, bench, rsr_b200
scene, size, workload = bench.make_scene(sys.argv[1])
W, H = size
gpu = rsr_b200.GPU(0)
host_out = [torch.empty((H, W), dtype=torch.int32).pin_memory().numpy().view(np.uint32) for _ in range(3)]
frames = []
for i in range(24):
    scene.record(gpu, size, host_out[i % 3], t=i / 60.0)
    frames.append(gpu.Finish())
for rec in frames[:2]:
    gpu.Submit(rec)
for i, rec in enumerate(frames[2:]):
    gpu.Submit(rec, sync=False)
    if i > 1: gpu.SyncFrame(2)
gpu.Sync()
gpu.close() if hasattr(gpu, "close") else None
del gpu
