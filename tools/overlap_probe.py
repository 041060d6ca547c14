#!/usr/bin/env python3
"""does a pinned D2H copy overlap a running kernel on this box? (events on the copy stream)"""
import torch
dev = torch.device("cuda", 0)
x = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
h = torch.empty(1920 * 1080 * 4, dtype=torch.uint8).pin_memory()
d = torch.empty(1920 * 1080 * 4, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
for label, busy in (("idle GPU", False), ("GPU busy with a 20 ms matmul chain", True)):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if busy:
        with torch.cuda.stream(s1):
            for _ in range(16):
                y = x @ x
    with torch.cuda.stream(s2):
        e0.record(s2)
        h.copy_(d, non_blocking=True)
        e1.record(s2)
    torch.cuda.synchronize()
    print(label, "copy ms", round(e0.elapsed_time(e1), 3))
print("asyncEngineCount", torch.cuda.get_device_properties(0).__repr__())
