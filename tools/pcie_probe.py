#!/usr/bin/env python3
"""Measures pinned host<->device copy bandwidth at the frame size of the e2e leg (1920x1080x4 B) and at 256 MiB:
the D2H figure caps the end-to-end frame rate (one resolved frame per step crosses PCIe)."""
import json
import torch

dev = torch.device("cuda", 0)
out = {}
for name, nbytes in (("frame_8.3MB", 1920 * 1080 * 4), ("256MiB", 256 << 20)):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for direction in ("d2h", "h2d"):
        for _ in range(3):
            (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        out[f"{direction}_{name}"] = {"ms": round(ms, 4), "GB/s": round(nbytes / ms / 1e6, 2)}
print(json.dumps(out))
