#!/usr/bin/env python3
"""timeline of the sub-frames of an 8K frame replayed on ONE GPU with frame overlap (RSRCU_TRACE=2): front end and tile
kernel start / end per unit.  python tools/c5_trace.py"""
import os
import sys

os.environ["RSRCU_TRACE"] = "2"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import rsr_b200  # noqa: E402
from rsr_b200 import scenes  # noqa: E402
from rsr_b200.subframes import SubframePlan  # noqa: E402

scene = scenes.BundledLikeScene()
gpu = rsr_b200.GPU(0)
gpu.set_overlap(True)
frame = torch.zeros((4320, 7680), dtype=torch.int32, device="cuda:0")
P = scenes.perspective(45.0, 7680 / 4320, 1.0, 400.0)
plan = SubframePlan(7680, 4320, 1, 1920, 1080)
retained = []
for sf in plan.subframes[:8]:
    scene.record(gpu, (sf.width, sf.height), None, t=0.0, static=True, proj=plan.projection(P, sf),
                 device_out=(frame.data_ptr() + 4 * (sf.y0 * 7680 + sf.x0), 7680))
    gpu.Submit(gpu.Finish())
    retained.append(gpu.Retain())
# trace frames 0..7 were the set-up submits; 8.. are replays
for rep in range(5):
    for fr in retained:
        gpu.Replay(fr)
gpu.Sync()
gpu.close()
