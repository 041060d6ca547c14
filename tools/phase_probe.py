#!/usr/bin/env python3
"""where a tile CTA's time goes (developer build with RSR_PHASE_PROF): python tools/phase_probe.py c2 [c4 ...]
phases: 0 prologue before the grid-dependency wait, 1 tile offsets + large-item scan, 2 load_chunk (list fetch + sort),
3 entry record fetch + wait for the other warps, 4 head publish, 5 clear / store commands, 6 batch cut (firstBad),
7 triangle setup, (7->8) draw_batch, 8 loop bookkeeping, 9 fragment count epilogue"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RSRCU_LIB"] = os.path.join(ROOT, "rsr_b200", "variants", "librsrcu_phase.so")
import bench  # noqa: E402
import rsr_b200  # noqa: E402

NAMES = ["prologue(pre-wait)", "tile offsets+large scan", "load_chunk", "entry fetch + warp wait", "head publish", "clear/store cmds",
         "batch cut + setup", "draw_batch (warp 0)", "loop end", "epilogue (stores of the last commands incl.)", "-"]
for name in sys.argv[1:] or ["c2"]:
    wl = bench.Workload(name)
    scene, size, workload = wl.scene, wl.sub_size, wl.name
    gpu = rsr_b200.GPU(0)
    scene.record(gpu, size, None, t=0.0, static=True)
    rec = gpu.Finish()
    for _ in range(3):
        gpu.Submit(rec)
    out = (C.c_ulonglong * 16)()
    gpu.L.rsrcu_debug_phase_cycles(out)
    n = 10
    for _ in range(n):
        gpu.Submit(rec)
    gpu.L.rsrcu_debug_phase_cycles(out)
    tot = sum(out) or 1
    tiles = ((size[0] + 31) // 32) * ((size[1] + 31) // 32)
    print(f"{workload}: {tot / n / tiles:.0f} cycles per tile CTA (thread 0)")
    for k, nm in enumerate(NAMES):
        print(f"   {nm:28s} {100 * out[k] / tot:5.1f}%  {out[k] / n / tiles:8.0f} cycles/CTA")
    k2 = (C.c_ulonglong * 16)()
    gpu.L.rsrcu_debug_k2_times(k2)
    gpu.Submit(rec)
    gpu.L.rsrcu_debug_k2_times(k2)
    print(f"   setup_kernel: first CTA start -> last CTA's ticket {(k2[1] - k2[0]) / 1e3:.1f} us, scan {(k2[2] - k2[1]) / 1e3:.1f} us, tile order {(k2[3] - k2[2]) / 1e3:.1f} us; last-indexed CTA started {(k2[4] - k2[0]) / 1e3:.1f} us after the first")
    print('   setup_kernel, slowest CTA per phase (thread 0), us: indices loaded, classified (+clip), counted, fenced, barrier (other warps):', [round(k2[8 + i] / 1e3, 1) for i in range(1, 6)])
    gpu.close()
