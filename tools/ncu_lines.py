#!/usr/bin/env python3
"""warp instructions executed and stall samples per CUDA source line of a captured kernel (needs -lineinfo + --import-source on):
python tools/ncu_lines.py gpurun_out/tile_c4.ncu-rep [N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = "?"
acc = {}
tot_i = tot_s = 0
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        si, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) <= ie or not r[0]:
        continue
    try:
        line = int(r[0]); s = int(r[si]); n = int(r[ie])
    except ValueError:
        continue
    e = acc.setdefault((cur_file, line), [0, 0, r[1].strip()[:110]])
    e[0] += n; e[1] += s
    tot_i += n; tot_s += s
print(f"total warp instructions {tot_i}, samples {tot_s}")
for (f, l), (n, s, src) in sorted(acc.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{100 * n / tot_i:5.2f}% inst {100 * s / max(tot_s, 1):5.2f}% smp  {f}:{l:<5d} {src}")
