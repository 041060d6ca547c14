#!/usr/bin/env python3
"""top SASS instructions of a captured kernel by warp-stall samples, with the dominant stall reason:
python tools/ncu_hot.py gpurun_out/tile_c4.ncu-rep [N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if len(r) > 3 and r[0] == "Address")
hdr = rows[hi]
si, ie, te = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for k, r in enumerate(rows[hi + 1:]):
    try:
        data.append((int(r[si]), int(r[ie]), int(r[te]), k, r))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data) or 1
toti = sum(d[1] for d in data) or 1
print(f"{len(data)} SASS instructions, {tot} samples, {toti} warp instructions executed")
agg = {}
for s, n, t, k, r in data:
    for i, h in stall_cols:
        try:
            agg[h] = agg.get(h, 0) + int(r[i])
        except ValueError:
            pass
print("stall totals:", ", ".join(f"{h[6:]} {100 * v / tot:.1f}%" for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
ops = {}
for s, n, t, k, r in data:
    op = r[1].split()[0] if not r[1].strip().startswith("@") else r[1].split()[1]
    op = op.split(".")[0]
    e = ops.setdefault(op, [0, 0])
    e[0] += n; e[1] += s
print("opcode mix (warp inst %, samples %):", ", ".join(f"{o} {100 * v[0] / toti:.1f}/{100 * v[1] / tot:.1f}" for o, v in sorted(ops.items(), key=lambda x: -x[1][0])[:24]))
for s, n, t, k, r in sorted(data, reverse=True)[:topn]:
    best = max(stall_cols, key=lambda c: int(r[c[0]] or 0))
    print(f"{100 * s / tot:5.2f}%  #{k:5d} exec {n:9d} thr/inst {t / max(n, 1):4.1f}  {best[1][6:]:12s} {r[1].strip()[:90]}")
