#!/usr/bin/env bash
# Runs ON the GPU box: parity tests, then the three single-GPU workloads (one JSON line each).
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in c2 c3 c4; do
  python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$w.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$w", "fps %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], "frac %.3f" % r["frac"], {k: round(v * 1000) for k, v in r["stage_ms"].items()})
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/bench_$w.err").read()[-2000:])
PY
done
