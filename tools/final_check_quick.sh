# Runs ON the GPU box: the whole GPU suite, smoke(), the default bench line (N = 1) and the reference arm
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^loading\|compiled in\|missing tileSize" | tail -4
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_final_bench_reference_c2.json 2>/dev/null
( time timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_final_bench_c2.json 2> gpurun_out/r2_final_bench_c2.err ) 2>&1 | grep real; tail -c 300 gpurun_out/r2_final_bench_c2.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_final_bench_c2.json") if l.startswith("{")][-1])
r=json.loads([l for l in open("gpurun_out/r2_final_bench_reference_c2.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"].get("repetitions_frames_per_s"), "ref", round(r["value"],1), "ratio", round(d["value"]/r["value"],1), "e2e ratio", round(d["e2e"]["value"]/r["value"],1))
print("parity", d["parity"]["diff_pixels"], "c4_4k frac", d["c4_4k"]["roofline"]["frac"], "c3_4k fps", d["c3_4k"]["frames_per_s"])
PY
