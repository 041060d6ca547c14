# Runs ON the GPU box: sanitizers over the small-frame script, the whole GPU suite, the default bench line (N = 1)
set -u
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/san_small_memcheck.log 2>&1; grep -E "ERROR SUMMARY|done" gpurun_out/san_small_memcheck.log | tail -2
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/san_small_racecheck.log 2>&1; grep -E "RACECHECK SUMMARY|done" gpurun_out/san_small_racecheck.log | tail -2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^loading\|compiled in\|missing tileSize" | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_final_bench_c2.json 2> gpurun_out/r2_final_bench_c2.err; tail -c 300 gpurun_out/r2_final_bench_c2.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_final_bench_reference_c2.json 2>/dev/null
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_final_bench_c2.json") if l.startswith("{")][-1])
r=json.loads([l for l in open("gpurun_out/r2_final_bench_reference_c2.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ref", round(r["value"],1), "ratio", round(d["value"]/r["value"],1), "e2e ratio", round(d["e2e"]["value"]/r["value"],1))
print("parity", d["parity"]["diff_pixels"], "c4_4k frac", d["c4_4k"]["roofline"]["frac"], "c3_4k fps", d["c3_4k"]["frames_per_s"])
print(json.dumps(d["bundled_scenes"]["scenes"]))
PY
