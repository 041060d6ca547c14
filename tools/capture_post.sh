set -u
timeout 300 python -m pytest tests/test_post_gpu.py tests/test_march_gpu.py -m gpu -q 2>&1 | tail -15
mkdir -p gpurun_out
timeout 420 python bench.py --workload f_rows > gpurun_out/r2b_bench_c2.json 2> gpurun_out/r2b_bench_c2.err
tail -c 600 gpurun_out/r2b_bench_c2.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2b_bench_c2.json") if l.startswith("{")][-1])

print(json.dumps(d.get("f_rows"), indent=1))
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_post.csv python tools/run_post.py 2 > gpurun_out/post_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"kawase_kernel|glow_kernel|march_kernel" -s 3 -c 4 -o gpurun_out/post_kernels python tools/run_post.py 3 > gpurun_out/post_full.log 2>&1
tail -3 gpurun_out/post_full.log
