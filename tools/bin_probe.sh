timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_scenes_gpu.py 2>&1 | tail -3
timeout 300 python bench.py --workload c3 --no-cpu-baseline --no-parity --steps 20 > gpurun_out/bin_probe_c3.json 2> gpurun_out/bin_probe_c3.err; tail -c 300 gpurun_out/bin_probe_c3.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bin_probe_c3.json") if l.startswith("{")][-1])
print("c3 fps", round(d["value"]), "ms", d["ms_per_step"], {k: round(v*1000) for k,v in d["roofline"]["stage_ms"].items()}, "flushed", d["flushed_frame"]["ms"])
PY
