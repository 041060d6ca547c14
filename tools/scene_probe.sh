timeout 300 python -m pytest tests/test_scenes_gpu.py -m gpu -q 2>&1 | grep -v "^loading\|compiled in\|missing tileSize" | tail -8
for th in 16 4; do
  timeout 100 python tools/scene_bench.py --lib dropin --scene instanced-cubes --seconds 2 --threads $th 2>/dev/null | tail -1
done
timeout 100 python tools/scene_bench.py --lib dropin --scene instanced-cubes --seconds 2 --threads 16 --static-assets 2>/dev/null | tail -1
timeout 100 python tools/scene_bench.py --lib ref --scene instanced-cubes --seconds 2 2>/dev/null | tail -1
