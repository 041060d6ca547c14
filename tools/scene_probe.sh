timeout 300 python -m pytest tests/test_dropin_gpu.py tests/test_scenes_gpu.py tests/test_cpp_host.py -m gpu -q 2>&1 | grep -v "^loading\|compiled in\|missing tileSize" | tail -8
for sc in tucker-and-dino colortest instanced-cubes; do
  timeout 100 python tools/scene_bench.py --lib dropin --scene $sc --seconds 2 2>/dev/null | tail -1
done
RSRCU_PIN_IN_PLACE=1 timeout 100 python tools/scene_bench.py --lib dropin --scene tucker-and-dino --seconds 2 2>/dev/null | tail -1
