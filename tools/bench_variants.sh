#!/usr/bin/env bash
# Runs ON the GPU box: bench each tuning variant (rsr_b200/variants/*.so) on the given workloads.
shopt -s nullglob
for lib in default rsr_b200/variants/*.so; do
  for w in "$@"; do
    if [ "$lib" = default ]; then unset RSRCU_LIB; else export RSRCU_LIB=$PWD/$lib; fi
    python bench.py --workload $w --no-cpu-baseline --no-parity --no-subrecords --e2e-frames 30 --sustained-seconds 0 --steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$lib','$w','fps %.0f'%d['value'],'frac %.3f'%r['frac'],{k:round(v*1000) for k,v in r['stage_ms'].items()})"
  done
done
