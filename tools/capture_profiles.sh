#!/usr/bin/env bash
# Runs ON the GPU box (via gpurun): ncu launch lists + one --set full capture of the dominant kernel per
# workload into gpurun_out/.  Summaries are then written locally with tools/ncu_summary.py into profiles/.
set -u
bash tools/capture_launches.sh
mkdir -p gpurun_out
for w in c2 c4 c3; do
  ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 2 -c 1 -o gpurun_out/tile_$w \
      python tools/run_frames.py $w 4 > gpurun_out/tile_$w.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:"setup_kernel|fill_kernel" -s 4 -c 2 -o gpurun_out/bin_c3 \
    python tools/run_frames.py c3 4 > gpurun_out/bin_c3.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > gpurun_out/smi.csv
