#!/usr/bin/env bash
# run locally after tools/capture_profiles.sh: gpurun_out/*.ncu-rep, launches_*.csv -> profiles/r1_*
set -eu
R=${1:-r1}
mkdir -p profiles
for w in c2 c3 c4; do
  cp gpurun_out/launches_$w.csv profiles/${R}_launches_$w.csv
  python tools/ncu_summary.py gpurun_out/tile_$w.ncu-rep > profiles/${R}_tile_kernel_$w.txt
done
python tools/ncu_summary.py gpurun_out/bin_c3.ncu-rep > profiles/${R}_bin_kernel_c3.txt
cp gpurun_out/smi.csv profiles/${R}_smi.csv
