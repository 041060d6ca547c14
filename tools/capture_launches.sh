#!/usr/bin/env bash
# Runs ON the GPU box: ncu launch lists (per-kernel device time of two whole frames) for the three single-GPU workloads.
set -u
mkdir -p gpurun_out
for w in c2 c4 c3; do
  per=5; [ $w = c3 ] && per=6
  ncu --metrics gpu__time_duration.sum --clock-control none -s $((per * 3)) -c $((per * 2)) --csv --log-file gpurun_out/launches_$w.csv \
      python tools/run_frames.py $w 6 > gpurun_out/launches_$w.log 2>&1
done
