#!/usr/bin/env bash
# Runs ON an N-GPU box: what bounds the end-to-end leg when N ranks share one host?  (a) pinned D2H bandwidth of every GPU
# at the same time, (b) the e2e loop with and without the read-back on every GPU at the same time, each pinned to its cores.
N=${1:-8}
CORES=$(nproc)
PER=$((CORES / N))
echo "== $N GPUs, $CORES cores, $PER per rank"
echo "== concurrent pinned copies (8.3 MB frames / 256 MiB), per GPU"
for i in $(seq 0 $((N - 1))); do
  CUDA_VISIBLE_DEVICES=$i taskset -c $((i * PER))-$((i * PER + PER - 1)) python tools/pcie_probe.py > /tmp/pcie_$i.log 2>&1 &
done
wait
for i in $(seq 0 $((N - 1))); do echo "gpu $i $(tail -1 /tmp/pcie_$i.log)"; done
echo "== concurrent e2e loops (read-back / device-only), per GPU"
for i in $(seq 0 $((N - 1))); do
  CUDA_VISIBLE_DEVICES=$i taskset -c $((i * PER))-$((i * PER + PER - 1)) python tools/e2e_probe.py c2 > /tmp/e2e_$i.log 2>&1 &
done
wait
for i in $(seq 0 $((N - 1))); do echo "gpu $i"; tail -2 /tmp/e2e_$i.log; done
echo "== one GPU alone, same pinning"
CUDA_VISIBLE_DEVICES=0 taskset -c 0-$((PER - 1)) python tools/e2e_probe.py c2 2>&1 | tail -2
nvidia-smi topo -m | head -14
