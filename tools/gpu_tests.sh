#!/bin/bash
# Runs each GPU test file in its own process (a CUDA fault in one cannot poison the others).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in "$@"; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q -s --no-header -p no:cacheprovider > gpurun_out/$b.log 2>&1
  echo "$b exit=$?" | tee -a gpurun_out/summary.txt
  tail -5 gpurun_out/$b.log
done
