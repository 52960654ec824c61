#!/bin/bash
# first GPU probe: curves + GEMM variants, each GEMM variant in its own process
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/probe1.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv >> $LOG 2>&1
echo "== curves" >> $LOG
timeout 600 python -m pytest tests/test_gpu_curves.py -m gpu -q -x 2>&1 | tail -15 >> $LOG
echo "== gemm variants" >> $LOG
for v in "256 256 128 0 0" "256 256 128 0 1" "256 256 128 1 1" "256 256 128 1 0" \
         "128 128 64 0 0" "128 128 64 0 1" "128 128 64 1 1" \
         "392 768 768 0 0" "1000 2304 768 0 0" "1000 768 2304 0 1" "768 2304 1000 1 1" \
         "392 768 256 0 0 bias_relu" "392 768 256 0 0 bias_res" "768 256 4096 1 1 fp32 4"; do
  echo "-- $v" >> $LOG
  timeout 120 python tools/gemm_selftest.py $v >> $LOG 2>&1
  echo "rc=$?" >> $LOG
done
tail -80 $LOG
