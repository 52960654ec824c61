#!/bin/bash
# round-1 re-entry check: GPU tests, kernel selftests of the fused LN-bwd/dropout/AdamW changes, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu2.log
LOG=gpurun_out/selftest2.log; : > $LOG
run() { echo "-- $*" >> $LOG; timeout 180 python tools/kernel_selftest.py "$@" >> $LOG 2>&1; echo "rc=$?" >> $LOG; }
run ln 1000 768
run ln 333 192
run ln 77 1024
run adamw
run attn 2 3 196 0.1 1
tail -20 $LOG
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench2.log 2>&1; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench2.log
