#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/attn_check.log; : > $LOG
run() { echo "-- $*" >> $LOG; timeout 120 python tools/kernel_selftest.py "$@" >> $LOG 2>&1; echo "rc=$?" >> $LOG; }
run attn 1 1 128 0.0 ${BWD:-1}
run attn 2 3 196 0.0 ${BWD:-1}
run attn 2 2 64 0.0 ${BWD:-1}
run attn 1 2 576 0.0 ${BWD:-1}
run attn 1 1 1024 0.0 ${BWD:-1}
run attn 3 2 208 0.0 ${BWD:-1}
run attn 2 1 209 0.0 ${BWD:-1}
run attn 2 3 196 0.1 1
run attn 1 2 576 0.1 1
cat $LOG
timeout 300 python tools/op_bench.py 256 attn 2>&1 | tee gpurun_out/op_bench_attn.log
