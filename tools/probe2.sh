#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/probe2.log
: > $LOG
run() { echo "-- $*" >> $LOG; timeout 180 python tools/kernel_selftest.py "$@" >> $LOG 2>&1; echo "rc=$?" >> $LOG; }
run ln 1000 768
run ln 333 192
run ln 77 1024
run adamw
run attn 1 1 128 0.0 0
run attn 1 1 128 0.0 1
run attn 2 3 196 0.0 1
run attn 2 2 64 0.0 1
run attn 1 2 576 0.0 1
run attn 2 3 196 0.1 1
run patch 3 3 224 16 1 768 hilbert fp32
run patch 3 3 224 16 1 384 hilbert bf16
run patch 5 3 32 4 1 192 z fp32
run patch 5 3 32 1 16 256 hilbert fp32
run patch 5 3 32 2 4 256 peano fp32
run patch 2 3 64 8 2 128 moore bf16
tail -70 $LOG
