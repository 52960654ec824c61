#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
python bench.py --steps 5 --warmup 3 --batch 64 --no-cpu-baseline > gpurun_out/bench_b64.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/bench_b64.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_b256.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/bench_b256.log
