#!/bin/bash
# GEMM / patch-embed epilogue check: correctness tests then per-op timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/op_bench.py 256 ${1:-gemm patch} > gpurun_out/op_bench.log 2>&1; echo "bench rc=$?"; cat gpurun_out/op_bench.log
