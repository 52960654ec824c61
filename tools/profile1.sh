#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 1400 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 160 -c 12 -o gpurun_out/prof_gemm_r1 $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu gemm rc=$?"
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_ -s 12 -c 4 -o gpurun_out/prof_attn_r1 $CMD > gpurun_out/ncu3.log 2>&1
echo "ncu attn rc=$?"
ls -la gpurun_out | tail -12
