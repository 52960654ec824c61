#!/bin/bash
# round-1 final profiles: launch list of one eager step + ncu --set full of the GEMM family and the attention kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph"
$CMD > gpurun_out/plain_r1c.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 1200 --csv --log-file gpurun_out/launches_r1c.csv $CMD > gpurun_out/ncu_r1c.log 2>&1
echo "ncu launches rc=$?"
$CMD > gpurun_out/plain_r1d.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 170 -c 13 -o gpurun_out/prof_gemm_r1_final $CMD > gpurun_out/ncu_r1d.log 2>&1
echo "ncu gemm rc=$?"
$CMD > gpurun_out/plain_r1e.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"attn_|patch_embed_fwd|layernorm" -s 30 -c 6 -o gpurun_out/prof_misc_r1_final $CMD > gpurun_out/ncu_r1e.log 2>&1
echo "ncu misc rc=$?"
python tools/launch_summary.py gpurun_out/launches_r1c.csv | head -24
