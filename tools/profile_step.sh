#!/bin/bash
# round-1 profiles: ncu launch list of one eager step (+ per-kernel summary) and ncu --set full captures of the GEMM
# family (PASS gemm) and of the non-GEMM kernels (PASS misc). Reports are summarised to text and deleted (gpurun copies
# back at most 64 MiB); set KEEP_REP=1 to keep them. usage: bash tools/profile_step.sh [launches] [gemm] [misc]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PASSES="${*:-launches gemm misc}"
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph"
for P in $PASSES; do
  $CMD > gpurun_out/plain_$P.log 2>&1 || { echo "plain run failed before pass $P"; exit 1; }
  case $P in
    launches)
      ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 1200 --csv --log-file gpurun_out/${PFX:-r1}_step_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
      echo "ncu launches rc=$?"
      python tools/launch_summary.py gpurun_out/${PFX:-r1}_step_launches.csv > gpurun_out/${PFX:-r1}_step_summary.txt; head -32 gpurun_out/${PFX:-r1}_step_summary.txt ;;
    gemm)
      ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 170 -c 13 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
      echo "ncu gemm rc=$?"
      python tools/ncu_summary.py gpurun_out/prof_gemm.ncu-rep > gpurun_out/${PFX:-r1}_gemm_full_capture.txt; cat gpurun_out/${PFX:-r1}_gemm_full_capture.txt
      [ -n "$KEEP_REP" ] || rm -f gpurun_out/prof_gemm.ncu-rep ;;
    misc)
      # first 22 non-GEMM launches of the 4th step: patch embed, LN / attention forward of the first layers ...
      ncu --set full --clock-control none --import-source on -k regex:"attn_|patch_embed_|patch_gather|layernorm|colsum_partial|adamw|softce|sumsq" -s 381 -c 22 -o gpurun_out/prof_misc_fwd $CMD > gpurun_out/ncu_misc_fwd.log 2>&1
      echo "ncu misc fwd rc=$?"
      python tools/ncu_summary.py gpurun_out/prof_misc_fwd.ncu-rep > gpurun_out/${PFX:-r1}_misc_full_capture.txt
      [ -n "$KEEP_REP" ] || rm -f gpurun_out/prof_misc_fwd.ncu-rep
      # ... and the tail of its backward + optimizer
      ncu --set full --clock-control none --import-source on -k regex:"attn_|patch_embed_|patch_gather|layernorm|colsum_partial|adamw|softce|sumsq" -s 440 -c 36 -o gpurun_out/prof_misc_bwd $CMD > gpurun_out/ncu_misc_bwd.log 2>&1
      echo "ncu misc bwd rc=$?"
      python tools/ncu_summary.py gpurun_out/prof_misc_bwd.ncu-rep | tail -n +2 >> gpurun_out/${PFX:-r1}_misc_full_capture.txt
      [ -n "$KEEP_REP" ] || rm -f gpurun_out/prof_misc_bwd.ncu-rep
      cat gpurun_out/${PFX:-r1}_misc_full_capture.txt ;;
  esac
done
