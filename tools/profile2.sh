#!/bin/bash
# launch list of one step (after the plain run exits 0), per B200_PROFILING.md
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_r1b.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 1200 --csv --log-file gpurun_out/launches_r1b.csv $CMD > gpurun_out/ncu_r1b.log 2>&1
echo "ncu launches rc=$?"
python tools/launch_summary.py gpurun_out/launches_r1b.csv | head -40
