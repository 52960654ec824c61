#!/bin/bash
# Last validation of a round on one B200: full GPU test suite, smoke, headline bench, ViT-L training bench, ncu launch
# list of one eager headline step. Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/fc_tests.log; cat $O/fc_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > $O/fc_bench_train.json 2> $O/fc_bench_train.err; cut -c1-400 $O/fc_bench_train.json
timeout 200 python bench.py --config vit_l16_384 --steps 8 --no-cpu-baseline > $O/fc_train_l.json 2> $O/fc_train_l.err; cut -c1-200 $O/fc_train_l.json
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 1200 --csv --log-file $O/fc_step_launches.csv $CMD > $O/ncu_launches.log 2>&1
python tools/launch_summary.py $O/fc_step_launches.csv > $O/fc_step_summary.txt; head -24 $O/fc_step_summary.txt
