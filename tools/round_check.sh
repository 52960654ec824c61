#!/bin/bash
# Round-end check on one B200: full GPU test suite, smoke, headline bench (+ reference arm), inference benches of the
# other BASELINE configs, ncu launch lists. Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/rc_tests.log; cat $O/rc_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > $O/rc_bench_train.json 2> $O/rc_bench_train.err; cut -c1-300 $O/rc_bench_train.json
timeout 200 python bench.py --mode infer --config vit_s16_224 --steps 20 > $O/rc_infer_s.json 2> $O/rc_infer_s.err
timeout 300 python bench.py --config vit_b16_1024 --steps 10 > $O/rc_infer_x4.json 2> $O/rc_infer_x4.err
timeout 200 python bench.py --config vit_b16_1024 --batch 16 --steps 5 --no-cpu-baseline > $O/rc_infer_x16.json 2> $O/rc_infer_x16.err
timeout 200 python bench.py --config vit_l16_384 --steps 8 --no-cpu-baseline > $O/rc_train_l.json 2> $O/rc_train_l.err
for f in rc_infer_s rc_infer_x4 rc_infer_x16 rc_train_l; do cut -c1-200 $O/$f.json; done
if [ -n "$NCU" ]; then
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph"
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 1200 --csv --log-file $O/rc_step_launches.csv $CMD > $O/ncu_launches.log 2>&1
  python tools/launch_summary.py $O/rc_step_launches.csv > $O/rc_step_summary.txt; head -24 $O/rc_step_summary.txt
  CMDX="python bench.py --config vit_b16_1024 --steps 1 --warmup 3 --no-cpu-baseline --no-graph"
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file $O/rc_infer_x_launches.csv $CMDX > $O/ncu_launches_x.log 2>&1
  python tools/launch_summary.py $O/rc_infer_x_launches.csv > $O/rc_infer_x_summary.txt; head -12 $O/rc_infer_x_summary.txt
fi
