#!/bin/bash
# Multi-GPU records of BASELINE.json's configurations on N GPUs of one box: bash tools/multi_gpu_records.sh N [what ...]
#   what in {train, curves, vitl, infer}; outputs gpurun_out/r2_<what>_<N>gpu.json (one JSON line each)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1; shift
WHAT="${*:-train curves vitl infer}"
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
run() { out=$1; shift; timeout 400 $R bench.py --gpus $N "$@" 2> gpurun_out/$out.err | grep "^{" > gpurun_out/$out.json; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/$out.json"))
    print("$out", round(d["value"], 1), d["unit"], round(d["ms_per_step"], 3), "ms/step", "exposed", d.get("allreduce_exposed_ms"), "dp_check", (d.get("dp_check") or {}).get("ok"))
except Exception as e:
    print("$out FAILED", e)
PY
}
for W in $WHAT; do
  case $W in
    train)  run r2_train_hilbert_${N}gpu --steps 10 --warmup 3 ;;
    curves) run r2_train_morton_${N}gpu --steps 10 --warmup 3 --curve morton --no-check-dp --no-exposed
            run r2_train_raster_${N}gpu --steps 10 --warmup 3 --curve raster --no-check-dp --no-exposed ;;
    vitl)   run r2_train_vitl_peano_${N}gpu --config vit_l16_384 --curve peano --steps 6 --warmup 3 --no-exposed
            run r2_train_vitl_hilbert_${N}gpu --config vit_l16_384 --curve hilbert --steps 6 --warmup 3 --no-check-dp --no-exposed ;;
    infer)  run r2_infer_1024_${N}gpu --config vit_b16_1024 --batch 16 --steps 5 --warmup 3
            run r2_infer_vits_${N}gpu --mode infer --config vit_s16_224 --steps 20 --warmup 3 ;;
  esac
done
