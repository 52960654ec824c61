#!/bin/bash
# compute-sanitizer pass over the kernel self-tests (ONE tool per gpurun call, as tools/B200 profiling recipe asks):
#   bash tools/sanitize.sh memcheck|racecheck|synccheck|initcheck      -> gpurun_out/r2_sanitizer_<tool>.log
# Small shapes: every kernel family (curves, GEMM epilogues + column sums, attention fwd/bwd with dropout, patch embed,
# LayerNorm, interp, optimizer, soft-target CE) runs once through the C ABI.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOL=${1:-racecheck}
python tools/sanitize_cases.py > gpurun_out/plain_sanitize.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_sanitize.log; exit 1; }
timeout ${SAN_TIMEOUT:-600} compute-sanitizer --tool $TOOL --print-limit 20 python tools/sanitize_cases.py > gpurun_out/r2_sanitizer_$TOOL.log 2>&1
echo "compute-sanitizer $TOOL rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SYNCCHECK SUMMARY|hazard|Invalid|out of bounds" gpurun_out/r2_sanitizer_$TOOL.log | head -20
