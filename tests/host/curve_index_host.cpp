// Test-only host build of csrc/curve_index.h: lets the CPU test-suite check the exact
// integer routines the CUDA kernel runs against the oracle without a GPU.
#include "curve_index.h"
extern "C" long long sfc_host_perm(int curve, int w, int h, long long* out_flat) {
  int64_t P; const int m = w > h ? w : h;
  const int order = sfc_order_for(curve, m, &P);
  long long cnt = 0;
  for (uint64_t d = 0; d < (uint64_t)(P * P); ++d) {
    int i, j; sfc_d2ij(curve, order, P, d, &i, &j);
    if (i >= 0 && i < w && j >= 0 && j < h) out_flat[cnt++] = (long long)i * h + j;
  }
  return cnt;
}
