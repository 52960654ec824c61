// K1 — curve permutation kernel (integer, sm_100a).
// Replaces the host recursion behind embed_and_prune_sfc
// (/root/reference/src/curves/space_filling_curves.py:471-491) and the flat-index step of the
// tokenizers (multi_hilbert.py:68-72): one thread per position d of the padded P x P curve computes
// its cell by integer descent (curve_index.h), flags in-domain cells, and a two-pass block scan
// compacts them in curve order into perm[rank] = i*h + j and inv[i*h + j] = rank.
#include "common.cuh"
#include "curve_index.h"
#include "sfcvit.h"

namespace {

constexpr int kThreads = 256;
constexpr int kItems = 8;                      // consecutive curve positions per thread
constexpr int kChunk = kThreads * kItems;      // positions per block

__device__ __forceinline__ int block_exclusive_scan(int v, int* total, int* smem /*[kThreads/32]*/) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) smem[wid] = incl;
  __syncthreads();
  int wsum = (lane < kThreads / 32) ? smem[lane] : 0;
  int wincl = wsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, wincl, o);
    if (lane >= o) wincl += t;
  }
  const int wexcl = __shfl_sync(0xffffffffu, wincl - wsum, wid);
  *total = __shfl_sync(0xffffffffu, wincl, kThreads / 32 - 1);
  __syncthreads();
  return wexcl + incl - v;
}

// pass 1: per-block number of in-domain cells
__global__ void __launch_bounds__(kThreads) curve_count_kernel(int curve, int order, long long P, int w, int h,
                                                               unsigned long long total, int* __restrict__ block_counts) {
  __shared__ int red[kThreads / 32];
  const unsigned long long base = (unsigned long long)blockIdx.x * kChunk + (unsigned long long)threadIdx.x * kItems;
  int c = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const unsigned long long d = base + k;
    if (d < total) {
      int i, j;
      sfc_d2ij(curve, order, P, d, &i, &j);
      c += (i < w && j < h) ? 1 : 0;
    }
  }
  int tot;
  block_exclusive_scan(c, &tot, red);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = tot;
}

// pass 2: block offset = sum of preceding block counts; recompute cells; scatter perm / inv
__global__ void __launch_bounds__(kThreads) curve_emit_kernel(int curve, int order, long long P, int w, int h,
                                                              unsigned long long total, const int* __restrict__ block_counts,
                                                              int* __restrict__ perm, int* __restrict__ inv) {
  __shared__ int red[kThreads / 32];
  __shared__ int s_off;
  int part = 0;
  for (int b = threadIdx.x; b < (int)blockIdx.x; b += kThreads) part += block_counts[b];
  int boff;
  block_exclusive_scan(part, &boff, red);
  if (threadIdx.x == 0) s_off = boff;
  __syncthreads();
  const int block_off = s_off;

  const unsigned long long base = (unsigned long long)blockIdx.x * kChunk + (unsigned long long)threadIdx.x * kItems;
  int flat[kItems];
  int c = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const unsigned long long d = base + k;
    flat[k] = -1;
    if (d < total) {
      int i, j;
      sfc_d2ij(curve, order, P, d, &i, &j);
      if (i < w && j < h) { flat[k] = i * h + j; ++c; }
    }
  }
  int tot;
  int rank = block_off + block_exclusive_scan(c, &tot, red);
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    if (flat[k] >= 0) {
      perm[rank] = flat[k];
      if (inv) inv[flat[k]] = rank;
      ++rank;
    }
  }
}

}  // namespace

extern "C" size_t sfc_curve_perm_scratch_bytes(int curve_id, int w, int h) {
  int64_t P;
  sfc_order_for(curve_id, w > h ? w : h, &P);
  const int64_t blocks = sfc_ceil_div64(P * P, kChunk);
  return (size_t)blocks * sizeof(int);
}

extern "C" int sfc_curve_perm(int curve_id, int w, int h, int32_t* perm_dev, int32_t* inv_dev, void* scratch_dev,
                              size_t scratch_bytes, cudaStream_t stream) {
  SFC_REQUIRE(curve_id >= SFC_HILBERT && curve_id <= SFC_RASTER, "sfc_curve_perm: unknown curve id %d", curve_id);
  SFC_REQUIRE(w >= 1 && h >= 1 && (int64_t)w * h < (1ll << 31), "sfc_curve_perm: bad grid %dx%d", w, h);
  SFC_REQUIRE(perm_dev != nullptr, "sfc_curve_perm: perm_dev is null");
  int64_t P;
  const int order = sfc_order_for(curve_id, w > h ? w : h, &P);
  SFC_REQUIRE(P <= 65536, "sfc_curve_perm: padded side %lld too large", (long long)P);
  const unsigned long long total = (unsigned long long)P * (unsigned long long)P;
  const int64_t blocks = sfc_ceil_div64((int64_t)total, kChunk);
  SFC_REQUIRE(scratch_dev != nullptr && scratch_bytes >= (size_t)blocks * sizeof(int),
              "sfc_curve_perm: scratch too small (%zu < %zu)", scratch_bytes, (size_t)blocks * sizeof(int));
  int* counts = static_cast<int*>(scratch_dev);
  curve_count_kernel<<<(unsigned)blocks, kThreads, 0, stream>>>(curve_id, order, P, w, h, total, counts);
  SFC_LAUNCH_OK();
  curve_emit_kernel<<<(unsigned)blocks, kThreads, 0, stream>>>(curve_id, order, P, w, h, total, counts, perm_dev, inv_dev);
  SFC_LAUNCH_OK();
  return 0;
}
