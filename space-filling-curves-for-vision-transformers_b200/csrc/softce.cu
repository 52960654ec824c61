// Soft-target cross entropy, forward and backward (sm_100a, memory-bound, tiny).
// Replaces the criterion the reference defines in /root/reference/main.py:45-51
//     loss = -(targets * log_softmax(inputs.float(), dim=-1)).sum(dim=-1).mean()
// and its autograd (six ATen launches per step: cast, log_softmax, mul, sum, mean, neg — plus their backward) by one
// kernel each way. One warp per row; the batch mean is reduced in a fixed order by the last block to finish (ticket),
// so the loss is bit-identical from run to run.
//   forward : lse[b] = logsumexp(x[b, :]), tsum[b] = sum_c t[b, c], loss = mean_b (lse[b] * tsum[b] - sum_c t[b, c] x[b, c])
//   backward: dx[b, c] = dloss / B * (exp(x[b, c] - lse[b]) * tsum[b] - t[b, c])
#include "common.cuh"
#include "sfcvit.h"

namespace {

constexpr int kWarps = 8;

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename T>
__global__ void __launch_bounds__(kWarps * 32) softce_fwd_kernel(const T* __restrict__ x, long long ldx, const float* __restrict__ t,
                                                                  long long ldt, int B, int C, float* __restrict__ loss,
                                                                  float* __restrict__ row_lse, float* __restrict__ row_tsum,
                                                                  float* __restrict__ partials, unsigned int* __restrict__ ticket) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ float red[kWarps];
  float block_loss = 0.f;
  for (int b = blockIdx.x * kWarps + warp; b < B; b += gridDim.x * kWarps) {      // warp-uniform
    const T* xr = x + (long long)b * ldx;
    const float* tr = t + (long long)b * ldt;
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, ldf<T>(xr + c));
    m = wmax(m);
    float se = 0.f, ts = 0.f, tx = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float xv = ldf<T>(xr + c), tv = __ldg(tr + c);
      se += __expf(xv - m);
      ts += tv;
      tx += tv * xv;
    }
    se = wsum(se); ts = wsum(ts); tx = wsum(tx);
    const float lse = m + __logf(se);
    if (lane == 0) {
      row_lse[b] = lse;
      row_tsum[b] = ts;
      block_loss += lse * ts - tx;
    }
  }
  if (lane == 0) red[warp] = block_loss;
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < kWarps; ++i) s += red[i];
    partials[blockIdx.x] = s;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x < 32) {
    float s = 0.f;
    for (unsigned int i = lane; i < gridDim.x; i += 32) s += __ldcg(partials + i);
    s = wsum(s);
    if (lane == 0) {
      loss[0] = s / (float)B;
      *ticket = 0u;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) softce_bwd_kernel(const T* __restrict__ x, long long ldx, const float* __restrict__ t, long long ldt,
                                                         const float* __restrict__ row_lse, const float* __restrict__ row_tsum,
                                                         const float* __restrict__ dloss, int B, int C, T* __restrict__ dx, long long lddx) {
  const float g = __ldg(dloss) / (float)B;
  const long long total = (long long)B * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / C), c = (int)(i % C);
    const float xv = ldf<T>(x + (long long)b * ldx + c);
    const float v = g * (__expf(xv - __ldg(row_lse + b)) * __ldg(row_tsum + b) - __ldg(t + (long long)b * ldt + c));
    if constexpr (sizeof(T) == 4) dx[(long long)b * lddx + c] = v;
    else dx[(long long)b * lddx + c] = __float2bfloat16(v);
  }
}

int fwd_grid(int B) {
  int g = (B + kWarps - 1) / kWarps;
  const int cap = 2 * sfc_num_sms();
  return g > cap ? cap : (g < 1 ? 1 : g);
}

}  // namespace

extern "C" size_t sfc_softce_scratch_bytes(void) { return 16 + sizeof(float) * 2 * (size_t)sfc_num_sms(); }

extern "C" int sfc_softce_fwd(const void* logits, int logits_fp32, long long ld_logits, const float* targets, long long ld_targets, int B,
                              int C, float* loss, float* row_lse, float* row_tsum, void* scratch, size_t scratch_bytes,
                              cudaStream_t stream) {
  SFC_REQUIRE(logits && targets && loss && row_lse && row_tsum && B > 0 && C > 0, "sfc_softce_fwd: bad arguments");
  SFC_REQUIRE(scratch && scratch_bytes >= sfc_softce_scratch_bytes(), "sfc_softce_fwd: scratch too small");
  unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch);
  float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + 16);
  const int grid = fwd_grid(B);
  if (logits_fp32)
    softce_fwd_kernel<float><<<grid, kWarps * 32, 0, stream>>>((const float*)logits, ld_logits, targets, ld_targets, B, C, loss, row_lse,
                                                                row_tsum, partials, ticket);
  else
    softce_fwd_kernel<__nv_bfloat16><<<grid, kWarps * 32, 0, stream>>>((const __nv_bfloat16*)logits, ld_logits, targets, ld_targets, B, C,
                                                                        loss, row_lse, row_tsum, partials, ticket);
  SFC_LAUNCH_OK();
  return 0;
}

extern "C" int sfc_softce_bwd(const void* logits, int logits_fp32, long long ld_logits, const float* targets, long long ld_targets,
                              const float* row_lse, const float* row_tsum, const float* dloss, int B, int C, void* dlogits,
                              long long ld_dlogits, cudaStream_t stream) {
  SFC_REQUIRE(logits && targets && row_lse && row_tsum && dloss && dlogits && B > 0 && C > 0, "sfc_softce_bwd: bad arguments");
  const long long total = (long long)B * C;
  long long blocks = (total + 255) / 256;
  if (blocks > 4ll * sfc_num_sms()) blocks = 4ll * sfc_num_sms();
  if (logits_fp32)
    softce_bwd_kernel<float><<<(unsigned)blocks, 256, 0, stream>>>((const float*)logits, ld_logits, targets, ld_targets, row_lse, row_tsum,
                                                                    dloss, B, C, (float*)dlogits, ld_dlogits);
  else
    softce_bwd_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, stream>>>((const __nv_bfloat16*)logits, ld_logits, targets, ld_targets,
                                                                            row_lse, row_tsum, dloss, B, C, (__nv_bfloat16*)dlogits,
                                                                            ld_dlogits);
  SFC_LAUNCH_OK();
  return 0;
}
