// K6 — fused gradient-norm + clip + AdamW over flat parameter buckets (sm_100a, memory-bound).
// Replaces torch.nn.utils.clip_grad_norm_(params, 1.0, foreach=False) + optim.AdamW.step()
// (/root/reference/src/training/train.py:165-166, main.py:288-289): one pass computes sum(g^2) of a flat
// gradient bucket, the clip coefficient min(1, max_norm / (||g|| + 1e-6)) is derived ON DEVICE (no host sync),
// and one pass applies the decoupled-weight-decay Adam update with the clipped (and 1/world averaged) gradient.
#include "common.cuh"
#include "sfcvit.h"

namespace {

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

template <typename T>
__global__ void __launch_bounds__(256) sumsq_kernel(const T* __restrict__ g, long long n, float* __restrict__ accum) {
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = to_f<T>(g[i]);
    s += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    s = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
    if (threadIdx.x == 0) atomicAdd(accum, s);
  }
}

// stats[0] = sum of squares of the *unscaled* gradient (all buckets); grad_scale multiplies g first (1/world).
template <typename T, typename S>
__global__ void __launch_bounds__(256) adamw_kernel(T* __restrict__ p, const T* __restrict__ g, S* __restrict__ m, S* __restrict__ v,
                                                    long long n, float lr, float beta1, float beta2, float eps, float wd,
                                                    float bc1, float bc2_sqrt, float grad_scale, float max_norm,
                                                    const float* __restrict__ stats) {
  float coef = grad_scale;
  if (max_norm > 0.f && stats) {
    const float total = sqrtf(stats[0]) * grad_scale;
    coef *= fminf(1.0f, max_norm / (total + 1e-6f));
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = to_f<T>(g[i]) * coef;
    float pi = to_f<T>(p[i]);
    float mi = to_f<S>(m[i]), vi = to_f<S>(v[i]);
    pi *= 1.0f - lr * wd;
    mi = beta1 * mi + (1.0f - beta1) * gi;
    vi = beta2 * vi + (1.0f - beta2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = from_f<T>(pi);
    m[i] = from_f<S>(mi);
    v[i] = from_f<S>(vi);
  }
}

int grid_for(long long n) {
  long long b = sfc_ceil_div64(n, 256 * 4);
  const long long cap = 8ll * sfc_num_sms();
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace

// accum[0] += sum(g^2); caller zeroes accum before the first bucket
extern "C" int sfc_grad_sumsq(const void* g, int g_fp32, long long n, float* accum, cudaStream_t stream) {
  SFC_REQUIRE(g && accum && n >= 0, "sfc_grad_sumsq: bad arguments");
  if (n == 0) return 0;
  if (g_fp32) sumsq_kernel<float><<<grid_for(n), 256, 0, stream>>>((const float*)g, n, accum);
  else sumsq_kernel<__nv_bfloat16><<<grid_for(n), 256, 0, stream>>>((const __nv_bfloat16*)g, n, accum);
  SFC_LAUNCH_OK();
  return 0;
}

extern "C" int sfc_adamw_step(void* p, const void* g, void* m, void* v, long long n, int param_fp32, int state_fp32, float lr,
                              float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                              float max_norm, const float* stats, cudaStream_t stream) {
  SFC_REQUIRE(p && g && m && v && n >= 0 && step >= 1, "sfc_adamw_step: bad arguments");
  if (n == 0) return 0;
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2s = sqrtf(1.0f - powf(beta2, (float)step));
  const int grid = grid_for(n);
#define ADAMW(T, S) adamw_kernel<T, S><<<grid, 256, 0, stream>>>((T*)p, (const T*)g, (S*)m, (S*)v, n, lr, beta1, beta2, eps, \
                                                                 weight_decay, bc1, bc2s, grad_scale, max_norm, stats)
  if (param_fp32 && state_fp32) ADAMW(float, float);
  else if (!param_fp32 && state_fp32) ADAMW(__nv_bfloat16, float);
  else if (!param_fp32 && !state_fp32) ADAMW(__nv_bfloat16, __nv_bfloat16);
  else { sfc_set_error("sfc_adamw_step: fp32 parameters with bf16 state is not supported"); return 2; }
#undef ADAMW
  SFC_LAUNCH_OK();
  return 0;
}
