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

// 8 consecutive elements per thread and iteration (one 16-byte vector of bf16, two of fp32)
template <typename T> struct Vec8;
template <> struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* f) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    f[0] = ptx::bf16_lo(u.x); f[1] = ptx::bf16_hi(u.x); f[2] = ptx::bf16_lo(u.y); f[3] = ptx::bf16_hi(u.y);
    f[4] = ptx::bf16_lo(u.z); f[5] = ptx::bf16_hi(u.z); f[6] = ptx::bf16_lo(u.w); f[7] = ptx::bf16_hi(u.w);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float* f) {
    uint4 u;
    u.x = ptx::pack_bf16(f[0], f[1]); u.y = ptx::pack_bf16(f[2], f[3]);
    u.z = ptx::pack_bf16(f[4], f[5]); u.w = ptx::pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <> struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float* f) {
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* f) {
    reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
};

// Deterministic: every block writes its partial, the last block to finish (ticket) adds them up in block order and
// adds the total to accum[0] — the same bits on every rank of a data-parallel job (an atomicAdd per block would make
// the clip coefficient, and with it the replicas' parameters, differ in the last place from rank to rank).
template <typename T>
__global__ void __launch_bounds__(256) sumsq_kernel(const T* __restrict__ g, long long n, float* __restrict__ accum,
                                                    unsigned int* __restrict__ ticket, float* __restrict__ partials) {
  float s = 0.f;
  const long long n8 = n / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    Vec8<T>::load(g + i * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += f[j] * f[j];
  }
  if (blockIdx.x == 0) {
    for (long long i = n8 * 8 + threadIdx.x; i < n; i += blockDim.x) {
      const float v = to_f<T>(g[i]);
      s += v * v;
    }
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
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
  }
  __shared__ bool last;
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float t = 0.f;
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) t += __ldcg(partials + i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    accum[0] += t;
    *ticket = 0u;                              // ready for the next launch on this scratch
  }
}

struct AdamArgs {
  float lr, beta1, beta2, eps, wd, bc1, bc2_sqrt, grad_scale, max_norm;
  const float* hyper;   // optional device block {lr, 1 - beta1^t, sqrt(1 - beta2^t)}: overrides lr / bc1 / bc2_sqrt (CUDA-graph replays)
};

__device__ __forceinline__ void adam_update(float& pi, float gi, float& mi, float& vi, const AdamArgs& a, float coef) {
  gi *= coef;
  pi *= 1.0f - a.lr * a.wd;
  mi = a.beta1 * mi + (1.0f - a.beta1) * gi;
  vi = a.beta2 * vi + (1.0f - a.beta2) * gi * gi;
  const float denom = sqrtf(vi) / a.bc2_sqrt + a.eps;
  pi -= (a.lr / a.bc1) * (mi / denom);
}

// stats[0] = sum of squares of the *unscaled* gradient (all buckets); grad_scale multiplies g first (1/world).
template <typename T, typename S>
__global__ void __launch_bounds__(256) adamw_kernel(T* __restrict__ p, const T* __restrict__ g, S* __restrict__ m, S* __restrict__ v,
                                                    long long n, AdamArgs a, const float* __restrict__ stats) {
  if (a.hyper != nullptr) { a.lr = __ldg(a.hyper); a.bc1 = __ldg(a.hyper + 1); a.bc2_sqrt = __ldg(a.hyper + 2); }
  float coef = a.grad_scale;
  if (a.max_norm > 0.f && stats) {
    const float total = sqrtf(stats[0]) * a.grad_scale;
    coef *= fminf(1.0f, a.max_norm / (total + 1e-6f));
  }
  const long long n8 = n / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float pf[8], gf[8], mf[8], vf[8];
    Vec8<T>::load(p + i * 8, pf);
    Vec8<T>::load(g + i * 8, gf);
    Vec8<S>::load(m + i * 8, mf);
    Vec8<S>::load(v + i * 8, vf);
#pragma unroll
    for (int j = 0; j < 8; ++j) adam_update(pf[j], gf[j], mf[j], vf[j], a, coef);
    Vec8<T>::store(p + i * 8, pf);
    Vec8<S>::store(m + i * 8, mf);
    Vec8<S>::store(v + i * 8, vf);
  }
  if (blockIdx.x == 0) {
    for (long long i = n8 * 8 + threadIdx.x; i < n; i += blockDim.x) {
      float pi = to_f<T>(p[i]), mi = to_f<S>(m[i]), vi = to_f<S>(v[i]);
      adam_update(pi, to_f<T>(g[i]), mi, vi, a, coef);
      p[i] = from_f<T>(pi);
      m[i] = from_f<S>(mi);
      v[i] = from_f<S>(vi);
    }
  }
}

__global__ void store_f32x4_kernel(float* dst, float a, float b, float c, float d) {
  dst[0] = a; dst[1] = b; dst[2] = c; dst[3] = d;
}

int grid_for(long long n) {
  long long b = sfc_ceil_div64(n, 256 * 8);
  const long long cap = 8ll * sfc_num_sms();
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" size_t sfc_grad_sumsq_scratch_bytes(void) { return 16 + sizeof(float) * 8 * (size_t)sfc_num_sms(); }

// accum[0] += sum(g^2); caller zeroes accum before the first bucket. scratch: sfc_grad_sumsq_scratch_bytes() bytes, zeroed
// ONCE by the caller (the kernel leaves its ticket at zero), private to the stream.
extern "C" int sfc_grad_sumsq(const void* g, int g_fp32, long long n, float* accum, void* scratch, size_t scratch_bytes,
                              cudaStream_t stream) {
  SFC_REQUIRE(g && accum && n >= 0, "sfc_grad_sumsq: bad arguments");
  SFC_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, "sfc_grad_sumsq: buffer must be 16-byte aligned");
  SFC_REQUIRE(scratch && scratch_bytes >= sfc_grad_sumsq_scratch_bytes(), "sfc_grad_sumsq: scratch too small");
  if (n == 0) return 0;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch);
  float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + 16);
  if (g_fp32) sumsq_kernel<float><<<grid_for(n), 256, 0, stream>>>((const float*)g, n, accum, ticket, partials);
  else sumsq_kernel<__nv_bfloat16><<<grid_for(n), 256, 0, stream>>>((const __nv_bfloat16*)g, n, accum, ticket, partials);
  SFC_LAUNCH_OK();
  return 0;
}

extern "C" int sfc_adamw_step(void* p, const void* g, void* m, void* v, long long n, int param_fp32, int state_fp32, float lr,
                              float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                              float max_norm, const float* stats, const float* hyper_dev, cudaStream_t stream) {
  SFC_REQUIRE(p && g && m && v && n >= 0 && (step >= 1 || hyper_dev), "sfc_adamw_step: bad arguments");
  SFC_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                reinterpret_cast<uintptr_t>(v)) & 15) == 0, "sfc_adamw_step: buffers must be 16-byte aligned");
  if (n == 0) return 0;
  const float bc1 = step >= 1 ? 1.0f - powf(beta1, (float)step) : 1.0f;
  const float bc2s = step >= 1 ? sqrtf(1.0f - powf(beta2, (float)step)) : 1.0f;
  const int grid = grid_for(n);
  AdamArgs a{lr, beta1, beta2, eps, weight_decay, bc1, bc2s, grad_scale, max_norm, hyper_dev};
#define ADAMW(T, S) adamw_kernel<T, S><<<grid, 256, 0, stream>>>((T*)p, (const T*)g, (S*)m, (S*)v, n, a, stats)
  if (param_fp32 && state_fp32) ADAMW(float, float);
  else if (!param_fp32 && state_fp32) ADAMW(__nv_bfloat16, float);
  else if (!param_fp32 && !state_fp32) ADAMW(__nv_bfloat16, __nv_bfloat16);
  else { sfc_set_error("sfc_adamw_step: fp32 parameters with bf16 state is not supported"); return 2; }
#undef ADAMW
  SFC_LAUNCH_OK();
  return 0;
}

// dst[0..3] = {a, b, c, d}, stream-ordered, values travel as kernel arguments (no host buffer whose lifetime the caller
// would have to manage while the stream runs ahead): how the host scheduler's lr / bias corrections reach a CUDA graph.
extern "C" int sfc_store_f32x4(float* dst, float a, float b, float c, float d, cudaStream_t stream) {
  SFC_REQUIRE(dst != nullptr, "sfc_store_f32x4: null destination");
  store_f32x4_kernel<<<1, 1, 0, stream>>>(dst, a, b, c, d);
  SFC_LAUNCH_OK();
  return 0;
}
