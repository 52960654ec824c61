// Shared GEMM epilogue (used by gemm.cu and patch_embed.cu): 32 consecutive accumulator columns of one
// output row are finished in registers and written to global memory.
// order: *alpha -> +bias[n] -> (store out_pre) -> act -> dropout -> aux (relu mask / gelu') -> +residual -> store
#pragma once
#include "common.cuh"
#include "sfcvit.h"

struct EpiParams {
  int N;                          // number of valid output columns
  const __nv_bfloat16* bias;      // [N] or null
  const __nv_bfloat16* residual;  // [*, ld_res] or null (added after activation)
  const __nv_bfloat16* aux;       // [*, ld_aux] or null
  void* out;                      // bf16 or fp32 [*, ld_out]
  __nv_bfloat16* out_pre;         // optional pre-activation copy (bf16, ld_out)
  long long ld_out, ld_res, ld_aux;
  long long split_stride;         // elements between split-K partial outputs (fp32)
  float alpha;
  int act;                        // SFC_ACT_*
  int aux_mode;                   // SFC_AUX_*
  int out_fp32;
  float drop_p;                   // dropout prob applied after activation (0 = off)
  unsigned long long drop_seed;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Counter-based dropout keep decision (same hash used by forward and backward): splitmix64 of (seed, index).
__device__ __forceinline__ bool drop_keep(unsigned long long seed, unsigned long long idx, float p) {
  unsigned long long z = seed + idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  const float u = (float)(unsigned)(z >> 40) * (1.0f / 16777216.0f);
  return u >= p;
}

// v[32]: raw accumulators of columns n0..n0+31 of one row. m_out: row in out / out_pre / aux; m_res: row in residual.
__device__ __forceinline__ void epi_apply_store(const EpiParams& p, float (&v)[32], long long m_out, long long m_res, int n0, int split) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
    const bool full = (n0 + 32 <= p.N);
    if (p.bias) {
      if (full && ((reinterpret_cast<uintptr_t>(p.bias + n0) & 15) == 0)) {
        const uint4* bp = reinterpret_cast<const uint4*>(p.bias + n0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 b = __ldg(bp + q);
          v[q * 8 + 0] += ptx::bf16_lo(b.x); v[q * 8 + 1] += ptx::bf16_hi(b.x);
          v[q * 8 + 2] += ptx::bf16_lo(b.y); v[q * 8 + 3] += ptx::bf16_hi(b.y);
          v[q * 8 + 4] += ptx::bf16_lo(b.z); v[q * 8 + 5] += ptx::bf16_hi(b.z);
          v[q * 8 + 6] += ptx::bf16_lo(b.w); v[q * 8 + 7] += ptx::bf16_hi(b.w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + j < p.N) v[j] += __bfloat162float(p.bias[n0 + j]);
      }
    }
    const bool out_vec = full && (p.ld_out % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    if (p.out_pre) {
      __nv_bfloat16* op = p.out_pre + m_out * p.ld_out + n0;
      if (out_vec && ((reinterpret_cast<uintptr_t>(p.out_pre) & 15) == 0)) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = ptx::pack_bf16(v[q * 8 + 0], v[q * 8 + 1]); o.y = ptx::pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
          o.z = ptx::pack_bf16(v[q * 8 + 4], v[q * 8 + 5]); o.w = ptx::pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
          reinterpret_cast<uint4*>(op)[q] = o;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + j < p.N) op[j] = __float2bfloat16(v[j]);
      }
    }
    if (p.act == SFC_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
    } else if (p.act == SFC_ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
    }
    if (p.drop_p > 0.0f) {
      const float sc = 1.0f / (1.0f - p.drop_p);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        v[j] = drop_keep(p.drop_seed, (unsigned long long)m_out * (unsigned long long)p.N + (unsigned long long)(n0 + j), p.drop_p) ? v[j] * sc : 0.0f;
    }
    if (p.aux_mode != SFC_AUX_NONE) {
      const __nv_bfloat16* ap = p.aux + m_out * p.ld_aux + n0;
      float a[32];
      if (full && (p.ld_aux % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.aux) & 15) == 0)) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 b = __ldg(reinterpret_cast<const uint4*>(ap) + q);
          a[q * 8 + 0] = ptx::bf16_lo(b.x); a[q * 8 + 1] = ptx::bf16_hi(b.x);
          a[q * 8 + 2] = ptx::bf16_lo(b.y); a[q * 8 + 3] = ptx::bf16_hi(b.y);
          a[q * 8 + 4] = ptx::bf16_lo(b.z); a[q * 8 + 5] = ptx::bf16_hi(b.z);
          a[q * 8 + 6] = ptx::bf16_lo(b.w); a[q * 8 + 7] = ptx::bf16_hi(b.w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) a[j] = (n0 + j < p.N) ? __bfloat162float(ap[j]) : 0.0f;
      }
      if (p.aux_mode == SFC_AUX_RELU_MASK) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = a[j] > 0.0f ? v[j] : 0.0f;
      } else {  // SFC_AUX_GELU_GRAD: aux holds the pre-activation
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= gelu_erf_grad(a[j]);
      }
    }
    if (p.residual) {
      const __nv_bfloat16* rp = p.residual + m_res * p.ld_res + n0;
      if (full && (p.ld_res % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0)) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 b = __ldg(reinterpret_cast<const uint4*>(rp) + q);
          v[q * 8 + 0] += ptx::bf16_lo(b.x); v[q * 8 + 1] += ptx::bf16_hi(b.x);
          v[q * 8 + 2] += ptx::bf16_lo(b.y); v[q * 8 + 3] += ptx::bf16_hi(b.y);
          v[q * 8 + 4] += ptx::bf16_lo(b.z); v[q * 8 + 5] += ptx::bf16_hi(b.z);
          v[q * 8 + 6] += ptx::bf16_lo(b.w); v[q * 8 + 7] += ptx::bf16_hi(b.w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + j < p.N) v[j] += __bfloat162float(rp[j]);
      }
    }
    if (p.out_fp32) {
      float* op = reinterpret_cast<float*>(p.out) + (long long)split * p.split_stride + m_out * p.ld_out + n0;
      if (full && (p.ld_out % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) && (p.split_stride % 4 == 0)) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          reinterpret_cast<float4*>(op)[q] = make_float4(v[q * 4 + 0], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + j < p.N) op[j] = v[j];
      }
    } else {
      __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + m_out * p.ld_out + n0;
      if (out_vec) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = ptx::pack_bf16(v[q * 8 + 0], v[q * 8 + 1]); o.y = ptx::pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
          o.z = ptx::pack_bf16(v[q * 8 + 4], v[q * 8 + 5]); o.w = ptx::pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
          reinterpret_cast<uint4*>(op)[q] = o;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + j < p.N) op[j] = __float2bfloat16(v[j]);
      }
    }
}
