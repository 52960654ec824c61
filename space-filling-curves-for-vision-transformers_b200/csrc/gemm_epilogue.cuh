// Shared GEMM epilogue (used by gemm.cu and patch_embed.cu): 32 consecutive accumulator columns of one
// output row are finished in registers and written to global memory.
// order: *alpha -> +bias[n] -> (store out_pre) -> act -> dropout -> aux (relu mask / gelu') -> +residual -> store
//
// The global operands of a chunk (bias / residual / aux, 64 B each per row) are fetched by epi_prefetch() one chunk
// AHEAD of their use, so that their latency overlaps the TMEM load + math + stores of the previous chunk (the
// epilogue has only one warp per SM sub-partition, i.e. no other warp to hide a dependent global load behind).
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

struct EpiRegs {                  // raw bf16x8 vectors of one 32-column chunk
  uint4 bias[4], res[4], aux[4];
  bool bias_vec, res_vec, aux_vec;   // operand was prefetched with vector loads (else: scalar path at use time)
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Counter-based dropout (same decisions in forward and backward, no mask tensor): one splitmix64 hash of
// (seed, index >> 2) yields four 16-bit uniform lanes; element `index` is kept iff its lane >= thr16 = p * 65536.
// The keep probability is exactly 1 - thr16 / 65536, and that value (not the nominal p) is used for the rescale.
__device__ __forceinline__ unsigned long long drop_hash4(unsigned long long seed, unsigned long long group) {
  unsigned long long z = seed + group * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ uint32_t drop_thr16(float p) { return (uint32_t)(p * 65536.0f); }
__device__ __forceinline__ float drop_inv_keep(float p) {
  return p > 0.f ? 65536.0f / (65536.0f - (float)drop_thr16(p)) : 1.0f;
}
__device__ __forceinline__ bool drop_keep(unsigned long long seed, unsigned long long idx, uint32_t thr16) {
  const unsigned long long h = drop_hash4(seed, idx >> 2);
  return ((uint32_t)(h >> (16 * (uint32_t)(idx & 3))) & 0xffffu) >= thr16;
}
// v[i] = keep(base + i) ? v[i] * inv_keep : 0 for NV consecutive elements (NV % 4 == 0)
template <int NV>
__device__ __forceinline__ void drop_apply(float* v, unsigned long long seed, unsigned long long base, float p) {
  const uint32_t thr = drop_thr16(p);
  const float sc = drop_inv_keep(p);
  if ((base & 3ull) == 0) {
#pragma unroll
    for (int g = 0; g < NV / 4; ++g) {
      const unsigned long long h = drop_hash4(seed, (base >> 2) + g);
#pragma unroll
      for (int k = 0; k < 4; ++k) v[g * 4 + k] = (((uint32_t)(h >> (16 * k)) & 0xffffu) >= thr) ? v[g * 4 + k] * sc : 0.f;
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = drop_keep(seed, base + i, thr) ? v[i] * sc : 0.f;
  }
}

__device__ __forceinline__ void epi_unpack8(const uint4& b, float* f) {
  f[0] = ptx::bf16_lo(b.x); f[1] = ptx::bf16_hi(b.x); f[2] = ptx::bf16_lo(b.y); f[3] = ptx::bf16_hi(b.y);
  f[4] = ptx::bf16_lo(b.z); f[5] = ptx::bf16_hi(b.z); f[6] = ptx::bf16_lo(b.w); f[7] = ptx::bf16_hi(b.w);
}

// Issues the global loads of chunk n0 (no use of the results here).
__device__ __forceinline__ void epi_prefetch(const EpiParams& p, EpiRegs& r, long long m_out, long long m_res, int n0, bool row_ok) {
  const bool full = row_ok && (n0 + 32 <= p.N);
  r.bias_vec = full && p.bias && ((reinterpret_cast<uintptr_t>(p.bias + n0) & 15) == 0);
  r.res_vec = full && p.residual && (p.ld_res % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0);
  r.aux_vec = full && p.aux_mode != SFC_AUX_NONE && (p.ld_aux % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.aux) & 15) == 0);
  if (r.bias_vec) {
    const uint4* bp = reinterpret_cast<const uint4*>(p.bias + n0);
#pragma unroll
    for (int q = 0; q < 4; ++q) r.bias[q] = __ldg(bp + q);
  }
  if (r.res_vec) {
    const uint4* rp = reinterpret_cast<const uint4*>(p.residual + m_res * p.ld_res + n0);
#pragma unroll
    for (int q = 0; q < 4; ++q) r.res[q] = __ldg(rp + q);
  }
  if (r.aux_vec) {
    const uint4* ap = reinterpret_cast<const uint4*>(p.aux + m_out * p.ld_aux + n0);
#pragma unroll
    for (int q = 0; q < 4; ++q) r.aux[q] = __ldg(ap + q);
  }
}

// v[32]: raw accumulators of columns n0..n0+31 of one row. m_out: row in out / out_pre / aux; m_res: row in residual.
__device__ __forceinline__ void epi_apply_store(const EpiParams& p, float (&v)[32], const EpiRegs& r, long long m_out, long long m_res,
                                                int n0, int split) {
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
  const bool full = (n0 + 32 <= p.N);
  if (p.bias) {
    if (r.bias_vec) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float f[8];
        epi_unpack8(r.bias[q], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[q * 8 + j] += f[j];
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
  if (p.drop_p > 0.0f)
    drop_apply<32>(v, p.drop_seed, (unsigned long long)m_out * (unsigned long long)p.N + (unsigned long long)n0, p.drop_p);
  if (p.aux_mode != SFC_AUX_NONE) {
    float a[32];
    if (r.aux_vec) {
#pragma unroll
      for (int q = 0; q < 4; ++q) epi_unpack8(r.aux[q], a + q * 8);
    } else {
      const __nv_bfloat16* ap = p.aux + m_out * p.ld_aux + n0;
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
    if (r.res_vec) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float f[8];
        epi_unpack8(r.res[q], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[q * 8 + j] += f[j];
      }
    } else {
      const __nv_bfloat16* rp = p.residual + m_res * p.ld_res + n0;
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

// Shared driver of the epilogue of one 128-row x BN-column accumulator tile: walks the column chunks with the
// operands of chunk c+1 in flight while chunk c is finished. `taddr` = TMEM address of this thread's lane, column 0.
template <int BN>
__device__ __forceinline__ void epi_tile(const EpiParams& p, uint32_t taddr, int n_base, long long m_out, long long m_res, bool row_ok,
                                         int split) {
  EpiRegs cur, nxt;
  epi_prefetch(p, cur, m_out, m_res, n_base, row_ok && n_base < p.N);
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    const int n0 = n_base + c * 32;
    if (n0 >= p.N) break;                      // warp-uniform
    uint32_t raw[32];
    ptx::tmem_ld_x32(taddr + c * 32, raw);
    const bool more = (c + 1 < BN / 32) && (n0 + 32 < p.N);
    if (more) epi_prefetch(p, nxt, m_out, m_res, n0 + 32, row_ok);
    ptx::tmem_ld_wait();
    if (row_ok) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
      epi_apply_store(p, v, cur, m_out, m_res, n0, split);
    }
    if (more) cur = nxt;
  }
}
