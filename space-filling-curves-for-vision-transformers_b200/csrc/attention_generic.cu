// K4g — generic-head-dimension attention (CUDA cores), forward and backward.
//
// The tcgen05 kernels of attention.cu are specialised for head_dim = 64 (every BASELINE.json model). The reference's own
// driver, however, builds VisionTransformer1D(embed 3 x 256 = 768, n_heads = 4) -> head_dim 192 on 64 tokens
// (/root/reference/main.py:269-282), so "main.py drives it unchanged" needs an attention for other head dimensions.
// Those shapes are small (N <= 128 tokens), so this path keeps one (image, head) per CTA entirely in shared memory and
// uses warp-per-row FMA code: same packed qkv layout, same lse output, same counter-based dropout as attention.cu.
//   forward : warp per query row: scores over keys (lanes = keys), softmax by shuffles, O = P V (lanes = features)
//   backward: pass 1, warp per query row -> dQ;  pass 2, warp per key row -> dK, dV (P recomputed from lse; no atomics)
#include "common.cuh"
#include "gemm_epilogue.cuh"   // DropKey, drop_keep
#include "sfcvit.h"

namespace {

constexpr int kGThreads = 256;
constexpr int kMaxKeysPerLane = 4;     // N <= 128
constexpr float kLog2eG = 1.4426950408889634f;

struct GAttnParams {
  int B, H, N, D, dh;
  float scale, drop_p;
  unsigned long long drop_seed;
  const unsigned long long* drop_epoch;
  const __nv_bfloat16* qkv;   // [B*N, 3D]
  __nv_bfloat16* out;         // fwd: O [B*N, D]
  float* lse;                 // [B, H, N]
  const __nv_bfloat16* o;     // bwd
  const __nv_bfloat16* dout;  // bwd [B*N, D]
  __nv_bfloat16* dqkv;        // bwd [B*N, 3D]
};

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// loads rows [0, N) x dh of one third of qkv (or of a [B*N, D] matrix when stride3 == false) into smem with row pitch ldp
__device__ __forceinline__ void load_rows(__nv_bfloat16* dst, int ldp, const __nv_bfloat16* src, long long row_stride, int N, int dh) {
  const int vec_per_row = dh / 8;
  for (int i = threadIdx.x; i < N * vec_per_row; i += blockDim.x) {
    const int r = i / vec_per_row, c = (i % vec_per_row) * 8;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + (long long)r * row_stride + c));
    __nv_bfloat16* d = dst + r * ldp + c;                // pitch is not 16-byte aligned: scalar pair stores
    reinterpret_cast<uint32_t*>(d)[0] = v.x; reinterpret_cast<uint32_t*>(d)[1] = v.y;
    reinterpret_cast<uint32_t*>(d)[2] = v.z; reinterpret_cast<uint32_t*>(d)[3] = v.w;
  }
}

__device__ __forceinline__ float dot_row(const __nv_bfloat16* a, const __nv_bfloat16* b, int dh) {
  float s = 0.f;
  for (int d = 0; d < dh; d += 2) {
    const uint32_t x = *reinterpret_cast<const uint32_t*>(a + d), y = *reinterpret_cast<const uint32_t*>(b + d);
    s = fmaf(ptx::bf16_lo(x), ptx::bf16_lo(y), s);
    s = fmaf(ptx::bf16_hi(x), ptx::bf16_hi(y), s);
  }
  return s;
}

__device__ __forceinline__ unsigned long long drop_index(const GAttnParams& p, int b, int h, int q, int key) {
  return (((unsigned long long)(b * p.H + h) * p.N + (unsigned long long)q) * (unsigned long long)((p.N + 15) & ~15)) + (unsigned long long)key;
}

// smem: Q, K, V (fwd) / + dO, O-free (bwd) as bf16 [N][ldp], ldp = dh + 2 (odd word pitch: conflict-free column walks)
__global__ void __launch_bounds__(kGThreads) attn_generic_fwd_kernel(const GAttnParams p) {
  extern __shared__ __align__(16) uint8_t smem_g[];
  const int N = p.N, dh = p.dh, ldp = dh + 2;
  __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(smem_g);
  __nv_bfloat16* sk = sq + N * ldp;
  __nv_bfloat16* sv = sk + N * ldp;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const __nv_bfloat16* base = p.qkv + (long long)b * N * 3 * p.D + h * dh;
  load_rows(sq, ldp, base, 3ll * p.D, N, dh);
  load_rows(sk, ldp, base + p.D, 3ll * p.D, N, dh);
  load_rows(sv, ldp, base + 2 * p.D, 3ll * p.D, N, dh);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const DropKey dkey = drop_key(p.drop_seed, p.drop_p, p.drop_epoch);
  const float sl2 = p.scale * kLog2eG;
  for (int q = warp; q < N; q += nwarps) {
    float s[kMaxKeysPerLane], mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < kMaxKeysPerLane; ++i) {
      const int j = lane + 32 * i;
      s[i] = j < N ? dot_row(sq + q * ldp, sk + j * ldp, dh) : -INFINITY;
      mx = fmaxf(mx, s[i]);
    }
    mx = warp_max(mx);
    float pr[kMaxKeysPerLane], sum = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxKeysPerLane; ++i) {
      const int j = lane + 32 * i;
      pr[i] = j < N ? exp2f((s[i] - mx) * sl2) : 0.f;
      sum += pr[i];
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int i = 0; i < kMaxKeysPerLane; ++i) {
      const int j = lane + 32 * i;
      pr[i] *= inv;
      if (p.drop_p > 0.f && j < N) pr[i] = drop_keep<16>(dkey, drop_index(p, b, h, q, j)) ? pr[i] * dkey.inv_keep : 0.f;
    }
    // O[q][d] = sum_j P[j] V[j][d]; lanes own feature columns d = lane, lane + 32, ...
    for (int d0 = 0; d0 < dh; d0 += 32) {
      const int d = d0 + lane;
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < kMaxKeysPerLane; ++i) {
        for (int jj = 0; jj < 32; ++jj) {
          const int j = jj + 32 * i;
          if (j >= N) break;
          const float pj = __shfl_sync(0xffffffffu, pr[i], jj);
          if (d < dh) acc = fmaf(pj, __bfloat162float(sv[j * ldp + d]), acc);
        }
      }
      if (d < dh) p.out[(long long)(b * N + q) * p.D + h * dh + d] = __float2bfloat16(acc);
    }
    if (lane == 0 && p.lse) p.lse[((long long)b * p.H + h) * N + q] = mx * p.scale + logf(sum);
  }
}

__global__ void __launch_bounds__(kGThreads) attn_generic_bwd_kernel(const GAttnParams p) {
  extern __shared__ __align__(16) uint8_t smem_g[];
  const int N = p.N, dh = p.dh, ldp = dh + 2;
  __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(smem_g);
  __nv_bfloat16* sk = sq + N * ldp;
  __nv_bfloat16* sv = sk + N * ldp;
  __nv_bfloat16* sdo = sv + N * ldp;
  float* slse = reinterpret_cast<float*>(sdo + N * ldp);      // [N] lse * log2e
  float* sdelta = slse + N;                                   // [N] rowsum(dO * O)
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const __nv_bfloat16* base = p.qkv + (long long)b * N * 3 * p.D + h * dh;
  load_rows(sq, ldp, base, 3ll * p.D, N, dh);
  load_rows(sk, ldp, base + p.D, 3ll * p.D, N, dh);
  load_rows(sv, ldp, base + 2 * p.D, 3ll * p.D, N, dh);
  load_rows(sdo, ldp, p.dout + (long long)b * N * p.D + h * dh, (long long)p.D, N, dh);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  __syncthreads();
  for (int q = warp; q < N; q += nwarps) {
    float dl = 0.f;
    for (int d = lane; d < dh; d += 32)
      dl += __bfloat162float(sdo[q * ldp + d]) * __bfloat162float(p.o[(long long)(b * N + q) * p.D + h * dh + d]);
    dl = warp_sum(dl);
    if (lane == 0) {
      sdelta[q] = dl;
      slse[q] = p.lse[((long long)b * p.H + h) * N + q] * kLog2eG;
    }
  }
  __syncthreads();
  const DropKey dkey = drop_key(p.drop_seed, p.drop_p, p.drop_epoch);
  const float sl2 = p.scale * kLog2eG;
  // dS[q][j] and the (dropped) P[q][j] of one (q, j) pair
  auto pair = [&](int q, int j, float& pd, float& ds) {
    const float pr = exp2f(dot_row(sq + q * ldp, sk + j * ldp, dh) * sl2 - slse[q]);
    float dp = dot_row(sdo + q * ldp, sv + j * ldp, dh);
    pd = pr;
    if (p.drop_p > 0.f) {
      const bool keep = drop_keep<16>(dkey, drop_index(p, b, h, q, j));
      pd = keep ? pr * dkey.inv_keep : 0.f;
      dp = keep ? dp * dkey.inv_keep : 0.f;
    }
    ds = pr * (dp - sdelta[q]) * p.scale;
  };
  // pass 1: dQ[q][d] = sum_j dS[q][j] K[j][d]
  for (int q = warp; q < N; q += nwarps) {
    float ds[kMaxKeysPerLane];
#pragma unroll
    for (int i = 0; i < kMaxKeysPerLane; ++i) {
      const int j = lane + 32 * i;
      float pd;
      ds[i] = 0.f;
      if (j < N) pair(q, j, pd, ds[i]);
    }
    for (int d0 = 0; d0 < dh; d0 += 32) {
      const int d = d0 + lane;
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < kMaxKeysPerLane; ++i) {
        for (int jj = 0; jj < 32; ++jj) {
          const int j = jj + 32 * i;
          if (j >= N) break;
          const float v = __shfl_sync(0xffffffffu, ds[i], jj);
          if (d < dh) acc = fmaf(v, __bfloat162float(sk[j * ldp + d]), acc);
        }
      }
      if (d < dh) p.dqkv[(long long)(b * N + q) * 3 * p.D + h * dh + d] = __float2bfloat16(acc);
    }
  }
  // pass 2: dV[j][d] = sum_q P[q][j] dO[q][d],  dK[j][d] = sum_q dS[q][j] Q[q][d]   (lanes = queries, then features)
  for (int j = warp; j < N; j += nwarps) {
    float pdv[kMaxKeysPerLane], dsv[kMaxKeysPerLane];
#pragma unroll
    for (int i = 0; i < kMaxKeysPerLane; ++i) {
      const int q = lane + 32 * i;
      pdv[i] = 0.f; dsv[i] = 0.f;
      if (q < N) pair(q, j, pdv[i], dsv[i]);
    }
    for (int d0 = 0; d0 < dh; d0 += 32) {
      const int d = d0 + lane;
      float av = 0.f, ak = 0.f;
#pragma unroll
      for (int i = 0; i < kMaxKeysPerLane; ++i) {
        for (int qq = 0; qq < 32; ++qq) {
          const int q = qq + 32 * i;
          if (q >= N) break;
          const float pv = __shfl_sync(0xffffffffu, pdv[i], qq);
          const float dv = __shfl_sync(0xffffffffu, dsv[i], qq);
          if (d < dh) {
            av = fmaf(pv, __bfloat162float(sdo[q * ldp + d]), av);
            ak = fmaf(dv, __bfloat162float(sq[q * ldp + d]), ak);
          }
        }
      }
      if (d < dh) {
        p.dqkv[(long long)(b * N + j) * 3 * p.D + p.D + h * dh + d] = __float2bfloat16(ak);
        p.dqkv[(long long)(b * N + j) * 3 * p.D + 2 * p.D + h * dh + d] = __float2bfloat16(av);
      }
    }
  }
}

}  // namespace

// Shapes served by this path: head_dim % 8 == 0, N <= 128 and the (image, head) working set within shared memory.
int sfc_attn_generic_supported(int N, int dh, bool bwd) {
  if (dh % 8 != 0 || dh < 8 || N < 1 || N > 32 * kMaxKeysPerLane) return 0;
  const size_t bytes = (size_t)(bwd ? 4 : 3) * N * (dh + 2) * 2 + (bwd ? 2 * N * 4 : 0);
  return bytes <= 220 * 1024;
}

int sfc_attn_generic_fwd(const void* qkv, void* out, float* lse, int B, int H, int N, int D, float scale, float drop_p,
                         unsigned long long drop_seed, cudaStream_t stream) {
  const int dh = D / H;
  GAttnParams p{};
  p.B = B; p.H = H; p.N = N; p.D = D; p.dh = dh; p.scale = scale; p.drop_p = drop_p; p.drop_seed = drop_seed;
  p.drop_epoch = sfc_dropout_epoch_ptr();
  p.qkv = (const __nv_bfloat16*)qkv; p.out = (__nv_bfloat16*)out; p.lse = lse;
  const int smem = 3 * N * (dh + 2) * 2;
  SFC_CUDA_OK(cudaFuncSetAttribute(attn_generic_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  attn_generic_fwd_kernel<<<B * H, kGThreads, smem, stream>>>(p);
  SFC_LAUNCH_OK();
  return 0;
}

int sfc_attn_generic_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int B, int H, int N,
                         int D, float scale, float drop_p, unsigned long long drop_seed, cudaStream_t stream) {
  const int dh = D / H;
  GAttnParams p{};
  p.B = B; p.H = H; p.N = N; p.D = D; p.dh = dh; p.scale = scale; p.drop_p = drop_p; p.drop_seed = drop_seed;
  p.drop_epoch = sfc_dropout_epoch_ptr();
  p.qkv = (const __nv_bfloat16*)qkv; p.o = (const __nv_bfloat16*)out; p.dout = (const __nv_bfloat16*)dout;
  p.lse = const_cast<float*>(lse); p.dqkv = (__nv_bfloat16*)dqkv;
  const int smem = 4 * N * (dh + 2) * 2 + 2 * N * 4;
  SFC_CUDA_OK(cudaFuncSetAttribute(attn_generic_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  attn_generic_bwd_kernel<<<B * H, kGThreads, smem, stream>>>(p);
  SFC_LAUNCH_OK();
  return 0;
}
