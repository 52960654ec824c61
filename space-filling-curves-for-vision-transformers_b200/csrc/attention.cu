// K4 — tcgen05 flash attention, forward and backward, head_dim = 64, non-causal, no mask (sm_100a).
//
// Replaces F.scaled_dot_product_attention inside nn.MultiheadAttention of the reference encoder
// (/root/reference/src/models/vit.py:197-206 -> torch TransformerEncoderLayer._sa_block; enable_flash_sdp at
// main.py:158-159) and the explicit softmax(QK^T)V of altvit.py:129-141.
//
// Data layout: the packed in-projection output qkv[B*N, 3*D] (bf16) is read IN PLACE through one TMA tensor
// map: Q/K/V tile of head h = 64-column box at column h*64 / D+h*64 / 2D+h*64. Outputs: O[B*N, D] (heads
// concatenated, ready for out_proj) and LSE[B, H, N] (fp32, natural log) for the backward pass.
//
// Forward, one CTA per (128-query tile, head, image): S = Q K^T (tcgen05, fp32 in TMEM), each of the 128
// threads owns one query row (TMEM lane == row, so the row max / sum need no shuffles), exp2 online softmax,
// P (bf16) is written to shared memory in the swizzled K-major layout and O_j = P V_j runs on the tensor core
// with V read as an MN-major operand (no transpose); the running O is rescaled in registers.
// Backward, one CTA per (128-key tile, head, image), loop over query tiles: S and dP = dO V^T on the tensor
// core, P / dS in registers -> shared memory, dV += P^T dO, dK += dS^T Q (MN-major A operands, accumulators
// resident in TMEM across the loop) and dQ_i = dS K (red.global.add into an fp32 accumulator).
#include "common.cuh"
#include "gemm_epilogue.cuh"   // DropKey, drop_apply
#include "sfcvit.h"

namespace {

constexpr int DH = 64;
constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int kTile = BQ * DH * 2;          // 16384 bytes: one 128 x 64 bf16 tile
constexpr float kLog2e = 1.4426950408889634f;

struct AttnParams {
  int B, H, N, D;
  float scale;
  float drop_p;
  unsigned long long drop_seed;
  __nv_bfloat16* out;      // fwd: O [B*N, D]
  float* lse;              // [B, H, N]
  // backward
  const __nv_bfloat16* o;  // [B*N, D]
  const __nv_bfloat16* dout;  // [B*N, D]
  __nv_bfloat16* dqkv;     // [B*N, 3D]
  float* dq_acc;           // [B*N, D] fp32, zero-initialised
};

// write 32 consecutive P values (columns c32*32 .. +31 of row r) as bf16 into the K-major SW128 operand buffer
__device__ __forceinline__ void store_p_chunk(uint8_t* buf, int r, int c32, const float* v) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int j8 = c32 * 4 + q;
    uint4 o;
    o.x = ptx::pack_bf16(v[q * 8 + 0], v[q * 8 + 1]); o.y = ptx::pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
    o.z = ptx::pack_bf16(v[q * 8 + 4], v[q * 8 + 5]); o.w = ptx::pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
    *reinterpret_cast<uint4*>(buf + (j8 >> 3) * kTile + r * 128 + (((j8 & 7) ^ (r & 7)) << 4)) = o;
  }
}

// ================================================= forward =================================================
// Persistent, warp-specialised. Work item = (image b, head h, query pair qp): 256 query rows = two 128-row tiles A
// and B, each owned by one softmax warpgroup (thread <-> query row <-> TMEM lane), so the tensor core, the TMA engine
// and the other warpgroup always have independent work while one warpgroup is inside its exp loop.
//   warp 0 : TMA producer (Q tiles, K / V tiles of `bkv` keys; runs ahead across items)
//   warp 1 : MMA issuer   (S_w = Q_w K^T into TMEM region w; O_w = P_w V into the same region once P_w is in smem)
//   warp 2 : TMEM allocator;  warps 4-7 : softmax warpgroup A;  warps 8-11 : softmax warpgroup B
// The key tile size is a runtime value (any multiple of 16 up to 208): N <= 208 (the 14 x 14 grid: 196 -> 208) is a
// single pass without online-softmax rescaling; longer sequences use equal tiles of <= 192 keys, double buffered.
constexpr int kFwdThreads = 384;

struct FwdLayout {          // runtime shared-memory map (bytes), all tile bases 1024-byte aligned
  int bkv, nkv, stages, p_atoms;
  int off_k, off_v, off_p, off_bar, total;
};

inline FwdLayout fwd_layout(int N) {
  FwdLayout L;
  if (N <= 208) { L.bkv = (N + 15) / 16 * 16; L.nkv = 1; L.stages = 1; }
  else {
    L.nkv = (N + 191) / 192;
    L.bkv = ((N + L.nkv - 1) / L.nkv + 15) / 16 * 16;
    L.stages = 2;
  }
  L.p_atoms = (L.bkv + 63) / 64;
  const int kv_bytes = (L.bkv * 128 + 1023) / 1024 * 1024;
  L.off_k = 2 * kTile;
  L.off_v = L.off_k + L.stages * kv_bytes;
  L.off_p = L.off_v + L.stages * kv_bytes;
  L.off_bar = L.off_p + 2 * L.p_atoms * kTile;
  L.total = L.off_bar + 32 * 8 + 16 + 1024;
  return L;
}

struct FwdBars {            // mbarrier indices
  static constexpr int q_full = 0, q_empty = 2, k_full = 4, k_empty = 6, v_full = 8, v_empty = 10, s_full = 12, p_full = 14,
                       o_full = 16, o_empty = 18, count = 20;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool kSingle>
__global__ void __launch_bounds__(kFwdThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv, const AttnParams p,
                const FwdLayout L) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bar);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + FwdBars::count);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int bkv = L.bkv, nkv = L.nkv, stages = L.stages;
  const int kv_bytes = (bkv * 128 + 1023) / 1024 * 1024;
  const int nqp = (p.N + 2 * BQ - 1) / (2 * BQ);
  const int n_items = p.B * p.H * nqp;

  if (tid == 0) {
    ptx::prefetch_tmap(&tmap_q);
    ptx::prefetch_tmap(&tmap_kv);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bars[FwdBars::q_full + i], 1);
      ptx::mbar_init(&bars[FwdBars::q_empty + i], 1);
      ptx::mbar_init(&bars[FwdBars::k_full + i], 1);
      ptx::mbar_init(&bars[FwdBars::k_empty + i], 1);
      ptx::mbar_init(&bars[FwdBars::v_full + i], 1);
      ptx::mbar_init(&bars[FwdBars::v_empty + i], 1);
      ptx::mbar_init(&bars[FwdBars::s_full + i], 1);
      ptx::mbar_init(&bars[FwdBars::p_full + i], 128);
      ptx::mbar_init(&bars[FwdBars::o_full + i], 1);
      ptx::mbar_init(&bars[FwdBars::o_empty + i], 128);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<512>(tmem_ptr);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (ptx::elect_one()) {
      int ii = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ii) {
        const int qp = item % nqp, h = (item / nqp) % p.H, b = item / (nqp * p.H);
        const int row0 = b * p.N;
        for (int w = 0; w < 2; ++w) {
          ptx::mbar_wait(&bars[FwdBars::q_empty + w], (ii & 1) ^ 1);
          ptx::mbar_expect_tx(&bars[FwdBars::q_full + w], kTile);
          ptx::tma_load_2d(&tmap_q, &bars[FwdBars::q_full + w], smem + w * kTile, h * DH, row0 + qp * 2 * BQ + w * BQ);
        }
        for (int j = 0; j < nkv; ++j) {
          const int g = ii * nkv + j;
          const int st = g % stages;
          const uint32_t ph = (uint32_t)((g / stages) & 1);
          ptx::mbar_wait(&bars[FwdBars::k_empty + st], ph ^ 1);
          ptx::mbar_expect_tx(&bars[FwdBars::k_full + st], (uint32_t)(bkv * 128));
          ptx::tma_load_2d(&tmap_kv, &bars[FwdBars::k_full + st], smem + L.off_k + st * kv_bytes, p.D + h * DH, row0 + j * bkv);
          ptx::mbar_wait(&bars[FwdBars::v_empty + st], ph ^ 1);
          ptx::mbar_expect_tx(&bars[FwdBars::v_full + st], (uint32_t)(bkv * 128));
          ptx::tma_load_2d(&tmap_kv, &bars[FwdBars::v_full + st], smem + L.off_v + st * kv_bytes, 2 * p.D + h * DH, row0 + j * bkv);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (ptx::elect_one()) {
      const uint32_t idesc_s = umma_idesc_bf16(BQ, bkv, false, false);
      const uint32_t idesc_o = umma_idesc_bf16(BQ, DH, false, true);
      const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const int total = my_items * nkv;
      auto issue_s = [&](int w, int g) {
        const int j = g % nkv, ii = g / nkv, st = g % stages;
        if (j == 0) ptx::mbar_wait(&bars[FwdBars::q_full + w], ii & 1);
        ptx::mbar_wait(&bars[FwdBars::k_full + st], (g / stages) & 1);
        ptx::mbar_wait(&bars[FwdBars::o_empty + w], (g & 1) ^ 1);          // TMEM region w free (O of step g-1 read out)
        ptx::tc_fence_after();
        const uint64_t dq = umma_smem_desc_sw128(ptx::smem_u32(smem + w * kTile), 0, 1024);
        const uint64_t dk = umma_smem_desc_sw128(ptx::smem_u32(smem + L.off_k + st * kv_bytes), 0, 1024);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          ptx::umma_f16(tmem_base + w * 256, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, k > 0 ? 1u : 0u);
        ptx::umma_commit(&bars[FwdBars::s_full + w]);
        if (w == 1) ptx::umma_commit(&bars[FwdBars::k_empty + st]);
        if (j == nkv - 1) ptx::umma_commit(&bars[FwdBars::q_empty + w]);
      };
      auto issue_pv = [&](int w, int g) {
        const int st = g % stages;
        ptx::mbar_wait(&bars[FwdBars::p_full + w], g & 1);
        ptx::mbar_wait(&bars[FwdBars::v_full + st], (g / stages) & 1);
        ptx::tc_fence_after();
        const uint32_t sp = ptx::smem_u32(smem + L.off_p + w * L.p_atoms * kTile);
        const uint32_t sv = ptx::smem_u32(smem + L.off_v + st * kv_bytes);
        for (int k = 0; k < bkv / 16; ++k) {
          const uint64_t da = umma_smem_desc_sw128(sp + (k >> 2) * kTile + (k & 3) * 32, 0, 1024);
          const uint64_t db = umma_smem_desc_sw128(sv + k * 2048, kTile, 1024);
          ptx::umma_f16(tmem_base + w * 256, da, db, idesc_o, k > 0 ? 1u : 0u);
        }
        ptx::umma_commit(&bars[FwdBars::o_full + w]);
        if (w == 1) ptx::umma_commit(&bars[FwdBars::v_empty + st]);
      };
      if (total > 0) {
        issue_s(0, 0);
        issue_s(1, 0);
        for (int g = 0; g < total; ++g) {
          issue_pv(0, g);
          if (g + 1 < total) issue_s(0, g + 1);
          issue_pv(1, g);
          if (g + 1 < total) issue_s(1, g + 1);
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== softmax warpgroups =====================
    const int w = (warp - 4) >> 2;                 // 0 = tile A, 1 = tile B
    const int quarter = warp & 3;
    const int lane = tid & 31;
    const int r = quarter * 32 + lane;             // row in tile == TMEM lane
    const uint32_t t_s = tmem_base + w * 256 + ((uint32_t)(quarter * 32) << 16);
    uint8_t* p_smem = smem + L.off_p + w * L.p_atoms * kTile;
    const float sl2 = p.scale * kLog2e;
    const DropKey dkey = drop_key(p.drop_seed, p.drop_p);
    const bool has_drop = p.drop_p > 0.f;
    int g = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int qp = item % nqp, h = (item / nqp) % p.H, b = item / (nqp * p.H);
      const int q_tile0 = qp * 2 * BQ + w * BQ;
      const int qi = q_tile0 + r;
      const bool warp_active = q_tile0 + quarter * 32 < p.N;     // warp-uniform: at least one valid query row
      float m_run = -INFINITY, l_run = 0.f;
      float o_acc[kSingle ? 1 : DH];
      if constexpr (!kSingle) {
#pragma unroll
        for (int d = 0; d < DH; ++d) o_acc[d] = 0.f;
      }
      for (int j = 0; j < nkv; ++j, ++g) {
        ptx::mbar_wait(&bars[FwdBars::s_full + w], g & 1);
        ptx::tc_fence_after();
        float alpha = 1.f;
        if (warp_active) {
          const int kv_valid = min(bkv, p.N - j * bkv);
          const int nch = (kv_valid + 31) / 32;
          float mx = -INFINITY;
#pragma unroll 1
          for (int c = 0; c < nch; ++c) {
            uint32_t rr[32];
            ptx::tmem_ld_x32(t_s + c * 32, rr);
            ptx::tmem_ld_wait();
            if (c * 32 + 32 <= kv_valid) {
#pragma unroll
              for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(rr[i]));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (c * 32 + i < kv_valid) mx = fmaxf(mx, __uint_as_float(rr[i]));
            }
          }
          const float m_new = fmaxf(m_run, mx);
          alpha = ex2_approx((m_run - m_new) * sl2);   // m_run = -inf on the first tile -> 0
          const float mb = m_new * sl2;
          float psum = 0.f;
          const unsigned long long drop_row = (((unsigned long long)(b * p.H + h) * p.N + (unsigned long long)qi) * p.N) + (unsigned long long)(j * bkv);
          const int nch_all = (bkv + 31) / 32;         // P columns read by the PV MMA: [0, bkv)
#pragma unroll 1
          for (int c = 0; c < nch_all; ++c) {
            float pv[32];
            if (c < nch) {
              uint32_t rr[32];
              ptx::tmem_ld_x32(t_s + c * 32, rr);
              ptx::tmem_ld_wait();
              if (c * 32 + 32 <= kv_valid) {
#pragma unroll
                for (int i = 0; i < 32; ++i) { pv[i] = ex2_approx(fmaf(__uint_as_float(rr[i]), sl2, -mb)); psum += pv[i]; }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  pv[i] = (c * 32 + i < kv_valid) ? ex2_approx(fmaf(__uint_as_float(rr[i]), sl2, -mb)) : 0.f;
                  psum += pv[i];
                }
              }
              if (has_drop) drop_apply<32>(pv, dkey, drop_row + (unsigned long long)(c * 32));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) pv[i] = 0.f;
            }
            store_p_chunk(p_smem, r, c, pv);
          }
          l_run = l_run * alpha + psum;
          m_run = m_new;
        }
        ptx::tc_fence_before();
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&bars[FwdBars::p_full + w]);
        ptx::mbar_wait(&bars[FwdBars::o_full + w], g & 1);
        ptx::tc_fence_after();
        if (warp_active) {
          if constexpr (kSingle) {
            // single key tile: O is final, normalise and store straight from TMEM
            const float inv_l = 1.0f / l_run;
#pragma unroll
            for (int c = 0; c < DH / 32; ++c) {
              uint32_t rr[32];
              ptx::tmem_ld_x32(t_s + c * 32, rr);
              ptx::tmem_ld_wait();
              if (qi < p.N) {
                __nv_bfloat16* op = p.out + (long long)(b * p.N + qi) * p.D + h * DH + c * 32;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  uint4 o;
                  o.x = ptx::pack_bf16(__uint_as_float(rr[q * 8 + 0]) * inv_l, __uint_as_float(rr[q * 8 + 1]) * inv_l);
                  o.y = ptx::pack_bf16(__uint_as_float(rr[q * 8 + 2]) * inv_l, __uint_as_float(rr[q * 8 + 3]) * inv_l);
                  o.z = ptx::pack_bf16(__uint_as_float(rr[q * 8 + 4]) * inv_l, __uint_as_float(rr[q * 8 + 5]) * inv_l);
                  o.w = ptx::pack_bf16(__uint_as_float(rr[q * 8 + 6]) * inv_l, __uint_as_float(rr[q * 8 + 7]) * inv_l);
                  reinterpret_cast<uint4*>(op)[q] = o;
                }
              }
            }
          } else {
#pragma unroll
            for (int c = 0; c < DH / 32; ++c) {
              uint32_t rr[32];
              ptx::tmem_ld_x32(t_s + c * 32, rr);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = fmaf(o_acc[c * 32 + i], alpha, __uint_as_float(rr[i]));
            }
          }
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&bars[FwdBars::o_empty + w]);
      }
      if (qi < p.N) {
        if constexpr (!kSingle) {
          const float inv_l = 1.0f / l_run;
          __nv_bfloat16* op = p.out + (long long)(b * p.N + qi) * p.D + h * DH;
#pragma unroll
          for (int q = 0; q < DH / 8; ++q) {
            uint4 o;
            o.x = ptx::pack_bf16(o_acc[q * 8 + 0] * inv_l, o_acc[q * 8 + 1] * inv_l);
            o.y = ptx::pack_bf16(o_acc[q * 8 + 2] * inv_l, o_acc[q * 8 + 3] * inv_l);
            o.z = ptx::pack_bf16(o_acc[q * 8 + 4] * inv_l, o_acc[q * 8 + 5] * inv_l);
            o.w = ptx::pack_bf16(o_acc[q * 8 + 6] * inv_l, o_acc[q * 8 + 7] * inv_l);
            reinterpret_cast<uint4*>(op)[q] = o;
          }
        }
        if (p.lse) p.lse[((long long)b * p.H + h) * p.N + qi] = m_run * p.scale + logf(l_run);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

// ================================================= backward =================================================
struct BwdSmem {
  static constexpr int kK = 0;
  static constexpr int kV = kK + kTile;
  static constexpr int kQ = kV + kTile;           // 2 buffers
  static constexpr int kDO = kQ + 2 * kTile;      // 2 buffers
  static constexpr int kP = kDO + 2 * kTile;      // 128 x 128 bf16 (rows = query, cols = key)
  static constexpr int kDS = kP + 2 * kTile;
  static constexpr int kBar = kDS + 2 * kTile;
  static constexpr int kTotal = kBar + 8 * 8 + 16 + 1024;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(128) attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv,
                                                       const __grid_constant__ CUtensorMap tmap_do, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_kv = reinterpret_cast<uint64_t*>(smem + BwdSmem::kBar);
  uint64_t* bar_qdo = bar_kv + 1;   // [2]
  uint64_t* bar_m1 = bar_kv + 3;
  uint64_t* bar_m2 = bar_kv + 4;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_kv + 5);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int jt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int kv0 = jt * BKV;
  const int nq = (p.N + BQ - 1) / BQ;
  const int row0 = b * p.N;

  if (tid == 0) {
    ptx::prefetch_tmap(&tmap_qkv);
    ptx::prefetch_tmap(&tmap_do);
    ptx::mbar_init(bar_kv, 1);
    ptx::mbar_init(&bar_qdo[0], 1);
    ptx::mbar_init(&bar_qdo[1], 1);
    ptx::mbar_init(bar_m1, 1);
    ptx::mbar_init(bar_m2, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc<512>(tmem_ptr);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + 128, tmem_dv = tmem_base + 256, tmem_dk = tmem_base + 320,
                 tmem_dq = tmem_base + 384;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;

  if (tid == 0) {
    ptx::mbar_expect_tx(bar_kv, 2 * kTile);
    ptx::tma_load_2d(&tmap_qkv, bar_kv, smem + BwdSmem::kK, p.D + h * DH, row0 + kv0);
    ptx::tma_load_2d(&tmap_qkv, bar_kv, smem + BwdSmem::kV, 2 * p.D + h * DH, row0 + kv0);
    ptx::mbar_expect_tx(&bar_qdo[0], 2 * kTile);
    ptx::tma_load_2d(&tmap_qkv, &bar_qdo[0], smem + BwdSmem::kQ, h * DH, row0);
    ptx::tma_load_2d(&tmap_do, &bar_qdo[0], smem + BwdSmem::kDO, h * DH, row0);
  }

  const float sl2 = p.scale * kLog2e;
  const int r = tid;
  const int kv_valid = min(BKV, p.N - kv0);
  const uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, false, false);     // S, dP : K-major x K-major
  const uint32_t idesc_t = umma_idesc_bf16(BKV, DH, true, true);       // dV, dK: MN-major A (P^T / dS^T), MN-major B
  const uint32_t idesc_q = umma_idesc_bf16(BQ, DH, false, true);       // dQ    : K-major A (dS), MN-major B (K)
  const DropKey dkey = drop_key(p.drop_seed, p.drop_p);
  const float inv_keep = dkey.inv_keep;

  ptx::mbar_wait(bar_kv, 0);
  for (int i = 0; i < nq; ++i) {
    const int buf = i & 1;
    const int q0 = i * BQ;
    if (tid == 0 && i + 1 < nq) {
      ptx::mbar_expect_tx(&bar_qdo[buf ^ 1], 2 * kTile);
      ptx::tma_load_2d(&tmap_qkv, &bar_qdo[buf ^ 1], smem + BwdSmem::kQ + (buf ^ 1) * kTile, h * DH, row0 + q0 + BQ);
      ptx::tma_load_2d(&tmap_do, &bar_qdo[buf ^ 1], smem + BwdSmem::kDO + (buf ^ 1) * kTile, h * DH, row0 + q0 + BQ);
    }
    ptx::mbar_wait(&bar_qdo[buf], (i >> 1) & 1);
    if (tid == 0) {
      ptx::tc_fence_after();
      const uint64_t dq = umma_smem_desc_sw128(ptx::smem_u32(smem + BwdSmem::kQ + buf * kTile), 0, 1024);
      const uint64_t dk = umma_smem_desc_sw128(ptx::smem_u32(smem + BwdSmem::kK), 0, 1024);
      const uint64_t ddo = umma_smem_desc_sw128(ptx::smem_u32(smem + BwdSmem::kDO + buf * kTile), 0, 1024);
      const uint64_t dv = umma_smem_desc_sw128(ptx::smem_u32(smem + BwdSmem::kV), 0, 1024);
#pragma unroll
      for (int k = 0; k < DH / 16; ++k) ptx::umma_f16(tmem_s, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < DH / 16; ++k) ptx::umma_f16(tmem_dp, ddo + (uint64_t)(k * 2), dv + (uint64_t)(k * 2), idesc_s, k > 0 ? 1u : 0u);
      ptx::umma_commit(bar_m1);
    }
    // per-row statistics (overlaps the MMAs): lse and delta = rowsum(dO * O)
    const int qi = q0 + r;
    const bool q_ok = qi < p.N;
    float lse2 = 0.f, delta = 0.f;
    if (q_ok) {
      lse2 = p.lse[((long long)b * p.H + h) * p.N + qi] * kLog2e;
      const uint4* orow = reinterpret_cast<const uint4*>(p.o + (long long)(row0 + qi) * p.D + h * DH);
      const uint4* drow = reinterpret_cast<const uint4*>(p.dout + (long long)(row0 + qi) * p.D + h * DH);
#pragma unroll
      for (int q = 0; q < DH / 8; ++q) {
        const uint4 a = __ldg(orow + q), d = __ldg(drow + q);
        delta += ptx::bf16_lo(a.x) * ptx::bf16_lo(d.x) + ptx::bf16_hi(a.x) * ptx::bf16_hi(d.x);
        delta += ptx::bf16_lo(a.y) * ptx::bf16_lo(d.y) + ptx::bf16_hi(a.y) * ptx::bf16_hi(d.y);
        delta += ptx::bf16_lo(a.z) * ptx::bf16_lo(d.z) + ptx::bf16_hi(a.z) * ptx::bf16_hi(d.z);
        delta += ptx::bf16_lo(a.w) * ptx::bf16_lo(d.w) + ptx::bf16_hi(a.w) * ptx::bf16_hi(d.w);
      }
    }
    ptx::mbar_wait(bar_m1, i & 1);
    ptx::tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BKV / 32; ++c) {
      uint32_t rs[32], rd[32];
      ptx::tmem_ld_x32(tmem_s + lane_off + c * 32, rs);
      ptx::tmem_ld_x32(tmem_dp + lane_off + c * 32, rd);
      ptx::tmem_ld_wait();
      float pv[32], dsv[32];
      const unsigned long long base = (((unsigned long long)(b * p.H + h) * p.N + (unsigned long long)qi) * p.N) + (unsigned long long)(kv0 + c * 32);
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const bool ok = q_ok && (c * 32 + e < kv_valid);
        const float pr = ok ? exp2f(__uint_as_float(rs[e]) * sl2 - lse2) : 0.f;
        float dp = __uint_as_float(rd[e]);
        float pd = pr;
        if (p.drop_p > 0.f) {
          const bool keep = drop_keep(dkey, base + e);
          pd = keep ? pr * inv_keep : 0.f;
          dp = keep ? dp * inv_keep : 0.f;
        }
        pv[e] = pd;
        dsv[e] = ok ? pr * (dp - delta) * p.scale : 0.f;
      }
      store_p_chunk(smem + BwdSmem::kP, r, c, pv);
      store_p_chunk(smem + BwdSmem::kDS, r, c, dsv);
    }
    ptx::tc_fence_before();
    ptx::fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      ptx::tc_fence_after();
      const uint32_t sp = ptx::smem_u32(smem + BwdSmem::kP);
      const uint32_t sds = ptx::smem_u32(smem + BwdSmem::kDS);
      const uint32_t sq = ptx::smem_u32(smem + BwdSmem::kQ + buf * kTile);
      const uint32_t sdo = ptx::smem_u32(smem + BwdSmem::kDO + buf * kTile);
      const uint32_t sk = ptx::smem_u32(smem + BwdSmem::kK);
      // dV[kv, d] += sum_q P[q, kv] dO[q, d]   ;   dK[kv, d] += sum_q dS[q, kv] Q[q, d]     (reduction over q rows)
#pragma unroll
      for (int k = 0; k < BQ / 16; ++k) {
        const uint64_t a_p = umma_smem_desc_sw128(sp + k * 2048, kTile, 1024);
        const uint64_t b_do = umma_smem_desc_sw128(sdo + k * 2048, kTile, 1024);
        ptx::umma_f16(tmem_dv, a_p, b_do, idesc_t, (i > 0 || k > 0) ? 1u : 0u);
      }
#pragma unroll
      for (int k = 0; k < BQ / 16; ++k) {
        const uint64_t a_ds = umma_smem_desc_sw128(sds + k * 2048, kTile, 1024);
        const uint64_t b_q = umma_smem_desc_sw128(sq + k * 2048, kTile, 1024);
        ptx::umma_f16(tmem_dk, a_ds, b_q, idesc_t, (i > 0 || k > 0) ? 1u : 0u);
      }
      // dQ[q, d] = sum_kv dS[q, kv] K[kv, d]
#pragma unroll
      for (int k = 0; k < BKV / 16; ++k) {
        const uint64_t a_ds = umma_smem_desc_sw128(sds + (k >> 2) * kTile + (k & 3) * 32, 0, 1024);
        const uint64_t b_k = umma_smem_desc_sw128(sk + k * 2048, kTile, 1024);
        ptx::umma_f16(tmem_dq, a_ds, b_k, idesc_q, k > 0 ? 1u : 0u);
      }
      ptx::umma_commit(bar_m2);
    }
    ptx::mbar_wait(bar_m2, i & 1);
    ptx::tc_fence_after();
#pragma unroll
    for (int c = 0; c < DH / 32; ++c) {
      uint32_t rr[32];
      ptx::tmem_ld_x32(tmem_dq + lane_off + c * 32, rr);
      ptx::tmem_ld_wait();
      if (q_ok) {
        float* dst = p.dq_acc + (long long)(row0 + qi) * p.D + h * DH + c * 32;
#pragma unroll
        for (int e = 0; e < 32; e += 4)
          red_add_v4(dst + e, __uint_as_float(rr[e]), __uint_as_float(rr[e + 1]), __uint_as_float(rr[e + 2]), __uint_as_float(rr[e + 3]));
      }
    }
    ptx::tc_fence_before();
  }

  // dV / dK: TMEM lane == key row of this tile
  const int kvi = kv0 + r;
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const uint32_t t = which == 0 ? tmem_dk : tmem_dv;
    __nv_bfloat16* dst = p.dqkv + (long long)(row0 + kvi) * (3 * p.D) + (which == 0 ? p.D : 2 * p.D) + h * DH;
#pragma unroll
    for (int c = 0; c < DH / 32; ++c) {
      uint32_t rr[32];
      ptx::tmem_ld_x32(t + lane_off + c * 32, rr);
      ptx::tmem_ld_wait();
      if (kvi < p.N) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = ptx::pack_bf16(__uint_as_float(rr[q * 8 + 0]), __uint_as_float(rr[q * 8 + 1]));
          o.y = ptx::pack_bf16(__uint_as_float(rr[q * 8 + 2]), __uint_as_float(rr[q * 8 + 3]));
          o.z = ptx::pack_bf16(__uint_as_float(rr[q * 8 + 4]), __uint_as_float(rr[q * 8 + 5]));
          o.w = ptx::pack_bf16(__uint_as_float(rr[q * 8 + 6]), __uint_as_float(rr[q * 8 + 7]));
          reinterpret_cast<uint4*>(dst + c * 32)[q] = o;
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

// dqkv[:, 0:D] = bf16(dq_acc)
__global__ void dq_finalize_kernel(const float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dqkv, long long rows, int D) {
  const long long total = rows * (D / 8);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / (D / 8);
    const int c = (int)(i % (D / 8)) * 8;
    const float4 a = *reinterpret_cast<const float4*>(dq_acc + m * D + c);
    const float4 bq = *reinterpret_cast<const float4*>(dq_acc + m * D + c + 4);
    uint4 o;
    o.x = ptx::pack_bf16(a.x, a.y); o.y = ptx::pack_bf16(a.z, a.w);
    o.z = ptx::pack_bf16(bq.x, bq.y); o.w = ptx::pack_bf16(bq.z, bq.w);
    *reinterpret_cast<uint4*>(dqkv + m * (3ll * D) + c) = o;
  }
}

int check_shape(int B, int H, int N, int D) {
  SFC_REQUIRE(B > 0 && H > 0 && N > 0 && D == H * DH, "attention: only head_dim = 64 is supported (D=%d, heads=%d)", D, H);
  SFC_REQUIRE((long long)B * N < (1ll << 31), "attention: too many rows");
  return 0;
}

}  // namespace

extern "C" int sfc_attn_fwd(const void* qkv, void* out, float* lse, int B, int H, int N, int D, float scale, float drop_p,
                            unsigned long long drop_seed, cudaStream_t stream) {
  if (int e = check_shape(B, H, N, D)) return e;
  SFC_REQUIRE(qkv && out, "sfc_attn_fwd: null pointer");
  SFC_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "sfc_attn_fwd: dropout p out of range");
  const FwdLayout L = fwd_layout(N);
  SFC_REQUIRE(L.total <= 232448, "sfc_attn_fwd: shared-memory plan of %d bytes exceeds 227 KB (N=%d)", L.total, N);
  CUtensorMap tq, tkv;
  if (int e = sfc_make_tmap_2d(&tq, qkv, 2, (uint64_t)3 * D, (uint64_t)B * N, (uint64_t)3 * D * 2, DH, BQ, true)) return e;
  if (int e = sfc_make_tmap_2d(&tkv, qkv, 2, (uint64_t)3 * D, (uint64_t)B * N, (uint64_t)3 * D * 2, DH, (uint32_t)L.bkv, true)) return e;
  AttnParams p{};
  p.B = B; p.H = H; p.N = N; p.D = D; p.scale = scale; p.drop_p = drop_p; p.drop_seed = drop_seed;
  p.out = (__nv_bfloat16*)out; p.lse = lse;
  static bool configured = false;
  if (!configured) {
    SFC_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    SFC_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    configured = true;
  }
  const long long items = (long long)B * H * ((N + 2 * BQ - 1) / (2 * BQ));
  const int grid = (int)(items < sfc_num_sms() ? items : sfc_num_sms());
  if (L.nkv == 1) attn_fwd_kernel<true><<<grid, kFwdThreads, L.total, stream>>>(tq, tkv, p, L);
  else attn_fwd_kernel<false><<<grid, kFwdThreads, L.total, stream>>>(tq, tkv, p, L);
  SFC_LAUNCH_OK();
  return 0;
}

extern "C" size_t sfc_attn_bwd_scratch_bytes(int B, int N, int D) { return (size_t)B * N * D * sizeof(float); }

// dq_acc scratch (fp32 [B*N, D]) is zeroed here (cudaMemsetAsync on the caller's stream).
extern "C" int sfc_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, void* scratch,
                            size_t scratch_bytes, int B, int H, int N, int D, float scale, float drop_p,
                            unsigned long long drop_seed, cudaStream_t stream) {
  if (int e = check_shape(B, H, N, D)) return e;
  SFC_REQUIRE(qkv && out && dout && lse && dqkv, "sfc_attn_bwd: null pointer");
  const size_t need = (size_t)B * N * D * sizeof(float);
  SFC_REQUIRE(scratch && scratch_bytes >= need, "sfc_attn_bwd: scratch too small (%zu < %zu)", scratch_bytes, need);
  SFC_CUDA_OK(cudaMemsetAsync(scratch, 0, need, stream));
  CUtensorMap tq, tdo;
  if (int e = sfc_make_tmap_2d(&tq, qkv, 2, (uint64_t)3 * D, (uint64_t)B * N, (uint64_t)3 * D * 2, DH, BQ, true)) return e;
  if (int e = sfc_make_tmap_2d(&tdo, dout, 2, (uint64_t)D, (uint64_t)B * N, (uint64_t)D * 2, DH, BQ, true)) return e;
  AttnParams p{};
  p.B = B; p.H = H; p.N = N; p.D = D; p.scale = scale; p.drop_p = drop_p; p.drop_seed = drop_seed;
  p.lse = const_cast<float*>(lse); p.o = (const __nv_bfloat16*)out; p.dout = (const __nv_bfloat16*)dout;
  p.dqkv = (__nv_bfloat16*)dqkv; p.dq_acc = (float*)scratch;
  static bool configured = false;
  if (!configured) {
    SFC_CUDA_OK(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::kTotal));
    configured = true;
  }
  dim3 grid((N + BKV - 1) / BKV, H, B);
  attn_bwd_kernel<<<grid, 128, BwdSmem::kTotal, stream>>>(tq, tdo, p);
  SFC_LAUNCH_OK();
  const long long rows = (long long)B * N;
  long long blocks = sfc_ceil_div64(rows * (D / 8), 256);
  const long long cap = 16ll * sfc_num_sms();
  if (blocks > cap) blocks = cap;
  dq_finalize_kernel<<<(unsigned)blocks, 256, 0, stream>>>((const float*)scratch, (__nv_bfloat16*)dqkv, rows, D);
  SFC_LAUNCH_OK();
  return 0;
}
