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
#include "gemm_epilogue.cuh"   // drop_keep
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
struct FwdSmem {
  static constexpr int kQ = 0;
  static constexpr int kK = kQ + kTile;          // 2 buffers
  static constexpr int kV = kK + 2 * kTile;      // 2 buffers
  static constexpr int kP = kV + 2 * kTile;      // 128 x 128 bf16
  static constexpr int kBar = kP + 2 * kTile;
  static constexpr int kTotal = kBar + 8 * 8 + 16;
};

__global__ void __launch_bounds__(128, 2) attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnParams p) {
  // no alignment slack here: two CTAs per SM need every byte of the 228 KB; the 1024-byte alignment that the
  // 128-byte swizzle needs comes from the declaration and is verified at run time.
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* smem = smem_dyn;
  if ((ptx::smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("sfcvit: attn_fwd dynamic smem not 1024-byte aligned\n");
    __trap();
  }
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + FwdSmem::kBar);
  uint64_t* bar_kv = bar_q + 1;   // [2]
  uint64_t* bar_s = bar_q + 3;
  uint64_t* bar_o = bar_q + 4;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_q + 5);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * BQ;
  const int nkv = (p.N + BKV - 1) / BKV;
  const int row0 = b * p.N;

  if (tid == 0) {
    ptx::prefetch_tmap(&tmap_qkv);
    ptx::mbar_init(bar_q, 1);
    ptx::mbar_init(&bar_kv[0], 1);
    ptx::mbar_init(&bar_kv[1], 1);
    ptx::mbar_init(bar_s, 1);
    ptx::mbar_init(bar_o, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc<256>(tmem_ptr);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;

  if (tid == 0) {
    ptx::mbar_expect_tx(bar_q, kTile);
    ptx::tma_load_2d(&tmap_qkv, bar_q, smem + FwdSmem::kQ, h * DH, row0 + q0);
    ptx::mbar_expect_tx(&bar_kv[0], 2 * kTile);
    ptx::tma_load_2d(&tmap_qkv, &bar_kv[0], smem + FwdSmem::kK, p.D + h * DH, row0);
    ptx::tma_load_2d(&tmap_qkv, &bar_kv[0], smem + FwdSmem::kV, 2 * p.D + h * DH, row0);
  }

  const float sl2 = p.scale * kLog2e;
  float m_run = -INFINITY, l_run = 0.f;
  float o_acc[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) o_acc[d] = 0.f;
  const int r = tid;                       // query row within the tile == TMEM lane
  const int qi = q0 + r;
  const uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, false, false);
  const uint32_t idesc_o = umma_idesc_bf16(BQ, DH, false, true);
  const DropKey dkey = drop_key(p.drop_seed, p.drop_p);
  ptx::mbar_wait(bar_q, 0);
  for (int j = 0; j < nkv; ++j) {
    const int buf = j & 1;
    if (tid == 0 && j + 1 < nkv) {
      ptx::mbar_expect_tx(&bar_kv[buf ^ 1], 2 * kTile);
      ptx::tma_load_2d(&tmap_qkv, &bar_kv[buf ^ 1], smem + FwdSmem::kK + (buf ^ 1) * kTile, p.D + h * DH, row0 + (j + 1) * BKV);
      ptx::tma_load_2d(&tmap_qkv, &bar_kv[buf ^ 1], smem + FwdSmem::kV + (buf ^ 1) * kTile, 2 * p.D + h * DH, row0 + (j + 1) * BKV);
    }
    ptx::mbar_wait(&bar_kv[buf], (j >> 1) & 1);
    if (tid == 0) {
      ptx::tc_fence_after();
      const uint64_t dq = umma_smem_desc_sw128(ptx::smem_u32(smem + FwdSmem::kQ), 0, 1024);
      const uint64_t dk = umma_smem_desc_sw128(ptx::smem_u32(smem + FwdSmem::kK + buf * kTile), 0, 1024);
#pragma unroll
      for (int k = 0; k < DH / 16; ++k) ptx::umma_f16(tmem_s, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, k > 0 ? 1u : 0u);
      ptx::umma_commit(bar_s);
    }
    ptx::mbar_wait(bar_s, j & 1);
    ptx::tc_fence_after();

    const int kv_valid = min(BKV, p.N - j * BKV);     // number of valid key columns in this tile (>= 1)
    // pass 1: row maximum
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < BKV / 32; ++c) {
      if (c * 32 >= kv_valid) break;
      uint32_t rr[32];
      ptx::tmem_ld_x32(tmem_s + lane_off + c * 32, rr);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c * 32 + i < kv_valid) mx = fmaxf(mx, __uint_as_float(rr[i]));
    }
    const float m_new = fmaxf(m_run, mx);
    const float alpha = exp2f((m_run - m_new) * sl2);   // m_run = -inf on the first tile -> 0
    const float mb = m_new * sl2;
    // pass 2: probabilities, row sum, bf16 P -> smem
    float psum = 0.f;
#pragma unroll 1
    for (int c = 0; c < BKV / 32; ++c) {
      float pv[32];
      if (c * 32 < kv_valid) {
        uint32_t rr[32];
        ptx::tmem_ld_x32(tmem_s + lane_off + c * 32, rr);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float e = (c * 32 + i < kv_valid) ? exp2f(__uint_as_float(rr[i]) * sl2 - mb) : 0.f;
          psum += e;
          pv[i] = e;
        }
        if (p.drop_p > 0.f) {
          const unsigned long long base = (((unsigned long long)(b * p.H + h) * p.N + (unsigned long long)qi) * p.N) + (unsigned long long)(j * BKV + c * 32);
          drop_apply<32>(pv, dkey, base);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) pv[i] = 0.f;
      }
      store_p_chunk(smem + FwdSmem::kP, r, c, pv);
    }
    l_run = l_run * alpha + psum;
    m_run = m_new;
    ptx::tc_fence_before();
    ptx::fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      ptx::tc_fence_after();
      const uint32_t sp = ptx::smem_u32(smem + FwdSmem::kP);
      const uint32_t sv = ptx::smem_u32(smem + FwdSmem::kV + buf * kTile);
#pragma unroll
      for (int k = 0; k < BKV / 16; ++k) {
        const uint64_t da = umma_smem_desc_sw128(sp + (k >> 2) * kTile + (k & 3) * 32, 0, 1024);
        const uint64_t db = umma_smem_desc_sw128(sv + k * 2048, kTile, 1024);
        ptx::umma_f16(tmem_o, da, db, idesc_o, k > 0 ? 1u : 0u);
      }
      ptx::umma_commit(bar_o);
    }
    ptx::mbar_wait(bar_o, j & 1);
    ptx::tc_fence_after();
#pragma unroll
    for (int c = 0; c < DH / 32; ++c) {
      uint32_t rr[32];
      ptx::tmem_ld_x32(tmem_o + lane_off + c * 32, rr);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = o_acc[c * 32 + i] * alpha + __uint_as_float(rr[i]);
    }
    ptx::tc_fence_before();
  }

  if (qi < p.N) {
    const float inv_l = 1.0f / l_run;
    __nv_bfloat16* op = p.out + (long long)(row0 + qi) * p.D + h * DH;
#pragma unroll
    for (int q = 0; q < DH / 8; ++q) {
      uint4 o;
      o.x = ptx::pack_bf16(o_acc[q * 8 + 0] * inv_l, o_acc[q * 8 + 1] * inv_l);
      o.y = ptx::pack_bf16(o_acc[q * 8 + 2] * inv_l, o_acc[q * 8 + 3] * inv_l);
      o.z = ptx::pack_bf16(o_acc[q * 8 + 4] * inv_l, o_acc[q * 8 + 5] * inv_l);
      o.w = ptx::pack_bf16(o_acc[q * 8 + 6] * inv_l, o_acc[q * 8 + 7] * inv_l);
      reinterpret_cast<uint4*>(op)[q] = o;
    }
    if (p.lse) p.lse[((long long)b * p.H + h) * p.N + qi] = m_run * p.scale + logf(l_run);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<256>(tmem_base);
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
  CUtensorMap tq;
  if (int e = sfc_make_tmap_2d(&tq, qkv, 2, (uint64_t)3 * D, (uint64_t)B * N, (uint64_t)3 * D * 2, DH, BQ, true)) return e;
  AttnParams p{};
  p.B = B; p.H = H; p.N = N; p.D = D; p.scale = scale; p.drop_p = drop_p; p.drop_seed = drop_seed;
  p.out = (__nv_bfloat16*)out; p.lse = lse;
  static bool configured = false;
  if (!configured) {
    SFC_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::kTotal));
    configured = true;
  }
  dim3 grid((N + BQ - 1) / BQ, H, B);
  attn_fwd_kernel<<<grid, 128, FwdSmem::kTotal, stream>>>(tq, p);
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
