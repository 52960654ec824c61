// K7 — linear resampling of a token stream along the sequence axis, written straight into a column slice of the
// feature-axis concatenation the hierarchical tokenizers feed to their fusion Linear
// (/root/reference/src/tokenizers/multiscale/multi_hilbert.py:30-40 and its morton/peano/moore/zigzag/onion copies:
//  F.interpolate(x.transpose(1, 2), size=n, mode="linear", align_corners=False).transpose(1, 2), then torch.cat(dim=-1)).
// Replaces two transposes + upsample_linear1d + cat (four passes over the stream) by one pass; equal lengths
// (main.py's [16, 4, 1]) degenerate to a strided copy into the slice.
//
// Index rule (ATen area_pixel_compute_source_index, align_corners = false): pos = max(0, (Ns / Nd) * (t + 0.5) - 0.5),
// i0 = floor(pos), i1 = i0 + (i0 < Ns - 1), w1 = pos - i0, w0 = 1 - w1, all in fp32.
// HBM-bound: reads <= 2 source rows (L2-resident neighbours) and writes one row per output token; one warp per row,
// 16-byte vectors. The backward is the transposed operator in gather form (deterministic, no atomics): one warp per
// SOURCE row walks the few output rows that reference it.
#include "common.cuh"
#include "sfcvit.h"

namespace {

constexpr int kWarps = 8;

__device__ __forceinline__ void interp_index(int t, float scale, int Ns, int& i0, int& i1, float& w0, float& w1) {
  float pos = scale * ((float)t + 0.5f) - 0.5f;
  pos = pos < 0.0f ? 0.0f : pos;
  i0 = (int)pos;
  if (i0 > Ns - 1) i0 = Ns - 1;
  i1 = i0 + (i0 < Ns - 1 ? 1 : 0);
  w1 = pos - (float)i0;
  w0 = 1.0f - w1;
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = ptx::bf16_lo(u.x); f[1] = ptx::bf16_hi(u.x); f[2] = ptx::bf16_lo(u.y); f[3] = ptx::bf16_hi(u.y);
  f[4] = ptx::bf16_lo(u.z); f[5] = ptx::bf16_hi(u.z); f[6] = ptx::bf16_lo(u.w); f[7] = ptx::bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = ptx::pack_bf16(f[0], f[1]); u.y = ptx::pack_bf16(f[2], f[3]);
  u.z = ptx::pack_bf16(f[4], f[5]); u.w = ptx::pack_bf16(f[6], f[7]);
  return u;
}

__global__ void __launch_bounds__(kWarps * 32)
interp_concat_fwd_kernel(const __nv_bfloat16* __restrict__ src, long long ld_src, int B, int Ns, int D,
                         __nv_bfloat16* __restrict__ dst, long long ld_dst, int Nd, float scale) {
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)B * Nd;
  for (long long row = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * kWarps) {
    const int b = (int)(row / Nd), t = (int)(row % Nd);
    int i0, i1;
    float w0, w1;
    interp_index(t, scale, Ns, i0, i1, w0, w1);
    const uint4* r0 = reinterpret_cast<const uint4*>(src + ((long long)b * Ns + i0) * ld_src);
    const uint4* r1 = reinterpret_cast<const uint4*>(src + ((long long)b * Ns + i1) * ld_src);
    uint4* o = reinterpret_cast<uint4*>(dst + row * ld_dst);
    for (int v = lane; v < D / 8; v += 32) {
      const uint4 a = __ldg(r0 + v);
      if (w1 == 0.0f) { o[v] = a; continue; }                 // exact copy (equal lengths, clamped ends)
      float fa[8], fb[8];
      unpack8(a, fa);
      unpack8(__ldg(r1 + v), fb);
#pragma unroll
      for (int e = 0; e < 8; ++e) fa[e] = w0 * fa[e] + w1 * fb[e];
      o[v] = pack8(fa);
    }
  }
}

__global__ void __launch_bounds__(kWarps * 32)
interp_concat_bwd_kernel(const __nv_bfloat16* __restrict__ ddst, long long ld_dst, int B, int Nd, int D,
                         __nv_bfloat16* __restrict__ dsrc, long long ld_src, int Ns, float scale) {
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)B * Ns;
  const float inv = 1.0f / scale;
  for (long long row = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * kWarps) {
    const int b = (int)(row / Ns), s = (int)(row % Ns);
    // output rows t with i0(t) == s or i1(t) == s have pos(t) in (s - 1, s + 1): a conservative window, filtered below
    int t_lo = (int)floorf(((float)s - 0.5f) * inv - 0.5f) - 1;
    int t_hi = (int)ceilf(((float)s + 1.5f) * inv - 0.5f) + 1;
    if (t_lo < 0 || s == 0) t_lo = 0;
    if (t_hi > Nd - 1) t_hi = Nd - 1;
    uint4* o = reinterpret_cast<uint4*>(dsrc + row * ld_src);
    for (int v = lane; v < D / 8; v += 32) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int t = t_lo; t <= t_hi; ++t) {
        int i0, i1;
        float w0, w1;
        interp_index(t, scale, Ns, i0, i1, w0, w1);
        const float w = (i0 == s ? w0 : 0.0f) + (i1 == s ? w1 : 0.0f);
        if (w == 0.0f) continue;
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(ddst + ((long long)b * Nd + t) * ld_dst) + v), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += w * f[e];
      }
      o[v] = pack8(acc);
    }
  }
}

int check_args(const void* a, long long lda, const void* b, long long ldb, int B, int Ns, int Nd, int D, const char* who) {
  SFC_REQUIRE(a && b, "%s: null pointer", who);
  SFC_REQUIRE(B > 0 && Ns > 0 && Nd > 0 && D > 0, "%s: bad shape", who);
  SFC_REQUIRE(D % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && lda >= D && ldb >= D, "%s: feature dim and row strides must be multiples of 8 elements", who);
  SFC_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0, "%s: pointers must be 16-byte aligned", who);
  return 0;
}

int grid_for(long long rows) {
  long long blocks = sfc_ceil_div64(rows, kWarps);
  const long long cap = 16ll * sfc_num_sms();
  return (int)(blocks < cap ? blocks : cap);
}

}  // namespace

extern "C" int sfc_interp_concat_fwd(const void* src, long long ld_src, int B, int Ns, int D, void* dst, long long ld_dst, int Nd,
                                     cudaStream_t stream) {
  if (int e = check_args(src, ld_src, dst, ld_dst, B, Ns, Nd, D, "sfc_interp_concat_fwd")) return e;
  interp_concat_fwd_kernel<<<grid_for((long long)B * Nd), kWarps * 32, 0, stream>>>(
      (const __nv_bfloat16*)src, ld_src, B, Ns, D, (__nv_bfloat16*)dst, ld_dst, Nd, (float)Ns / (float)Nd);
  SFC_LAUNCH_OK();
  return 0;
}

extern "C" int sfc_interp_concat_bwd(const void* ddst, long long ld_dst, int B, int Nd, int D, void* dsrc, long long ld_src, int Ns,
                                     cudaStream_t stream) {
  if (int e = check_args(ddst, ld_dst, dsrc, ld_src, B, Ns, Nd, D, "sfc_interp_concat_bwd")) return e;
  interp_concat_bwd_kernel<<<grid_for((long long)B * Ns), kWarps * 32, 0, stream>>>(
      (const __nv_bfloat16*)ddst, ld_dst, B, Nd, D, (__nv_bfloat16*)dsrc, ld_src, Ns, (float)Ns / (float)Nd);
  SFC_LAUNCH_OK();
  return 0;
}
