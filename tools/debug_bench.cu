// Micro-benchmarks of on-chip rates used to size the kernels (not part of the product path; exported for tools/).
#include "common.cuh"
#include "sfcvit.h"

namespace {

// every warp streams its 32 TMEM lanes x 512 columns `iters` times with tcgen05.ld.32x32b.x32
__global__ void __launch_bounds__(512, 1) tmem_ld_bench_kernel(int iters, long long* cycles, float* sink) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) ptx::tmem_alloc<512>(&tmem_ptr);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t base = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int c = 0; c < 512; c += 64) {
      uint32_t r0[32], r1[32];
      ptx::tmem_ld_x32(base + c, r0);
      ptx::tmem_ld_x32(base + c + 32, r1);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc += __uint_as_float(r0[i] & 0x3f800000u) + __uint_as_float(r1[i] & 0x3f800000u);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_ptr);
  }
}

// exp2 + fma stream: `iters` x 64 MUFU.EX2 per thread
__global__ void __launch_bounds__(512, 1) ex2_bench_kernel(int iters, long long* cycles, float* sink) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = 0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = exp2f(x[i] * 0.5f - 1.0f);
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 123.456f) sink[0] = s;
}

}  // namespace

// which: 0 = TMEM load (bytes per iteration per CTA = warps * 32 lanes * 512 cols * 4 B), 1 = exp2 (64 per thread per iter)
extern "C" int sfc_debug_bench(int which, int threads, int iters, long long* cycles_dev, float* sink_dev, cudaStream_t stream) {
  SFC_REQUIRE(threads % 128 == 0 && threads >= 128 && threads <= 512, "sfc_debug_bench: threads must be 128..512, multiple of 128");
  if (which == 0) tmem_ld_bench_kernel<<<sfc_num_sms(), threads, 0, stream>>>(iters, cycles_dev, sink_dev);
  else ex2_bench_kernel<<<sfc_num_sms(), threads, 0, stream>>>(iters, cycles_dev, sink_dev);
  SFC_LAUNCH_OK();
  return 0;
}
