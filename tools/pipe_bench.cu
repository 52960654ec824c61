// Dispatch / pipe micro-benchmark for sizing the attention element-wise passes (not part of the product library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_bench tools/pipe_bench.cu && tools/pipe_bench
// Every variant runs `iters` iterations of a fixed instruction mix on 8 independent register chains per thread, one CTA
// per SM, and prints cycles per iteration for 1 / 2 / 4 warps per scheduler: does MUFU.EX2 overlap with FMA- and
// ALU-pipe work issued by the same scheduler, what do packed f32x2 / f16x2 forms cost.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int NFMA, int NINT, bool MUFU, int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, long long* cycles, float* sink) {
  float x[8], y[8], w[8];
  uint32_t z[8];
  unsigned long long q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = -0.001f * (threadIdx.x + i); y[i] = 0.5f + i; w[i] = 1.0f + i; z[i] = threadIdx.x * 7 + i;
    q[i] = ((unsigned long long)__float_as_uint(0.25f + i) << 32) | __float_as_uint(0.5f + i);
  }
  const float a = 0.999f, b = 0.001f;
  const uint32_t thr = 0x19999999u;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MUFU) {
        if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(z[i]));
        if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(z[i]));
      }
#pragma unroll
      for (int j = 0; j < NFMA; ++j) {
        if (MODE == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q[(i + j) & 7]) : "l"(q[(i + j + 1) & 7]), "l"(q[(i + j + 2) & 7]));
        else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y[(i + j) & 7]) : "f"(a), "f"(b));
      }
#pragma unroll
      for (int j = 0; j < NINT; ++j) {
        asm volatile("{ .reg .pred p; mad.lo.u32 %0, %0, 747796405, 12345; setp.ge.u32 p, %0, %2; selp.f32 %1, %1, 0f00000000, p; }"
                     : "+r"(z[(i + j) & 7]), "+f"(w[(i + j) & 7]) : "r"(thr));
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + y[i] + w[i] + (float)z[i] + (float)(q[i] & 0xff);
  if (s == 123.456f) sink[0] = s;
}

template <int NFMA, int NINT, bool MUFU, int MODE>
void run(const char* name) {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4);
  const int iters = 2000;
  printf("%-44s", name);
  for (int threads : {128, 256, 512}) {
    k<NFMA, NINT, MUFU, MODE><<<148, threads>>>(iters, cyc, sink);
    k<NFMA, NINT, MUFU, MODE><<<148, threads>>>(iters, cyc, sink);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double m = 0; for (int i = 0; i < 148; ++i) m += h[i];
    printf("  %d warps/sched: %7.1f clk/iter", threads / 128, m / 148 / iters);
  }
  printf("\n");
  cudaFree(cyc); cudaFree(sink);
}

int main() {
  printf("per iteration and thread: 8 x [MUFU?] + 8 x NFMA fma + 8 x NINT (mad, setp, selp)\n");
  run<0, 0, true, 0>("8 ex2.f32");
  run<1, 0, false, 0>("8 fma");
  run<4, 0, false, 0>("32 fma");
  run<1, 0, true, 0>("8 ex2.f32 + 8 fma");
  run<3, 0, true, 0>("8 ex2.f32 + 24 fma");
  run<7, 0, true, 0>("8 ex2.f32 + 56 fma");
  run<0, 1, false, 0>("8 (mad,setp,selp)");
  run<0, 1, true, 0>("8 ex2.f32 + 8 (mad,setp,selp)");
  run<2, 1, true, 0>("8 ex2.f32 + 16 fma + 8 (mad,setp,selp)");
  run<0, 0, true, 1>("8 ex2.f16x2 (16 exps)");
  run<0, 0, true, 2>("8 ex2.bf16x2 (16 exps)");
  run<2, 0, true, 1>("8 ex2.f16x2 + 16 fma");
  run<4, 0, false, 3>("32 fma.f32x2 (64 fmas)");
  run<1, 0, false, 3>("8 fma.f32x2 (16 fmas)");
  return 0;
}
