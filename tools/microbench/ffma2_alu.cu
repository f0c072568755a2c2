// Does FFMA2 free issue slots for other pipes?  16 FMAs (scalar) or 8 FFMA2 (packed) + A ALU-pipe ops + M MUFU per round.
#include <cstdio>
#include <cuda_runtime.h>

template <int A, int M, bool PACKED>
__global__ void mix_kernel(float* out, int iters, float a, float b) {
  float2 x[8];
  float mn[8];
  float mu[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = make_float2(threadIdx.x * 1e-3f + 2 * i, threadIdx.x * 1e-3f + 2 * i + 1); mn[i] = threadIdx.x * 0.5f + i; }
#pragma unroll
  for (int i = 0; i < 4; ++i) mu[i] = threadIdx.x * 1e-2f + i;
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (PACKED) x[i] = __ffma2_rn(x[i], a2, b2);
      else { x[i].x = __fmaf_rn(x[i].x, a, b); x[i].y = __fmaf_rn(x[i].y, a, b); }
    }
#pragma unroll
    for (int i = 0; i < A; ++i) mn[i % 8] = fminf(fmaxf(mn[i % 8], b), a + float(i));   // 2 FMNMX (ALU pipe) each
#pragma unroll
    for (int i = 0; i < M; ++i) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(mu[i % 4])); mu[i % 4] = y * a; }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y + mn[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) s += mu[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_it(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) launch();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

template <int A, int M>
void run(int warps) {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  const int iters = 20000, threads = warps * 32;
  float ts = time_it([&] { mix_kernel<A, M, false><<<148, threads>>>(out, iters, 0.999f, 1e-3f); });
  float tp = time_it([&] { mix_kernel<A, M, true><<<148, threads>>>(out, iters, 0.999f, 1e-3f); });
  // cycles per round per SMSP-warp: time * clock / (iters * warps_per_smsp)
  const double clk = 1.92e9, wps = warps / 4.0;
  printf("16 FMA + %2d FMNMX + %d MUFU(+FMUL) per round, %2d warps/SM: scalar %.3f ms (%.1f cyc/round/warp) packed %.3f ms (%.1f cyc/round/warp) speedup %.2fx\n",
         2 * A, M, warps, ts, ts * 1e-3 * clk / iters / wps, tp, tp * 1e-3 * clk / iters / wps, ts / tp);
  cudaFree(out);
}

int main() {
  for (int w : {16}) {
    run<0, 0>(w);
    run<4, 0>(w);
    run<8, 0>(w);
    run<4, 1>(w);
    run<4, 2>(w);
    run<2, 2>(w);
    run<0, 2>(w);
  }
  return 0;
}
