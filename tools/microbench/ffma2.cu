// Microbenchmark: does packed fp32x2 (FFMA2 / FMUL2 / FADD2, sm_100) relieve an issue-bound FP32 + MUFU mix?
// Each thread runs ITER rounds of K independent FMA chains (+ optional MUFU per round), scalar vs packed.
#include <cstdio>
#include <cuda_runtime.h>

template <int K, int MUFU>
__global__ void scalar_kernel(float* out, int iters, float a, float b) {
  float x[2 * K];
#pragma unroll
  for (int i = 0; i < 2 * K; ++i) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 2 * K; ++i) x[i] = __fmaf_rn(x[i], a, b);
#pragma unroll
    for (int i = 0; i < MUFU; ++i) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[i])); x[i] = y; }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 2 * K; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int K, int MUFU>
__global__ void packed_kernel(float* out, int iters, float a, float b) {
  float2 x[K];
#pragma unroll
  for (int i = 0; i < K; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + 2 * i, threadIdx.x * 1e-3f + 2 * i + 1);
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < K; ++i) x[i] = __ffma2_rn(x[i], a2, b2);
#pragma unroll
    for (int i = 0; i < MUFU; ++i) {
      float y;
      float& r = (i & 1) ? x[i / 2].y : x[i / 2].x;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(r));
      r = y;
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < K; ++i) s += x[i].x + x[i].y;
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

template <int K, int MUFU>
void run(int warps_per_sm) {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  const int iters = 20000, threads = warps_per_sm * 32;
  float ts = time_it([&] { scalar_kernel<K, MUFU><<<148, threads>>>(out, iters, 0.999f, 1e-3f); });
  float tp = time_it([&] { packed_kernel<K, MUFU><<<148, threads>>>(out, iters, 0.999f, 1e-3f); });
  const double fma = 2.0 * K * iters * 148.0 * threads;
  printf("K=%d (FMA chains %d) MUFU/round=%d warps/SM=%2d : scalar %.3f ms (%.1f TFMA/s) packed %.3f ms (%.1f TFMA/s) speedup %.2fx\n", K, 2 * K, MUFU,
         warps_per_sm, ts, fma / ts * 1e-9, tp, fma / tp * 1e-9, ts / tp);
  cudaFree(out);
}

int main() {
  for (int w : {8, 16, 32}) {
    run<8, 0>(w);
    run<8, 2>(w);
    run<8, 4>(w);
  }
  return 0;
}
