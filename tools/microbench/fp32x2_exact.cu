// Are the packed FP32x2 instructions (FFMA2 / FMUL2 / FADD2) bit-identical to their scalar counterparts?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp32x2_exact fp32x2_exact.cu && ./fp32x2_exact
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ unsigned rng(unsigned& s) { s = s * 1664525u + 1013904223u; return s; }
__device__ float rnd_float(unsigned& s, int mode) {
  unsigned r = rng(s);
  if (mode == 0) return __uint_as_float((r & 0x807fffffu) | ((100u + (rng(s) % 56u)) << 23));   // wide exponent range, normal
  if (mode == 1) return (float(r >> 8) * (1.0f / 16777216.0f));                                   // [0,1)
  return __uint_as_float(r & 0x80ffffffu);                                                        // subnormal / tiny
}

__global__ void k(unsigned long long* bad, int mode) {
  unsigned s = blockIdx.x * 9781u + threadIdx.x * 6271u + 12345u + mode;
  unsigned long long nf = 0, nm = 0, na = 0;
  for (int i = 0; i < 4096; ++i) {
    float a0 = rnd_float(s, mode), a1 = rnd_float(s, mode), b0 = rnd_float(s, mode), b1 = rnd_float(s, mode);
    float c0 = rnd_float(s, mode), c1 = rnd_float(s, mode);
    if (i & 1) { c0 = -a0 * b0 * (1.0f + 1e-6f); }     // cancellation
    float2 f = __ffma2_rn(make_float2(a0, a1), make_float2(b0, b1), make_float2(c0, c1));
    float2 m = __fmul2_rn(make_float2(a0, a1), make_float2(b0, b1));
    float2 d = __fadd2_rn(make_float2(a0, a1), make_float2(c0, c1));
    nf += (__float_as_uint(f.x) != __float_as_uint(__fmaf_rn(a0, b0, c0))) + (__float_as_uint(f.y) != __float_as_uint(__fmaf_rn(a1, b1, c1)));
    nm += (__float_as_uint(m.x) != __float_as_uint(__fmul_rn(a0, b0))) + (__float_as_uint(m.y) != __float_as_uint(__fmul_rn(a1, b1)));
    na += (__float_as_uint(d.x) != __float_as_uint(__fadd_rn(a0, c0))) + (__float_as_uint(d.y) != __float_as_uint(__fadd_rn(a1, c1)));
  }
  atomicAdd(bad + 0, nf); atomicAdd(bad + 1, nm); atomicAdd(bad + 2, na);
}

int main() {
  unsigned long long* bad;
  cudaMallocManaged(&bad, 24);
  for (int mode = 0; mode < 3; ++mode) {
    bad[0] = bad[1] = bad[2] = 0;
    k<<<148, 256>>>(bad, mode);
    cudaDeviceSynchronize();
    printf("mode %d: of %llu pairs: ffma2 mismatches %llu, fmul2 %llu, fadd2 %llu\n", mode, 148ull * 256 * 4096 * 2, bad[0], bad[1], bad[2]);
  }
  return 0;
}
