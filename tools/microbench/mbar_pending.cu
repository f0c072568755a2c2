// What does mbarrier.pending_count report for the state mbarrier.arrive returns?  (PTX: the state BEFORE the arrival.)
// One CTA, an mbarrier with 4 expected arrivals; thread 0 arrives 8 times (two phases) and prints the pending count each
// arrival saw.  Expected "4 3 2 1 4 3 2 1": the last arrival of a phase sees 1 — what the self-fed ring keys on
// (svbrdf_kernels.cu, mbar_arrive_is_last).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/mbar_pending tools/microbench/mbar_pending.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned* out) {
  __shared__ unsigned long long bar;
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(&bar));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(4u) : "memory");
    for (int i = 0; i < 8; ++i) {
      unsigned long long st;
      unsigned c;
      asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(st) : "r"(a) : "memory");
      asm volatile("mbarrier.pending_count.b64 %0, %1;" : "=r"(c) : "l"(st));
      out[i] = c;
    }
  }
}
int main() {
  unsigned* d;
  unsigned h[8];
  cudaMalloc(&d, sizeof(h));
  k<<<1, 32>>>(d);
  if (cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) { printf("error\n"); return 1; }
  printf("pending:");
  for (int i = 0; i < 8; ++i) printf(" %u", h[i]);
  printf("\n");
  return 0;
}
