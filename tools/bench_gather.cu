// bench_gather.cu -- ceiling for the emit kernel's memory pattern: out[i] = in[perm[i]] for 80-byte records
// (16-byte pieces, 5 lanes per record, streaming stores), no checksum, no framing.  Prints ms and GB/s moved.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bench_gather tools/bench_gather.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint4 ldg_stream_v4(const void *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream_v4(void *p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__global__ void k_fill_perm(uint32_t *perm, uint32_t n, uint32_t mul) {
  // a bijection on [0, n) for n = 10^8: i * mul mod n with gcd(mul, n) = 1
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    perm[i] = (uint32_t)((i * (uint64_t)mul) % n);
}

template <int UNROLL>
__global__ void __launch_bounds__(256) k_gather(const uint8_t *__restrict__ in, const uint32_t *__restrict__ perm,
                                                uint8_t *__restrict__ out, uint64_t npieces) {
  // piece q = 5 * record + c
  const uint64_t stride = (uint64_t)gridDim.x * 256 * UNROLL;
  for (uint64_t base = (uint64_t)blockIdx.x * 256 * UNROLL; base < npieces; base += stride) {
    uint4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      uint64_t q = base + threadIdx.x + u * 256;
      if (q < npieces) {
        uint32_t r = (uint32_t)(q / 5), c = (uint32_t)(q - 5ull * r);
        v[u] = ldg_stream_v4(in + (uint64_t)perm[r] * 80 + 16 * c);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      uint64_t q = base + threadIdx.x + u * 256;
      if (q < npieces) stg_stream_v4(out + q * 16, v[u]);
    }
  }
}

int main(int argc, char **argv) {
  const uint32_t n = argc > 1 ? (uint32_t)atoll(argv[1]) : 100000000u;
  uint8_t *in, *out;
  uint32_t *perm;
  cudaMalloc(&in, (size_t)n * 80);
  cudaMalloc(&out, (size_t)n * 80);
  cudaMalloc(&perm, (size_t)n * 4);
  cudaMemset(in, 1, (size_t)n * 80);
  k_fill_perm<<<148 * 8, 256>>>(perm, n, 48271u * 7919u + 2u * 3u * 0u + 0u | 1u);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const uint64_t npieces = (uint64_t)n * 5;
  for (int ctas = 2; ctas <= 8; ctas += 2) {
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      k_gather<5><<<148 * ctas, 256>>>(in, perm, out, npieces);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep) printf("gather-copy 80 B records, unroll 5, %d CTAs/SM: %.3f ms, %.0f GB/s (84 B/rec read + 80 B/rec write)\n", ctas, ms,
                      (double)n * 164 / (ms * 1e-3) / 1e9);
    }
  }
  // sequential copy of the same volume for reference
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(e0);
    cudaMemcpyAsync(out, in, (size_t)n * 80, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep) printf("cudaMemcpy D2D %.3f ms, %.0f GB/s\n", ms, (double)n * 160 / (ms * 1e-3) / 1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
