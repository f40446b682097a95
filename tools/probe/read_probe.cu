// read_probe.cu -- how fast can the K1 access pattern READ 5 GB (no stores, trivial arithmetic)?
// Same work assignment as pack_kernel: CTA = one contiguous run of 32 tiles of 256 chunks, one
// 256-bit load per thread and pass. Variants: plain loop, register prefetch, 2 chunks per thread.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void ld256(const uint4* p, uint32_t* w) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
               : "l"(p));
}

template <int MODE, int WORK>
__global__ void __launch_bounds__(256) read_kernel(const uint4* __restrict__ a, int64_t n_chunks, uint32_t* out) {
  const int lane = threadIdx.x & 31;
  const int64_t c_lo = (int64_t)blockIdx.x * (32 * 256);
  const int64_t c_end = c_lo + 32 * 256 < n_chunks ? c_lo + 32 * 256 : n_chunks;
  uint32_t acc = 0;
  if (MODE == 0) {
    for (int64_t c0 = c_lo + (threadIdx.x & ~31); c0 < c_end; c0 += 256) {
      uint32_t w[8];
      ld256(a + 2 * (c0 + lane), w);
      uint32_t x = w[0] ^ w[1] ^ w[2] ^ w[3] ^ w[4] ^ w[5] ^ w[6] ^ w[7];
#pragma unroll 1
      for (int k = 0; k < WORK; ++k) x = x * 0x9E3779B1u + (x >> 7);  // dependent filler (FMA pipe)
      acc ^= x;
    }
  } else {
    uint32_t w[8], nw[8];
    int64_t c0 = c_lo + (threadIdx.x & ~31);
    ld256(a + 2 * (c0 + lane), w);
    for (; c0 < c_end; c0 += 256) {
      if (c0 + 256 < c_end) ld256(a + 2 * (c0 + 256 + lane), nw);
      uint32_t x = w[0] ^ w[1] ^ w[2] ^ w[3] ^ w[4] ^ w[5] ^ w[6] ^ w[7];
#pragma unroll 1
      for (int k = 0; k < WORK; ++k) x = x * 0x9E3779B1u + (x >> 7);
      acc ^= x;
#pragma unroll
      for (int k = 0; k < 8; ++k) w[k] = nw[k];
    }
  }
  if (acc == 0x12345678u) out[0] = acc;
}

int main() {
  const int64_t bytes = 5009ll * 1000448;  // config 2's slot space
  const int64_t n_chunks = bytes / 32 / 256 * 256;
  uint4* a;
  uint32_t* out;
  cudaMalloc(&a, n_chunks * 32);
  cudaMalloc(&out, 4);
  cudaMemset(a, 0x41, n_chunks * 32);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const unsigned blocks = (unsigned)((n_chunks + 32 * 256 - 1) / (32 * 256));
  auto run = [&](const char* name, void (*k)(const uint4*, int64_t, uint32_t*)) {
    float best = 1e9f;
    for (int r = 0; r < 6; ++r) {
      cudaEventRecord(e0);
      k<<<blocks, 256>>>(a, n_chunks, out);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (r && ms < best) best = ms;
    }
    printf("%-44s %.3f ms  %.2f TB/s\n", name, best, n_chunks * 32.0 / best / 1e9);
  };
  run("plain loop, no work", read_kernel<0, 0>);
  run("prefetch, no work", read_kernel<1, 0>);
  run("plain loop, 40 dependent IMADs per chunk", read_kernel<0, 40>);
  run("prefetch, 40 dependent IMADs per chunk", read_kernel<1, 40>);
  run("plain loop, 100 dependent IMADs per chunk", read_kernel<0, 100>);
  run("prefetch, 100 dependent IMADs per chunk", read_kernel<1, 100>);
  return 0;
}
