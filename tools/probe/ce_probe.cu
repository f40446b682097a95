// build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/probe/ce_probe tools/probe/ce_probe.cu
// ce_probe.cu -- does a small copy on stream B wait behind a large same-direction copy on
// stream A? (decides how the streamed search moves its small per-group metadata)
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__global__ void copy_k(uint4* dst, const uint4* src, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  const size_t big = 1ull << 30, small = 64 << 10;
  void *hb, *hs, *hs2, *db, *ds, *db2;
  CK(cudaHostAlloc(&hb, big, cudaHostAllocDefault));
  CK(cudaHostAlloc(&hs, small, cudaHostAllocDefault));
  CK(cudaHostAlloc(&hs2, big, cudaHostAllocDefault));
  CK(cudaMalloc(&db, big)); CK(cudaMalloc(&ds, small)); CK(cudaMalloc(&db2, big));
  cudaStream_t a, b, c;
  CK(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&c, cudaStreamNonBlocking));
  for (int rep = 0; rep < 3; ++rep) {
    double t0 = now();
    CK(cudaMemcpyAsync(db, hb, big, cudaMemcpyHostToDevice, a));
    CK(cudaStreamSynchronize(a));
    double t1 = now();
    printf("H2D 1 GiB alone: %.2f ms (%.1f GB/s)\n", t1 - t0, big / (t1 - t0) / 1e6);
    t0 = now();
    CK(cudaMemcpyAsync(hs2, db2, big, cudaMemcpyDeviceToHost, c));
    CK(cudaStreamSynchronize(c));
    t1 = now();
    printf("D2H 1 GiB alone: %.2f ms (%.1f GB/s)\n", t1 - t0, big / (t1 - t0) / 1e6);
    t0 = now();
    CK(cudaMemcpyAsync(db, hb, big, cudaMemcpyHostToDevice, a));
    CK(cudaMemcpyAsync(hs2, db2, big, cudaMemcpyDeviceToHost, c));
    CK(cudaStreamSynchronize(a)); CK(cudaStreamSynchronize(c));
    t1 = now();
    printf("H2D + D2H 1 GiB each, concurrent: %.2f ms\n", t1 - t0);
    // small H2D on b while big H2D on a
    t0 = now();
    CK(cudaMemcpyAsync(db, hb, big, cudaMemcpyHostToDevice, a));
    CK(cudaMemcpyAsync(ds, hs, small, cudaMemcpyHostToDevice, b));
    CK(cudaStreamSynchronize(b));
    t1 = now();
    CK(cudaStreamSynchronize(a));
    double t2 = now();
    printf("small H2D (memcpy) behind big H2D: small done after %.3f ms, big after %.2f ms\n", t1 - t0, t2 - t0);
    // small D2H on b while big D2H on c
    t0 = now();
    CK(cudaMemcpyAsync(hs2, db2, big, cudaMemcpyDeviceToHost, c));
    CK(cudaMemcpyAsync(hs, ds, small, cudaMemcpyDeviceToHost, b));
    CK(cudaStreamSynchronize(b));
    t1 = now();
    CK(cudaStreamSynchronize(c));
    t2 = now();
    printf("small D2H (memcpy) behind big D2H: small done after %.3f ms, big after %.2f ms\n", t1 - t0, t2 - t0);
    // small H2D by a kernel reading pinned host memory
    t0 = now();
    CK(cudaMemcpyAsync(db, hb, big, cudaMemcpyHostToDevice, a));
    copy_k<<<8, 256, 0, b>>>((uint4*)ds, (const uint4*)hs, small / 16);
    CK(cudaStreamSynchronize(b));
    t1 = now();
    CK(cudaStreamSynchronize(a));
    t2 = now();
    printf("small H2D (kernel, zero-copy) beside big H2D: small done after %.3f ms, big after %.2f ms\n", t1 - t0, t2 - t0);
    t0 = now();
    CK(cudaMemcpyAsync(hs2, db2, big, cudaMemcpyDeviceToHost, c));
    copy_k<<<8, 256, 0, b>>>((uint4*)hs, (const uint4*)ds, small / 16);
    CK(cudaStreamSynchronize(b));
    t1 = now();
    CK(cudaStreamSynchronize(c));
    t2 = now();
    printf("small D2H (kernel, zero-copy) beside big D2H: small done after %.3f ms, big after %.2f ms\n", t1 - t0, t2 - t0);
    // two big H2D on two streams: serialized or shared?
    t0 = now();
    CK(cudaMemcpyAsync(db, hb, big / 2, cudaMemcpyHostToDevice, a));
    CK(cudaMemcpyAsync((char*)db + big / 2, (char*)hb + big / 2, big / 2, cudaMemcpyHostToDevice, b));
    CK(cudaStreamSynchronize(a));
    t1 = now();
    CK(cudaStreamSynchronize(b));
    t2 = now();
    printf("two 512 MiB H2D on two streams: first done %.2f ms, second %.2f ms\n", t1 - t0, t2 - t0);
    // big zero-copy kernel bandwidth (SM-driven H2D)
    t0 = now();
    copy_k<<<148 * 4, 256, 0, b>>>((uint4*)db, (const uint4*)hb, big / 16);
    CK(cudaStreamSynchronize(b));
    t1 = now();
    printf("1 GiB H2D by kernel (zero-copy): %.2f ms (%.1f GB/s)\n", t1 - t0, big / (t1 - t0) / 1e6);
  }
  return 0;
}
