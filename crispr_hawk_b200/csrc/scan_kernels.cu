// scan_kernels.cu -- K1 (pack) for sm_100a: ASCII haplotype texts -> bit-sliced IUPAC planes,
// case plane and its nz summary. Integer / bitwise work on HBM-resident bit planes: no tensor
// cores. K2 (the PAM scan) lives in scan2_kernels.cu. See DESIGN.md section 3.
#include <cuda_runtime.h>
#include <stdint.h>

#include "hawk_core.h"
#include "hawk_kernels.h"

namespace hawk {

// ------------------------------------------------------------------ K1: pack
// 32 ASCII bytes of one chunk: one 256-bit load (sm_100 has LDG.256; a warp instruction then
// covers 1 KB of consecutive text) when the text is 32-byte aligned, else two 128-bit loads.
template <bool WIDE>
__device__ __forceinline__ void load_chunk_text(const uint4* __restrict__ ascii, int64_t c, uint32_t* w) {
  if (WIDE) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(ascii + 2 * c));
  } else {
    const uint4 a = __ldg(&ascii[2 * c]), b = __ldg(&ascii[2 * c + 1]);
    w[0] = a.x, w[1] = a.y, w[2] = a.z, w[3] = a.w;
    w[4] = b.x, w[5] = b.y, w[6] = b.z, w[7] = b.w;
  }
}

#ifndef HAWK_PACK_RUN
#define HAWK_PACK_RUN 32
#endif
constexpr int64_t PACK_RUN = HAWK_PACK_RUN;  // tiles of 256 chunks per CTA

// One thread per chunk: 32 ASCII bytes -> {A,C,G,T} plane words + case word.
template <bool WIDE>
__global__ void __launch_bounds__(256) pack_kernel(const uint4* __restrict__ ascii,
                                                   int64_t n_chunks, uint4* __restrict__ q,
                                                   uint32_t* __restrict__ v, uint32_t* __restrict__ nz,
                                                   unsigned long long* __restrict__ bad) {
  // a warp packs 32 consecutive chunks per pass (chunk index = multiple of 32 + lane), so one
  // ballot gives the 32 "chunk holds a variant base" bits of a whole nz word
  const int lane = threadIdx.x & 31;
  // every CTA packs one contiguous run of PACK_RUN tiles of 256 chunks (256 KB of text), and
  // consecutive CTAs take consecutive runs: at any moment the CTAs in flight work inside one
  // moving window of ~150 MB instead of striding over the whole slot space (a grid-stride loop
  // over 5 GB ran 10 % slower: 1.42 -> 1.27 ms)
  const int64_t c_lo = (int64_t)blockIdx.x * (PACK_RUN * 256);
  const int64_t c_end = c_lo + PACK_RUN * 256 < n_chunks ? c_lo + PACK_RUN * 256 : n_chunks;
  for (int64_t c0 = c_lo + (threadIdx.x & ~31); c0 < c_end; c0 += 256) {
    const int64_t c = c0 + lane;
    uint32_t vw = 0;
    if (c < n_chunks) {
      uint32_t w[8];
      load_chunk_text<WIDE>(ascii, c, w);
      const PackedChunk o = pack_chunk_v3(w);
      q[c] = make_uint4(o.a, o.c, o.g, o.t);
      v[c] = o.v;
      vw = o.v;
      if (o.invalid) atomicMin(bad, (unsigned long long)(c * 32 + (__ffs(o.invalid) - 1)));
    }
    const uint32_t word = __ballot_sync(0xFFFFFFFFu, vw != 0);
    if (lane == 0) nz[c0 >> 5] = word;
  }
}

}  // namespace hawk

// ------------------------------------------------------------------ launcher
using namespace hawk;

extern "C" int hawk_pack_dev(void* stream, const uint8_t* d_ascii, int64_t total_slots, void* d_q,
                             uint32_t* d_v, uint32_t* d_nz, int64_t* d_bad) {
  if (total_slots < 0 || (total_slots % HAWK_CHUNK) != 0)
    return hawk_fail(HAWK_EINVAL, "hawk_pack_dev: total_slots must be a multiple of 32");
  if (((uintptr_t)d_ascii & 15) || ((uintptr_t)d_q & 15))
    return hawk_fail(HAWK_EINVAL, "hawk_pack_dev: buffers must be 16-byte aligned");
  int64_t n_chunks = total_slots / HAWK_CHUNK;
  if (n_chunks == 0) return HAWK_OK;
  const int64_t blocks = (n_chunks + PACK_RUN * 256 - 1) / (PACK_RUN * 256);
  if (((uintptr_t)d_ascii & 31) == 0)
    pack_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)d_ascii, n_chunks, (uint4*)d_q, d_v, d_nz, (unsigned long long*)d_bad);
  else
    pack_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)d_ascii, n_chunks, (uint4*)d_q, d_v, d_nz, (unsigned long long*)d_bad);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "pack_kernel launch");
}

