// synth_kernels.cu -- haplotype materialisation on the device (first piece of the
// "next" row N1): reference text + per-haplotype sorted, non-overlapping
// edit lists -> haplotype texts in the slot layout, with the reference's conventions
// (haplotype.py:106-121,185-252): copied bases keep the reference's case (upper), every
// ALT allele character is written lower-case, SNV = 1 base, insertion = anchor + inserted
// bases, deletion = the 1-base anchor.
//
// Used by bench.py / tests to build the BASELINE.json workloads (5,009 x 1 Mb, 5,000 x
// 50 Mb shards) without ever holding them on the host.
#include <cuda_runtime.h>
#include <stdint.h>

#include "hawk_kernels.h"

namespace hawk {

struct MaterializeArgs {
  const uint8_t* ref;        // reference text (ASCII), ref_len bytes
  int64_t ref_len;
  const int64_t* edit_off;   // n_hap + 1
  const int32_t* edit_pos;   // reference index of the anchor base
  const int32_t* edit_reflen;  // reference bases consumed (1, or 1 + k for a k-base deletion)
  const int32_t* edit_altlen;  // bases written (1, or 1 + k for a k-base insertion)
  const int64_t* edit_altoff;  // offset of the ALT allele text in alt_pool
  const int32_t* edit_outpos;  // haplotype index where the edit's ALT text starts
  const uint8_t* alt_pool;   // upper- or lower-case ASCII; written lower-case
  const int64_t* slot_off;   // n_hap + 1
  const int32_t* len;        // haplotype lengths
  int32_t n_hap;
  uint8_t* out;              // slot space
};

__global__ void __launch_bounds__(256) materialize_kernel(const __grid_constant__ MaterializeArgs A,
                                                          int64_t n_chunks) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chunks) return;
  // haplotype of chunk c (slot_off is chunk aligned, so a chunk never spans two)
  const int64_t slot = c * 32;
  int32_t lo = 0, hi = A.n_hap;
  while (hi - lo > 1) {
    int32_t m = (lo + hi) >> 1;
    if (A.slot_off[m] <= slot) lo = m; else hi = m;
  }
  const int32_t h = lo;
  const int64_t j0 = slot - A.slot_off[h];  // haplotype index of the chunk's first slot
  const int32_t L = A.len[h];
  const int64_t e0 = A.edit_off[h], e1 = A.edit_off[h + 1];
  // last edit with outpos <= j0
  int64_t el = e0 - 1;
  {
    int64_t l = e0, r = e1;
    while (l < r) {
      int64_t m = (l + r) >> 1;
      if (A.edit_outpos[m] <= j0) l = m + 1; else r = m;
    }
    el = l - 1;
  }
  alignas(16) uint8_t buf[32];
#pragma unroll 4
  for (int i = 0; i < 32; ++i) {
    int64_t j = j0 + i;
    uint8_t ch = 0;
    if (j >= 0 && j < L) {
      while (el + 1 < e1 && A.edit_outpos[el + 1] <= j) ++el;
      if (el < e0) {
        ch = A.ref[j];
      } else {
        int64_t d = j - A.edit_outpos[el];
        if (d < A.edit_altlen[el]) {
          ch = A.alt_pool[A.edit_altoff[el] + d] | 0x20;  // ALT allele text is lower-case
        } else {
          ch = A.ref[(int64_t)A.edit_pos[el] + A.edit_reflen[el] + (d - A.edit_altlen[el])];
        }
      }
    }
    buf[i] = ch;
  }
  uint4* dst = reinterpret_cast<uint4*>(A.out + A.slot_off[h] + j0);
  dst[0] = reinterpret_cast<const uint4*>(buf)[0];
  dst[1] = reinterpret_cast<const uint4*>(buf)[1];
}

}  // namespace hawk

using namespace hawk;

extern "C" int hawk_materialize_dev(void* stream, const uint8_t* d_ref, int64_t ref_len,
                                    const int64_t* d_edit_off, const int32_t* d_edit_pos,
                                    const int32_t* d_edit_reflen, const int32_t* d_edit_altlen,
                                    const int64_t* d_edit_altoff, const int32_t* d_edit_outpos,
                                    const uint8_t* d_alt_pool, const int64_t* d_slot_off,
                                    const int32_t* d_len, int32_t n_hap, int64_t total_slots,
                                    uint8_t* d_ascii_out) {
  const int64_t n_chunks = total_slots / HAWK_CHUNK;
  if (n_hap <= 0 || n_chunks <= 0) return HAWK_OK;
  if ((uintptr_t)d_ascii_out & 15) return hawk_fail(HAWK_EINVAL, "hawk_materialize_dev: output must be 16-byte aligned");
  MaterializeArgs A{d_ref, ref_len, d_edit_off, d_edit_pos, d_edit_reflen, d_edit_altlen, d_edit_altoff,
                    d_edit_outpos, d_alt_pool, d_slot_off, d_len, n_hap, d_ascii_out};
  int64_t blocks = (n_chunks + 255) / 256;
  materialize_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(A, n_chunks);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "materialize_kernel launch");
}
