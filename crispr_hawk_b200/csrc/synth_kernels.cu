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

// 32 bytes of `src` starting at an arbitrary byte address (readable up to the next 4-byte
// boundary past src + 32): nine aligned words, realigned with byte permutes
__device__ __forceinline__ void load32_unaligned(const uint8_t* src, uint32_t out[8]) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)3);
  const uint32_t sel = 0x3210u + 0x1111u * (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3);
  uint32_t prev = __ldg(w);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint32_t next = __ldg(w + k + 1);
    out[k] = __byte_perm(prev, next, sel);
    prev = next;
  }
}

// One 32-byte output chunk of haplotype h starting at haplotype index j0, byte by byte
// (chunks that hold ALT text). `el` = last edit with outpos <= j0, or e0 - 1.
__device__ __forceinline__ void chunk_bytewise(const MaterializeArgs& A, int32_t h, int64_t j0, int64_t el, uint4* dst) {
  const int32_t L = A.len[h];
  const int64_t e0 = A.edit_off[h], e1 = A.edit_off[h + 1];
  uint32_t w[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    uint32_t word = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t j = j0 + 4 * k + b;
      uint32_t ch = 0;
      if (j < L) {
        while (el + 1 < e1 && A.edit_outpos[el + 1] <= j) ++el;
        if (el < e0) {
          ch = A.ref[j];
        } else {
          const int64_t d = j - A.edit_outpos[el];
          if (d < A.edit_altlen[el]) ch = A.alt_pool[A.edit_altoff[el] + d] | 0x20u;  // ALT text is lower-case
          else ch = A.ref[(int64_t)A.edit_pos[el] + A.edit_reflen[el] + (d - A.edit_altlen[el])];
        }
      }
      word |= ch << (8 * b);
    }
    w[k] = word;
  }
  dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
  dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// Pass 1, grid (chunks of the longest haplotype / 256, haplotypes): every chunk that holds no
// ALT text is one realigned 32-byte block copy of the reference; gaps are zeroed; chunks with
// ALT text are left to pass 2.
__global__ void __launch_bounds__(256) materialize_kernel(const __grid_constant__ MaterializeArgs A) {
  for (int32_t h = blockIdx.y; h < A.n_hap; h += gridDim.y) {
    const int64_t s0 = A.slot_off[h] - HAWK_SLOT_GAP;  // the gap in front belongs to this haplotype's grid row
    const int64_t s1 = A.slot_off[h + 1] - HAWK_SLOT_GAP + (h + 1 == A.n_hap ? HAWK_SLOT_GAP : 0);
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t slot = s0 + c * 32;
    if (slot >= s1) continue;
    const int64_t j0 = slot - A.slot_off[h];
    const int32_t L = A.len[h];
    uint4* dst = reinterpret_cast<uint4*>(A.out + slot);
    if (j0 < 0 || j0 >= L) {
      dst[0] = dst[1] = make_uint4(0, 0, 0, 0);
      continue;
    }
    const int64_t e0 = A.edit_off[h], e1 = A.edit_off[h + 1];
    int64_t l = e0, r = e1;  // last edit with outpos <= j0
    while (l < r) {
      const int64_t m = (l + r) >> 1;
      if (A.edit_outpos[m] <= j0) l = m + 1; else r = m;
    }
    const int64_t el = l - 1;
    int64_t shift = 0;  // reference index = haplotype index + shift behind the ALT text of edit el
    bool plain = true;
    if (el >= e0) {
      const int64_t after = (int64_t)A.edit_outpos[el] + A.edit_altlen[el];
      shift = (int64_t)A.edit_pos[el] + A.edit_reflen[el] - after;
      plain = j0 >= after;
    }
    plain = plain && (el + 1 >= e1 || A.edit_outpos[el + 1] >= j0 + 32);
    if (!plain) continue;  // pass 2
    if (j0 + 32 <= L) {
      uint32_t w[8];
      load32_unaligned(A.ref + j0 + shift, w);
      dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
      dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    } else {
      chunk_bytewise(A, h, j0, el, dst);  // last, partial chunk of the haplotype
    }
  }
}

// Pass 2, one thread per edit: the chunks its ALT text touches, byte by byte. Two edits in
// one chunk both write the same bytes.
__global__ void __launch_bounds__(128) materialize_edits_kernel(const __grid_constant__ MaterializeArgs A,
                                                                int64_t e_lo, int64_t e_hi) {
  const int64_t e = e_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= e_hi) return;
  int32_t lo = 0, hi = A.n_hap;  // haplotype of edit e
  while (hi - lo > 1) {
    const int32_t m = (lo + hi) >> 1;
    if (A.edit_off[m] <= e) lo = m; else hi = m;
  }
  const int32_t h = lo;
  const int64_t e0 = A.edit_off[h];
  const int64_t first = A.edit_outpos[e] >> 5, last = ((int64_t)A.edit_outpos[e] + A.edit_altlen[e] - 1) >> 5;
  for (int64_t c = first; c <= last; ++c) {
    const int64_t j0 = c * 32;
    int64_t el = e;  // last edit with outpos <= j0
    while (el >= e0 && A.edit_outpos[el] > j0) --el;
    chunk_bytewise(A, h, j0, el, reinterpret_cast<uint4*>(A.out + A.slot_off[h] + j0));
  }
}

// ------------------------------------------------------------------ edit lists -> geometry
// One CTA per haplotype walks its edits in tiles of 256 with a running carry:
//   pass 0 (segs == null): validates the edits, writes edit_outpos, the haplotype length and
//          its number of posmap segments;
//   pass 1: writes the run-length posmap segments (haplotype.py:138-159 conventions) at
//          seg_off[h]: (0, region_start, step 1), then per insertion (outpos + 1, anchor, step 0)
//          and (outpos + altlen, anchor + 1, step 1), per deletion (outpos + 1, anchor + reflen, 1).
struct DeriveArgs {
  const int64_t* edit_off;
  const int32_t* pos;
  const int32_t* reflen;
  const int32_t* altlen;
  const int64_t* altoff;
  int64_t ref_len, alt_pool_len;
  int32_t region_start;
  int32_t* outpos;      // pass 0
  int32_t* edit_hap;    // pass 0, optional: haplotype of every edit
  int32_t* len;         // pass 0
  int32_t* seg_count;   // pass 0
  int32_t* bad;         // pass 0: smallest haplotype index with an invalid edit list, else INT_MAX
  const int64_t* seg_off;  // pass 1
  int32_t* seg_rel;
  int32_t* seg_gen;
  uint8_t* seg_step;
};

__global__ void __launch_bounds__(256) derive_kernel(const __grid_constant__ DeriveArgs A, int pass) {
  __shared__ int64_t s_shift[256];
  __shared__ int32_t s_segs[256];
  __shared__ int64_t carry_shift;
  __shared__ int32_t carry_segs;
  const int h = blockIdx.x, tid = threadIdx.x;
  const int64_t e0 = A.edit_off[h], e1 = A.edit_off[h + 1];
  if (tid == 0) {
    carry_shift = 0;
    carry_segs = 1;  // segment 0 = (0, region_start, step 1)
    if (pass == 1) {
      const int64_t s0 = A.seg_off[h];
      A.seg_rel[s0] = 0;
      A.seg_gen[s0] = A.region_start;
      A.seg_step[s0] = 1;
    }
  }
  __syncthreads();
  for (int64_t base = e0; base < e1; base += 256) {
    const int64_t e = base + tid;
    int64_t d = 0;
    int32_t ns = 0, p = 0, rl = 1, al = 1;
    if (e < e1) {
      p = A.pos[e];
      rl = A.reflen[e];
      al = A.altlen[e];
      d = (int64_t)al - rl;
      ns = al > 1 ? 2 : (rl > 1 ? 1 : 0);
      if (pass == 0) {
        const int64_t prev_end = e > e0 ? (int64_t)A.pos[e - 1] + A.reflen[e - 1] : 0;
        const bool ok = rl >= 1 && al >= 1 && !(rl > 1 && al > 1) && p >= prev_end && (int64_t)p + rl <= A.ref_len &&
                        A.altoff[e] >= 0 && A.altoff[e] + al <= A.alt_pool_len;
        if (!ok) atomicMin(A.bad, h);
      }
    }
    s_shift[tid] = d;
    s_segs[tid] = ns;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {  // inclusive scans over the tile
      const int64_t a = tid >= o ? s_shift[tid - o] : 0;
      const int32_t b = tid >= o ? s_segs[tid - o] : 0;
      __syncthreads();
      s_shift[tid] += a;
      s_segs[tid] += b;
      __syncthreads();
    }
    if (e < e1) {
      const int64_t op = (int64_t)p + carry_shift + (s_shift[tid] - d);
      if (pass == 0) {
        A.outpos[e] = (int32_t)op;
        if (A.edit_hap) A.edit_hap[e] = h;
      } else if (ns) {
        const int64_t s = A.seg_off[h] + carry_segs + (s_segs[tid] - ns);
        const int32_t anchor = A.region_start + p;
        if (al > 1) {
          A.seg_rel[s] = (int32_t)(op + 1);
          A.seg_gen[s] = anchor;
          A.seg_step[s] = 0;
          A.seg_rel[s + 1] = (int32_t)(op + al);
          A.seg_gen[s + 1] = anchor + 1;
          A.seg_step[s + 1] = 1;
        } else {
          A.seg_rel[s] = (int32_t)(op + 1);
          A.seg_gen[s] = anchor + rl;
          A.seg_step[s] = 1;
        }
      }
    }
    __syncthreads();
    if (tid == 255) {
      carry_shift += s_shift[255];
      carry_segs += s_segs[255];
    }
    __syncthreads();
  }
  if (pass == 0 && tid == 0) {
    A.len[h] = (int32_t)(A.ref_len + carry_shift);
    A.seg_count[h] = carry_segs;
  }
}

int launch_derive(cudaStream_t st, int32_t n_hap, const int64_t* edit_off, const int32_t* pos, const int32_t* reflen,
                  const int32_t* altlen, const int64_t* altoff, int64_t ref_len, int64_t alt_pool_len,
                  int32_t region_start, int32_t* outpos, int32_t* len, int32_t* seg_count, int32_t* bad,
                  const int64_t* seg_off, int32_t* seg_rel, int32_t* seg_gen, uint8_t* seg_step, int pass,
                  int32_t* edit_hap) {
  if (n_hap <= 0) return HAWK_OK;
  DeriveArgs A{edit_off, pos, reflen, altlen, altoff, ref_len, alt_pool_len, region_start, outpos, edit_hap, len, seg_count,
               bad, seg_off, seg_rel, seg_gen, seg_step};
  derive_kernel<<<(unsigned)n_hap, 256, 0, st>>>(A, pass);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "derive_kernel launch");
}

}  // namespace hawk

using namespace hawk;

// edits [e_lo, e_hi) must be those of the n_hap haplotypes d_edit_off describes (the whole table:
// 0 .. n_edits; one haplotype of a larger table: its own range, with d_edit_off / d_len pointing at it)
int hawk_materialize_range(cudaStream_t stream, const uint8_t* d_ref, int64_t ref_len, const int64_t* d_edit_off,
                           const int32_t* d_edit_pos, const int32_t* d_edit_reflen, const int32_t* d_edit_altlen,
                           const int64_t* d_edit_altoff, const int32_t* d_edit_outpos, const uint8_t* d_alt_pool,
                           const int64_t* d_slot_off, const int32_t* d_len, int32_t n_hap, int64_t total_slots,
                           int64_t e_lo, int64_t e_hi, int32_t max_len, uint8_t* d_ascii_out) {
  if (n_hap <= 0 || total_slots <= 0) return HAWK_OK;
  if ((uintptr_t)d_ascii_out & 15) return hawk_fail(HAWK_EINVAL, "hawk_materialize_dev: output must be 16-byte aligned");
  MaterializeArgs A{d_ref, ref_len, d_edit_off, d_edit_pos, d_edit_reflen, d_edit_altlen, d_edit_altoff,
                    d_edit_outpos, d_alt_pool, d_slot_off, d_len, n_hap, d_ascii_out};
  // grid row = haplotype: its leading gap + its padded slots (+ the trailing gap of the last one)
  const int64_t row_chunks = (((int64_t)max_len + HAWK_SLOT_ALIGN - 1) / HAWK_SLOT_ALIGN * HAWK_SLOT_ALIGN + 2 * HAWK_SLOT_GAP) / 32;
  dim3 grid((unsigned)((row_chunks + 255) / 256), (unsigned)(n_hap < 65535 ? n_hap : 65535));
  materialize_kernel<<<grid, 256, 0, stream>>>(A);
  hawk_note_launch(1);
  if (e_hi > e_lo) {
    materialize_edits_kernel<<<(unsigned)((e_hi - e_lo + 127) / 128), 128, 0, stream>>>(A, e_lo, e_hi);
    hawk_note_launch(1);
  }
  return hawk_check_cuda(cudaGetLastError(), "materialize kernels launch");
}

extern "C" int hawk_materialize_dev(void* stream, const uint8_t* d_ref, int64_t ref_len,
                                    const int64_t* d_edit_off, const int32_t* d_edit_pos,
                                    const int32_t* d_edit_reflen, const int32_t* d_edit_altlen,
                                    const int64_t* d_edit_altoff, const int32_t* d_edit_outpos,
                                    const uint8_t* d_alt_pool, const int64_t* d_slot_off,
                                    const int32_t* d_len, int32_t n_hap, int64_t total_slots,
                                    int64_t n_edits, int32_t max_len, uint8_t* d_ascii_out) {
  return hawk_materialize_range((cudaStream_t)stream, d_ref, ref_len, d_edit_off, d_edit_pos, d_edit_reflen, d_edit_altlen,
                                d_edit_altoff, d_edit_outpos, d_alt_pool, d_slot_off, d_len, n_hap, total_slots, 0, n_edits,
                                max_len, d_ascii_out);
}
