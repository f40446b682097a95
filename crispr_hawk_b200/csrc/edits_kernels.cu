// edits_kernels.cu -- N1 (haplotypes given as reference + edit lists, haplotypes.py:716-751):
// the bit planes of a batch built WHERE THEY ARE READ, without ever writing the haplotype texts.
//
// A phased / variant-free search reads a haplotype's planes in two places only: every chunk of
// a REF haplotype, and -- for the others -- the chunks within a few chunks of a variant (lower-
// case) base: candidates are chunks next to a variant chunk, their windows and the rows' texts
// reach `reach` chunks further (hawk_core.h, fused_kernels.cu). A haplotype that is the
// reference plus ~10^3 edits therefore needs ~10^3 x (2 reach + 1) chunks of planes, not its
// 31,000: per edit a handful of threads assemble those chunks from the REFERENCE's planes
// (packed once, 0.6 MB, L2-resident) by funnel shifts -- between two edits a haplotype is the
// reference at a constant offset -- and from the ALT texts. 5,009 haplotypes x 1 Mb: 4.2 M edits
// x 7 chunks x 20 B = 0.6 GB written instead of 5 GB of text written, read again and packed
// into 3.1 GB of planes.
//
// Conventions of the texts these planes stand for (haplotype.py:106-121, 185-252, as in
// synth_kernels.cu): copied bases keep the reference's case, every ALT character is lower-case
// (SNV: 1 base; insertion: anchor + inserted bases; deletion: the 1-base anchor).
#include <cuda_runtime.h>
#include <stdint.h>

#include "hawk_core.h"
#include "hawk_kernels.h"
#include "hawk_post.h"

namespace hawk {

constexpr int GAP_CHUNKS = HAWK_SLOT_GAP / HAWK_CHUNK;

// ALT pool characters must be IUPAC letters (either case): smallest offending offset, else untouched
__global__ void pool_check_kernel(const uint8_t* __restrict__ pool, int64_t n, unsigned long long* __restrict__ bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t e = iupac_entry(pool[i]);
  if (pool[i] == 0 || (e & 0x80)) atomicMin(bad, (unsigned long long)i);
}

// haplotypes without edits are the reference itself: their planes are a copy of the reference's
// (leading gap, padding and trailing gap zeroed)
__global__ void __launch_bounds__(256) edits_plain_kernel(const uint4* __restrict__ ref_q, const uint32_t* __restrict__ ref_v,
                                                          int64_t ref_chunks, const int32_t* __restrict__ plain,
                                                          const int64_t* __restrict__ slot_off, uint4* __restrict__ q,
                                                          uint32_t* __restrict__ v) {
  const int32_t h = plain[blockIdx.y];
  const int64_t chunk0 = slot_off[h] >> 5;
  const int64_t terr = (slot_off[h + 1] - slot_off[h]) >> 5;  // padded text + trailing gap
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x - GAP_CHUNKS; c < terr; c += (int64_t)gridDim.x * blockDim.x) {
    const bool in = c >= 0 && c < ref_chunks;
    q[chunk0 + c] = in ? ref_q[c] : make_uint4(0u, 0u, 0u, 0u);
    v[chunk0 + c] = in ? ref_v[c] : 0u;
  }
}

struct EditWindowArgs {
  const uint4* ref_q;  // reference planes, chunk r = reference index >> 5; one zero chunk of slack behind
  const int64_t* edit_off;
  const int32_t* pos;
  const int32_t* reflen;
  const int32_t* altlen;
  const int64_t* altoff;
  const int32_t* outpos;
  const int32_t* edit_hap;  // haplotype of every edit (derive_kernel, pass 0)
  const uint8_t* pool;
  const int64_t* slot_off;
  const int32_t* len;
  int32_t n_hap;
  int64_t n_edits;
  uint4* q;
  uint32_t* v;
  uint32_t* nz;
  int32_t reach;  // chunks either side of an edit's own chunks
};

// One thread per (edit, k): chunk first - reach + k of the edit's window (and every 2 reach + 1
// chunks after it, for ALT texts that span more than one chunk). A chunk two windows share belongs to the
// earlier edit. The reference holds no lower-case base here (the caller checked), so case bits
// come from ALT text only.
__global__ void __launch_bounds__(128) edit_windows_kernel(const __grid_constant__ EditWindowArgs A) {
  const int KW = 2 * A.reach + 1;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t e = idx / KW;
  const int k = (int)(idx - e * KW);
  if (e >= A.n_edits) return;
  const int32_t h = __ldg(&A.edit_hap[e]);
  const int64_t e0 = A.edit_off[h], e1 = A.edit_off[h + 1];
  const int32_t L = A.len[h];
  const int64_t chunk0 = A.slot_off[h] >> 5;
  const int64_t terr = (A.slot_off[h + 1] - A.slot_off[h]) >> 5;
  const int64_t first = A.outpos[e] >> 5, last = ((int64_t)A.outpos[e] + A.altlen[e] - 1) >> 5;
  const int64_t owned_before = e > e0 ? (((int64_t)A.outpos[e - 1] + A.altlen[e - 1] - 1) >> 5) + A.reach : INT64_MIN;
  for (int64_t c = first - A.reach + k; c <= last + A.reach; c += KW) {
    if (c >= first && c <= last) atomicOr(&A.nz[(chunk0 + c) >> 5], 1u << ((chunk0 + c) & 31));
    if (c <= owned_before || c < -GAP_CHUNKS || c >= terr) continue;
    const int64_t j0 = c * 32;
    uint32_t pa = 0, pc = 0, pg = 0, pt = 0, pv = 0;
    if (j0 >= 0 && j0 < L) {
      const int64_t end = j0 + 32 < L ? j0 + 32 : L;
      int64_t el = e;  // last edit with outpos <= j0, or e0 - 1
      while (el >= e0 && A.outpos[el] > j0) --el;
      while (el + 1 < e1 && A.outpos[el + 1] <= j0) ++el;
      int64_t j = j0;
      while (j < end) {
        const int64_t alt_end = el >= e0 ? (int64_t)A.outpos[el] + A.altlen[el] : INT64_MIN;
        if (j < alt_end) {  // inside the ALT text of edit el
          const int64_t stop = alt_end < end ? alt_end : end;
          const uint8_t* src = A.pool + A.altoff[el] + (j - A.outpos[el]);
          for (; j < stop; ++j, ++src) {
            const uint32_t n = iupac_entry(*src) & 0xFu, bit = (uint32_t)(j - j0);
            pa |= (n & 1u) << bit;
            pc |= ((n >> 1) & 1u) << bit;
            pg |= ((n >> 2) & 1u) << bit;
            pt |= ((n >> 3) & 1u) << bit;
            pv |= 1u << bit;
          }
        } else {  // reference bases at a constant offset, up to the next edit
          const int64_t shift = el >= e0 ? (int64_t)A.pos[el] + A.reflen[el] - alt_end : 0;
          const int64_t nxt = el + 1 < e1 ? (int64_t)A.outpos[el + 1] : INT64_MAX;
          const int64_t stop = nxt < end ? nxt : end;
          const int n = (int)(stop - j);
          const int64_t r = j + shift;
          const uint4 q0 = __ldg(&A.ref_q[r >> 5]), q1 = __ldg(&A.ref_q[(r >> 5) + 1]);
          const uint32_t sh = (uint32_t)(r & 31), m = n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u), at = (uint32_t)(j - j0);
          pa |= (funnel_r(q0.x, q1.x, sh) & m) << at;
          pc |= (funnel_r(q0.y, q1.y, sh) & m) << at;
          pg |= (funnel_r(q0.z, q1.z, sh) & m) << at;
          pt |= (funnel_r(q0.w, q1.w, sh) & m) << at;
          j = stop;
        }
        while (el + 1 < e1 && A.outpos[el + 1] <= j) ++el;
      }
    }
    A.q[chunk0 + c] = make_uint4(pa, pc, pg, pt);
    A.v[chunk0 + c] = pv;
  }
}

int launch_pool_check(cudaStream_t st, const uint8_t* pool, int64_t n, unsigned long long* bad) {
  if (n <= 0) return HAWK_OK;
  pool_check_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pool, n, bad);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "pool_check_kernel launch");
}

int launch_edits_plain(cudaStream_t st, const void* ref_q, const uint32_t* ref_v, int64_t ref_chunks, const int32_t* plain,
                       int32_t n_plain, const int64_t* slot_off, void* q, uint32_t* v) {
  if (n_plain <= 0) return HAWK_OK;
  int64_t bx = (ref_chunks + 2 * GAP_CHUNKS + 4 + 255) / 256;
  if (bx > 1024) bx = 1024;
  edits_plain_kernel<<<dim3((unsigned)bx, (unsigned)n_plain), 256, 0, st>>>((const uint4*)ref_q, ref_v, ref_chunks, plain,
                                                                            slot_off, (uint4*)q, v);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "edits_plain_kernel launch");
}

int launch_edit_windows(cudaStream_t st, const void* ref_q, const int64_t* edit_off, const int32_t* pos, const int32_t* reflen,
                        const int32_t* altlen, const int64_t* altoff, const int32_t* outpos, const int32_t* edit_hap,
                        const uint8_t* pool, const int64_t* slot_off, const int32_t* len, int32_t n_hap, int64_t n_edits,
                        void* q, uint32_t* v, uint32_t* nz, int32_t reach) {
  if (n_edits <= 0 || n_hap <= 0) return HAWK_OK;
  EditWindowArgs A{(const uint4*)ref_q, edit_off, pos, reflen, altlen, altoff, outpos, edit_hap, pool, slot_off, len, n_hap,
                   n_edits, (uint4*)q, v, nz, reach};
  const int64_t threads = n_edits * (2 * reach + 1);
  edit_windows_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(A);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "edit_windows_kernel launch");
}

}  // namespace hawk
