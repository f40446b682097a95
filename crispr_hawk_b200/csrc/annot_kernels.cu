// annot_kernels.cu -- N2 of the scope table: the post-search pure functions of the reference's
// annotation.py on the device-resident guide table (phased / variant-free searches).
//
//   annot_variants  polish_guide_variants (annotation.py:246-281): which of the haplotype's
//                   variants are visible in a guide. One thread per row walks the core's G + P
//                   bases: genomic coordinate through the run-length posmap (segment pointer
//                   advanced incrementally), variant at that coordinate by a monotone walk of
//                   the haplotype's sorted variant table, then _check_insertion / _check_snv
//                   (:197-243) on the window text. Two passes (count, exclusive scan, write)
//                   give a CSR list of variant indices per row.
//   annot_text      reverse_guides (:27-51, guide.py:245-255): rows of strand 1 are replaced by
//                   their IUPAC-aware reverse complement (utils.py:46-79, case kept);
//                   gc_content (:513-541): G+C+S and A+C+G+T+S+W counts of the guide part
//                   (Bio.SeqUtils.gc_fraction, ambiguous="remove"), PAM excluded.
#include <cuda_runtime.h>
#include <stdint.h>

#include "hawk_core.h"
#include "hawk_kernels.h"
#include "hawk_post.h"

namespace hawk {

__device__ __forceinline__ bool is_upper(uint8_t c) { return c >= 'A' && c <= 'Z'; }
__device__ __forceinline__ uint8_t to_upper(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }

struct AnnotVarArgs {
  BatchView B;
  ScanConst K;
  VariantView V;
  const int32_t* hap;
  const uint8_t* strand;
  const int32_t* pos;
  const int32_t* stop;
  const uint8_t* text;
  int32_t text_stride;
  int64_t n;
  uint32_t* cnt;          // pass 0
  const uint64_t* off;    // pass 1
  int32_t* idx;           // pass 1: variant index inside its haplotype's list
  int32_t* flags;         // [0] != 0: the reference's assert in _find_insertion_stop would fire
};

template <int PASS>
__global__ void __launch_bounds__(256) annot_variants_kernel(const __grid_constant__ AnnotVarArgs A) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A.n) return;
  const int32_t h = A.hap[r];
  const int s = A.strand[r];
  const int32_t pivot = A.pos[r] + A.K.geom[s].c0;
  const int32_t stop = A.stop[r];
  const uint8_t* core = A.text + (size_t)r * A.text_stride + HAWK_GUIDESEQPAD;
  const int C = A.K.C;
  const int64_t v0 = A.V.var_off[h], v1 = A.V.var_off[h + 1];
  uint32_t found = 0;
  int64_t out = PASS ? (int64_t)A.off[r] : 0;
  if (v1 > v0) {
    // segment holding the core's first base
    const int64_t s0 = A.B.seg_off[h], s1 = A.B.seg_off[h + 1];
    int64_t k = s0, hi = s1;
    while (hi - k > 1) {
      const int64_t mid = (k + hi) >> 1;
      if (A.B.seg_rel[mid] <= pivot) k = mid; else hi = mid;
    }
    int32_t p = A.B.seg_gen[k] + (A.B.seg_step[k] ? (pivot - A.B.seg_rel[k]) : 0);
    // first variant at or after the core's first coordinate
    int64_t j = v0, jh = v1;
    while (j < jh) {
      const int64_t mid = (j + jh) >> 1;
      if (A.V.var_pos[mid] + A.V.pos_base < p) j = mid + 1; else jh = mid;
    }
    int64_t last = -1;
    for (int i = 0; i < C && j < v1; ++i) {
      const int32_t idx = pivot + i;
      while (k + 1 < s1 && A.B.seg_rel[k + 1] <= idx) ++k;
      p = A.B.seg_gen[k] + (A.B.seg_step[k] ? (idx - A.B.seg_rel[k]) : 0);
      while (j < v1 && A.V.var_pos[j] + A.V.pos_base < p) ++j;
      int offset = 0;  // annotation.py:264: reset per base, carried over the variants of one base
      for (int64_t jj = j; jj < v1 && A.V.var_pos[jj] + A.V.pos_base == p; ++jj) {
        const int32_t rl = A.V.var_reflen[jj], al = A.V.var_altlen[jj];
        const uint8_t* alt = A.V.alt_pool + A.V.var_altoff[jj];
        const bool is_snv = rl == al;
        if (!is_snv) offset = rl < al ? al - rl : 0;
        const int seglen = (i + offset + 1 <= C ? offset + 1 : C - i);
        const uint8_t* seg = core + i;
        bool ok = false;
        if (!is_snv) {  // _check_insertion (:197-226)
          if (i == 0) {
            int up = -1;  // _find_insertion_stop: first upper-case character, 0 when none
            for (int t = 0; t < seglen; ++t)
              if (is_upper(seg[t])) { up = t; break; }
            if (up == 0) atomicExch(&A.flags[0], 1);  // the reference asserts here
            const int kk = up < 0 ? 0 : up;
            bool e = kk <= al;
            for (int t = 0; e && t < kk; ++t) e = alt[al - kk + t] == to_upper(seg[t]);
            ok = e;
          }
          if (!ok && p == stop) {
            bool e = seglen <= al;
            for (int t = 0; e && t < seglen; ++t) e = alt[t] == to_upper(seg[t]);
            ok = e;
          }
        }
        if (!ok) {  // _check_snv (:229-243): all lower-case and equal to the ALT allele
          bool e = seglen == al;
          for (int t = 0; e && t < seglen; ++t) e = !is_upper(seg[t]) && alt[t] == to_upper(seg[t]);
          ok = e;
        }
        if (ok && jj != last) {  // a set: the bases of one insertion share the anchor's coordinate
          last = jj;
          if (PASS) A.idx[out++] = (int32_t)(jj - v0);
          ++found;
        }
      }
    }
  }
  if (!PASS) A.cnt[r] = found;
}

struct AnnotTextArgs {
  ScanConst K;
  const uint8_t* strand;
  const uint8_t* text;
  int32_t text_stride;
  int64_t n;
  uint8_t* rc_text;
  int32_t* gc_num;
  int32_t* gc_den;
};

// complement of one IUPAC letter, case kept (utils.py:46-79); other bytes unchanged
__device__ __forceinline__ uint8_t rc_char(uint8_t c) {
  const uint8_t u = to_upper(c);
  uint8_t o;
  switch (u) {
    case 'A': o = 'T'; break;
    case 'C': o = 'G'; break;
    case 'G': o = 'C'; break;
    case 'T': o = 'A'; break;
    case 'U': o = 'A'; break;
    case 'R': o = 'Y'; break;
    case 'Y': o = 'R'; break;
    case 'M': o = 'K'; break;
    case 'K': o = 'M'; break;
    case 'H': o = 'D'; break;
    case 'D': o = 'H'; break;
    case 'B': o = 'V'; break;
    case 'V': o = 'B'; break;
    default: o = u; break;  // N, S, W
  }
  return (uint8_t)(o | (c & 0x20));
}

// one thread per 16-byte word of the output text (coalesced 128-bit stores; the source bytes of
// a row are shared by its words through L1); the row's first word also counts G/C
__global__ void __launch_bounds__(256) annot_text_kernel(const __grid_constant__ AnnotTextArgs A) {
  const int wpr = A.text_stride >> 4;  // words per row
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = t / wpr;
  if (r >= A.n) return;
  const int w = (int)(t - r * wpr);
  const int s = A.strand[r];
  const int W = A.K.C + 2 * HAWK_GUIDESEQPAD;
  const uint8_t* src = A.text + (size_t)r * A.text_stride;
  uint32_t word[4] = {0, 0, 0, 0};
#pragma unroll
  for (int b = 0; b < 16; ++b) {
    const int i = 16 * w + b;
    uint8_t c = 0;
    if (i < W) c = s ? rc_char(src[W - 1 - i]) : src[i];
    word[b >> 2] |= (uint32_t)c << (8 * (b & 3));
  }
  reinterpret_cast<uint4*>(A.rc_text + (size_t)r * A.text_stride)[w] = make_uint4(word[0], word[1], word[2], word[3]);
  if (w == 0) {
    // guide part of the forward core: right' = right XOR strand (search_guides.py:538)
    const bool rp = (A.K.right != 0) != (s == 1);
    const int g0 = HAWK_GUIDESEQPAD + (rp ? A.K.P : 0);
    int num = 0, den = 0;
    for (int i = 0; i < A.K.G; ++i) {
      const uint8_t u = to_upper(src[g0 + i]);
      const int gc = (u == 'G') | (u == 'C') | (u == 'S');
      num += gc;
      den += gc | (u == 'A') | (u == 'T') | (u == 'W');
    }
    A.gc_num[r] = num;
    A.gc_den[r] = den;
  }
}

int launch_annot_variants(cudaStream_t st, const BatchView& B, const ScanConst& K, const VariantView& V,
                          const int32_t* hap, const uint8_t* strand, const int32_t* pos, const int32_t* stop,
                          const uint8_t* text, int32_t text_stride, int64_t n, uint32_t* cnt, const uint64_t* off,
                          int32_t* idx, int32_t* flags, int pass) {
  if (n <= 0) return HAWK_OK;
  AnnotVarArgs A{B, K, V, hap, strand, pos, stop, text, text_stride, n, cnt, off, idx, flags};
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (pass == 0) annot_variants_kernel<0><<<blocks, 256, 0, st>>>(A);
  else annot_variants_kernel<1><<<blocks, 256, 0, st>>>(A);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "annot_variants_kernel launch");
}

int launch_annot_text(cudaStream_t st, const ScanConst& K, const uint8_t* strand, const uint8_t* text,
                      int32_t text_stride, int64_t n, uint8_t* rc_text, int32_t* gc_num, int32_t* gc_den) {
  if (n <= 0) return HAWK_OK;
  AnnotTextArgs A{K, strand, text, text_stride, n, rc_text, gc_num, gc_den};
  const int64_t words = n * (text_stride >> 4);
  annot_text_kernel<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(A);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "annot_text_kernel launch");
}

}  // namespace hawk
