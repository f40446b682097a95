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
  const uint8_t* core = A.text + (size_t)r * A.text_stride + HAWK_GUIDESEQPAD;
  int64_t out = PASS ? (int64_t)A.off[r] : 0;
  bool would_assert = false;
  const uint32_t found = annot_row_variants(
      A.B, A.K, A.V, A.hap[r], A.strand[r], A.pos[r], A.stop[r], core,
      [&](int32_t j) {
        if (PASS) A.idx[out++] = j;
      },
      &would_assert);
  if (would_assert) atomicExch(&A.flags[0], 1);
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
  for (int b = 0; b < 16; ++b) word[b >> 2] |= (uint32_t)annot_text_byte(src, W, s, 16 * w + b) << (8 * (b & 3));
  reinterpret_cast<uint4*>(A.rc_text + (size_t)r * A.text_stride)[w] = make_uint4(word[0], word[1], word[2], word[3]);
  if (w == 0) annot_gc_counts(A.K, s, src, &A.gc_num[r], &A.gc_den[r]);
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
