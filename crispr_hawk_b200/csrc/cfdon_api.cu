// cfdon_api.cu -- N4: the CFDon specificity score of every guide row against the REF guide of its
// (start, strand) key (scoring.py:303-387 cfdon_score / group_guides_position,
// scores/crisprhawk_scores.py:65-87 cfdon, scores/cfdscore/cfdscore.py:53-95 compute_cfd) on the
// device-resident table of a phased / variant-free hawk_search. The grouping the reference builds
// with a dict is the table's own bucket column (first row of the key); the REF guide of a key, if
// any, IS that first row when REF is haplotype 0 -- REF rows are emitted first. The mismatch /
// PAM factor tables are the caller's (the reference loads them from its model files at run time).
#include <cuda_runtime.h>
#include <stdint.h>

#include "hawk_core.h"
#include "hawk_kernels.h"
#include "hawk_host.h"

namespace hawk {

__global__ void cfdon_kernel(const int32_t* __restrict__ hap, const uint8_t* __restrict__ strand,
                             const uint32_t* __restrict__ bucket, const uint8_t* __restrict__ text, int32_t text_stride,
                             int W, int G, int P, int right, const uint8_t* __restrict__ is_ref, int32_t n_hap,
                             const double* __restrict__ mm, const double* __restrict__ pam2, int64_t n,
                             double* __restrict__ out, unsigned long long* __restrict__ bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t b = bucket[i];
  double v = __longlong_as_double(0x7FF8000000000000ll);  // no REF guide at this key: NaN (crisprhawk_scores.py:81-82)
  if ((int64_t)b < n && (uint32_t)hap[b] < (uint32_t)n_hap && is_ref[hap[b]]) {
    if (!cfdon_row(text + (int64_t)b * text_stride, text + i * (int64_t)text_stride, W, G, P, right, strand[i], mm, pam2, &v))
      atomicMin(bad, (unsigned long long)i);
  }
  out[i] = v;
}

}  // namespace hawk

using namespace hawk;

extern "C" int hawk_result_cfdon(hawk_result* r, const uint8_t* is_ref, int32_t n_hap, const double* mm, const double* pam2,
                                 double* scores, int64_t* bad_row) {
  if (!r || !r->is_table || !mm || !pam2 || n_hap <= 0 || !is_ref || (r->n_guides > 0 && !scores))
    return hawk_fail(HAWK_EINVAL, "hawk_result_cfdon: needs the table of a phased / variant-free hawk_search and both tables");
  if (bad_row) *bad_row = -1;
  if (n_hap > 0 && !is_ref[0]) {
    for (int32_t h = 1; h < n_hap; ++h)
      if (is_ref[h])
        return hawk_fail(HAWK_EINVAL, "hawk_result_cfdon: REF must be haplotype 0 (its guide is looked up as the first row of a key)");
  }
  hawk_ctx* c = r->ctx;
  CKCUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const int64_t n = r->n_guides;
  if (n == 0) return HAWK_OK;
  DevBuf d_ref, d_mm, d_pam, d_out, d_bad;
  CK(upload(c, d_ref, is_ref, (size_t)n_hap));
  CK(upload(c, d_mm, mm, 20 * 16 * sizeof(double)));
  CK(upload(c, d_pam, pam2, 16 * sizeof(double)));
  CK(d_out.alloc(c, (size_t)n * 8));
  const unsigned long long none = ~0ull;
  CK(upload(c, d_bad, &none, 8));
  cfdon_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
      r->hap.as<int32_t>(), r->strand.as<uint8_t>(), r->bucket.as<uint32_t>(), r->text.as<uint8_t>(), r->text_stride, r->window,
      r->params.guide_len, r->params.pam_len, r->params.right, d_ref.as<uint8_t>(), n_hap, d_mm.as<double>(), d_pam.as<double>(), n,
      d_out.as<double>(), d_bad.as<unsigned long long>());
  hawk_note_launch(1);
  CK(hawk_check_cuda(cudaGetLastError(), "cfdon_kernel launch"));
  c->d2h_bytes += n * 8 + 8;
  CKCUDA(cudaMemcpyAsync(scores, d_out.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  unsigned long long bad = none;
  CKCUDA(cudaMemcpyAsync(&bad, d_bad.p, 8, cudaMemcpyDeviceToHost, st));
  CKCUDA(cudaStreamSynchronize(st));
  if (bad != none) {
    if (bad_row) *bad_row = (int64_t)bad;
    return hawk_fail(HAWK_ECFD, "hawk_result_cfdon: row %llu: a mismatch or PAM letter outside A, C, G, T, or a key the "
                     "tables do not hold (the reference raises KeyError there)", bad);
  }
  return HAWK_OK;
}
