// collapse_api.cu -- N2, the row collapse of the report (reports._collapse_report_entries,
// reports.py:958-1008, for the score-free column set): guide rows that agree in
// (start, stop, strand, origin, guide + PAM text -- case included) become one report row.
// gc_content is a function of the text, chr / pam_class / target are per region, so those are
// the distinguishing columns of the reference's groupby key.
//
//   collapse_key   per row: k_hi = start << 32 | stop; k_lo = 64-bit hash of (strand, origin, the
//                  G + P core characters of the window text)
//   sort           two stable LSD radix sorts (cub::DeviceRadixSort, library code: the one sort of
//                  the package, off the hot path) -> rows ordered by (k_hi, k_lo), ties in
//                  emission order, which is the order pandas' "first" aggregations see
//   collapse_head  head[k] = row k of that order starts a new group; equal hashes are confirmed
//                  on the actual bytes (a collision is reported, never merged)
// The host (crispr_hawk_b200/report_rows.py) orders the groups the way pandas sorts the groupby
// keys and joins the per-group strings (sample sets, haplotype ids): string assembly only.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/device/device_radix_sort.cuh>

#include "hawk_host.h"

namespace hawk {

__device__ __forceinline__ uint64_t mix_u64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

__global__ void collapse_key_kernel(const int32_t* __restrict__ hap, const uint8_t* __restrict__ strand,
                                    const int32_t* __restrict__ start, const int32_t* __restrict__ stop,
                                    const uint8_t* __restrict__ text, int32_t text_stride, int32_t core_len,
                                    const uint8_t* __restrict__ is_ref, int32_t n_hap, int64_t n, uint64_t* __restrict__ k_hi,
                                    uint64_t* __restrict__ k_lo, uint32_t* __restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  k_hi[i] = ((uint64_t)(uint32_t)start[i] << 32) | (uint32_t)stop[i];
  const uint8_t* core = text + i * (int64_t)text_stride + HAWK_GUIDESEQPAD;
  const uint32_t hp = (uint32_t)hap[i];  // a haplotype index the caller's is_ref does not cover counts as not REF
  uint64_t h = mix_u64(0x9E3779B97F4A7C15ull ^ ((uint64_t)strand[i] << 1) ^ (uint64_t)(hp < (uint32_t)n_hap ? is_ref[hp] : 0));
  for (int j = 0; j < core_len; j += 8) {
    uint64_t w = 0;
    for (int k = 0; k < 8 && j + k < core_len; ++k) w |= (uint64_t)core[j + k] << (8 * k);
    h = mix_u64(h ^ w);
  }
  k_lo[i] = h;
  idx[i] = (uint32_t)i;
}

__global__ void collapse_head_kernel(const uint32_t* __restrict__ perm, const uint64_t* __restrict__ k_hi,
                                     const uint64_t* __restrict__ k_lo, const int32_t* __restrict__ hap,
                                     const uint8_t* __restrict__ strand, const uint8_t* __restrict__ text,
                                     int32_t text_stride, int32_t core_len, const uint8_t* __restrict__ is_ref, int32_t n_hap,
                                     int64_t n, uint8_t* __restrict__ head, int* __restrict__ collision) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  if (k == 0) {
    head[0] = 1;
    return;
  }
  // k_hi / k_lo are in sorted order already (sorted keys); perm maps to the rows
  if (k_hi[k] != k_hi[k - 1] || k_lo[k] != k_lo[k - 1]) {
    head[k] = 1;
    return;
  }
  const uint32_t a = perm[k], b = perm[k - 1];
  const uint32_t ha = (uint32_t)hap[a], hb = (uint32_t)hap[b];
  bool same = strand[a] == strand[b] &&
              (ha < (uint32_t)n_hap ? is_ref[ha] : 0) == (hb < (uint32_t)n_hap ? is_ref[hb] : 0);
  const uint8_t* ca = text + (int64_t)a * text_stride + HAWK_GUIDESEQPAD;
  const uint8_t* cb = text + (int64_t)b * text_stride + HAWK_GUIDESEQPAD;
  for (int j = 0; j < core_len && same; ++j) same = ca[j] == cb[j];
  head[k] = same ? 0 : 1;
  if (!same) atomicExch(collision, 1);
}

}  // namespace hawk

using namespace hawk;

static void hawk_collapse_gather(cudaStream_t st, const uint32_t* idx, const uint64_t* src, uint64_t* dst, int64_t n);

extern "C" int hawk_result_collapse(hawk_result* r, const uint8_t* is_ref, int32_t n_hap, uint32_t* perm, uint8_t* head,
                                    int32_t* collision) {
  if (!r || !r->is_table || (r->n_guides > 0 && (!is_ref || !perm || !head)) || n_hap <= 0)
    return hawk_fail(HAWK_EINVAL, "hawk_result_collapse: needs the table of a phased / variant-free hawk_search");
  hawk_ctx* c = r->ctx;
  CKCUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const int64_t n = r->n_guides;
  if (collision) *collision = 0;
  if (n == 0) return HAWK_OK;
  const int core_len = r->params.guide_len + r->params.pam_len;
  DevBuf d_ref, khi[2], klo[2], idx[2], d_head, d_col, d_tmp;
  CK(upload(c, d_ref, is_ref, (size_t)n_hap));
  for (int k = 0; k < 2; ++k) {
    CK(khi[k].alloc(c, (size_t)n * 8));
    CK(klo[k].alloc(c, (size_t)n * 8));
    CK(idx[k].alloc(c, (size_t)n * 4));
  }
  CK(d_head.alloc(c, (size_t)n));
  CK(d_col.alloc(c, 4, true));
  const unsigned blocks = (unsigned)((n + 255) / 256);
  collapse_key_kernel<<<blocks, 256, 0, st>>>(r->hap.as<int32_t>(), r->strand.as<uint8_t>(), r->start.as<int32_t>(),
                                              r->stop.as<int32_t>(), r->text.as<uint8_t>(), r->text_stride, core_len,
                                              d_ref.as<uint8_t>(), n_hap, n, khi[0].as<uint64_t>(), klo[0].as<uint64_t>(),
                                              idx[0].as<uint32_t>());
  hawk_note_launch(1);
  CK(hawk_check_cuda(cudaGetLastError(), "collapse_key_kernel launch"));
  // LSD: by the hash first, then (stably) by (start, stop); the other key rides along as a second
  // value sort with the same permutation
  size_t tmp_bytes = 0, need = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, need, klo[0].as<uint64_t>(), klo[1].as<uint64_t>(), idx[0].as<uint32_t>(),
                                  idx[1].as<uint32_t>(), (int)n, 0, 64, st);
  tmp_bytes = need;
  CK(d_tmp.alloc(c, tmp_bytes));
  if (n >= 0x7FFFFFFFll) return hawk_fail(HAWK_ECAPACITY, "hawk_result_collapse: more than 2^31 rows");
  // pass 1: sort (k_lo, idx) by k_lo
  CKCUDA(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, klo[0].as<uint64_t>(), klo[1].as<uint64_t>(),
                                         idx[0].as<uint32_t>(), idx[1].as<uint32_t>(), (int)n, 0, 64, st));
  // pass 2: (start, stop) in pass-1 order, stable sort by it, carrying the permutation
  hawk_collapse_gather(st, idx[1].as<uint32_t>(), khi[0].as<uint64_t>(), khi[1].as<uint64_t>(), n);
  CKCUDA(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, khi[1].as<uint64_t>(), khi[0].as<uint64_t>(),
                                         idx[1].as<uint32_t>(), idx[0].as<uint32_t>(), (int)n, 0, 64, st));
  // khi[0] = sorted (start, stop), idx[0] = the final permutation; the hashes in that order are
  // gathered from klo[0], which still holds them in row order
  hawk_collapse_gather(st, idx[0].as<uint32_t>(), klo[0].as<uint64_t>(), klo[1].as<uint64_t>(), n);
  collapse_head_kernel<<<blocks, 256, 0, st>>>(idx[0].as<uint32_t>(), khi[0].as<uint64_t>(), klo[1].as<uint64_t>(),
                                               r->hap.as<int32_t>(), r->strand.as<uint8_t>(), r->text.as<uint8_t>(),
                                               r->text_stride, core_len, d_ref.as<uint8_t>(), n_hap, n, d_head.as<uint8_t>(),
                                               d_col.as<int>());
  hawk_note_launch(1);
  CK(hawk_check_cuda(cudaGetLastError(), "collapse_head_kernel launch"));
  c->d2h_bytes += n * 5 + 4;
  CKCUDA(cudaMemcpyAsync(perm, idx[0].p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  CKCUDA(cudaMemcpyAsync(head, d_head.p, (size_t)n, cudaMemcpyDeviceToHost, st));
  int col = 0;
  CKCUDA(cudaMemcpyAsync(&col, d_col.p, 4, cudaMemcpyDeviceToHost, st));
  CKCUDA(cudaStreamSynchronize(st));
  if (collision) *collision = col;
  return HAWK_OK;
}

namespace hawk {
__global__ void collapse_gather_kernel(const uint32_t* __restrict__ idx, const uint64_t* __restrict__ src,
                                       uint64_t* __restrict__ dst, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}
}  // namespace hawk

static void hawk_collapse_gather(cudaStream_t st, const uint32_t* idx, const uint64_t* src, uint64_t* dst, int64_t n) {
  if (n <= 0 || !src) return;
  hawk::collapse_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(idx, src, dst, n);
  hawk_note_launch(1);
}
