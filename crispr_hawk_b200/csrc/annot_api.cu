// annot_api.cu -- host layer of N2 (include/hawkscan.h: hawk_batch_set_variants,
// hawk_result_annotate, hawk_result_fetch_variants): the reference's post-search pure
// functions on guides (annotation.py:563-572) evaluated on the device-resident guide table.
#include <cuda_runtime.h>
#include <stdint.h>

#include "hawk_core.h"
#include "hawk_kernels.h"
#include "hawk_post.h"
#include "hawk_host.h"

using namespace hawk;

extern "C" int hawk_batch_set_variants(hawk_batch* b, const int64_t* var_off, const int32_t* var_pos,
                                       const int32_t* var_reflen, const int32_t* var_altlen,
                                       const int64_t* var_altoff, const uint8_t* alt_pool, int64_t alt_pool_len) {
  if (!b || !var_off) return hawk_fail(HAWK_EINVAL, "hawk_batch_set_variants: bad arguments");
  hawk_ctx* c = b->ctx;
  CKCUDA(cudaSetDevice(c->device));
  const int64_t n = b->n_hap ? var_off[b->n_hap] : 0;
  if (n < 0 || (n > 0 && (!var_pos || !var_reflen || !var_altlen || !var_altoff || !alt_pool)))
    return hawk_fail(HAWK_EINVAL, "hawk_batch_set_variants: bad arguments");
  for (int32_t h = 0; h < b->n_hap; ++h) {
    if (var_off[h + 1] < var_off[h]) return hawk_fail(HAWK_EINVAL, "hawk_batch_set_variants: var_off must ascend");
    for (int64_t j = var_off[h]; j < var_off[h + 1]; ++j) {
      if (j > var_off[h] && var_pos[j] < var_pos[j - 1])
        return hawk_fail(HAWK_EINVAL, "hawk_batch_set_variants: variants of haplotype %d are not sorted by position", h);
      if (j > var_off[h] && var_pos[j] == var_pos[j - 1])
        return hawk_fail(HAWK_EINVAL, "hawk_batch_set_variants: haplotype %d has two variants at position %d (the reference's "
                         "own annotation depends on Python's set order there: annotate such a list with its functions)", h, var_pos[j]);
      if (var_reflen[j] < 1 || var_altlen[j] < 1 || var_altoff[j] < 0 || var_altoff[j] + var_altlen[j] > alt_pool_len)
        return hawk_fail(HAWK_EINVAL, "hawk_batch_set_variants: bad allele of variant %lld", (long long)j);
    }
  }
  const int64_t zero64 = 0;
  const int32_t zero32 = 0;
  const uint8_t zero8 = 0;
  const size_t m = (size_t)(n > 0 ? n : 1);
  CK(upload(c, b->var_off, b->n_hap ? (const void*)var_off : (const void*)&zero64, (size_t)(b->n_hap + 1) * 8));
  CK(upload(c, b->var_pos, n ? (const void*)var_pos : (const void*)&zero32, m * 4));
  CK(upload(c, b->var_rl, n ? (const void*)var_reflen : (const void*)&zero32, m * 4));
  CK(upload(c, b->var_al, n ? (const void*)var_altlen : (const void*)&zero32, m * 4));
  CK(upload(c, b->var_ao, n ? (const void*)var_altoff : (const void*)&zero64, m * 8));
  CK(upload(c, b->var_pool, alt_pool_len > 0 ? alt_pool : &zero8, (size_t)(alt_pool_len > 0 ? alt_pool_len : 1)));
  CKCUDA(cudaStreamSynchronize(c->stream));
  b->var_pos_base = 0;
  b->has_variants = true;
  return HAWK_OK;
}

extern "C" int hawk_result_annotate(hawk_result* r, hawk_batch* b, uint8_t* rc_text, int32_t* gc_num,
                                    int32_t* gc_den, int64_t* gv_off, int64_t* gv_total) {
  if (!r || !b) return hawk_fail(HAWK_EINVAL, "hawk_result_annotate: bad arguments");
  if (!r->is_table) return hawk_fail(HAWK_EINVAL, "hawk_result_annotate: needs the result of a phased / variant-free hawk_search");
  if (r->ctx != b->ctx) return hawk_fail(HAWK_EINVAL, "hawk_result_annotate: result and batch belong to different contexts");
  if (!b->has_posmap) return hawk_fail(HAWK_EINVAL, "hawk_result_annotate: the batch has no coordinate maps");
  const bool want_var = gv_off != nullptr;
  if (want_var && !b->has_variants)
    return hawk_fail(HAWK_EINVAL, "hawk_result_annotate: call hawk_batch_set_variants first");
  hawk_ctx* c = r->ctx;
  CKCUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const int64_t n = r->n_guides;
  r->gv_total = 0;
  if (gv_total) *gv_total = 0;
  if (want_var) gv_off[0] = 0;
  if (n == 0) return HAWK_OK;
  const ScanConst K = make_scan_const(r->params, 0);
  if (rc_text || gc_num || gc_den) {
    DevBuf d_rc, d_num, d_den;
    CK(d_rc.alloc(c, (size_t)n * r->text_stride));
    CK(d_num.alloc(c, (size_t)n * 4));
    CK(d_den.alloc(c, (size_t)n * 4));
    CK(launch_annot_text(st, K, r->strand.as<uint8_t>(), r->text.as<uint8_t>(), r->text_stride, n, d_rc.as<uint8_t>(),
                         d_num.as<int32_t>(), d_den.as<int32_t>()));
    if (rc_text) CKCUDA(cudaMemcpyAsync(rc_text, d_rc.p, (size_t)n * r->text_stride, cudaMemcpyDeviceToHost, st));
    if (gc_num) CKCUDA(cudaMemcpyAsync(gc_num, d_num.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (gc_den) CKCUDA(cudaMemcpyAsync(gc_den, d_den.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    c->d2h_bytes += (int64_t)n * ((rc_text ? r->text_stride : 0) + (gc_num ? 4 : 0) + (gc_den ? 4 : 0));
    CKCUDA(cudaStreamSynchronize(st));
  }
  if (!want_var) return HAWK_OK;
  BatchView B{};
  B.n_hap = b->n_hap;
  B.seg_off = b->seg_off.as<int64_t>();
  B.seg_rel = b->seg_rel.as<int32_t>();
  B.seg_gen = b->seg_gen.as<int32_t>();
  B.seg_step = b->seg_step.as<uint8_t>();
  VariantView V{b->var_off.as<int64_t>(), b->var_pos.as<int32_t>(), b->var_rl.as<int32_t>(), b->var_al.as<int32_t>(),
                b->var_ao.as<int64_t>(), b->var_pool.as<uint8_t>(), b->var_pos_base};
  DevBuf d_cnt, d_off, d_tiles, d_flags;
  CK(d_cnt.alloc(c, (size_t)n * 4));
  CK(d_off.alloc(c, (size_t)(n + 1) * 8));
  CK(d_tiles.alloc(c, ((size_t)scan_tiles(n) + 2) * 8));
  CK(d_flags.alloc(c, 16, true));
  CK(launch_annot_variants(st, B, K, V, r->hap.as<int32_t>(), r->strand.as<uint8_t>(), r->pos.as<int32_t>(),
                           r->stop.as<int32_t>(), r->text.as<uint8_t>(), r->text_stride, n, d_cnt.as<uint32_t>(), nullptr,
                           nullptr, d_flags.as<int32_t>(), 0));
  CK(exclusive_scan_u32(st, d_cnt.as<uint32_t>(), n, d_off.as<uint64_t>(), d_tiles.as<uint64_t>(),
                        d_off.as<uint64_t>() + n));
  uint64_t total = 0;
  CK(c->small_d2h_sync(&total, d_off.as<uint64_t>() + n, 8));
  CK(r->gv_idx.alloc(c, (size_t)(total ? total : 1) * 4));
  CK(launch_annot_variants(st, B, K, V, r->hap.as<int32_t>(), r->strand.as<uint8_t>(), r->pos.as<int32_t>(),
                           r->stop.as<int32_t>(), r->text.as<uint8_t>(), r->text_stride, n, nullptr, d_off.as<uint64_t>(),
                           r->gv_idx.as<int32_t>(), d_flags.as<int32_t>(), 1));
  CKCUDA(cudaMemcpyAsync(gv_off, d_off.p, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, st));
  c->d2h_bytes += (int64_t)(n + 1) * 8;
  int32_t flags[4] = {0, 0, 0, 0};
  CK(c->small_d2h_sync(flags, d_flags.p, 16));
  if (flags[0])
    return hawk_fail(HAWK_EASSERT, "hawk_result_annotate: an indel's anchor is upper-case at the first base of a guide; "
                     "the reference asserts here (annotation.py:191, _find_insertion_stop)");
  r->gv_total = (int64_t)total;
  if (gv_total) *gv_total = (int64_t)total;
  return HAWK_OK;
}

extern "C" int hawk_result_fetch_variants(hawk_result* r, int32_t* gv_idx) {
  if (!r) return hawk_fail(HAWK_EINVAL, "hawk_result_fetch_variants: null result");
  if (r->gv_total == 0) return HAWK_OK;
  if (!gv_idx) return hawk_fail(HAWK_EINVAL, "hawk_result_fetch_variants: null output");
  hawk_ctx* c = r->ctx;
  CKCUDA(cudaSetDevice(c->device));
  CKCUDA(cudaMemcpyAsync(gv_idx, r->gv_idx.p, (size_t)r->gv_total * 4, cudaMemcpyDeviceToHost, c->stream));
  c->d2h_bytes += r->gv_total * 4;
  CKCUDA(cudaStreamSynchronize(c->stream));
  return HAWK_OK;
}
