// stream_api.cu -- the streamed search of the host layer: one call takes the haplotypes of a
// region from HOST memory to the guide table in HOST memory, with the PCIe traffic of both
// directions overlapped (see include/hawkscan.h, hawk_search_stream).
//
// The haplotypes other than REF are cut into groups. Group g is searched as a batch of its own
// -- REF + the group's block, so the redundancy filter (search_guides.py:340-369) has its REF
// partners -- by the unchanged pipeline of api.cu, while the bulk copy of group g + 1's input
// runs on a second stream and the guide rows of group g - 1 leave on a third. The reference's
// emission order is haplotype-major (search_guides.py:530-547), so the table is the
// concatenation of the groups' tables; the REF rows are emitted by the first group only, and
// the first-seen bucket ids (:306-337) come from one (start, strand) table shared by all groups
// (a row's bucket is the smallest global row index with its key; that row lies in the same or
// an earlier group, so the id is final when the group's rows are written).
//
// Small transfers (metadata, totals) never touch the copy engines (hawk_ctx::small_h2d): the
// bulk copies own them for the whole call.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <functional>
#include <vector>

#include "hawk_core.h"
#include "hawk_kernels.h"
#include "hawk_post.h"
#include "hawk_host.h"

using namespace hawk;

extern "C" int32_t hawk_table_text_stride(int32_t pam_len, int32_t guide_len) {
  const int W = pam_len + guide_len + 2 * HAWK_GUIDESEQPAD;
  return (W + 15) / 16 * 16;
}

namespace {

struct Group {
  int32_t lo, hi;  // caller's haplotype indices [lo, hi) (REF, when it leads the batch, is added to every group)
};

struct Plan {
  int32_t ref_h = -1;       // caller's REF haplotype, -1: none
  bool ref_first = false;   // REF is haplotype 0: every group is REF + a block
  std::vector<Group> groups;
};

// REF first (or absent) -> blocks of roughly equal slot count; anything else -> one group that
// is the caller's batch as it stands
int make_plan(const int64_t* slot_off, int32_t n_hap, const uint8_t* is_ref, int32_t n_groups, int64_t group_bytes,
              Plan& P) {
  int32_t n_ref = 0;
  for (int32_t h = 0; h < n_hap; ++h)
    if (is_ref[h]) {
      if (P.ref_h < 0) P.ref_h = h;
      ++n_ref;
    }
  if (n_ref > 1)
    return hawk_fail(HAWK_EDUPREF, "hawk_search_stream: %d haplotypes are labelled REF; the reference aborts on "
                     "the duplicate REF guides this produces (search_guides.py:328-334)", n_ref);
  P.ref_first = P.ref_h == 0;
  const int32_t first = P.ref_first ? 1 : 0;
  if (P.ref_h > 0 || n_hap - first <= 1) {
    P.groups.push_back(Group{first, n_hap});
    return HAWK_OK;
  }
  const int64_t total = slot_off[n_hap] - slot_off[first];
  int64_t k = n_groups > 0 ? n_groups : (total + group_bytes - 1) / group_bytes;
  if (k < 1) k = 1;
  if (k > n_hap - first) k = n_hap - first;
  if (k > 256) k = 256;
  int32_t lo = first;
  for (int64_t g = 0; g < k && lo < n_hap; ++g) {
    const int64_t want = slot_off[first] + total * (g + 1) / k;
    int32_t hi = lo + 1;
    while (hi < n_hap && slot_off[hi + 1] <= want) ++hi;
    if (g == k - 1) hi = n_hap;
    P.groups.push_back(Group{lo, hi});
    lo = hi;
  }
  return HAWK_OK;
}

// genomic coordinate range of the whole batch (same rule as hawk_batch_set_posmap)
int coord_range(const int32_t* len, int32_t n_hap, const int64_t* seg_off, const int32_t* seg_rel,
                const int32_t* seg_gen, const uint8_t* seg_step, int32_t* gmin_out, int32_t* gmax_out) {
  int64_t gmin = INT64_MAX, gmax = INT64_MIN;
  for (int32_t h = 0; h < n_hap; ++h) {
    const int64_t s0 = seg_off[h], s1 = seg_off[h + 1];
    if (s1 <= s0) return hawk_fail(HAWK_EINVAL, "hawk_search_stream: haplotype %d has no posmap segment", h);
    for (int64_t k = s0; k < s1; ++k) {
      const int64_t rel_end = k + 1 < s1 ? seg_rel[k + 1] : len[h];
      const int64_t lo = seg_gen[k], hi = seg_gen[k] + (seg_step[k] && rel_end > seg_rel[k] ? rel_end - seg_rel[k] - 1 : 0);
      if (lo < gmin) gmin = lo;
      if (hi > gmax) gmax = hi;
    }
  }
  if (gmin > gmax) gmin = 0, gmax = -1;
  if (gmin < INT32_MIN || gmax > INT32_MAX) return hawk_fail(HAWK_EINVAL, "hawk_search_stream: coordinates exceed 32 bits");
  *gmin_out = (int32_t)gmin;
  *gmax_out = (int32_t)gmax;
  return HAWK_OK;
}

int ensure_streams(hawk_ctx* c) {
  if (!c->h2d_stream) CKCUDA(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
  if (!c->d2h_stream) CKCUDA(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
  return HAWK_OK;
}

struct EventPool {
  std::vector<cudaEvent_t> ev;
  ~EventPool() {
    for (auto e : ev) cudaEventDestroy(e);
  }
  int get(cudaEvent_t* out) {
    cudaEvent_t e;
    CKCUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ev.push_back(e);
    *out = e;
    return HAWK_OK;
  }
};

struct Totals {
  int64_t n_guides = 0, n_hits[2] = {0, 0}, scanned_bp = 0;
};

// The group loop shared by the two entry points. `prefetch(g)` queues the bulk input copy of
// group g on the context's H2D stream (may be empty); `make_batch(g, &batch)` builds the
// group's packed batch on the compute stream (after the prefetch event) with its coordinate
// maps attached.
int run_groups(hawk_ctx* c, const Plan& P, const hawk_params* params, const int32_t* scan_start,
               const int32_t* scan_stop, const uint8_t* is_ref, int32_t gmin, int32_t gmax,
               const std::function<int(size_t)>& prefetch,
               const std::function<int(size_t, hawk_batch**)>& make_batch, const hawk_table_out* out, Totals& T) {
  cudaStream_t st = c->stream;
  const int32_t stride = hawk_table_text_stride(params->pam_len, params->guide_len);
  if (out && out->text && out->text_stride != stride)
    return hawk_fail(HAWK_EINVAL, "hawk_search_stream: text_stride must be %d for this PAM / guide length", stride);
  const int64_t key_span = gmax >= gmin ? ((int64_t)gmax - gmin + 1) * 2 : 2;
  if (key_span > HAWK_DIRECT_KEY_SPAN)
    return hawk_fail(HAWK_EINVAL, "hawk_search_stream: coordinate range too wide for the shared first-seen table "
                     "(use hawk_batch_create + hawk_search)");
  CK(ensure_streams(c));
  DevBuf key_table;
  CK(key_table.alloc(c, (size_t)key_span * 4));
  CKCUDA(cudaMemsetAsync(key_table.p, 0xFF, (size_t)key_span * 4, st));
  EventPool events;
  const size_t n_groups = P.groups.size();
  hawk_result* pending = nullptr;  // previous group's table: its rows may still be leaving
  cudaEvent_t pending_ev = nullptr;
  int64_t ref_hits[2] = {0, 0};
  bool overflow = false;
  int rc = HAWK_OK;
  if (n_groups) rc = prefetch(0);
  for (size_t g = 0; g < n_groups && rc == HAWK_OK; ++g) {
    const Group& G = P.groups[g];
    if (g + 1 < n_groups && (rc = prefetch(g + 1))) break;
    hawk_batch* b = nullptr;
    if ((rc = make_batch(g, &b))) break;
    // the group's per-haplotype search arguments
    const int32_t n_local = (P.ref_first ? 1 : 0) + (G.hi - G.lo);
    std::vector<int32_t> a(n_local), e(n_local);
    std::vector<uint8_t> ir(n_local);
    int32_t k = 0;
    if (P.ref_first) {
      a[0] = scan_start[0], e[0] = scan_stop[0], ir[0] = 1;
      k = 1;
    }
    for (int32_t h = G.lo; h < G.hi; ++h, ++k) a[k] = scan_start[h], e[k] = scan_stop[h], ir[k] = is_ref[h];
    StreamLink link;
    link.row_base = T.n_guides;
    link.drop_ref = P.ref_first && g > 0;
    link.ref_local = P.ref_first ? 0 : -1;
    link.ref_global = 0;
    link.hap_add = P.ref_first ? G.lo - 1 : G.lo;
    link.key_table = key_table.as<uint32_t>();
    link.key_min = gmin;
    hawk_result* r = nullptr;
    rc = hawk_search_impl(c, b, params, a.data(), e.data(), ir.data(), &link, &r);
    hawk_batch_destroy(b);
    if (rc) break;
    if (g == 0) ref_hits[0] = r->ref_hits[0], ref_hits[1] = r->ref_hits[1];
    for (int s = 0; s < 2; ++s) T.n_hits[s] += r->n_hits[s] - (link.drop_ref ? ref_hits[s] : 0);
    T.scanned_bp += r->scanned_bp - (link.drop_ref && e[0] > a[0] ? (int64_t)e[0] - (a[0] < 0 ? 0 : a[0]) : 0);
    const int64_t n = r->n_guides, base = T.n_guides;
    T.n_guides += n;
    // the previous group's rows: the compute stream may reuse their memory once they have left
    if (pending) {
      cudaStreamWaitEvent(st, pending_ev, 0);
      hawk_result_destroy(pending);
      pending = nullptr;
    }
    if (out && T.n_guides > out->capacity) overflow = true;  // keep counting, stop copying
    if (out && !overflow && n > 0) {
      cudaStream_t ds = c->d2h_stream;
      const size_t m = (size_t)n;
      cudaError_t ce = cudaSuccess;
      auto cp = [&](void* dst, const void* src, size_t bytes) {
        if (dst && ce == cudaSuccess) {
          ce = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ds);
          c->d2h_bytes += (int64_t)bytes;
        }
      };
      cp(out->hap ? out->hap + base : nullptr, r->hap.p, m * 4);
      cp(out->strand ? out->strand + base : nullptr, r->strand.p, m);
      cp(out->pos ? out->pos + base : nullptr, r->pos.p, m * 4);
      cp(out->start ? out->start + base : nullptr, r->start.p, m * 4);
      cp(out->stop ? out->stop + base : nullptr, r->stop.p, m * 4);
      cp(out->bucket ? out->bucket + base : nullptr, r->bucket.p, m * 4);
      cp(out->text ? out->text + (size_t)base * stride : nullptr, r->text.p, m * (size_t)stride);
      if ((rc = hawk_check_cuda(ce, "guide table D2H"))) {
        hawk_result_destroy(r);
        break;
      }
      if ((rc = events.get(&pending_ev)) || (rc = hawk_check_cuda(cudaEventRecord(pending_ev, ds), "cudaEventRecord"))) {
        cudaStreamSynchronize(ds);
        hawk_result_destroy(r);
        break;
      }
      pending = r;
    } else {
      hawk_result_destroy(r);
    }
  }
  // drain: input copies that were queued for a group never reached, rows still leaving
  cudaStreamSynchronize(c->h2d_stream);
  cudaError_t ce = cudaStreamSynchronize(c->d2h_stream);
  if (pending) hawk_result_destroy(pending);
  cudaStreamSynchronize(st);
  if (rc == HAWK_OK) rc = hawk_check_cuda(ce, "guide table D2H");
  if (rc == HAWK_OK && overflow)
    rc = hawk_fail(HAWK_ECAPACITY, "hawk_search_stream: %lld guide rows, output capacity %lld",
                   (long long)T.n_guides, (long long)out->capacity);
  return rc;
}

void store_totals(const Totals& T, int64_t* n_guides, int64_t* n_hits, int64_t* scanned_bp) {
  if (n_guides) *n_guides = T.n_guides;
  if (n_hits) n_hits[0] = T.n_hits[0], n_hits[1] = T.n_hits[1];
  if (scanned_bp) *scanned_bp = T.scanned_bp;
}

}  // namespace

// host-only: the group plan of the streamed search (for callers sizing their buffers, and for
// the CPU tests): groups [lo[g], hi[g]) of haplotype indices; returns their number, or a
// negative error code
extern "C" int32_t hawk_stream_plan(const int64_t* slot_off, int32_t n_hap, const uint8_t* is_ref, int32_t n_groups,
                                    int32_t* lo, int32_t* hi, int32_t capacity) {
  if (n_hap < 0 || (n_hap > 0 && (!slot_off || !is_ref))) return hawk_fail(HAWK_EINVAL, "hawk_stream_plan: bad arguments");
  if (n_hap == 0) return 0;
  Plan P;
  const int rc = make_plan(slot_off, n_hap, is_ref, n_groups, 192ll << 20, P);
  if (rc != HAWK_OK) return rc;
  if ((int64_t)P.groups.size() > capacity) return hawk_fail(HAWK_ECAPACITY, "hawk_stream_plan: %zu groups", P.groups.size());
  for (size_t g = 0; g < P.groups.size(); ++g) {
    if (lo) lo[g] = P.groups[g].lo;
    if (hi) hi[g] = P.groups[g].hi;
  }
  return (int32_t)P.groups.size();
}

extern "C" int hawk_search_stream(hawk_ctx* c, const uint8_t* ascii, const int64_t* slot_off, const int32_t* len,
                                  int32_t n_hap, const int64_t* seg_off, const int32_t* seg_rel,
                                  const int32_t* seg_gen, const uint8_t* seg_step, const hawk_params* params,
                                  const int32_t* scan_start, const int32_t* scan_stop, const uint8_t* is_ref,
                                  int32_t n_groups, const hawk_table_out* out, int64_t* n_guides, int64_t* n_hits,
                                  int64_t* scanned_bp, int64_t* bad_slot) {
  if (!c || !params || n_hap < 0 ||
      (n_hap > 0 && (!ascii || !slot_off || !len || !seg_off || !seg_rel || !seg_gen || !seg_step || !scan_start ||
                     !scan_stop || !is_ref)))
    return hawk_fail(HAWK_EINVAL, "hawk_search_stream: bad arguments");
  if (params->flags & HAWK_F_UNPHASED)
    return hawk_fail(HAWK_EINVAL, "hawk_search_stream: phased / variant-free searches only (use hawk_search)");
  CKCUDA(cudaSetDevice(c->device));
  if (bad_slot) *bad_slot = -1;
  Totals T;
  if (n_hap == 0) {
    store_totals(T, n_guides, n_hits, scanned_bp);
    return HAWK_OK;
  }
  std::vector<int64_t> expect(n_hap + 1);
  CK(hawk_layout(len, n_hap, expect.data(), nullptr));
  for (int32_t h = 0; h <= n_hap; ++h)
    if (slot_off[h] != expect[h])
      return hawk_fail(HAWK_EINVAL, "hawk_search_stream: slot_off[%d] does not follow hawk_layout", h);
  Plan P;
  CK(make_plan(slot_off, n_hap, is_ref, n_groups, 192ll << 20, P));
  int32_t gmin = 0, gmax = -1;
  CK(coord_range(len, n_hap, seg_off, seg_rel, seg_gen, seg_step, &gmin, &gmax));
  CK(ensure_streams(c));
  // a group's texts on the device: [leading gap + REF region][the block's regions], which is a
  // valid slot space of its own because hawk_layout is translation invariant
  const int64_t head = P.ref_first ? slot_off[1] : HAWK_SLOT_GAP;
  int64_t max_bytes = 0;
  for (auto& G : P.groups) max_bytes = std::max(max_bytes, head + slot_off[G.hi] - slot_off[G.lo]);
  CKCUDA(cudaStreamSynchronize(c->stream));  // the staging blocks go to another stream: no pending users
  DevBuf staging[2];
  CK(staging[0].alloc(c, (size_t)max_bytes));
  if (P.groups.size() > 1) CK(staging[1].alloc(c, (size_t)max_bytes));
  EventPool events;
  std::vector<cudaEvent_t> ready(P.groups.size(), nullptr);

  auto prefetch = [&](size_t g) -> int {
    const Group& G = P.groups[g];
    uint8_t* d = staging[g & 1].as<uint8_t>();
    cudaStream_t hs = c->h2d_stream;
    // head: the caller's slot space up to the end of REF (or just its leading gap)
    CKCUDA(cudaMemcpyAsync(d, ascii, (size_t)head, cudaMemcpyHostToDevice, hs));
    CKCUDA(cudaMemcpyAsync(d + head, ascii + slot_off[G.lo], (size_t)(slot_off[G.hi] - slot_off[G.lo]),
                           cudaMemcpyHostToDevice, hs));
    c->h2d_bytes += head + (slot_off[G.hi] - slot_off[G.lo]);
    CK(events.get(&ready[g]));
    CKCUDA(cudaEventRecord(ready[g], hs));
    return HAWK_OK;
  };

  auto make_batch = [&](size_t g, hawk_batch** out_b) -> int {
    const Group& G = P.groups[g];
    const int32_t n_ref = P.ref_first ? 1 : 0, n_local = n_ref + (G.hi - G.lo);
    std::vector<int32_t> l(n_local);
    if (n_ref) l[0] = len[0];
    for (int32_t h = G.lo; h < G.hi; ++h) l[n_ref + h - G.lo] = len[h];
    std::vector<int64_t> so(n_local + 1);
    CK(hawk_layout(l.data(), n_local, so.data(), nullptr));
    CKCUDA(cudaStreamWaitEvent(c->stream, ready[g], 0));
    hawk_batch* b = nullptr;
    int64_t bad = -1;
    int rc = batch_create_impl(c, staging[g & 1].as<uint8_t>(), true, so.data(), l.data(), n_local, &b, &bad);
    if (rc == HAWK_EIUPAC && bad_slot) {
      // back to the caller's slot numbering
      *bad_slot = bad < head ? bad : bad - head + slot_off[G.lo];
    }
    CK(rc);
    // the group's run-length coordinate maps: REF's segments, then the block's (contiguous)
    std::vector<int64_t> go(n_local + 1);
    std::vector<int32_t> rel, gen;
    std::vector<uint8_t> step;
    go[0] = 0;
    auto add = [&](int32_t h, int32_t k) {
      rel.insert(rel.end(), seg_rel + seg_off[h], seg_rel + seg_off[h + 1]);
      gen.insert(gen.end(), seg_gen + seg_off[h], seg_gen + seg_off[h + 1]);
      step.insert(step.end(), seg_step + seg_off[h], seg_step + seg_off[h + 1]);
      go[k + 1] = go[k] + (seg_off[h + 1] - seg_off[h]);
    };
    if (n_ref) add(0, 0);
    for (int32_t h = G.lo; h < G.hi; ++h) add(h, n_ref + h - G.lo);
    rc = hawk_batch_set_posmap(b, go.data(), rel.data(), gen.data(), step.data());
    if (rc) {
      hawk_batch_destroy(b);
      return rc;
    }
    *out_b = b;
    return HAWK_OK;
  };

  c->bulk_h2d = true;
  int rc = run_groups(c, P, params, scan_start, scan_stop, is_ref, gmin, gmax, prefetch, make_batch, out, T);
  c->bulk_h2d = false;
  store_totals(T, n_guides, n_hits, scanned_bp);
  return rc;
}

extern "C" int hawk_search_stream_edits(hawk_ctx* c, const uint8_t* ref_ascii, int64_t ref_len, int32_t region_start,
                                        int32_t n_hap, const int64_t* edit_off, const int32_t* edit_pos,
                                        const int32_t* edit_reflen, const int32_t* edit_altlen,
                                        const int64_t* edit_altoff, const uint8_t* alt_pool, int64_t alt_pool_len,
                                        const hawk_params* params, const int32_t* scan_start,
                                        const int32_t* scan_stop, const uint8_t* is_ref, int32_t n_groups,
                                        const hawk_table_out* out, int64_t* n_guides, int64_t* n_hits,
                                        int64_t* scanned_bp) {
  if (!c || !params || !ref_ascii || ref_len <= 0 || n_hap < 0 ||
      (n_hap > 0 && (!edit_off || !scan_start || !scan_stop || !is_ref)))
    return hawk_fail(HAWK_EINVAL, "hawk_search_stream_edits: bad arguments");
  if (params->flags & HAWK_F_UNPHASED)
    return hawk_fail(HAWK_EINVAL, "hawk_search_stream_edits: phased / variant-free searches only");
  CKCUDA(cudaSetDevice(c->device));
  Totals T;
  if (n_hap == 0) {
    store_totals(T, n_guides, n_hits, scanned_bp);
    return HAWK_OK;
  }
  // group sizes by haplotype count: the texts are about ref_len each
  std::vector<int64_t> pseudo(n_hap + 1);
  for (int32_t h = 0; h <= n_hap; ++h) pseudo[h] = (int64_t)h * ref_len;
  Plan P;
  // fewer, larger groups than the text path: only edit lists go up and the planes are built around
  // the edits alone (edits_kernels.cu), so the call is bound by the rows leaving; a group only has
  // to be small enough for the first rows to leave early (measured, 5 G haplotype-bp: 2 - 4 groups
  // without the window text, 4 - 8 with it; every group costs ~0.5 ms of host round trips)
  CK(make_plan(pseudo.data(), n_hap, is_ref, n_groups, 1280ll << 20, P));
  if (P.ref_first && edit_off[1] != edit_off[0])
    return hawk_fail(HAWK_EINVAL, "hawk_search_stream_edits: the REF haplotype must have no edits");

  auto prefetch = [&](size_t) -> int { return HAWK_OK; };  // only edit lists cross PCIe: nothing bulky

  auto make_batch = [&](size_t g, hawk_batch** out_b) -> int {
    const Group& G = P.groups[g];
    const int32_t n_ref = P.ref_first ? 1 : 0, n_local = n_ref + (G.hi - G.lo);
    std::vector<int64_t> eo(n_local + 1);
    const int64_t e0 = edit_off[G.lo];
    eo[0] = 0;
    if (n_ref) eo[1] = 0;
    for (int32_t h = G.lo; h < G.hi; ++h) eo[n_ref + h - G.lo + 1] = edit_off[h + 1] - e0;
    return hawk_batch_create_from_edits(c, ref_ascii, ref_len, region_start, n_local, eo.data(), edit_pos + e0,
                                        edit_reflen + e0, edit_altlen + e0, edit_altoff + e0, alt_pool, alt_pool_len,
                                        out_b, nullptr);
  };

  // edits never create coordinates outside the reference
  int rc = run_groups(c, P, params, scan_start, scan_stop, is_ref, region_start, (int32_t)(region_start + ref_len - 1),
                      prefetch, make_batch, out, T);
  store_totals(T, n_guides, n_hits, scanned_bp);
  return rc;
}
