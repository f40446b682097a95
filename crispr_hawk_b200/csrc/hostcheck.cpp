// hostcheck.cpp -- TEST SUPPORT, not a product path.
//
// Compiles the __host__ __device__ core of the kernels (hawk_core.h) for the
// CPU and drives it serially, so the `-m "not gpu"` tests can check the exact
// bit logic the CUDA kernels execute (pack, per-chunk scan with fused filters,
// run-length posmap, REF-core comparison, IUPAC resolution) against the oracle
// in a container without a GPU. The block-level machinery of the kernels
// (look-back, scans, merge, hashing) is replaced by trivial serial code here and
// is only exercised by the GPU tests. libhawkscan.so never links or loads this.
#include <stdint.h>
#include <limits>
#include <string.h>

#include <algorithm>
#include <map>
#include <vector>

#include "hawk_core.h"

using namespace hawk;

extern "C" {

// pack_chunk_v3 (what the kernels run) against pack_chunk on n_chunks * 32 arbitrary bytes: returns the
// number of chunks whose `invalid` words differ, or -- on chunks without an invalid byte -- whose
// planes / case word differ
int64_t hawkcheck_pack_v3_diff(const uint8_t* bytes, int64_t n_chunks) {
  int64_t bad = 0;
  for (int64_t c = 0; c < n_chunks; ++c) {
    uint32_t w[8];
    memcpy(w, bytes + 32 * c, 32);
    const hawk::PackedChunk x = hawk::pack_chunk(w), y = hawk::pack_chunk_v3(w);
    if (x.invalid != y.invalid) ++bad;
    else if (!x.invalid && (x.a != y.a || x.c != y.c || x.g != y.g || x.t != y.t || x.v != y.v)) ++bad;
  }
  return bad;
}

// planes_to_chars32 against the letters it must give back: pack_chunk on n_chunks * 32 valid bytes
// (IUPAC letters of either case), then planes -> characters with the first `keep` of them valid;
// returns the number of chunks that do not come back as the input (zero bytes past `keep`)
int64_t hawkcheck_chars32_diff(const uint8_t* bytes, int64_t n_chunks, int32_t keep) {
  int64_t bad = 0;
  const uint32_t valid = keep >= 32 ? 0xFFFFFFFFu : ((1u << keep) - 1u);
  for (int64_t c = 0; c < n_chunks; ++c) {
    uint32_t w[8], out[8];
    memcpy(w, bytes + 32 * c, 32);
    const hawk::PackedChunk x = hawk::pack_chunk(w);
    hawk::planes_to_chars32(x.a, x.c, x.g, x.t, x.v, valid, out);
    uint8_t got[32];
    memcpy(got, out, 32);
    for (int i = 0; i < 32; ++i) {
      uint8_t want = bytes[32 * c + i];
      if (want == 0) want = '?';  // an unused slot has nibble 0: the table's '?'
      if (i >= keep) want = 0;
      if (got[i] != want) { ++bad; break; }
    }
  }
  return bad;
}

// K1 on the CPU: ascii slot space -> planes; returns first invalid slot or -1
int64_t hawkcheck_pack(const uint8_t* ascii, int64_t total_slots, uint32_t* q /*4 per chunk*/,
                       uint32_t* v) {
  int64_t bad = -1;
  for (int64_t c = 0; c < total_slots / 32; ++c) {
    uint32_t words[8];
    memcpy(words, ascii + c * 32, 32);
    PackedChunk o = pack_chunk(words);
    q[4 * c + 0] = o.a;
    q[4 * c + 1] = o.c;
    q[4 * c + 2] = o.g;
    q[4 * c + 3] = o.t;
    v[c] = o.v;
    if (o.invalid && bad < 0) bad = c * 32 + __builtin_ctz(o.invalid);
  }
  return bad;
}

struct CheckTable {
  std::vector<int32_t> hap, pos, start, stop;
  std::vector<uint8_t> strand;
  std::vector<int64_t> bucket;
  std::vector<uint8_t> text;
  std::vector<uint64_t> hits[2];
  int window = 0;
  int err = 0;
};

static BatchView make_view(const uint32_t* q, const uint32_t* v, const int64_t* slot_off,
                           const int32_t* len, const int32_t* a, const int32_t* b,
                           const uint8_t* is_ref, int32_t n_hap, const int64_t* seg_off,
                           const int32_t* seg_rel, const int32_t* seg_gen, const uint8_t* seg_step,
                           const int64_t* va_off, const int32_t* va_idx, const int64_t* va_ent_off,
                           const uint8_t* va_ref) {
  BatchView B{};
  B.q = (const Planes*)q;
  B.v = v;
  B.slot_off = slot_off;
  B.len = len;
  B.scan_start = a;
  B.scan_stop = b;
  B.is_ref = is_ref;
  B.n_hap = n_hap;
  B.seg_off = seg_off;
  B.seg_rel = seg_rel;
  B.seg_gen = seg_gen;
  B.seg_step = seg_step;
  B.va_off = va_off;
  B.va_idx = va_idx;
  B.va_ent_off = va_ent_off;
  B.va_ref = va_ref;
  return B;
}

static void scan_all(const BatchView& B, const ScanConst& K, std::vector<uint64_t> hits[2]) {
  for (int32_t h = 0; h < B.n_hap; ++h) {
    HapScan H = load_hap_scan(B, K, h);
    if (H.b <= H.a) continue;
    for (int64_t c = H.a >> 5; c < ((int64_t)H.b + 31) >> 5; ++c) {
      uint32_t out[2], raw[2];
      scan_chunk(B, K, H, c, out, raw);
      for (int s = 0; s < 2; ++s) {
        uint32_t bits = K.raw ? raw[s] : out[s];
        while (bits) {
          int bit = __builtin_ctz(bits);
          bits &= bits - 1;
          hits[s].push_back(((uint64_t)(uint32_t)h << 32) | (uint64_t)(c * 32 + bit));
        }
      }
    }
  }
}

void* hawkcheck_search(const uint32_t* q, const uint32_t* v, const int64_t* slot_off,
                       const int32_t* len, const int32_t* a, const int32_t* b,
                       const uint8_t* is_ref, int32_t n_hap, const int64_t* seg_off,
                       const int32_t* seg_rel, const int32_t* seg_gen, const uint8_t* seg_step,
                       const int64_t* va_off, const int32_t* va_idx, const int64_t* va_ent_off,
                       const uint8_t* va_ref, const hawk_params* params, int raw) {
  CheckTable* T = new CheckTable();
  BatchView B = make_view(q, v, slot_off, len, a, b, is_ref, n_hap, seg_off, seg_rel, seg_gen,
                          seg_step, va_off, va_idx, va_ent_off, va_ref);
  ScanConst K = make_scan_const(*params, raw);
  const int W = K.C + 2 * HAWK_GUIDESEQPAD;
  T->window = W;
  // the coarse segment index rows_fast uses on the device (api.cu ensure_seg_index), so that the
  // CPU tests walk the indexed search of row_coords too
  std::vector<int32_t> seg_idx;
  if (seg_off && n_hap > 0) {
    int32_t max_len = 0;
    for (int32_t h = 0; h < n_hap; ++h) max_len = len[h] > max_len ? len[h] : max_len;
    const int32_t stride = (max_len >> HAWK_SEG_IDX_SHIFT) + 2;
    seg_idx.resize((size_t)stride * n_hap);
    for (int32_t h = 0; h < n_hap; ++h)
      for (int32_t k = 0; k < stride; ++k) seg_idx[(size_t)h * stride + k] = seg_index_entry(seg_off, seg_rel, h, k);
    B.seg_idx = seg_idx.data();
    B.seg_idx_stride = stride;
  }
  scan_all(B, K, T->hits);
  if (raw) return T;
  int32_t ref_h = -1;
  for (int32_t h = 0; h < n_hap; ++h)
    if (is_ref[h]) { ref_h = h; break; }
  // per stream: rows (hap,pos,start,stop,text,keep)
  struct Row { int32_t h, pos, start, stop; int s; std::vector<uint8_t> text; };
  std::vector<Row> rows[2];
  for (int s = 0; s < 2; ++s) {
    const std::vector<uint64_t>& recs = T->hits[s];
    int64_t rlo = 0, rhi = 0;
    if (ref_h >= 0) {
      uint64_t klo = (uint64_t)(uint32_t)ref_h << 32, khi = ((uint64_t)(uint32_t)ref_h + 1) << 32;
      rlo = std::lower_bound(recs.begin(), recs.end(), klo) - recs.begin();
      rhi = std::lower_bound(recs.begin(), recs.end(), khi) - recs.begin();
    }
    auto ref_pivot = [&](int32_t start) -> int32_t {
      if (ref_h < 0) return -1;
      int64_t i = find_ref_partner(B, K, recs.data(), rlo, rhi, ref_h, s, start);
      if (i >= rhi) return -1;
      int32_t rpivot = (int32_t)(recs[i] & 0xFFFFFFFFu) + K.geom[s].c0;
      int32_t rstart = posmap_eval(seg_rel, seg_gen, seg_step, seg_off[ref_h], seg_off[ref_h + 1], rpivot);
      return rstart == start ? rpivot : -1;
    };
    for (size_t i = 0; i < recs.size(); ++i) {
      int32_t h = (int32_t)(recs[i] >> 32), pos = (int32_t)(recs[i] & 0xFFFFFFFFu);
      RowCoords rc = row_coords(B, K, h, pos, s);
      int64_t chunk0 = slot_off[h] >> 5;
      int32_t w0 = pos + K.geom[s].w0;
      bool rp = K.geom[s].c0 == 0;
      int k0 = rp ? HAWK_GUIDESEQPAD : W - HAWK_GUIDESEQPAD - K.P;
      if (!K.unphased) {
        if (ref_h >= 0 && !is_ref[h]) {
          int32_t rpivot = ref_pivot(rc.start);
          if (rpivot >= 0 && cores_equal(B.q, chunk0, rc.pivot, slot_off[ref_h] >> 5, rpivot, K.C)) continue;
        }
        Row r{h, pos, rc.start, rc.stop, s, std::vector<uint8_t>(W)};
        for (int j = 0; j < W; ++j) {
          char ch = nibble_letter(nibble_at(B.q, chunk0, w0 + j));
          r.text[j] = (uint8_t)(lower_at(B.v, chunk0, w0 + j) ? ch + 32 : ch);
        }
        rows[s].push_back(std::move(r));
        continue;
      }
      // unphased: resolve
      std::vector<Column> cols(W);
      uint64_t prod = 1;
      bool ok = true;
      for (int j = 0; j < W && ok; ++j) {
        uint32_t code = (j >= k0 && j < k0 + K.P) ? K.pat[s][j - k0] : 0u;
        if (!load_column(B, h, chunk0, w0 + j, code, cols[j])) { T->err = HAWK_EALLELES; ok = false; break; }
        prod *= column_count(cols[j]);
      }
      if (!ok) return T;
      for (uint64_t t = 0; t < prod; ++t) {
        Row r{h, pos, rc.start, rc.stop, s, std::vector<uint8_t>(W)};
        uint64_t x = t;
        for (int j = W - 1; j >= 0; --j) {
          uint32_t cc = column_count(cols[j]);
          r.text[j] = (uint8_t)column_char(B, cols[j], (uint32_t)(x % cc));
          x /= cc;
        }
        if (ref_h >= 0 && !is_ref[h]) {
          int32_t rpivot = ref_pivot(rc.start);
          if (rpivot >= 0) {
            bool same = true;
            for (int j = 0; j < K.C && same; ++j)
              same = (iupac_entry(r.text[HAWK_GUIDESEQPAD + j]) & 15u) ==
                     nibble_at(B.q, slot_off[ref_h] >> 5, rpivot + j);
            if (same) continue;
          }
        }
        rows[s].push_back(std::move(r));
      }
    }
  }
  // emission order: haplotype, strand, then stream order
  std::vector<const Row*> order;
  size_t i0 = 0, i1 = 0;
  while (i0 < rows[0].size() || i1 < rows[1].size()) {
    bool take0 = i1 >= rows[1].size() || (i0 < rows[0].size() && rows[0][i0].h <= rows[1][i1].h);
    order.push_back(take0 ? &rows[0][i0++] : &rows[1][i1++]);
  }
  std::map<std::pair<int32_t, int>, int64_t> first;
  for (size_t i = 0; i < order.size(); ++i) {
    const Row& r = *order[i];
    T->hap.push_back(r.h);
    T->strand.push_back((uint8_t)r.s);
    T->pos.push_back(r.pos);
    T->start.push_back(r.start);
    T->stop.push_back(r.stop);
    auto it = first.emplace(std::make_pair(r.start, r.s), (int64_t)i).first;
    T->bucket.push_back(it->second);
    T->text.insert(T->text.end(), r.text.begin(), r.text.end());
  }
  return T;
}

int64_t hawkcheck_n(void* t) { return (int64_t)((CheckTable*)t)->hap.size(); }
int64_t hawkcheck_nhits(void* t, int s) { return (int64_t)((CheckTable*)t)->hits[s].size(); }
int hawkcheck_err(void* t) { return ((CheckTable*)t)->err; }
int hawkcheck_window(void* t) { return ((CheckTable*)t)->window; }
void hawkcheck_fetch(void* tp, int32_t* hap, uint8_t* strand, int32_t* pos, int32_t* start,
                     int32_t* stop, int64_t* bucket, uint8_t* text) {
  CheckTable* t = (CheckTable*)tp;
  size_t n = t->hap.size();
  if (!n) return;
  memcpy(hap, t->hap.data(), n * 4);
  memcpy(strand, t->strand.data(), n);
  memcpy(pos, t->pos.data(), n * 4);
  memcpy(start, t->start.data(), n * 4);
  memcpy(stop, t->stop.data(), n * 4);
  memcpy(bucket, t->bucket.data(), n * 8);
  memcpy(text, t->text.data(), t->text.size());
}
void hawkcheck_fetch_hits(void* tp, int s, uint64_t* out) {
  CheckTable* t = (CheckTable*)tp;
  if (!t->hits[s].empty()) memcpy(out, t->hits[s].data(), t->hits[s].size() * 8);
}
void hawkcheck_free(void* t) { delete (CheckTable*)t; }


// N2 on the CPU: the kernels' own per-row logic (annot_row_variants, annot_text_byte,
// annot_gc_counts of hawk_core.h) over a guide table, serially. gv_off has n + 1 entries, gv_idx
// room for n * (G + P) indices; returns the number of variant references or -1 when the
// reference's assert would fire.
int64_t hawkcheck_annotate(const int64_t* seg_off, const int32_t* seg_rel, const int32_t* seg_gen,
                           const uint8_t* seg_step, const int64_t* var_off, const int32_t* var_pos,
                           const int32_t* var_reflen, const int32_t* var_altlen, const int64_t* var_altoff,
                           const uint8_t* alt_pool, const hawk_params* params, int64_t n, const int32_t* hap,
                           const uint8_t* strand, const int32_t* pos, const int32_t* stop, const uint8_t* text,
                           int32_t text_stride, uint8_t* rc_text, int32_t* gc_num, int32_t* gc_den, int64_t* gv_off,
                           int32_t* gv_idx) {
  BatchView B{};
  B.seg_off = seg_off;
  B.seg_rel = seg_rel;
  B.seg_gen = seg_gen;
  B.seg_step = seg_step;
  const ScanConst K = make_scan_const(*params, 0);
  const VariantView V{var_off, var_pos, var_reflen, var_altlen, var_altoff, alt_pool, 0};
  const int W = K.C + 2 * HAWK_GUIDESEQPAD;
  int64_t out = 0;
  bool would_assert = false;
  gv_off[0] = 0;
  for (int64_t r = 0; r < n; ++r) {
    const uint8_t* src = text + (size_t)r * text_stride;
    annot_row_variants(B, K, V, hap[r], strand[r], pos[r], stop[r], src + HAWK_GUIDESEQPAD,
                       [&](int32_t j) { gv_idx[out++] = j; }, &would_assert);
    gv_off[r + 1] = out;
    for (int i = 0; i < text_stride; ++i) rc_text[(size_t)r * text_stride + i] = annot_text_byte(src, W, strand[r], i);
    annot_gc_counts(K, strand[r], src, &gc_num[r], &gc_den[r]);
  }
  return would_assert ? -1 : out;
}

// N4: cfdon_row over a whole table (the loop of cfdon_kernel); returns the first row where the
// reference would raise KeyError, or -1
int64_t hawkcheck_cfdon(const int32_t* hap, const uint8_t* strand, const uint32_t* bucket, const uint8_t* text,
                        int32_t text_stride, int32_t W, int32_t G, int32_t P, int32_t right, const uint8_t* is_ref,
                        const double* mm, const double* pam2, int64_t n, double* out) {
  int64_t bad = -1;
  for (int64_t i = 0; i < n; ++i) {
    const uint32_t b = bucket[i];
    double v = std::numeric_limits<double>::quiet_NaN();
    if (is_ref[hap[b]] &&
        !hawk::cfdon_row(text + (int64_t)b * text_stride, text + i * (int64_t)text_stride, W, G, P, right, strand[i], mm, pam2, &v) &&
        bad < 0)
      bad = i;
    out[i] = v;
  }
  return bad;
}

// N4, second half: feature_byte / onehot_channel over a whole table (the loops of
// featurize_kernel); returns the first row with a letter outside A, C, G, T, or -1
int64_t hawkcheck_featurize(const uint8_t* strand, const uint8_t* text, int32_t text_stride, int32_t W, int32_t lead,
                            int64_t n, uint8_t* kmers, float* onehot) {
  const int L = hawk::feature_len(W, lead);
  int64_t bad = -1;
  for (int64_t i = 0; i < n; ++i) {
    const uint8_t* src = text + i * (int64_t)text_stride;
    for (int j = 0; j < L; ++j) {
      const uint8_t u = hawk::feature_byte(src, W, strand[i], lead, j);
      const int ch = hawk::onehot_channel(u);
      if (kmers) kmers[i * L + j] = u;
      if (onehot)
        for (int k = 0; k < 4; ++k) onehot[(i * 4 + k) * L + j] = ch == k ? 1.0f : 0.0f;
      if (ch < 0 && bad < 0) bad = i;
    }
  }
  return bad;
}

}  // extern "C"
