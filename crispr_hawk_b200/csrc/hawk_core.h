// hawk_core.h -- bit-level building blocks shared by the CUDA kernels and the
// host-side self-check harness (hostcheck.cpp). Everything here is
// __host__ __device__ so the exact code the kernels run can be exercised on a
// CPU against the oracle before it ever reaches a GPU.
//
// Layout recap (include/hawkscan.h): one chunk = 32 consecutive base slots;
// planes q[chunk] = {A,C,G,T} (bit i of each word belongs to slot 32*chunk+i),
// v[chunk] = lower-case ("variant base") bits.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HAWK_HD __host__ __device__ __forceinline__
#define HAWK_UNROLL _Pragma("unroll")
#else
#define HAWK_HD inline
#define HAWK_UNROLL
#endif

namespace hawk {

struct Planes {
  uint32_t a, c, g, t;
};

// ---- ASCII -> 4-bit IUPAC mask (encoder.py:18-34) -------------------------------
// entry: bits 0-3 nibble, bit 4 lower-case, 0x80 invalid (non-IUPAC, non-NUL)
HAWK_HD uint8_t iupac_entry(uint8_t ch) {
  if (ch == 0) return 0;  // unused slot
  uint8_t lower = (ch >= 'a' && ch <= 'z') ? 0x10 : 0;
  uint8_t up = lower ? (uint8_t)(ch - 32) : ch;
  uint8_t n;
  switch (up) {
    case 'A': n = 1; break;
    case 'C': n = 2; break;
    case 'G': n = 4; break;
    case 'T': n = 8; break;
    case 'R': n = 5; break;
    case 'Y': n = 10; break;
    case 'S': n = 6; break;
    case 'W': n = 9; break;
    case 'K': n = 12; break;
    case 'M': n = 3; break;
    case 'B': n = 14; break;
    case 'D': n = 13; break;
    case 'H': n = 11; break;
    case 'V': n = 7; break;
    case 'N': n = 15; break;
    default: return 0x80;
  }
  return (uint8_t)(n | lower);
}

// K1 inner step: 4 table entries (one per ASCII byte, iupac_entry format) packed in a
// word -> bit `b` of every entry gathered into 4 consecutive bits.
// (y & 0x01010101) * 0x01020408 moves bits 0,8,16,24 to bits 24..27; bits 28..31 stay 0.
HAWK_HD uint32_t gather_entry_bit(uint32_t entries, int b) {
  return (((entries >> b) & 0x01010101u) * 0x01020408u) >> 24;
}

struct PackedChunk {
  uint32_t a, c, g, t, v, invalid;
};

// 32 ASCII bytes (8 little-endian words) -> plane words; `entry(byte)` is the LUT
template <class F>
HAWK_HD PackedChunk pack_chunk(const uint32_t* words, F&& entry) {
  PackedChunk o{0, 0, 0, 0, 0, 0};
  HAWK_UNROLL
  for (int k = 0; k < 8; ++k) {
    uint32_t x = words[k];
    uint32_t e = (uint32_t)entry(x & 0xFF) | ((uint32_t)entry((x >> 8) & 0xFF) << 8) |
                 ((uint32_t)entry((x >> 16) & 0xFF) << 16) | ((uint32_t)entry(x >> 24) << 24);
    int sh = 4 * k;
    o.a |= gather_entry_bit(e, 0) << sh;
    o.c |= gather_entry_bit(e, 1) << sh;
    o.g |= gather_entry_bit(e, 2) << sh;
    o.t |= gather_entry_bit(e, 3) << sh;
    o.v |= gather_entry_bit(e, 4) << sh;
    o.invalid |= gather_entry_bit(e, 7) << sh;
  }
  return o;
}

// nibble -> IUPAC letter (inverse of the table above), upper-case
HAWK_HD char nibble_letter(uint32_t n) {
  // "?ACMGRSVTWYHKDBN" as two 64-bit immediates (no local array, no stack frame)
  n &= 15;
  const uint64_t t0 = ((uint64_t)'?') | ((uint64_t)'A' << 8) | ((uint64_t)'C' << 16) |
                      ((uint64_t)'M' << 24) | ((uint64_t)'G' << 32) | ((uint64_t)'R' << 40) |
                      ((uint64_t)'S' << 48) | ((uint64_t)'V' << 56);
  const uint64_t t1 = ((uint64_t)'T') | ((uint64_t)'W' << 8) | ((uint64_t)'Y' << 16) |
                      ((uint64_t)'H' << 24) | ((uint64_t)'K' << 32) | ((uint64_t)'D' << 40) |
                      ((uint64_t)'B' << 48) | ((uint64_t)'N' << 56);
  return (char)(((n & 8) ? t1 : t0) >> ((n & 7) * 8));
}

HAWK_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, sh);
#else
  sh &= 31;
  return sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
#endif
}

HAWK_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}

// bits i of a 32-slot chunk starting at p0 with lo <= p0 + i < hi
HAWK_HD uint32_t interval_mask(int64_t lo, int64_t hi, int64_t p0) {
  int64_t l = lo - p0, h = hi - p0;
  if (l < 0) l = 0;
  if (h > 32) h = 32;
  if (h <= l) return 0u;
  uint32_t upper = (h == 32) ? 0xFFFFFFFFu : ((1u << h) - 1u);
  uint32_t lower = (1u << l) - 1u;  // l < 32 here
  return upper & ~lower;
}

// plane union selected by an IUPAC nibble: bit i set <=> nibble(slot i) & code != 0
HAWK_HD uint32_t select_planes(const Planes& p, uint32_t code) {
  uint32_t r = 0;
  if (code & 1) r |= p.a;
  if (code & 2) r |= p.c;
  if (code & 4) r |= p.g;
  if (code & 8) r |= p.t;
  return r;
}

// search_guides.py:32-46 (match) for the 32 positions of one chunk at once:
// bit i of the result <=> for every k < P: pattern[k] & nibble(p0 + i + k) != 0.
// `cur` / `nxt` are the planes of the chunk and of the following chunk (P <= 16
// never reaches further). Pattern nibbles equal to 15 (N) match every real base.
HAWK_HD uint32_t match_chunk(const Planes& cur, const Planes& nxt, const uint8_t* pattern, int P) {
  uint32_t m = 0xFFFFFFFFu;
  for (int k = 0; k < P; ++k) {
    uint32_t code = pattern[k];
    if (code == 15u) continue;
    uint32_t lo = select_planes(cur, code), hi = select_planes(nxt, code);
    m &= funnel_r(lo, hi, (uint32_t)k);
  }
  return m;
}

// Per-strand geometry of search_guides.py:134-160, :395-420, :372-392.
// rp = right' (right XOR strand, :538). Window [pos + w0, pos + w1), core
// [pos + c0, pos + c0 + C), pivot = pos + c0.
struct StrandGeom {
  int32_t c0;   // core start relative to pos: 0 (right') or -G
  int32_t w0;   // window start relative to pos
  int32_t lo;   // smallest admissible pos (window inside the haplotype)
  int32_t hi_sub;  // admissible pos <= len - hi_sub
  int32_t stop_off;  // genomic stop = posmap[pos + stop_off] (:278-279)
};

HAWK_HD StrandGeom strand_geom(int G, int P, bool rp, bool unphased, int pad) {
  StrandGeom s;
  if (rp) {
    s.c0 = 0;
    s.w0 = -pad;
    s.lo = pad;                                   // pos - PAD >= 0
    s.hi_sub = G + P + pad + (unphased ? 1 : 0);  // pos + G + P + PAD <= len (< len when unphased)
    s.stop_off = G + P;
  } else {
    s.c0 = -G;
    s.w0 = -G - pad;
    s.lo = G + pad;      // pos - G - PAD >= 0 (is_pamhit_valid is the same bound)
    s.hi_sub = P + pad;  // pos + P + PAD <= len
    s.stop_off = P;
  }
  return s;
}

// bit x of a bit-vector stored as 32-bit words, via accessor f(word_index)
template <class F>
HAWK_HD uint32_t bits32_at(F&& word, int64_t bit) {
  int64_t w = bit >> 5;  // arithmetic shift: floor for negatives
  uint32_t sh = (uint32_t)(bit & 31);
  uint32_t lo = word(w);
  if (sh == 0) return lo;
  return funnel_r(lo, word(w + 1), sh);
}

// search_guides.py:468-471 for 32 positions at once: bit i <=> some variant
// (lower-case) bit inside [p0 + i + c0, p0 + i + c0 + C). `vword(k)` returns
// case word k of the haplotype (0 outside it).
template <class F>
HAWK_HD uint32_t core_variant_mask(F&& vword, int64_t p0, int c0, int C) {
  // sliding OR by doubling over a 32-bit lane view: S_w(x) = OR v[x .. x+w)
  // evaluated for x = p0 + c0 + i, i in [0,32). Needs bits up to x + C - 1.
  // Work on an explicit window of words starting at the word holding p0 + c0.
  int64_t b0 = p0 + c0;
  int64_t w0 = b0 >> 5;
  uint32_t sh = (uint32_t)(b0 & 31);
  uint32_t out = 0;
  int nwords = (int)((sh + 31 + C + 31) >> 5);  // words covering [b0, b0 + 32 + C - 1)
  uint32_t prev = vword(w0);
  // out[i] = OR_{j<C} bit(b0 + i + j): walk the words once, OR-ing the C shifted views
  // that fall into each word pair. Equivalent to a dense loop over j but touches
  // each word pair once.
  int j = 0;
  for (int k = 0; k < nwords && j < C; ++k) {
    uint32_t next = vword(w0 + k + 1);
    // shifts s = sh + j - 32k must lie in [0, 32)
    for (; j < C; ++j) {
      int s = (int)sh + j - 32 * k;
      if (s >= 32) break;
      out |= funnel_r(prev, next, (uint32_t)s);
    }
    prev = next;
  }
  return out;
}

// haplotype.py:90-104,138-159 posmap, run-length encoded (hawkscan.h):
// evaluate posmap(i) for one haplotype's segment range [s0, s1).
HAWK_HD int32_t posmap_eval(const int32_t* seg_rel, const int32_t* seg_gen, const uint8_t* seg_step,
                            int64_t s0, int64_t s1, int32_t i) {
  // last segment with seg_rel <= i
  int64_t lo = s0, hi = s1;  // invariant: seg_rel[lo] <= i (seg_rel[s0] == 0)
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (seg_rel[mid] <= i) lo = mid; else hi = mid;
  }
  return seg_gen[lo] + (seg_step[lo] ? (i - seg_rel[lo]) : 0);
}

// extract nibble of slot `pos` of a haplotype whose first chunk is chunk0
HAWK_HD uint32_t nibble_at(const Planes* q, int64_t chunk0, int64_t pos) {
  const Planes& p = q[chunk0 + (pos >> 5)];
  uint32_t b = (uint32_t)(pos & 31);
  return ((p.a >> b) & 1u) | (((p.c >> b) & 1u) << 1) | (((p.g >> b) & 1u) << 2) |
         (((p.t >> b) & 1u) << 3);
}
HAWK_HD uint32_t lower_at(const uint32_t* v, int64_t chunk0, int64_t pos) {
  return (v[chunk0 + (pos >> 5)] >> (pos & 31)) & 1u;
}

// search_guides.py:356-369: upper-cased core equality == nibble equality over C slots
HAWK_HD bool cores_equal(const Planes* q, int64_t chunk0_a, int64_t pos_a, int64_t chunk0_b,
                         int64_t pos_b, int C) {
  for (int done = 0; done < C; done += 32) {
    int n = C - done < 32 ? C - done : 32;
    uint32_t keep = n == 32 ? 0xFFFFFFFFu : ((1u << n) - 1u);
    uint32_t diff = 0;
    for (int plane = 0; plane < 4; ++plane) {
      auto wa = [&](int64_t w) { return (&q[chunk0_a + w].a)[plane]; };
      auto wb = [&](int64_t w) { return (&q[chunk0_b + w].a)[plane]; };
      diff |= bits32_at(wa, pos_a + done) ^ bits32_at(wb, pos_b + done);
    }
    if (diff & keep) return false;
  }
  return true;
}

}  // namespace hawk

// =============================================================================
// Views over one packed batch (device or host pointers) and the per-chunk scan.
// =============================================================================
#include "../../include/hawkscan.h"

namespace hawk {

struct BatchView {
  const Planes* q;            // planes, one per chunk
  const uint32_t* v;          // case bits, one word per chunk
  const int64_t* slot_off;    // n_hap + 1
  const int32_t* len;         // n_hap
  const int32_t* scan_start;  // n_hap (search_guides.py:49-84)
  const int32_t* scan_stop;   // n_hap
  const uint8_t* is_ref;      // n_hap (samples == "REF")
  int32_t n_hap;
  // coordinate maps (may be null for scan-only use)
  const int64_t* seg_off;
  const int32_t* seg_rel;
  const int32_t* seg_gen;
  const uint8_t* seg_step;
  // variant alleles (unphased only, may be null)
  const int64_t* va_off;
  const int32_t* va_idx;
  const int64_t* va_ent_off;
  const uint8_t* va_ref;
};

struct ScanConst {
  int32_t P, G, C;      // C = G + P (core length)
  int32_t right;        // --right
  int32_t unphased;     // HAWK_F_UNPHASED
  int32_t raw;          // 1: pam_search semantics (no in-range / REF-core filter)
  uint8_t pat[2][HAWK_MAX_PAM];  // [0] forward PAM, [1] reverse complement
  StrandGeom geom[2];   // per strand (right' = right XOR strand)
};

HAWK_HD ScanConst make_scan_const(const hawk_params& p, int raw) {
  ScanConst k;
  k.P = p.pam_len;
  k.G = p.guide_len;
  k.C = p.pam_len + p.guide_len;
  k.right = p.right ? 1 : 0;
  k.unphased = (p.flags & HAWK_F_UNPHASED) ? 1 : 0;
  k.raw = raw;
  for (int i = 0; i < HAWK_MAX_PAM; ++i) {
    k.pat[0][i] = p.pam_fwd[i];
    k.pat[1][i] = p.pam_rc[i];
  }
  for (int s = 0; s < 2; ++s)
    k.geom[s] = strand_geom(k.G, k.P, (k.right != 0) != (s == 1), k.unphased != 0, HAWK_GUIDESEQPAD);
  return k;
}

// per-haplotype scalars the scan needs, loaded once per span
struct HapScan {
  int64_t chunk0;   // first chunk of the haplotype
  int32_t len, a, b;
  int32_t nchunks;  // chunks holding its bases
  int32_t is_ref;
  int64_t lo[2], hi[2];  // admissible position interval per strand, [lo, hi)
};

HAWK_HD HapScan load_hap_scan(const BatchView& B, const ScanConst& K, int32_t h) {
  HapScan H;
  H.chunk0 = B.slot_off[h] >> 5;
  H.len = B.len[h];
  H.a = B.scan_start[h];
  int32_t b = B.scan_stop[h];
  int32_t bmax = H.len - K.P + 1;  // a PAM must fit inside the haplotype
  H.b = b < bmax ? b : bmax;
  if (H.a < 0) H.a = 0;
  H.nchunks = (H.len + 31) >> 5;
  H.is_ref = B.is_ref[h];
  for (int s = 0; s < 2; ++s) {
    int64_t lo = H.a, hi = H.b;
    if (!K.raw) {
      int64_t glo = K.geom[s].lo, ghi = (int64_t)H.len - K.geom[s].hi_sub + 1;
      if (glo > lo) lo = glo;
      if (ghi < hi) hi = ghi;
    }
    H.lo[s] = lo;
    H.hi[s] = hi;
  }
  return H;
}

// The whole per-position work of pam_search + the fused filters for the 32
// positions of chunk `c` (haplotype-relative) : out[s] = surviving hit bits of
// strand s, raw[s] = PAM matches inside the scan interval before the filters.
HAWK_HD void scan_chunk(const BatchView& B, const ScanConst& K, const HapScan& H, int64_t c,
                        uint32_t out[2], uint32_t raw[2]) {
  int64_t p0 = c << 5;
  uint32_t inscan = interval_mask(H.a, H.b, p0);
  out[0] = out[1] = raw[0] = raw[1] = 0;
  if (!inscan) return;
  uint32_t cand[2] = {interval_mask(H.lo[0], H.hi[0], p0), interval_mask(H.lo[1], H.hi[1], p0)};
  auto vword = [&](int64_t w) -> uint32_t {
    return (w < 0 || w >= H.nchunks) ? 0u : B.v[H.chunk0 + w];
  };
  if (!K.raw && !H.is_ref) {
    // search_guides.py:468-471: a non-REF hit survives only if its core holds a
    // variant base; skip the plane loads when no variant bit is in reach.
    for (int s = 0; s < 2; ++s)
      if (cand[s]) cand[s] &= core_variant_mask(vword, p0, K.geom[s].c0, K.C);
    if (!(cand[0] | cand[1])) {
      // raw counts still need the match when requested; they are only reported
      // for REF haplotypes / raw mode, so nothing more to do here.
      return;
    }
  }
  Planes cur = B.q[H.chunk0 + c], nxt = B.q[H.chunk0 + c + 1];
  uint32_t mf = match_chunk(cur, nxt, K.pat[0], K.P);
  uint32_t mr = match_chunk(cur, nxt, K.pat[1], K.P);
  raw[0] = mf & inscan;
  raw[1] = mr & inscan;
  out[0] = mf & cand[0];
  out[1] = mr & cand[1];
}

// ---- per-hit row geometry -------------------------------------------------------
struct RowCoords {
  int32_t pivot;  // first core index (search_guides.py:301)
  int32_t start, stop;  // adjust_guide_position, :260-280
};

HAWK_HD RowCoords row_coords(const BatchView& B, const ScanConst& K, int32_t h, int32_t pos, int s) {
  RowCoords r;
  const StrandGeom& g = K.geom[s];
  r.pivot = pos + g.c0;
  int64_t s0 = B.seg_off[h], s1 = B.seg_off[h + 1];
  r.start = posmap_eval(B.seg_rel, B.seg_gen, B.seg_step, s0, s1, r.pivot);
  r.stop = posmap_eval(B.seg_rel, B.seg_gen, B.seg_step, s0, s1, pos + g.stop_off);
  return r;
}

// Find the REF row (records of haplotype `ref_h`, ascending pos, in
// recs[lo, hi)) whose genomic start equals `start`; returns its index or -1.
// The start of a REF row is monotone in pos, so binary search applies.
HAWK_HD int64_t find_ref_partner(const BatchView& B, const ScanConst& K, const uint64_t* recs,
                                 int64_t lo, int64_t hi, int32_t ref_h, int s, int32_t start) {
  int64_t s0 = B.seg_off[ref_h], s1 = B.seg_off[ref_h + 1];
  int c0 = K.geom[s].c0;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    int32_t pos = (int32_t)(recs[mid] & 0xFFFFFFFFu);
    int32_t st = posmap_eval(B.seg_rel, B.seg_gen, B.seg_step, s0, s1, pos + c0);
    if (st < start) lo = mid + 1; else hi = mid;
  }
  return lo;  // caller checks lo < end and start equality
}

// ---- unphased resolution (search_guides.py:175-257) -------------------------------
// Column descriptor of one window position.
struct Column {
  uint32_t nib;      // IUPAC nibble of the haplotype base
  uint32_t lower;    // haplotype base is lower-case
  uint32_t allowed;  // bases (one-hot subset of nib) allowed: nib & PAM pattern on PAM columns
  int64_t e0;        // first allele entry (ambiguous columns)
  int32_t m;         // number of allele entries (0 => column is a single concrete base)
};

// returns false when an ambiguity code has no variant_alleles entry (KeyError)
HAWK_HD bool load_column(const BatchView& B, int32_t h, int64_t chunk0, int32_t idx,
                         uint32_t pam_code /* 0 = not a PAM column */, Column& c) {
  c.nib = nibble_at(B.q, chunk0, idx);
  c.lower = lower_at(B.v, chunk0, idx);
  c.m = 0;
  c.e0 = 0;
  c.allowed = pam_code ? (c.nib & pam_code) : c.nib;
  if (popc32(c.nib) <= 1) return true;
  // binary search the haplotype's allele sites for idx
  int64_t lo = B.va_off[h], hi = B.va_off[h + 1];
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (B.va_idx[mid] < idx) lo = mid + 1; else hi = mid;
  }
  if (lo >= B.va_off[h + 1] || B.va_idx[lo] != idx) return false;
  c.e0 = B.va_ent_off[lo];
  c.m = (int32_t)(B.va_ent_off[lo + 1] - c.e0);
  return true;
}

// number of candidate characters of a column that survive the PAM filter
HAWK_HD uint32_t column_count(const Column& c) {
  if (c.m == 0) return c.allowed ? 1u : 0u;
  return (uint32_t)popc32(c.allowed) * (uint32_t)c.m;
}

// t-th surviving candidate, in the reference's order: bases ascending A,C,G,T
// (utils.py:82-98), then allele entries (search_guides.py:207-213); upper-case
// iff the base equals the entry's REF allele.
HAWK_HD char column_char(const BatchView& B, const Column& c, uint32_t t) {
  if (c.m == 0) {
    char ch = nibble_letter(c.nib);
    return c.lower ? (char)(ch + 32) : ch;
  }
  uint32_t bi = t / (uint32_t)c.m, e = t % (uint32_t)c.m;
  uint32_t base = 0, allowed = c.allowed;
  for (uint32_t k = 0; k <= bi; ++k) {  // bi-th set bit of `allowed`
    base = allowed & (~allowed + 1u);
    allowed &= allowed - 1u;
  }
  char ch = nibble_letter(base);
  return (B.va_ref[c.e0 + e] == base) ? ch : (char)(ch + 32);
}

}  // namespace hawk
