/*
 * scan_oracle.c -- TEST INFRASTRUCTURE ONLY: plain-C restatement of the
 * reference's guide-discovery scan for sizes the Python oracle cannot reach.
 *
 * Follows /root/reference/src/crisprhawk (v0.2.2) function by function, scalar
 * and position by position like the reference's own loops:
 *   oracle_encode      encoder.py:48-57      char -> IUPAC nibble (upper-cased)
 *   match_at           search_guides.py:32-46   every PAM nibble intersects the base
 *   oracle_search      :87-131 scan both strands over [scan_start, scan_stop),
 *                      :395-420 in-range, :134-160 window, :468-471 REF-core filter,
 *                      :260-280 genomic start/stop through the posmap,
 *                      :306-369 remove_redundant_guides (first-seen bucket order)
 *   resolve_hit        :372-392 is_pamhit_valid, :175-257 _decode_iupac / resolve_guide /
 *                      _valid_guide (unphased: variants_present and not phased, :473-479)
 *
 * Parity status: PINNED -- tests/test_oracle_c.py checks this file against the
 * reference-generated golden vectors (tests/golden) and against the Python
 * oracle. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product never does.
 *
 * Haplotypes arrive in the same flat form the product's C-ABI uses: ASCII slot
 * space + slot_off/len + run-length posmap segments, so both sides can be fed
 * from one generator. Threads: OpenMP over haplotypes (the reference itself is
 * single-threaded; `threads` = 1 reproduces that).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PAD 10 /* guide.py:21 GUIDESEQPAD */

static uint8_t NIB[256];
static int nib_ready = 0;

static void init_nib(void) {
  if (nib_ready) return;
  memset(NIB, 0, sizeof NIB);
  const char *letters = "ACGTRYSWKMBDHVN";
  const uint8_t vals[] = {1, 2, 4, 8, 5, 10, 6, 9, 12, 3, 14, 13, 11, 7, 15};
  for (int i = 0; letters[i]; ++i) {
    NIB[(uint8_t)letters[i]] = vals[i];
    NIB[(uint8_t)(letters[i] + 32)] = vals[i]; /* nt.upper(), encoder.py:52 */
  }
  nib_ready = 1;
}

/* encoder.py:48-57; returns index of the first non-IUPAC character or -1 */
int64_t oracle_encode(const uint8_t *text, int64_t n, uint8_t *bits) {
  init_nib();
  int64_t bad = -1;
  for (int64_t i = 0; i < n; ++i) {
    bits[i] = NIB[text[i]];
    if (!bits[i] && bad < 0) bad = i;
  }
  return bad;
}

static inline int match_at(const uint8_t *pat, int P, const uint8_t *bits, int64_t pos) {
  for (int i = 0; i < P; ++i)
    if (pat[i] && !(pat[i] & bits[pos + i])) return 0;
  return 1;
}

typedef struct {
  int32_t hap, pos, start, stop;
  uint8_t strand;
  int64_t txt; /* unphased: offset of the resolved string in the haplotype's text pool, else -1 */
} row_t;

typedef struct {
  row_t *rows;
  int64_t n, cap;
  uint64_t *hits[2]; /* raw pam_search hits (hap << 32 | pos) */
  int64_t nh[2], caph[2];
  uint8_t *txt; /* resolved strings (unphased), W bytes each */
  int64_t ntxt, captxt;
  int status;
} vec_t;

static void push_row(vec_t *v, row_t r) {
  if (v->n == v->cap) {
    v->cap = v->cap ? v->cap * 2 : 1024;
    v->rows = (row_t *)realloc(v->rows, (size_t)v->cap * sizeof(row_t));
  }
  v->rows[v->n++] = r;
}
static void push_hit(vec_t *v, int s, uint64_t rec) {
  if (v->nh[s] == v->caph[s]) {
    v->caph[s] = v->caph[s] ? v->caph[s] * 2 : 1024;
    v->hits[s] = (uint64_t *)realloc(v->hits[s], (size_t)v->caph[s] * 8);
  }
  v->hits[s][v->nh[s]++] = rec;
}

typedef struct {
  int64_t n_rows;
  int32_t *hap, *pos, *start, *stop;
  uint8_t *strand;
  uint8_t *text;
  int window;
  int64_t n_hits[2];
  uint64_t *hits[2];
  int64_t scanned_bp;
  int status; /* 0 ok; 1 KeyError: ambiguity code without variant_alleles entry (:207-213);
                 2 duplicate REF guide (:328-334); 3 expansion of one hit above 2^24 strings */
} oracle_table;

#define ORACLE_MAX_EXPANSION (1ull << 24)

static uint8_t *push_text(vec_t *v, int W) {
  if (v->ntxt + W > v->captxt) {
    v->captxt = v->captxt ? v->captxt * 2 : 4096;
    while (v->captxt < v->ntxt + W) v->captxt *= 2;
    v->txt = (uint8_t *)realloc(v->txt, (size_t)v->captxt);
  }
  uint8_t *p = v->txt + v->ntxt;
  v->ntxt += W;
  return p;
}

/* variant_alleles of one haplotype (haplotype.py:287-291), flattened like the C-ABI's
 * hawk_batch_set_alleles: sites [va_off[h], va_off[h+1]) sorted by va_idx, site j owns entries
 * [va_ent_off[j], va_ent_off[j+1]); va_ref[e] = nibble of a one-base REF allele, else 0 */
typedef struct {
  const int64_t *va_off;
  const int32_t *va_idx;
  const int64_t *va_ent_off;
  const uint8_t *va_ref;
} alleles_t;

/* _decode_iupac (:175-213) for one window column: candidate characters in the reference's
 * order -- bases ascending A, C, G, T (utils.py:82-98), then the site's allele entries;
 * upper-case iff the base equals the entry's REF allele. `pam_code` != 0 keeps only the bases
 * _valid_guide (:163-169) would accept in this PAM column. Returns the count, -1 on KeyError. */
static int decode_column(uint8_t ch, int32_t idx, int32_t h, const alleles_t *A, uint8_t pam_code,
                         uint8_t *cand /* <= 4 * entries */, int cap) {
  const uint8_t nib = NIB[ch];
  if (__builtin_popcount(nib) <= 1) {
    if (pam_code && !(nib & pam_code)) return 0;
    cand[0] = ch; /* case preserved (:205-206) */
    return 1;
  }
  int64_t lo = A->va_off[h], hi = A->va_off[h + 1];
  while (lo < hi) {
    int64_t m = (lo + hi) / 2;
    if (A->va_idx[m] < idx) lo = m + 1; else hi = m;
  }
  if (lo >= A->va_off[h + 1] || A->va_idx[lo] != idx) return -1;
  const int64_t e0 = A->va_ent_off[lo], e1 = A->va_ent_off[lo + 1];
  static const char base_letter[4] = {'A', 'C', 'G', 'T'};
  int n = 0;
  for (int b = 0; b < 4; ++b) {
    if (!(nib & (1 << b))) continue;
    if (pam_code && !(pam_code & (1 << b))) continue;
    for (int64_t e = e0; e < e1; ++e) {
      if (n >= cap) return -2;
      cand[n++] = (uint8_t)(A->va_ref[e] == (1 << b) ? base_letter[b] : base_letter[b] + 32);
    }
  }
  return n;
}

static int32_t posmap_at(const int32_t *rel, const int32_t *gen, const uint8_t *step, int64_t s0,
                         int64_t s1, int32_t i) {
  int64_t lo = s0, hi = s1;
  while (hi - lo > 1) {
    int64_t m = (lo + hi) / 2;
    if (rel[m] <= i) lo = m; else hi = m;
  }
  return gen[lo] + (step[lo] ? i - rel[lo] : 0);
}

/* bucket bookkeeping for remove_redundant_guides: open-addressing map key -> first row */
typedef struct {
  uint64_t *keys;
  int64_t *first, *refrow;
  uint64_t mask;
} bmap_t;
static uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}
static int64_t bmap_slot(bmap_t *m, uint64_t key) {
  uint64_t s = mix(key) & m->mask;
  while (m->keys[s] != ~0ull && m->keys[s] != key) s = (s + 1) & m->mask;
  return (int64_t)s;
}

/* resolve_guide (:216-257) for one hit: every concrete window string whose PAM slice still
 * matches, in itertools.product order (last column fastest), appended to the haplotype's text
 * pool. Returns the number of strings, -1 on KeyError, -3 above ORACLE_MAX_EXPANSION. */
static int64_t resolve_hit(vec_t *v, const uint8_t *text, int64_t w0, int W, int32_t h, const alleles_t *A,
                           const uint8_t *pattern, int P, int rp) {
  enum { MAXC = 64 };
  uint8_t cand[256][MAXC];
  int cnt[256];
  const int k0 = rp ? PAD : W - PAD - P; /* :252 */
  uint64_t total = 1;
  for (int j = 0; j < W; ++j) {
    const uint8_t code = (j >= k0 && j < k0 + P) ? pattern[j - k0] : 0;
    const int n = decode_column(text[w0 + j], (int32_t)(w0 + j), h, A, code, cand[j], MAXC);
    if (n == -1) return -1;
    if (n == -2) return -3;
    cnt[j] = n;
    total *= (uint64_t)n;
    if (total > ORACLE_MAX_EXPANSION) return -3;
  }
  if (total == 0) return 0;
  int digit[256];
  memset(digit, 0, sizeof(int) * (size_t)W);
  for (uint64_t t = 0; t < total; ++t) {
    uint8_t *dst = push_text(v, W);
    for (int j = 0; j < W; ++j) dst[j] = cand[j][digit[j]];
    for (int j = W - 1; j >= 0; --j) { /* odometer, last column fastest */
      if (++digit[j] < cnt[j]) break;
      digit[j] = 0;
    }
  }
  return (int64_t)total;
}

oracle_table *oracle_search2(const uint8_t *ascii, const int64_t *slot_off, const int32_t *len,
                             const int32_t *scan_start, const int32_t *scan_stop,
                             const uint8_t *is_ref, int32_t n_hap, const int64_t *seg_off,
                             const int32_t *seg_rel, const int32_t *seg_gen, const uint8_t *seg_step,
                             const uint8_t *pam_fwd, const uint8_t *pam_rc, int P, int G, int right,
                             int threads, int raw_only, int unphased, const int64_t *va_off,
                             const int32_t *va_idx, const int64_t *va_ent_off, const uint8_t *va_ref) {
  init_nib();
  const int W = G + P + 2 * PAD;
  if (W > 256) return NULL;
  const alleles_t A = {va_off, va_idx, va_ent_off, va_ref};
  vec_t *per = (vec_t *)calloc((size_t)(n_hap > 0 ? n_hap : 1), sizeof(vec_t));
  int64_t scanned = 0;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : scanned)
  for (int32_t h = 0; h < n_hap; ++h) {
    const int64_t L = len[h];
    const uint8_t *text = ascii + slot_off[h];
    uint8_t *bits = (uint8_t *)malloc((size_t)(L > 0 ? L : 1));
    oracle_encode(text, L, bits); /* crisprhawk.py:70-75 */
    vec_t *v = &per[h];
    int64_t a = scan_start[h], b = scan_stop[h];
    if (b > a) scanned += b - a;
    /* pam_search: forward and reverse pattern at the same forward index (:94-98) */
    for (int64_t pos = a; pos < b; ++pos) {
      if (match_at(pam_fwd, P, bits, pos)) push_hit(v, 0, ((uint64_t)(uint32_t)h << 32) | (uint64_t)pos);
      if (match_at(pam_rc, P, bits, pos)) push_hit(v, 1, ((uint64_t)(uint32_t)h << 32) | (uint64_t)pos);
    }
    if (!raw_only) {
      for (int s = 0; s < 2 && !v->status; ++s) { /* retrieve_guides per strand, :530-547 */
        const int rp = s == 1 ? !right : right; /* :538 */
        const uint8_t *pattern = s == 0 ? pam_fwd : pam_rc; /* :166 */
        for (int64_t k = 0; k < v->nh[s] && !v->status; ++k) {
          int64_t pos = (int64_t)(v->hits[s][k] & 0xFFFFFFFFu);
          int64_t w0 = rp ? pos - PAD : pos - G - PAD;
          int64_t w1 = rp ? pos + G + P + PAD : pos + P + PAD;
          if (w0 < 0 || w1 > L) continue; /* is_pamhit_in_range :414-420 */
          if (!is_ref[h]) { /* :468-471 core.isupper() */
            int has_lower = 0;
            for (int64_t j = w0 + PAD; j < w1 - PAD; ++j)
              if (text[j] >= 'a' && text[j] <= 'z') { has_lower = 1; break; }
            if (!has_lower) continue;
          }
          int64_t n_str = 1, txt0 = -1;
          if (unphased) { /* :473-479 */
            const int valid = rp ? (pos + G + P + PAD < L) : (pos - G - PAD >= 0); /* :372-392 */
            if (!valid) continue;
            txt0 = v->ntxt;
            n_str = resolve_hit(v, text, w0, W, h, &A, pattern, P, rp);
            if (n_str < 0) {
              v->status = n_str == -1 ? 1 : 3;
              break;
            }
          }
          int32_t pivot = (int32_t)(rp ? pos : pos - G);
          row_t r;
          r.hap = h;
          r.pos = (int32_t)pos;
          r.strand = (uint8_t)s;
          r.start = posmap_at(seg_rel, seg_gen, seg_step, seg_off[h], seg_off[h + 1], pivot);
          r.stop = posmap_at(seg_rel, seg_gen, seg_step, seg_off[h], seg_off[h + 1],
                             (int32_t)(rp ? pos + G + P : pos + P));
          for (int64_t t = 0; t < n_str; ++t) {
            r.txt = unphased ? txt0 + t * W : -1;
            push_row(v, r);
          }
        }
      }
    }
    free(bits);
  }
  oracle_table *T = (oracle_table *)calloc(1, sizeof(oracle_table));
  T->window = W;
  T->scanned_bp = scanned;
  for (int32_t h = 0; h < n_hap; ++h)
    if (per[h].status && !T->status) T->status = per[h].status;
  /* raw hits, concatenated in haplotype order */
  for (int s = 0; s < 2; ++s) {
    int64_t n = 0;
    for (int32_t h = 0; h < n_hap; ++h) n += per[h].nh[s];
    T->n_hits[s] = n;
    T->hits[s] = (uint64_t *)malloc((size_t)(n ? n : 1) * 8);
    int64_t o = 0;
    for (int32_t h = 0; h < n_hap; ++h) {
      if (per[h].nh[s]) memcpy(T->hits[s] + o, per[h].hits[s], (size_t)per[h].nh[s] * 8);
      o += per[h].nh[s];
    }
  }
  /* emission order (haplotype, strand, position, product) == per-haplotype row order */
  int64_t n_all = 0;
  if (!T->status)
    for (int32_t h = 0; h < n_hap; ++h) n_all += per[h].n;
  row_t *all = (row_t *)malloc((size_t)(n_all ? n_all : 1) * sizeof(row_t));
  if (!T->status) {
    int64_t o = 0;
    for (int32_t h = 0; h < n_hap; ++h) {
      if (per[h].n) memcpy(all + o, per[h].rows, (size_t)per[h].n * sizeof(row_t));
      o += per[h].n;
    }
  }
#define ROW_TEXT(g) ((g)->txt >= 0 ? per[(g)->hap].txt + (g)->txt \
                                   : ascii + slot_off[(g)->hap] + (((g)->strand == 1 ? !right : right) ? (g)->pos - PAD : (g)->pos - G - PAD))
  /* remove_redundant_guides (:306-369): bucket by (start, strand) in first-seen order */
  bmap_t m;
  uint64_t size = 1024;
  while (size < (uint64_t)n_all * 2) size <<= 1;
  m.mask = size - 1;
  m.keys = (uint64_t *)malloc(size * 8);
  m.first = (int64_t *)malloc(size * 8);
  m.refrow = (int64_t *)malloc(size * 8);
  memset(m.keys, 0xFF, size * 8);
  for (int64_t i = 0; i < n_all; ++i) {
    uint64_t key = ((uint64_t)(uint32_t)all[i].start << 1) | all[i].strand;
    int64_t s = bmap_slot(&m, key);
    if (m.keys[s] == ~0ull) {
      m.keys[s] = key;
      m.first[s] = i;
      m.refrow[s] = -1;
    }
    if (is_ref[all[i].hap]) {
      if (m.refrow[s] < 0) m.refrow[s] = i;
      else if (unphased) T->status = 2; /* :328-334 (the phased callers never feed two REFs) */
    }
  }
  uint8_t *keep = (uint8_t *)malloc((size_t)(n_all ? n_all : 1));
  int64_t *bucket = (int64_t *)malloc((size_t)(n_all ? n_all : 1) * 8);
  int64_t n_keep = 0;
  for (int64_t i = 0; i < n_all; ++i) {
    uint64_t key = ((uint64_t)(uint32_t)all[i].start << 1) | all[i].strand;
    int64_t s = bmap_slot(&m, key);
    bucket[i] = m.first[s];
    keep[i] = 1;
    int64_t rr = m.refrow[s];
    if (rr >= 0 && !is_ref[all[i].hap]) {
      /* upper-cased core equality, :360-367 */
      const uint8_t *gt = ROW_TEXT(&all[i]) + PAD, *rt = ROW_TEXT(&all[rr]) + PAD;
      int same = 1;
      for (int j = 0; j < G + P; ++j)
        if ((gt[j] & 0xDF) != (rt[j] & 0xDF)) { same = 0; break; }
      if (same) keep[i] = 0;
    }
    n_keep += keep[i];
  }
  /* final order: buckets by first appearance, members in emission order == stable sort by bucket.
   * bucket ids are row indices < n_all: counting placement via prefix of bucket sizes */
  int64_t *cnt = (int64_t *)calloc((size_t)(n_all + 1), 8);
  for (int64_t i = 0; i < n_all; ++i)
    if (keep[i]) cnt[bucket[i] + 1]++;
  for (int64_t i = 0; i < n_all; ++i) cnt[i + 1] += cnt[i];
  T->n_rows = n_keep;
  size_t nr = (size_t)(n_keep ? n_keep : 1);
  T->hap = (int32_t *)malloc(nr * 4);
  T->pos = (int32_t *)malloc(nr * 4);
  T->start = (int32_t *)malloc(nr * 4);
  T->stop = (int32_t *)malloc(nr * 4);
  T->strand = (uint8_t *)malloc(nr);
  T->text = (uint8_t *)malloc(nr * (size_t)W);
  for (int64_t i = 0; i < n_all; ++i) {
    if (!keep[i]) continue;
    int64_t o = cnt[bucket[i]]++;
    const row_t *g = &all[i];
    T->hap[o] = g->hap;
    T->pos[o] = g->pos;
    T->start[o] = g->start;
    T->stop[o] = g->stop;
    T->strand[o] = g->strand;
    memcpy(T->text + (size_t)o * W, ROW_TEXT(g), (size_t)W); /* :134-160 / resolved string */
  }
#undef ROW_TEXT
  free(cnt); free(keep); free(bucket); free(m.keys); free(m.first); free(m.refrow); free(all);
  for (int32_t h = 0; h < n_hap; ++h) { free(per[h].rows); free(per[h].hits[0]); free(per[h].hits[1]); free(per[h].txt); }
  free(per);
  return T;
}

/* phased / variant-free search (the original entry point) */
oracle_table *oracle_search(const uint8_t *ascii, const int64_t *slot_off, const int32_t *len,
                            const int32_t *scan_start, const int32_t *scan_stop,
                            const uint8_t *is_ref, int32_t n_hap, const int64_t *seg_off,
                            const int32_t *seg_rel, const int32_t *seg_gen, const uint8_t *seg_step,
                            const uint8_t *pam_fwd, const uint8_t *pam_rc, int P, int G, int right,
                            int threads, int raw_only) {
  return oracle_search2(ascii, slot_off, len, scan_start, scan_stop, is_ref, n_hap, seg_off, seg_rel, seg_gen,
                        seg_step, pam_fwd, pam_rc, P, G, right, threads, raw_only, 0, NULL, NULL, NULL, NULL);
}

void oracle_table_free(oracle_table *T) {
  if (!T) return;
  free(T->hap); free(T->pos); free(T->start); free(T->stop); free(T->strand); free(T->text);
  free(T->hits[0]); free(T->hits[1]);
  free(T);
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
