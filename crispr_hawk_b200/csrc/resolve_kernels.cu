// resolve_kernels.cu -- the unphased guide-table pipeline (variants_present and not phased):
// is_pamhit_valid + resolve_guide + _decode_iupac + _valid_guide (search_guides.py:163-257,
// :372-392), genomic coordinates (:260-280), remove_redundant_guides (:340-369) and the
// emission-order merge (:530-547), in two passes over the HIT stream -- the resolved strings
// are written exactly once, straight into their final table rows:
//
//   resolve_count   per hit: start / stop through the run-length posmap; the ambiguous columns
//                   of its window are read off a bit mask ("IUPAC code with more than one base"
//                   = two or more planes set), so the work is proportional to the ambiguity
//                   codes in the window, not to its length; number of strings = product of the
//                   per-column candidate counts (PAM columns pre-filtered by the pattern, which
//                   gives the same set and order as itertools.product + _valid_guide); REF
//                   partner by direct lookup (REF coordinates are linear), and in closed form
//                   the number of strings whose upper-cased core equals the REF guide's (those
//                   remove_redundant_guides drops) -> kept rows per hit, per-block sums
//   blk_prefix64    exclusive prefix of the block sums (one CTA per strand)
//   hap_offsets     kept rows of each strand stream that precede a haplotype
//   resolve_write   per hit: enumerates its strings in product order (last column fastest),
//                   skips the dropped ones, writes the table rows (window text rebuilt from the
//                   planes once per hit, the ambiguous columns patched per string) and marks
//                   the first-seen table of the (start, strand) key
// The older one-thread-per-string kernels of post_kernels.cu remain for guide geometries
// outside the scan's fast form (windows longer than 80 characters).
#include <cuda_runtime.h>
#include <stdint.h>

#include "hawk_core.h"
#include "hawk_kernels.h"
#include "hawk_post.h"

namespace hawk {

constexpr int RES_T = 128;   // hits per block (count and write must agree)

// window bits [w0, w0 + 96) of a per-chunk bit plane accessor, as three words
template <class F>
__device__ __forceinline__ void window96(F&& word, int64_t chunk0, int32_t w0, uint32_t out[3]) {
  const int64_t c = chunk0 + (w0 >> 5);  // arithmetic shift: the zero gap covers w0 < 0
  const uint32_t sh = (uint32_t)(w0 & 31);
  const uint32_t x0 = word(c), x1 = word(c + 1), x2 = word(c + 2), x3 = word(c + 3);
  out[0] = funnel_r(x0, x1, sh);
  out[1] = funnel_r(x1, x2, sh);
  out[2] = funnel_r(x2, x3, sh);
}

__device__ __forceinline__ uint32_t ambiguous_bits(const Planes& p) {
  // two or more of the four planes set
  return (p.a & p.c) | (p.g & p.t) | ((p.a | p.c) & (p.g | p.t));
}

struct ResStrand {
  const uint64_t* recs;
  int64_t n;
  const uint32_t* ref_bm;
  int32_t* start;
  int32_t* stop;
  uint32_t* cnt;      // kept rows per hit
  int32_t* rpivot;    // two words per hit: [REF partner's core start (REF-relative) when some string can equal
                      // it, else -1 | the variant_alleles site of the first ambiguous column relative to the
                      // haplotype's first site, else -1 -- so that the second pass need not search for it again]
  uint64_t* blk_sum;
};

struct ResCountArgs {
  BatchView B;
  ScanConst K;
  ResStrand S[2];
  int32_t ref_h, ref_g0, ref_len;
  int* err;
};

// Column `j` of the window starting at haplotype index w0 (load_column of hawk_core.h with a
// cursor): the ambiguous columns of one window are visited in ascending order and every one of
// them is a site of the haplotype's variant_alleles table, so the table is binary-searched once
// per hit (`site` < 0) and walked forward afterwards. Returns false on a missing entry.
__device__ __forceinline__ bool decode_amb_column(const BatchView& B, const ScanConst& K, int32_t h, int64_t chunk0,
                                                  int32_t w0, int j, int s, int W, Column& c, uint32_t& cnt,
                                                  int64_t& site) {
  const bool rp = K.geom[s].c0 == 0;
  const int k0 = rp ? HAWK_GUIDESEQPAD : W - HAWK_GUIDESEQPAD - K.P;  // search_guides.py:252
  const uint32_t pam_code = (j >= k0 && j < k0 + K.P) ? K.pat[s][j - k0] : 0u;
  const int32_t idx = w0 + j;
  c.nib = nibble_at(B.q, chunk0, idx);
  c.lower = lower_at(B.v, chunk0, idx);
  c.allowed = pam_code ? (c.nib & pam_code) : c.nib;
  const int64_t end = B.va_off[h + 1];
  if (site < 0) {
    int64_t lo = B.va_off[h], hi = end;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (B.va_idx[mid] < idx) lo = mid + 1; else hi = mid;
    }
    site = lo;
  } else {
    while (site < end && B.va_idx[site] < idx) ++site;
  }
  if (site >= end || B.va_idx[site] != idx) return false;
  c.e0 = B.va_ent_off[site];
  c.m = (int32_t)(B.va_ent_off[site + 1] - c.e0);
  cnt = (uint32_t)popc32(c.allowed) * (uint32_t)c.m;
  return true;
}

// the same without a cursor (columns visited in any order)
__device__ __forceinline__ bool decode_amb_column(const BatchView& B, const ScanConst& K, int32_t h, int64_t chunk0,
                                                  int32_t w0, int j, int s, int W, Column& c, uint32_t& cnt) {
  int64_t site = -1;
  return decode_amb_column(B, K, h, chunk0, w0, j, s, W, c, cnt, site);
}

// Candidates of an ambiguous column as packed characters (<= 8), in the reference's order -- bases
// ascending A, C, G, T, then the site's allele entries; upper-case iff the base equals the entry's
// REF allele (search_guides.py:207-213) -- and the set of candidates whose base is `refnib`.
__device__ __forceinline__ void column_candidates(const BatchView& B, const Column& c, uint32_t refnib, uint64_t& chars,
                                                  uint32_t& eq) {
  chars = 0;
  eq = 0;
  uint32_t d = 0, allowed = c.allowed;
  while (allowed) {
    const uint32_t base = allowed & (~allowed + 1u);
    allowed &= allowed - 1u;
    const uint32_t up = (uint32_t)(uint8_t)nibble_letter(base);
    for (int32_t e = 0; e < c.m; ++e, ++d) {
      const uint32_t ch = (B.va_ref[c.e0 + e] == base) ? up : up + 32u;
      chars |= (uint64_t)ch << (8 * d);
      if (base == refnib) eq |= 1u << d;
    }
  }
}

__global__ void __launch_bounds__(RES_T) resolve_count_kernel(const __grid_constant__ ResCountArgs A) {
  __shared__ uint64_t red[RES_T / 32];
  const int s = blockIdx.x & 1;
  const int64_t blk = blockIdx.x >> 1;
  const ResStrand S = A.S[s];
  if (blk * RES_T >= S.n) return;
  const int64_t i = blk * RES_T + threadIdx.x;
  const BatchView& B = A.B;
  const ScanConst& K = A.K;
  const int W = K.C + 2 * HAWK_GUIDESEQPAD;
  uint64_t kept = 0;
  if (i < S.n) {
    const uint64_t rec = S.recs[i];
    const int32_t h = (int32_t)(rec >> 32), pos = (int32_t)(rec & 0xFFFFFFFFu);
    const RowCoords rc = row_coords(B, K, h, pos, s);
    S.start[i] = rc.start;
    S.stop[i] = rc.stop;
    const int64_t chunk0 = B.slot_off[h] >> 5;
    const int32_t w0 = pos + K.geom[s].w0;
    uint32_t amb[3];
    window96([&](int64_t c) { return ambiguous_bits(B.q[c]); }, chunk0, w0, amb);
    if (W < 96) amb[W >> 5] &= (1u << (W & 31)) - 1u;
    if (W <= 64) amb[2] = 0;
    if (W <= 32) amb[1] = 0;
    // REF partner (remove_redundant_guides keys on (start, strand); REF coordinates are linear)
    int32_t rpivot = -1;
    const bool is_ref = B.is_ref[h] != 0;
    if (!is_ref && A.ref_h >= 0) {
      const int32_t rp = rc.start - A.ref_g0, rpos = rp - K.geom[s].c0;
      if (rp >= 0 && rpos >= 0 && rpos < A.ref_len && ((S.ref_bm[rpos >> 5] >> (rpos & 31)) & 1u)) rpivot = rp;
    }
    // non-ambiguous core columns must equal the REF core for any string to be redundant
    uint64_t total = 1, equal = 0;
    const int64_t rchunk0 = A.ref_h >= 0 ? (B.slot_off[A.ref_h] >> 5) : 0;
    if (rpivot >= 0) {
      equal = 1;
      // compare the cores nibble by nibble where the haplotype is unambiguous
      for (int done = 0; done < K.C; done += 32) {
        const int n = K.C - done < 32 ? K.C - done : 32;
        const uint32_t keep = n == 32 ? 0xFFFFFFFFu : ((1u << n) - 1u);
        const int64_t ba = (int64_t)rc.pivot + done, bb = (int64_t)rpivot + done;
        const Planes a0 = B.q[chunk0 + (ba >> 5)], a1 = B.q[chunk0 + (ba >> 5) + 1];
        const Planes b0 = B.q[rchunk0 + (bb >> 5)], b1 = B.q[rchunk0 + (bb >> 5) + 1];
        const uint32_t sa = (uint32_t)(ba & 31), sb = (uint32_t)(bb & 31);
        const uint32_t diff = (funnel_r(a0.a, a1.a, sa) ^ funnel_r(b0.a, b1.a, sb)) | (funnel_r(a0.c, a1.c, sa) ^ funnel_r(b0.c, b1.c, sb)) |
                              (funnel_r(a0.g, a1.g, sa) ^ funnel_r(b0.g, b1.g, sb)) | (funnel_r(a0.t, a1.t, sa) ^ funnel_r(b0.t, b1.t, sb));
        // ambiguity bits of the same core positions: window bit = PAD + done + x
        const int wb = HAWK_GUIDESEQPAD + done;
        const uint32_t ab = funnel_r(amb[wb >> 5], (wb >> 5) + 1 < 3 ? amb[(wb >> 5) + 1] : 0u, (uint32_t)(wb & 31));
        if (diff & ~ab & keep) equal = 0;
      }
    }
    int64_t site = -1, first_site = -1;
    for (int w = 0; w < 3 && total; ++w) {
      uint32_t bits = amb[w];
      while (bits) {
        const int j = 32 * w + __ffs(bits) - 1;
        bits &= bits - 1;
        Column c;
        uint32_t cnt;
        if (!decode_amb_column(B, K, h, chunk0, w0, j, s, W, c, cnt, site)) {
          atomicExch(A.err, HAWK_EALLELES);
          total = 0;
          break;
        }
        if (first_site < 0) first_site = site;
        total *= cnt;
        if (total > HAWK_MAX_EXPANSION) {
          atomicExch(A.err, HAWK_ECAPACITY);
          total = 0;
          break;
        }
        if (equal) {
          const int cj = j - HAWK_GUIDESEQPAD;  // core column?
          if (cj >= 0 && cj < K.C) {
            const uint32_t refnib = nibble_at(B.q, rchunk0, (int64_t)rpivot + cj);
            equal *= (c.allowed & refnib) ? (uint64_t)c.m : 0ull;  // the candidates of that base: one per allele entry
          } else {
            equal *= cnt;
          }
        }
      }
    }
    if (total == 0) equal = 0;
    kept = total - equal;
    S.cnt[i] = (uint32_t)kept;
    S.rpivot[2 * i] = equal ? rpivot : -1;
    S.rpivot[2 * i + 1] = first_site < 0 ? -1 : (int32_t)(first_site - B.va_off[h]);
  }
  uint64_t x = kept;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t t = 0;
    for (int k = 0; k < RES_T / 32; ++k) t += red[k];
    S.blk_sum[blk] = t;
  }
}

__global__ void __launch_bounds__(1024) blk_prefix64_kernel(const uint64_t* __restrict__ in0, const uint64_t* __restrict__ in1,
                                                            int64_t n0, int64_t n1, uint64_t* __restrict__ out0,
                                                            uint64_t* __restrict__ out1, uint64_t* totals) {
  __shared__ uint64_t part[1024];
  const int s = blockIdx.x;
  const uint64_t* in = s ? in1 : in0;
  uint64_t* out = s ? out1 : out0;
  const int64_t n = s ? n1 : n0;
  const int tid = threadIdx.x;
  const int64_t per = (n + 1023) / 1024;
  const int64_t lo = (int64_t)tid * per, hi = lo + per < n ? lo + per : n;
  uint64_t sum = 0;
  for (int64_t j = lo; j < hi; ++j) sum += in[j];
  part[tid] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    uint64_t y = tid >= o ? part[tid - o] : 0;
    __syncthreads();
    part[tid] += y;
    __syncthreads();
  }
  uint64_t run = part[tid] - sum;
  for (int64_t j = lo; j < hi; ++j) {
    const uint64_t v = in[j];
    out[j] = run;
    run += v;
  }
  if (tid == 1023) totals[s] = part[1023];
}

// kb[s][h] = kept rows of stream s whose haplotype is < h  (h = 0 .. n_hap)
__global__ void hap_offsets_cnt_kernel(const uint64_t* __restrict__ recs0, const uint64_t* __restrict__ recs1, int64_t n0,
                                       int64_t n1, const uint32_t* __restrict__ cnt0, const uint32_t* __restrict__ cnt1,
                                       const uint64_t* __restrict__ base0, const uint64_t* __restrict__ base1,
                                       const uint64_t* __restrict__ totals, int32_t n_hap, uint64_t* __restrict__ kb) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * (int64_t)(n_hap + 1)) return;
  const int s = (int)(t / (n_hap + 1));
  const int32_t h = (int32_t)(t % (n_hap + 1));
  const uint64_t* recs = s ? recs1 : recs0;
  const int64_t n = s ? n1 : n0;
  const uint32_t* cnt = s ? cnt1 : cnt0;
  const uint64_t* base = s ? base1 : base0;
  const uint64_t key = (uint64_t)(uint32_t)h << 32;
  int64_t lo = 0, hi = n;  // first record with haplotype >= h
  while (lo < hi) {
    const int64_t m = (lo + hi) >> 1;
    if (recs[m] < key) lo = m + 1; else hi = m;
  }
  uint64_t r;
  if (lo >= n) {
    r = totals[s];
  } else {
    const int64_t b0 = lo / RES_T * RES_T;
    r = base[lo / RES_T];
    for (int64_t j = b0; j < lo; ++j) r += cnt[j];
  }
  kb[(size_t)s * (n_hap + 1) + h] = r;
}

struct ResWriteStrand {
  const uint64_t* recs;
  int64_t n;
  const int32_t* start;
  const int32_t* stop;
  const uint32_t* cnt;
  const int32_t* rpivot;
  const uint64_t* blk_base;
  const uint64_t* kb_other;
};

struct ResWriteArgs {
  BatchView B;
  ScanConst K;
  ResWriteStrand S[2];
  int32_t ref_h;
  int32_t text_stride;
  int32_t* o_hap;
  uint8_t* o_strand;
  int32_t* o_pos;
  int32_t* o_start;
  int32_t* o_stop;
  uint8_t* o_text;
  uint32_t* key_table;
  int32_t key_min;
};

// Per-hit descriptor of the row-balanced writer, in shared memory: up to RES_DCOLS ambiguous
// columns, each with <= 8 candidate characters.
constexpr int RES_DCOLS = 8;
struct HitDesc {
  uint64_t chars[RES_DCOLS];  // candidate characters of column a, 8 bits each
  uint8_t j[RES_DCOLS];       // window column
  uint8_t cnt[RES_DCOLS];     // candidates
  uint8_t eq[RES_DCOLS];      // bit d: candidate d equals the REF base there (pad columns: all candidates)
  uint32_t inv[RES_DCOLS];    // ceil(2^32 / cnt): t / cnt = umulhi(t, inv) for t < 2^24, cnt <= 8
  uint32_t total;             // strings of the hit (0: the hit takes the one-thread path below)
  uint32_t drops;             // 1: a REF partner exists and some string equals it
  int32_t n_amb;
  int32_t h, pos, start, stop;
  uint64_t f;                 // first table row of the hit
};

// Row-balanced writer: a warp takes 32 consecutive hits; every lane describes its own hit in
// shared memory (window text from the planes, candidate characters of the ambiguous columns),
// then the warp's strings -- product index r over all its hits -- are dealt to the lanes round
// robin: owner hit by a 5-step search in the warp's prefix of string counts, digits from the
// mixed-radix index (last column fastest), and the row's place from the number of redundant
// strings before it, which is a rank in the product of the per-column "equals REF" digit sets.
// Hits with more than RES_DCOLS ambiguous columns or a site with > 8 candidates (rare) are
// written by their own lane afterwards, one string at a time.
template <int N16>  // text_stride / 16: 3..5
__global__ void __launch_bounds__(RES_T) resolve_write_kernel(const __grid_constant__ ResWriteArgs A) {
  constexpr int NW = RES_T / 32;
  __shared__ uint64_t wsum[NW];
  __shared__ HitDesc desc[NW][32];
  __shared__ uint4 base_txt[NW][32][N16];
  __shared__ uint32_t tot_excl[NW][33];
  const int s = blockIdx.x & 1;
  const int64_t blk = blockIdx.x >> 1;
  const ResWriteStrand S = A.S[s];
  if (blk * RES_T >= S.n) return;
  const int64_t i = blk * RES_T + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const BatchView& B = A.B;
  const ScanConst& K = A.K;
  const int W = K.C + 2 * HAWK_GUIDESEQPAD;
  const uint32_t mine = i < S.n ? S.cnt[i] : 0u;
  // exclusive scan of the kept counts over the block
  uint64_t incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint64_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= d) incl += y;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  uint64_t wbase = 0;
  for (int k = 0; k < warp; ++k) wbase += wsum[k];

  HitDesc& D = desc[warp][lane];
  D.total = 0;
  D.n_amb = 0;
  D.drops = 0;
  bool slow = false;
  uint32_t amb[3] = {0u, 0u, 0u};
  int64_t chunk0 = 0, rchunk0 = 0;
  int32_t w0 = 0, rpivot = -1;
  if (mine != 0) {
    const uint64_t rec = S.recs[i];
    const int32_t h = (int32_t)(rec >> 32), pos = (int32_t)(rec & 0xFFFFFFFFu);
    D.h = h;
    D.pos = pos;
    D.start = S.start[i];
    D.stop = S.stop[i];
    D.f = S.blk_base[blk] + wbase + (incl - mine) + S.kb_other[h + (s == 1 ? 1 : 0)];
    if (A.key_table) atomicMin(&A.key_table[((uint32_t)(D.start - A.key_min) << 1) | (uint32_t)s], (uint32_t)D.f);
    chunk0 = B.slot_off[h] >> 5;
    w0 = pos + K.geom[s].w0;
    // window text of the haplotype itself (codes as they stand)
    const int64_t c = chunk0 + (w0 >> 5);
    const uint32_t sh = (uint32_t)(w0 & 31);
    const uint4 q0 = *reinterpret_cast<const uint4*>(&B.q[c]), q1 = *reinterpret_cast<const uint4*>(&B.q[c + 1]),
                q2 = *reinterpret_cast<const uint4*>(&B.q[c + 2]), q3 = *reinterpret_cast<const uint4*>(&B.q[c + 3]);
    const uint32_t v0 = B.v[c], v1 = B.v[c + 1], v2 = B.v[c + 2], v3 = B.v[c + 3];
    uint32_t pa[3], pc[3], pg[3], pt[3], pv[3];
    pa[0] = funnel_r(q0.x, q1.x, sh), pa[1] = funnel_r(q1.x, q2.x, sh), pa[2] = funnel_r(q2.x, q3.x, sh);
    pc[0] = funnel_r(q0.y, q1.y, sh), pc[1] = funnel_r(q1.y, q2.y, sh), pc[2] = funnel_r(q2.y, q3.y, sh);
    pg[0] = funnel_r(q0.z, q1.z, sh), pg[1] = funnel_r(q1.z, q2.z, sh), pg[2] = funnel_r(q2.z, q3.z, sh);
    pt[0] = funnel_r(q0.w, q1.w, sh), pt[1] = funnel_r(q1.w, q2.w, sh), pt[2] = funnel_r(q2.w, q3.w, sh);
    pv[0] = funnel_r(v0, v1, sh), pv[1] = funnel_r(v1, v2, sh), pv[2] = funnel_r(v2, v3, sh);
#pragma unroll
    for (int w = 0; w < 3; ++w) amb[w] = (pa[w] & pc[w]) | (pg[w] & pt[w]) | ((pa[w] | pc[w]) & (pg[w] | pt[w]));
    if (W < 96) amb[W >> 5] &= (1u << (W & 31)) - 1u;
    if (W <= 64) amb[2] = 0;
    if (W <= 32) amb[1] = 0;
#pragma unroll
    for (int g = 0; 2 * g < N16; ++g) {  // 32 window columns at a time (planes_to_chars32, hawk_core.h)
      const int left = W - 32 * g;
      uint32_t wd[8];
      planes_to_chars32(pa[g], pc[g], pg[g], pt[g], pv[g], left >= 32 ? 0xFFFFFFFFu : (left > 0 ? (1u << left) - 1u : 0u), wd);
      base_txt[warp][lane][2 * g] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
      if (2 * g + 1 < N16) base_txt[warp][lane][2 * g + 1] = make_uint4(wd[4], wd[5], wd[6], wd[7]);
    }
    // ambiguous columns, first column first (= most significant digit of the product index)
    rpivot = S.rpivot[2 * i];
    rchunk0 = A.ref_h >= 0 ? (B.slot_off[A.ref_h] >> 5) : 0;
    uint64_t total = 1;
    int n_amb = 0;
    const int32_t site_rel = S.rpivot[2 * i + 1];  // found by resolve_count: the walk starts there
    int64_t site = site_rel < 0 ? -1 : B.va_off[h] + site_rel;
    for (int w = 0; w < 3; ++w) {
      uint32_t bits = amb[w];
      while (bits) {
        const int j = 32 * w + __ffs(bits) - 1;
        bits &= bits - 1;
        Column c2;
        uint32_t cnt;
        decode_amb_column(B, K, h, chunk0, w0, j, s, W, c2, cnt, site);
        total *= cnt;
        if (n_amb < RES_DCOLS && cnt <= 8) {
          const int cj = j - HAWK_GUIDESEQPAD;
          const bool core = rpivot >= 0 && cj >= 0 && cj < K.C;
          uint64_t chars;
          uint32_t eq;
          column_candidates(B, c2, core ? nibble_at(B.q, rchunk0, (int64_t)rpivot + cj) : 0u, chars, eq);
          if (!core) eq = (1u << cnt) - 1u;  // outside the core every candidate "equals"
          D.chars[n_amb] = chars;
          D.j[n_amb] = (uint8_t)j;
          D.cnt[n_amb] = (uint8_t)cnt;
          D.eq[n_amb] = (uint8_t)eq;
          D.inv[n_amb] = (uint32_t)((0x100000000ull + cnt - 1) / cnt);
        } else {
          slow = true;
        }
        ++n_amb;
      }
    }
    D.n_amb = n_amb;
    D.drops = rpivot >= 0 ? 1u : 0u;
    D.total = slow ? 0u : (uint32_t)total;
  }
  // the warp's strings, dealt to the lanes
  uint32_t tincl = D.total;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, tincl, d);
    if (lane >= d) tincl += y;
  }
  tot_excl[warp][lane] = tincl - D.total;
  if (lane == 31) tot_excl[warp][32] = tincl;
  __syncwarp();
  const uint32_t T = tot_excl[warp][32];
  for (uint32_t r = lane; r < T; r += 32) {
    // owner: last hit with tot_excl <= r
    int lo = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1)
      if (tot_excl[warp][lo + step] <= r) lo += step;
    const HitDesc& H = desc[warp][lo];
    uint32_t t = r - tot_excl[warp][lo];
    // digits (last column fastest) and, when a REF partner exists, the redundancy rank
    uint32_t digit[RES_DCOLS];
#pragma unroll
    for (int a = RES_DCOLS - 1; a >= 0; --a) {
      if (a < H.n_amb) {
        const uint32_t c = H.cnt[a], q = c == 1 ? t : __umulhi(t, H.inv[a]);  // 2^32 / 1 does not fit
        digit[a] = t - q * c;
        t = q;
      } else {
        digit[a] = 0;
      }
    }
    uint32_t dropped_before = 0;
    bool drop = false;
    if (H.drops) {
      uint32_t suffix[RES_DCOLS + 1];  // product of |eq| over the columns behind a
      suffix[RES_DCOLS] = 1;
#pragma unroll
      for (int a = RES_DCOLS - 1; a >= 0; --a) suffix[a] = suffix[a + 1] * (a < H.n_amb ? (uint32_t)__popc(H.eq[a]) : 1u);
      drop = true;
#pragma unroll
      for (int a = 0; a < RES_DCOLS; ++a) {
        if (a < H.n_amb && drop) {
          const uint32_t e = H.eq[a];
          dropped_before += (uint32_t)__popc(e & ((1u << digit[a]) - 1u)) * suffix[a + 1];
          if (!((e >> digit[a]) & 1u)) drop = false;
        }
      }
    }
    if (drop) continue;  // redundant with the REF guide (search_guides.py:356-369)
    const uint64_t row = H.f + (r - tot_excl[warp][lo]) - dropped_before;
    A.o_hap[row] = H.h;
    A.o_strand[row] = (uint8_t)s;
    A.o_pos[row] = H.pos;
    A.o_start[row] = H.start;
    A.o_stop[row] = H.stop;
    uint8_t* const dst8 = A.o_text + row * (uint64_t)A.text_stride;
    uint4* const dst = reinterpret_cast<uint4*>(dst8);
#pragma unroll
    for (int p = 0; p < N16; ++p) dst[p] = base_txt[warp][lo][p];
#pragma unroll
    for (int a = 0; a < RES_DCOLS; ++a)
      if (a < H.n_amb) dst8[H.j[a]] = (uint8_t)(H.chars[a] >> (8 * digit[a]));  // same thread, program order
  }
  if (!slow || mine == 0) return;
  // rare shapes (more than RES_DCOLS ambiguous columns, or a site with more than 8 candidates):
  // one string at a time by the hit's own lane, columns re-derived per string
  {
    const int32_t h = D.h, pos = D.pos, st = D.start, sp = D.stop;
    uint64_t total = 1;
    for (int w = 0; w < 3; ++w) {
      uint32_t bits = amb[w];
      while (bits) {
        const int j = 32 * w + __ffs(bits) - 1;
        bits &= bits - 1;
        Column c2;
        uint32_t cnt;
        decode_amb_column(B, K, h, chunk0, w0, j, s, W, c2, cnt);
        total *= cnt;
      }
    }
    uint8_t* const text0 = A.o_text + D.f * (uint64_t)A.text_stride;
    uint64_t k = 0;
    for (uint64_t t = 0; t < total; ++t) {
      // two passes over the columns: decide first, write only a string that stays (a redundant
      // one must not touch the slot behind the hit's last row: it belongs to the next hit)
      bool same = rpivot >= 0;
      uint8_t* dst = text0 + k * (uint64_t)A.text_stride;
      for (int pass = same ? 0 : 1; pass < 2; ++pass) {
        if (pass == 1) {
          if (same) break;
#pragma unroll
          for (int p = 0; p < N16; ++p) reinterpret_cast<uint4*>(dst)[p] = base_txt[warp][lane][p];
        }
        uint64_t rem = t;
        for (int w = 2; w >= 0; --w) {
          uint32_t bits = amb[w];
          while (bits) {
            const int j = 32 * w + 31 - __clz(bits);  // last column first: least significant digit
            bits &= ~(1u << (j & 31));
            Column c2;
            uint32_t cnt;
            decode_amb_column(B, K, h, chunk0, w0, j, s, W, c2, cnt);
            const char ch = column_char(B, c2, (uint32_t)(rem % cnt));
            rem /= cnt;
            if (pass == 1) {
              dst[j] = (uint8_t)ch;
            } else {
              const int cj = j - HAWK_GUIDESEQPAD;
              if (same && cj >= 0 && cj < K.C)
                same = (iupac_entry((uint8_t)ch) & 15u) == nibble_at(B.q, rchunk0, (int64_t)rpivot + cj);
            }
          }
        }
      }
      if (!same) {
        const uint64_t row = D.f + k;
        A.o_hap[row] = h;
        A.o_strand[row] = (uint8_t)s;
        A.o_pos[row] = pos;
        A.o_start[row] = st;
        A.o_stop[row] = sp;
        ++k;
      }
    }
  }
}

int64_t resolve_blocks(int64_t n) { return (n + RES_T - 1) / RES_T; }

int launch_resolve_count(cudaStream_t st, const BatchView& B, const ScanConst& K, const uint64_t* const recs[2],
                         const int64_t n[2], const RefInfo& ref, const uint32_t* const ref_bm[2], int32_t* const start[2],
                         int32_t* const stop[2], uint32_t* const cnt[2], int32_t* const rpivot[2],
                         uint64_t* const blk_sum[2], int* err) {
  const int64_t nb = resolve_blocks(n[0] > n[1] ? n[0] : n[1]);
  if (nb <= 0) return HAWK_OK;
  ResCountArgs A{};
  A.B = B;
  A.K = K;
  for (int s = 0; s < 2; ++s) A.S[s] = ResStrand{recs[s], n[s], ref_bm[s], start[s], stop[s], cnt[s], rpivot[s], blk_sum[s]};
  A.ref_h = ref.h;
  A.ref_g0 = ref.g0;
  A.ref_len = ref.len;
  A.err = err;
  resolve_count_kernel<<<(unsigned)(2 * nb), RES_T, 0, st>>>(A);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "resolve_count_kernel launch");
}

int launch_blk_prefix64(cudaStream_t st, const uint64_t* const sum[2], const int64_t n_blk[2], uint64_t* const base[2],
                        uint64_t* totals) {
  blk_prefix64_kernel<<<2, 1024, 0, st>>>(sum[0], sum[1], n_blk[0], n_blk[1], base[0], base[1], totals);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "blk_prefix64_kernel launch");
}

int launch_hap_offsets_cnt(cudaStream_t st, const uint64_t* const recs[2], const int64_t n[2], const uint32_t* const cnt[2],
                           const uint64_t* const base[2], const uint64_t* totals, int32_t n_hap, uint64_t* kb) {
  const int64_t nt = 2 * (int64_t)(n_hap + 1);
  hap_offsets_cnt_kernel<<<(unsigned)((nt + 127) / 128), 128, 0, st>>>(recs[0], recs[1], n[0], n[1], cnt[0], cnt[1], base[0],
                                                                      base[1], totals, n_hap, kb);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "hap_offsets_cnt_kernel launch");
}

int launch_resolve_write(cudaStream_t st, const BatchView& B, const ScanConst& K, const uint64_t* const recs[2],
                         const int64_t n[2], const int32_t* const start[2], const int32_t* const stop[2],
                         const uint32_t* const cnt[2], const int32_t* const rpivot[2], const uint64_t* const blk_base[2],
                         const uint64_t* const kb_other[2], int32_t ref_h, int32_t text_stride, int32_t* o_hap,
                         uint8_t* o_strand, int32_t* o_pos, int32_t* o_start, int32_t* o_stop, uint8_t* o_text,
                         uint32_t* key_table, int32_t key_min) {
  const int64_t nb = resolve_blocks(n[0] > n[1] ? n[0] : n[1]);
  if (nb <= 0) return HAWK_OK;
  ResWriteArgs A{};
  A.B = B;
  A.K = K;
  for (int s = 0; s < 2; ++s)
    A.S[s] = ResWriteStrand{recs[s], n[s], start[s], stop[s], cnt[s], rpivot[s], blk_base[s], kb_other[s]};
  A.ref_h = ref_h;
  A.text_stride = text_stride;
  A.o_hap = o_hap;
  A.o_strand = o_strand;
  A.o_pos = o_pos;
  A.o_start = o_start;
  A.o_stop = o_stop;
  A.o_text = o_text;
  A.key_table = key_table;
  A.key_min = key_min;
  const unsigned grid = (unsigned)(2 * nb);
  switch (text_stride / 16) {
    case 3: resolve_write_kernel<3><<<grid, RES_T, 0, st>>>(A); break;
    case 4: resolve_write_kernel<4><<<grid, RES_T, 0, st>>>(A); break;
    case 5: resolve_write_kernel<5><<<grid, RES_T, 0, st>>>(A); break;
    default: return hawk_fail(HAWK_EINVAL, "resolve_write: text stride %d outside the fast form", text_stride);
  }
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "resolve_write_kernel launch");
}

}  // namespace hawk
