// post_kernels.cu -- everything downstream of the PAM scan that the reference
// does per hit in retrieve_guides / remove_redundant_guides
// (search_guides.py:340-369, :423-507): genomic coordinates through the
// run-length posmap, REF-core redundancy, unphased IUPAC resolution, window
// extraction, emission-order merge of the two strand streams and first-seen
// bucket ids. All kernels are simple gathers over the (small) hit stream.
#include <cuda_runtime.h>
#include <stdint.h>

#include "hawk_core.h"
#include "hawk_kernels.h"
#include "hawk_post.h"

namespace hawk {

static inline unsigned grid_for(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > (1ll << 30)) b = 1ll << 30;
  return (unsigned)b;
}

// ---------------------------------------------------------------- REF ranges
// recs are sorted by (hap, pos): rows of haplotype ref_h form one range.
// first index with recs[i] >= key, searched by a whole warp: 32 probes per step cut the range
// 33-fold, so 23 dependent loads of a binary search over ~10^7 records become 5
__device__ __forceinline__ int64_t warp_lower_bound(const uint64_t* __restrict__ recs, int64_t n, uint64_t key) {
  const int lane = threadIdx.x & 31;
  int64_t lo = 0, hi = n;  // answer in [lo, hi]
  while (hi - lo > 32) {
    const int64_t w = hi - lo;
    const int64_t p = lo + (w * (lane + 1)) / 33;  // lo < p < hi, ascending over the lanes
    const uint32_t less = __ballot_sync(0xFFFFFFFFu, recs[p] < key);  // 1..1 0..0
    const int c = __popc(less);
    const int64_t p_lo = c > 0 ? lo + (w * c) / 33 : lo - 1, p_hi = c < 32 ? lo + (w * (c + 1)) / 33 : hi;
    lo = p_lo + 1;
    hi = p_hi;
  }
  const uint32_t less = __ballot_sync(0xFFFFFFFFu, lo + lane < hi && recs[lo + lane] < key);
  return lo + __popc(less);
}

__global__ void ref_range_kernel(const uint64_t* recs0, int64_t n0, const uint64_t* recs1,
                                 int64_t n1, int32_t ref_h, int64_t* out) {
  const int w = threadIdx.x >> 5;  // warp = (strand, bound)
  const int s = w >> 1;
  const uint64_t* recs = s ? recs1 : recs0;
  const int64_t n = s ? n1 : n0;
  int64_t r = 0;
  if (ref_h >= 0) r = warp_lower_bound(recs, n, ((uint64_t)(uint32_t)ref_h + (uint64_t)(w & 1)) << 32);
  if ((threadIdx.x & 31) == 0) out[w] = r;
}

// REF partner of a row with genomic `start` on strand s; returns the REF core
// start index (relative to the REF haplotype) or -1.
__device__ __forceinline__ int32_t ref_partner_pivot(const BatchView& B, const ScanConst& K,
                                                     const uint64_t* recs, const int64_t* ref_range,
                                                     int32_t ref_h, int s, int32_t start) {
  int64_t lo = ref_range[2 * s], hi = ref_range[2 * s + 1];
  int64_t i = find_ref_partner(B, K, recs, lo, hi, ref_h, s, start);
  if (i >= hi) return -1;
  int32_t rpos = (int32_t)(recs[i] & 0xFFFFFFFFu);
  int32_t rpivot = rpos + K.geom[s].c0;
  int32_t rstart = posmap_eval(B.seg_rel, B.seg_gen, B.seg_step, B.seg_off[ref_h],
                               B.seg_off[ref_h + 1], rpivot);
  return rstart == start ? rpivot : -1;
}

// ---------------------------------------------------------------- rows (per hit)
__global__ void rows_kernel(BatchView B, ScanConst K, const uint64_t* __restrict__ recs, int64_t n,
                            int s, int32_t ref_h, const int64_t* __restrict__ ref_range, int dedup,
                            int32_t* __restrict__ start, int32_t* __restrict__ stop,
                            uint8_t* __restrict__ keep) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t rec = recs[i];
  int32_t h = (int32_t)(rec >> 32), pos = (int32_t)(rec & 0xFFFFFFFFu);
  RowCoords rc = row_coords(B, K, h, pos, s);
  start[i] = rc.start;
  stop[i] = rc.stop;
  uint8_t k = 1;
  if (dedup && ref_h >= 0 && !B.is_ref[h]) {
    // remove_redundant_guides (:356-369): a non-REF guide whose upper-cased core
    // equals the REF guide's at the same (start, strand) is dropped
    int32_t rpivot = ref_partner_pivot(B, K, recs, ref_range, ref_h, s, rc.start);
    if (rpivot >= 0 &&
        cores_equal(B.q, B.slot_off[h] >> 5, rc.pivot, B.slot_off[ref_h] >> 5, rpivot, K.C))
      k = 0;
  }
  keep[i] = k;
}

// ---------------------------------------------------------------- unphased resolution
__device__ __forceinline__ uint32_t pam_code_of_column(const ScanConst& K, int s, int j, int W) {
  bool rp = K.geom[s].c0 == 0;
  int k0 = rp ? HAWK_GUIDESEQPAD : W - HAWK_GUIDESEQPAD - K.P;  // search_guides.py:252
  return (j >= k0 && j < k0 + K.P) ? K.pat[s][j - k0] : 0u;
}

__global__ void expand_count_kernel(BatchView B, ScanConst K, const uint64_t* __restrict__ recs,
                                    int64_t n, int s, uint64_t* __restrict__ cnt, int* err) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t rec = recs[i];
  int32_t h = (int32_t)(rec >> 32), pos = (int32_t)(rec & 0xFFFFFFFFu);
  const int W = K.C + 2 * HAWK_GUIDESEQPAD;
  int64_t chunk0 = B.slot_off[h] >> 5;
  int32_t w0 = pos + K.geom[s].w0;
  uint64_t prod = 1;
  for (int j = 0; j < W; ++j) {
    Column c;
    if (!load_column(B, h, chunk0, w0 + j, pam_code_of_column(K, s, j, W), c)) {
      atomicExch(err, HAWK_EALLELES);
      prod = 0;
      break;
    }
    prod *= column_count(c);
    if (prod > HAWK_MAX_EXPANSION) {
      atomicExch(err, HAWK_ECAPACITY);
      prod = 0;
      break;
    }
  }
  cnt[i] = prod;
}

// one thread per resolved string: decode the mixed-radix index (last column
// fastest == itertools.product order, search_guides.py:246-251)
__global__ void expand_write_kernel(BatchView B, ScanConst K, const uint64_t* __restrict__ recs,
                                    int64_t n_hits, int s, const uint64_t* __restrict__ off,
                                    int64_t n_rows, int32_t ref_h,
                                    const int64_t* __restrict__ ref_range,
                                    const int32_t* __restrict__ start, uint8_t* __restrict__ text,
                                    int64_t* __restrict__ row_hit, uint8_t* __restrict__ keep) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  // hit owning row r: last i with off[i] <= r
  int64_t lo = 0, hi = n_hits;
  while (hi - lo > 1) {
    int64_t m = (lo + hi) >> 1;
    if (off[m] <= (uint64_t)r) lo = m; else hi = m;
  }
  int64_t i = lo;
  uint64_t t = (uint64_t)r - off[i];
  uint64_t rec = recs[i];
  int32_t h = (int32_t)(rec >> 32), pos = (int32_t)(rec & 0xFFFFFFFFu);
  const int W = K.C + 2 * HAWK_GUIDESEQPAD;
  int64_t chunk0 = B.slot_off[h] >> 5;
  int32_t w0 = pos + K.geom[s].w0;
  uint8_t* dst = text + r * W;
  for (int j = W - 1; j >= 0; --j) {
    Column c;
    load_column(B, h, chunk0, w0 + j, pam_code_of_column(K, s, j, W), c);
    uint32_t cc = column_count(c);
    uint32_t d = (uint32_t)(t % cc);
    t /= cc;
    dst[j] = (uint8_t)column_char(B, c, d);
  }
  row_hit[r] = i;
  uint8_t k = 1;
  if (ref_h >= 0 && !B.is_ref[h]) {
    int32_t rpivot = ref_partner_pivot(B, K, recs, ref_range, ref_h, s, start[i]);
    if (rpivot >= 0) {
      int64_t rchunk0 = B.slot_off[ref_h] >> 5;
      bool same = true;
      for (int j = 0; j < K.C && same; ++j)
        same = (iupac_entry(dst[HAWK_GUIDESEQPAD + j]) & 15u) == nibble_at(B.q, rchunk0, rpivot + j);
      if (same) k = 0;
    }
  }
  keep[r] = k;
}

// ---------------------------------------------------------------- exclusive scan (u64)
constexpr int SCAN_T = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_T * SCAN_ITEMS;

template <class T>
__global__ void tile_sum_kernel(const T* __restrict__ in, int64_t n, uint64_t* __restrict__ sums) {
  __shared__ uint64_t sh[SCAN_T / 32];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  uint64_t x = 0;
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + (int64_t)k * SCAN_T + threadIdx.x;
    if (i < n) x += (uint64_t)in[i];
  }
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t t = 0;
    for (int k = 0; k < SCAN_T / 32; ++k) t += sh[k];
    sums[blockIdx.x] = t;
  }
}

// single CTA: in-place exclusive scan of the tile sums; total -> sums[n_tiles]
__global__ void tile_scan_kernel(uint64_t* sums, int64_t n_tiles, uint64_t* total_out) {
  __shared__ uint64_t sh[1024];
  __shared__ uint64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < n_tiles; base += 1024) {
    int64_t i = base + threadIdx.x;
    uint64_t x = i < n_tiles ? sums[i] : 0;
    sh[threadIdx.x] = x;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      uint64_t y = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += y;
      __syncthreads();
    }
    uint64_t incl = sh[threadIdx.x], c = carry;
    if (i < n_tiles) sums[i] = c + incl - x;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    sums[n_tiles] = carry;
    if (total_out) *total_out = carry;
  }
}

template <class T>
__global__ void tile_apply_kernel(const T* __restrict__ in, int64_t n,
                                  const uint64_t* __restrict__ sums, uint64_t* __restrict__ out) {
  // thread t owns SCAN_ITEMS consecutive items so the order is preserved
  __shared__ uint64_t sh[SCAN_T / 32];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  uint64_t vals[SCAN_ITEMS], mine = 0;
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    vals[k] = (base + k < n) ? (uint64_t)in[base + k] : 0;
    mine += vals[k];
  }
  uint64_t incl = mine;
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    uint64_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) sh[warp] = incl;
  __syncthreads();
  uint64_t wbase = 0;
  for (int k = 0; k < warp; ++k) wbase += sh[k];
  uint64_t run = sums[blockIdx.x] + wbase + incl - mine;
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) out[base + k] = run;
    run += vals[k];
  }
}

template <class T>
static int exclusive_scan_impl(cudaStream_t st, const T* in, int64_t n, uint64_t* out,
                               uint64_t* tile_sums /* n_tiles + 1 */, uint64_t* total_out = nullptr) {
  int64_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (n_tiles < 1) n_tiles = 1;
  tile_sum_kernel<T><<<(unsigned)n_tiles, SCAN_T, 0, st>>>(in, n, tile_sums);
  hawk_note_launch(1);
  tile_scan_kernel<<<1, 1024, 0, st>>>(tile_sums, n_tiles, total_out);
  hawk_note_launch(1);
  tile_apply_kernel<T><<<(unsigned)n_tiles, SCAN_T, 0, st>>>(in, n, tile_sums, out);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "exclusive_scan launch");
}

int64_t scan_tiles(int64_t n) {
  int64_t t = (n + SCAN_TILE - 1) / SCAN_TILE;
  return t < 1 ? 1 : t;
}
int exclusive_scan_u8(cudaStream_t st, const uint8_t* in, int64_t n, uint64_t* out, uint64_t* tile_sums) {
  return exclusive_scan_impl<uint8_t>(st, in, n, out, tile_sums);
}
int exclusive_scan_u64(cudaStream_t st, const uint64_t* in, int64_t n, uint64_t* out, uint64_t* tile_sums) {
  return exclusive_scan_impl<uint64_t>(st, in, n, out, tile_sums);
}
int exclusive_scan_u32(cudaStream_t st, const uint32_t* in, int64_t n, uint64_t* out, uint64_t* tile_sums,
                       uint64_t* total_out) {
  return exclusive_scan_impl<uint32_t>(st, in, n, out, tile_sums, total_out);
}

// ---------------------------------------------------------------- final gather
// Merge the two strand streams into the reference's emission order
// (haplotype, strand, position, expansion; search_guides.py:530-547).
struct GatherArgs {
  BatchView B;
  ScanConst K;
  // per stream
  const uint64_t* recs[2];
  const int64_t* row_hit[2];   // null: row == hit (phased)
  const uint8_t* keep[2];
  const uint64_t* kept_excl[2];
  const uint8_t* text_pre[2];  // unphased: resolved strings per row; null: extract from planes
  const int32_t* start[2];
  const int32_t* stop[2];
  int64_t n_rows[2];
  uint64_t kept_total[2];
  // outputs
  int32_t* o_hap;
  uint8_t* o_strand;
  int32_t* o_pos;
  int32_t* o_start;
  int32_t* o_stop;
  uint8_t* o_text;
  int32_t text_stride;
};

__device__ __forceinline__ int32_t row_hap(const GatherArgs& A, int t, int64_t row) {
  int64_t hit = A.row_hit[t] ? A.row_hit[t][row] : row;
  return (int32_t)(A.recs[t][hit] >> 32);
}

__global__ void gather_kernel(const __grid_constant__ GatherArgs A, int s) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A.n_rows[s] || !A.keep[s][r]) return;
  int64_t hit = A.row_hit[s] ? A.row_hit[s][r] : r;
  uint64_t rec = A.recs[s][hit];
  int32_t h = (int32_t)(rec >> 32), pos = (int32_t)(rec & 0xFFFFFFFFu);
  // rows of the other stream emitted before this one: hap < h (strand 0) or hap <= h (strand 1)
  int t = 1 - s;
  int64_t lo = 0, hi = A.n_rows[t];
  while (lo < hi) {
    int64_t m = (lo + hi) >> 1;
    int32_t hm = row_hap(A, t, m);
    bool before = s == 0 ? (hm < h) : (hm <= h);
    if (before) lo = m + 1; else hi = m;
  }
  uint64_t other = lo < A.n_rows[t] ? A.kept_excl[t][lo] : A.kept_total[t];
  uint64_t f = A.kept_excl[s][r] + other;
  A.o_hap[f] = h;
  A.o_strand[f] = (uint8_t)s;
  A.o_pos[f] = pos;
  A.o_start[f] = A.start[s][hit];
  A.o_stop[f] = A.stop[s][hit];
  const int W = A.K.C + 2 * HAWK_GUIDESEQPAD;
  uint8_t* dst = A.o_text + f * (uint64_t)A.text_stride;
  for (int j = W; j < A.text_stride; ++j) dst[j] = 0;
  if (A.text_pre[s]) {
    const uint8_t* src = A.text_pre[s] + r * W;
    for (int j = 0; j < W; ++j) dst[j] = src[j];
  } else {
    // extract_guide_sequence (:134-160): rebuild the padded window text from planes + case bits
    int64_t chunk0 = A.B.slot_off[h] >> 5;
    int32_t w0 = pos + A.K.geom[s].w0;
    for (int j = 0; j < W; ++j) {
      char ch = nibble_letter(nibble_at(A.B.q, chunk0, w0 + j));
      dst[j] = (uint8_t)(lower_at(A.B.v, chunk0, w0 + j) ? ch + 32 : ch);
    }
  }
}

// ---------------------------------------------------------------- first-seen buckets
// bucket id of a row = smallest emission index among the rows sharing its
// (start, strand) key (group_guides_position, :306-337: dict insertion order).
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

__global__ void bucket_insert_kernel(const int32_t* __restrict__ start,
                                     const uint8_t* __restrict__ strand, int64_t n,
                                     unsigned long long* keys, unsigned long long* vals,
                                     uint64_t mask) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long key = ((unsigned long long)(uint32_t)start[i] << 1) | strand[i];
  uint64_t slot = mix64(key) & mask;
  for (;;) {
    unsigned long long old = atomicCAS(&keys[slot], ~0ull, key);
    if (old == ~0ull || old == key) {
      atomicMin(&vals[slot], (unsigned long long)i);
      return;
    }
    slot = (slot + 1) & mask;
  }
}

__global__ void bucket_lookup_kernel(const int32_t* __restrict__ start,
                                     const uint8_t* __restrict__ strand, int64_t n,
                                     const unsigned long long* keys, const unsigned long long* vals,
                                     uint64_t mask, uint32_t* __restrict__ bucket) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long key = ((unsigned long long)(uint32_t)start[i] << 1) | strand[i];
  uint64_t slot = mix64(key) & mask;
  while (keys[slot] != key) slot = (slot + 1) & mask;
  bucket[i] = (uint32_t)vals[slot];
}

// ---------------------------------------------------------------- nibble export
__global__ void export_nibbles_kernel(const Planes* __restrict__ q, const uint32_t* __restrict__ v,
                                      int64_t chunk0, int32_t len, uint8_t* __restrict__ nib,
                                      uint8_t* __restrict__ lower) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  nib[i] = (uint8_t)nibble_at(q, chunk0, i);
  if (lower) lower[i] = (uint8_t)lower_at(v, chunk0, i);
}

// ---------------------------------------------------------------- launch wrappers
int launch_ref_range(cudaStream_t st, const uint64_t* r0, int64_t n0, const uint64_t* r1, int64_t n1,
                     int32_t ref_h, int64_t* out) {
  ref_range_kernel<<<1, 128, 0, st>>>(r0, n0, r1, n1, ref_h, out);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "ref_range_kernel launch");
}

int launch_rows(cudaStream_t st, const BatchView& B, const ScanConst& K, const uint64_t* recs,
                int64_t n, int s, int32_t ref_h, const int64_t* ref_range, int dedup,
                int32_t* start, int32_t* stop, uint8_t* keep) {
  if (n <= 0) return HAWK_OK;
  rows_kernel<<<grid_for(n, 128), 128, 0, st>>>(B, K, recs, n, s, ref_h, ref_range, dedup, start,
                                                stop, keep);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "rows_kernel launch");
}

int launch_expand_count(cudaStream_t st, const BatchView& B, const ScanConst& K,
                        const uint64_t* recs, int64_t n, int s, uint64_t* cnt, int* err) {
  if (n <= 0) return HAWK_OK;
  expand_count_kernel<<<grid_for(n, 128), 128, 0, st>>>(B, K, recs, n, s, cnt, err);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "expand_count_kernel launch");
}

int launch_expand_write(cudaStream_t st, const BatchView& B, const ScanConst& K,
                        const uint64_t* recs, int64_t n_hits, int s, const uint64_t* off,
                        int64_t n_rows, int32_t ref_h, const int64_t* ref_range,
                        const int32_t* start, uint8_t* text, int64_t* row_hit, uint8_t* keep) {
  if (n_rows <= 0) return HAWK_OK;
  expand_write_kernel<<<grid_for(n_rows, 128), 128, 0, st>>>(B, K, recs, n_hits, s, off, n_rows,
                                                            ref_h, ref_range, start, text, row_hit,
                                                            keep);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "expand_write_kernel launch");
}

int launch_gather(cudaStream_t st, const GatherLaunch& g) {
  GatherArgs A;
  A.B = g.B;
  A.K = g.K;
  for (int s = 0; s < 2; ++s) {
    A.recs[s] = g.recs[s];
    A.row_hit[s] = g.row_hit[s];
    A.keep[s] = g.keep[s];
    A.kept_excl[s] = g.kept_excl[s];
    A.text_pre[s] = g.text_pre[s];
    A.start[s] = g.start[s];
    A.stop[s] = g.stop[s];
    A.n_rows[s] = g.n_rows[s];
    A.kept_total[s] = g.kept_total[s];
  }
  A.o_hap = g.o_hap;
  A.o_strand = g.o_strand;
  A.o_pos = g.o_pos;
  A.o_start = g.o_start;
  A.o_stop = g.o_stop;
  A.o_text = g.o_text;
  A.text_stride = g.text_stride;
  for (int s = 0; s < 2; ++s) {
    if (g.n_rows[s] <= 0) continue;
    gather_kernel<<<grid_for(g.n_rows[s], 128), 128, 0, st>>>(A, s);
    hawk_note_launch(1);
  }
  return hawk_check_cuda(cudaGetLastError(), "gather_kernel launch");
}

int launch_buckets(cudaStream_t st, const int32_t* start, const uint8_t* strand, int64_t n,
                   unsigned long long* keys, unsigned long long* vals, uint64_t table_size,
                   uint32_t* bucket) {
  if (n <= 0) return HAWK_OK;
  bucket_insert_kernel<<<grid_for(n, 256), 256, 0, st>>>(start, strand, n, keys, vals,
                                                         table_size - 1);
  hawk_note_launch(1);
  bucket_lookup_kernel<<<grid_for(n, 256), 256, 0, st>>>(start, strand, n, keys, vals,
                                                         table_size - 1, bucket);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "bucket kernels launch");
}

int launch_export_nibbles(cudaStream_t st, const void* q, const uint32_t* v, int64_t chunk0,
                          int32_t len, uint8_t* nib, uint8_t* lower) {
  if (len <= 0) return HAWK_OK;
  export_nibbles_kernel<<<grid_for(len, 256), 256, 0, st>>>((const Planes*)q, v, chunk0, len, nib,
                                                            lower);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "export_nibbles_kernel launch");
}

}  // namespace hawk
