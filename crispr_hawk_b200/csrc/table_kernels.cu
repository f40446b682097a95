// table_kernels.cu -- the phased / variant-free guide-table pipeline downstream of the
// scan (retrieve_guides + remove_redundant_guides, search_guides.py:340-369, :423-507,
// without the unphased resolve_guide branch, which lives in post_kernels.cu):
//
//   ref_bitmap   one bit per REF position and strand that holds a REF guide
//   rows         per hit: genomic start/stop through the run-length posmap (:260-280), REF
//                partner by direct lookup (REF coordinates are linear) + upper-cased core
//                comparison on the nibble planes (:356-369) -> keep flag, kept rows per block
//   blk_prefix   exclusive prefix of the per-block kept counts (single CTA)
//   hap_offsets  rows of each strand stream that precede a haplotype (for the emission-order
//                merge of the two streams, :530-547)
//   gather       surviving rows -> table columns in emission order, padded window text
//                rebuilt from the planes (:134-160) with 128-bit stores, first-seen bucket id
//                of the (start, strand) key by atomicMin on a direct-address table (:306-337)
//   bucket_read  bucket id column
#include <cuda_runtime.h>
#include <stdint.h>

#include "hawk_core.h"
#include "hawk_kernels.h"
#include "hawk_post.h"

namespace hawk {

constexpr int ROW_T = 256;  // rows per block in rows / gather (the two must agree)

// ---------------------------------------------------------------- REF guide bitmap
__global__ void ref_bitmap_kernel(const uint64_t* __restrict__ recs0, const uint64_t* __restrict__ recs1,
                                  const int64_t* __restrict__ ref_range, uint32_t* __restrict__ bm0,
                                  uint32_t* __restrict__ bm1) {
  const int s = blockIdx.y;
  const uint64_t* recs = s ? recs1 : recs0;
  uint32_t* bm = s ? bm1 : bm0;
  const int64_t lo = ref_range[2 * s], hi = ref_range[2 * s + 1];
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t pos = (uint32_t)(recs[i] & 0xFFFFFFFFu);
    atomicOr(&bm[pos >> 5], 1u << (pos & 31));
  }
}

// ---------------------------------------------------------------- coarse posmap index
// rows_fast finds a hit's posmap segment by binary search: ~8 dependent loads per hit on a
// haplotype with a few hundred segments, the longest chain of the kernel. One thread per
// (haplotype, 4 kb bucket) does that search once per batch; a hit then needs one lookup and a
// search over the segments of its own bucket (usually one or two).
__global__ void seg_index_kernel(const int64_t* __restrict__ seg_off, const int32_t* __restrict__ seg_rel, int32_t n_hap,
                                 int32_t stride, int32_t* __restrict__ idx) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)n_hap * stride) return;
  const int32_t h = (int32_t)(t / stride), k = (int32_t)(t - (int64_t)h * stride);
  idx[t] = seg_index_entry(seg_off, seg_rel, h, k);
}

int launch_seg_index(cudaStream_t st, const int64_t* seg_off, const int32_t* seg_rel, int32_t n_hap, int32_t stride,
                     int32_t* idx) {
  const int64_t n = (int64_t)n_hap * stride;
  if (n <= 0) return HAWK_OK;
  seg_index_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(seg_off, seg_rel, n_hap, stride, idx);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "seg_index_kernel launch");
}

// ---------------------------------------------------------------- rows
// Both strands in one launch, blocks interleaved (even = strand 0, odd = strand 1): the two hit
// streams are sorted by (haplotype, position) and about equally dense, so block k of either
// stream works on the same stretch of the same haplotype and the second reader of a plane
// sector finds it in L2 instead of DRAM.
struct RowsStrand {
  const uint64_t* recs;
  int64_t n;
  const uint32_t* ref_bm;     // REF guide bitmap of this strand (ref_linear)
  int32_t* start;
  int32_t* stop;
  uint8_t* keep;
  uint32_t* blk_cnt;
};

struct RowsArgs {
  BatchView B;
  ScanConst K;
  RowsStrand S[2];
  int32_t ref_h;         // -1: no REF haplotype
  int32_t ref_linear;    // REF posmap is one step-1 segment: posmap(i) = ref_g0 + i
  int32_t ref_g0, ref_len;
  const int64_t* ref_range;   // REF record range per strand (general REF posmap)
  int32_t drop_ref;           // streamed search, later groups: REF rows serve as partners only
};

__global__ void __launch_bounds__(ROW_T) rows_fast_kernel(const __grid_constant__ RowsArgs A) {
  const int s = blockIdx.x & 1;
  const int64_t blk = blockIdx.x >> 1;
  const RowsStrand S = A.S[s];
  if (blk * ROW_T >= S.n) return;  // the shorter stream has fewer blocks (uniform over the CTA)
  const int64_t i = blk * ROW_T + threadIdx.x;
  int k = 0;
  if (i < S.n) {
    const uint64_t rec = S.recs[i];
    const int32_t h = (int32_t)(rec >> 32), pos = (int32_t)(rec & 0xFFFFFFFFu);
    const RowCoords rc = row_coords(A.B, A.K, h, pos, s);
    S.start[i] = rc.start;
    S.stop[i] = rc.stop;
    k = 1;
    if (A.drop_ref && A.B.is_ref[h]) {
      k = 0;
    } else if (A.ref_h >= 0 && !A.B.is_ref[h]) {
      // remove_redundant_guides (:356-369): a non-REF guide whose upper-cased core equals the
      // REF guide's at the same (start, strand) is dropped
      int32_t rpivot = -1;
      if (A.ref_linear) {
        const int32_t rp = rc.start - A.ref_g0, rpos = rp - A.K.geom[s].c0;
        if (rp >= 0 && rpos >= 0 && rpos < A.ref_len && ((S.ref_bm[rpos >> 5] >> (rpos & 31)) & 1u)) rpivot = rp;
      } else {
        const int64_t lo = A.ref_range[2 * s], hi = A.ref_range[2 * s + 1];
        const int64_t j = find_ref_partner(A.B, A.K, S.recs, lo, hi, A.ref_h, s, rc.start);
        if (j < hi) {
          const int32_t rp = (int32_t)(S.recs[j] & 0xFFFFFFFFu) + A.K.geom[s].c0;
          if (posmap_eval(A.B.seg_rel, A.B.seg_gen, A.B.seg_step, A.B.seg_off[A.ref_h], A.B.seg_off[A.ref_h + 1], rp) ==
              rc.start)
            rpivot = rp;
        }
      }
      if (rpivot >= 0 &&
          cores_equal(A.B.q, A.B.slot_off[h] >> 5, rc.pivot, A.B.slot_off[A.ref_h] >> 5, rpivot, A.K.C))
        k = 0;
    }
    S.keep[i] = (uint8_t)k;
  }
  const int cnt = __syncthreads_count(k);
  if (threadIdx.x == 0) S.blk_cnt[blk] = (uint32_t)cnt;
}

// ---------------------------------------------------------------- block prefix (one CTA per strand)
__global__ void __launch_bounds__(1024) blk_prefix_kernel(const uint32_t* __restrict__ in0, const uint32_t* __restrict__ in1,
                                                          int64_t n0, int64_t n1, uint64_t* __restrict__ out0,
                                                          uint64_t* __restrict__ out1, uint64_t* totals) {
  __shared__ uint64_t part[1024];
  const int s = blockIdx.x;
  const uint32_t* in = s ? in1 : in0;
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
    out[j] = run;
    run += in[j];
  }
  if (tid == 1023) totals[s] = part[1023];
}

// ---------------------------------------------------------------- per-haplotype offsets
// kb[s][h] = kept rows of stream s whose haplotype is < h  (h = 0 .. n_hap)
__global__ void hap_offsets_kernel(const uint64_t* __restrict__ recs0, const uint64_t* __restrict__ recs1,
                                   int64_t n0, int64_t n1, const uint8_t* __restrict__ keep0,
                                   const uint8_t* __restrict__ keep1, const uint64_t* __restrict__ base0,
                                   const uint64_t* __restrict__ base1, const uint64_t* __restrict__ totals,
                                   int32_t n_hap, uint64_t* __restrict__ kb) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * (int64_t)(n_hap + 1)) return;
  const int s = (int)(t / (n_hap + 1));
  const int32_t h = (int32_t)(t % (n_hap + 1));
  const uint64_t* recs = s ? recs1 : recs0;
  const int64_t n = s ? n1 : n0;
  const uint8_t* keep = s ? keep1 : keep0;
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
    const int64_t b0 = lo / ROW_T * ROW_T;
    r = base[lo / ROW_T];
    for (int64_t j = b0; j < lo; ++j) r += keep[j];
  }
  kb[(size_t)s * (n_hap + 1) + h] = r;
}

// ---------------------------------------------------------------- gather
struct GatherStrand {
  const uint64_t* recs;
  const uint8_t* keep;
  const uint64_t* blk_base;
  const int32_t* start;
  const int32_t* stop;
  const uint64_t* kb_other;  // kb of the other stream, n_hap + 1 entries
  int64_t n;
};

struct GatherFastArgs {  // both strands in one launch, blocks interleaved (see RowsArgs)
  BatchView B;
  ScanConst K;
  GatherStrand S[2];
  int32_t text_stride;  // bytes per text row, multiple of 16
  int32_t* o_hap;
  uint8_t* o_strand;
  int32_t* o_pos;
  int32_t* o_start;
  int32_t* o_stop;
  uint8_t* o_text;
  uint32_t* key_table;  // direct-address first-seen table, or null
  int32_t key_min;      // smallest genomic coordinate of the batch
  RowMap rm;            // streamed search: global row numbering / haplotype indices
};

__global__ void __launch_bounds__(ROW_T) gather_fast_kernel(const __grid_constant__ GatherFastArgs A) {
  __shared__ uint32_t warp_cnt[ROW_T / 32];
  const int s = blockIdx.x & 1;
  const int64_t blk = blockIdx.x >> 1;
  const GatherStrand S = A.S[s];
  if (blk * ROW_T >= S.n) return;
  const int64_t i = blk * ROW_T + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool k = i < S.n && S.keep[i];
  const uint32_t bal = __ballot_sync(0xFFFFFFFFu, k);
  if (lane == 0) warp_cnt[warp] = __popc(bal);
  __syncthreads();
  if (!k) return;
  uint32_t rank = __popc(bal & ((1u << lane) - 1u));
  for (int w = 0; w < warp; ++w) rank += warp_cnt[w];
  const uint64_t rec = S.recs[i];
  const int32_t h = (int32_t)(rec >> 32), pos = (int32_t)(rec & 0xFFFFFFFFu);
  // rows of the other stream emitted before this one: haplotype < h (strand 0) or <= h (strand 1)
  const uint64_t f = S.blk_base[blk] + rank + S.kb_other[h + (s == 1 ? 1 : 0)];
  const int32_t st = S.start[i];
  A.o_hap[f] = h == A.rm.ref_local ? A.rm.ref_global : h + A.rm.hap_add;
  A.o_strand[f] = (uint8_t)s;
  A.o_pos[f] = pos;
  A.o_start[f] = st;
  A.o_stop[f] = S.stop[i];
  if (A.key_table) atomicMin(&A.key_table[((uint32_t)(st - A.key_min) << 1) | (uint32_t)s], (uint32_t)(f + A.rm.row_base));
  // extract_guide_sequence (:134-160): padded window text from planes + case bits
  const int W = A.K.C + 2 * HAWK_GUIDESEQPAD;
  const int64_t chunk0 = A.B.slot_off[h] >> 5;
  const int32_t w0 = pos + A.K.geom[s].w0;
  uint4* dst = reinterpret_cast<uint4*>(A.o_text + f * (uint64_t)A.text_stride);
  for (int j0 = 0; j0 < A.text_stride; j0 += 32) {
    const int32_t o = w0 + j0;
    const uint32_t sh = (uint32_t)(o & 31);
    const uint4 q0 = *reinterpret_cast<const uint4*>(&A.B.q[chunk0 + (o >> 5)]);
    const uint4 q1 = *reinterpret_cast<const uint4*>(&A.B.q[chunk0 + (o >> 5) + 1]);
    const uint32_t v0 = A.B.v[chunk0 + (o >> 5)], v1 = A.B.v[chunk0 + (o >> 5) + 1];
    const uint32_t pa = funnel_r(q0.x, q1.x, sh), pc = funnel_r(q0.y, q1.y, sh), pg = funnel_r(q0.z, q1.z, sh),
                   pt = funnel_r(q0.w, q1.w, sh), pv = funnel_r(v0, v1, sh);
    const int left = W - j0;  // > 0: the stride rounds W up to a multiple of 16
    uint32_t w[8];
    planes_to_chars32(pa, pc, pg, pt, pv, left >= 32 ? 0xFFFFFFFFu : ((1u << left) - 1u), w);
    dst[j0 >> 4] = make_uint4(w[0], w[1], w[2], w[3]);
    if (j0 + 16 < A.text_stride) dst[(j0 >> 4) + 1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
}

__global__ void bucket_read_kernel(const int32_t* __restrict__ start, const uint8_t* __restrict__ strand,
                                   const uint64_t* __restrict__ totals, const uint32_t* __restrict__ key_table,
                                   int32_t key_min, uint32_t* __restrict__ bucket) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)(totals[0] + totals[1])) return;
  bucket[i] = key_table[((uint32_t)(start[i] - key_min) << 1) | strand[i]];
}

// first-seen table of a finished table (final merge of per-rank tables): key -> smallest row
// a start outside [key_min, key_min + key_span) (or a strand other than 0 / 1) is reported through
// `bad` instead of indexing the table out of bounds
__global__ void first_seen_mark_kernel(const int32_t* __restrict__ start, const uint8_t* __restrict__ strand, int64_t n,
                                       int32_t key_min, int64_t key_span, uint32_t* __restrict__ key_table,
                                       uint32_t* __restrict__ bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t k = (int64_t)start[i] - key_min;
  if (k < 0 || k >= key_span || strand[i] > 1) {
    atomicOr(bad, 1u);
    return;
  }
  atomicMin(&key_table[((uint32_t)k << 1) | strand[i]], (uint32_t)i);
}

__global__ void first_seen_read_kernel(const int32_t* __restrict__ start, const uint8_t* __restrict__ strand, int64_t n,
                                       int32_t key_min, int64_t key_span, const uint32_t* __restrict__ key_table,
                                       uint32_t* __restrict__ bucket) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t k = (int64_t)start[i] - key_min;
  bucket[i] = (k < 0 || k >= key_span || strand[i] > 1) ? 0xFFFFFFFFu : key_table[((uint32_t)k << 1) | strand[i]];
}

// ---------------------------------------------------------------- launch wrappers
static inline unsigned blocks_for(int64_t n, int t) {
  int64_t b = (n + t - 1) / t;
  return (unsigned)(b < 1 ? 1 : b);
}

int64_t row_blocks(int64_t n) { return (n + ROW_T - 1) / ROW_T; }

int launch_ref_bitmap(cudaStream_t st, const uint64_t* r0, const uint64_t* r1, const int64_t* ref_range,
                      uint32_t* bm0, uint32_t* bm1) {
  ref_bitmap_kernel<<<dim3(296, 2), 256, 0, st>>>(r0, r1, ref_range, bm0, bm1);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "ref_bitmap_kernel launch");
}

int launch_rows_fast(cudaStream_t st, const BatchView& B, const ScanConst& K, const uint64_t* const recs[2],
                     const int64_t n[2], const RefInfo& ref, const uint32_t* const ref_bm[2], const int64_t* ref_range,
                     int32_t drop_ref, int32_t* const start[2], int32_t* const stop[2], uint8_t* const keep[2],
                     uint32_t* const blk_cnt[2]) {
  const int64_t nb = row_blocks(n[0] > n[1] ? n[0] : n[1]);
  if (nb <= 0) return HAWK_OK;
  RowsArgs A{};
  A.B = B;
  A.K = K;
  for (int s = 0; s < 2; ++s) A.S[s] = RowsStrand{recs[s], n[s], ref_bm[s], start[s], stop[s], keep[s], blk_cnt[s]};
  A.ref_h = ref.h;
  A.ref_linear = ref.linear;
  A.ref_g0 = ref.g0;
  A.ref_len = ref.len;
  A.ref_range = ref_range;
  A.drop_ref = drop_ref;
  rows_fast_kernel<<<(unsigned)(2 * nb), ROW_T, 0, st>>>(A);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "rows_fast_kernel launch");
}

int launch_blk_prefix(cudaStream_t st, const uint32_t* const cnt[2], const int64_t n_blk[2], uint64_t* const base[2],
                      uint64_t* totals) {
  blk_prefix_kernel<<<2, 1024, 0, st>>>(cnt[0], cnt[1], n_blk[0], n_blk[1], base[0], base[1], totals);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "blk_prefix_kernel launch");
}

int launch_hap_offsets(cudaStream_t st, const uint64_t* r0, const uint64_t* r1, int64_t n0, int64_t n1,
                       const uint8_t* k0, const uint8_t* k1, const uint64_t* b0, const uint64_t* b1,
                       const uint64_t* totals, int32_t n_hap, uint64_t* kb) {
  const int64_t n = 2 * (int64_t)(n_hap + 1);
  hap_offsets_kernel<<<blocks_for(n, 128), 128, 0, st>>>(r0, r1, n0, n1, k0, k1, b0, b1, totals, n_hap, kb);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "hap_offsets_kernel launch");
}

int launch_gather_fast(cudaStream_t st, const BatchView& B, const ScanConst& K, const uint64_t* const recs[2],
                       const uint8_t* const keep[2], const uint64_t* const blk_base[2], const int32_t* const start[2],
                       const int32_t* const stop[2], const uint64_t* const kb_other[2], const int64_t n[2],
                       int32_t text_stride, int32_t* o_hap, uint8_t* o_strand, int32_t* o_pos, int32_t* o_start,
                       int32_t* o_stop, uint8_t* o_text, uint32_t* key_table, int32_t key_min, const RowMap& rm) {
  const int64_t nb = row_blocks(n[0] > n[1] ? n[0] : n[1]);
  if (nb <= 0) return HAWK_OK;
  GatherFastArgs A{};
  A.B = B;
  A.K = K;
  for (int s = 0; s < 2; ++s) A.S[s] = GatherStrand{recs[s], keep[s], blk_base[s], start[s], stop[s], kb_other[s], n[s]};
  A.text_stride = text_stride;
  A.o_hap = o_hap;
  A.o_strand = o_strand;
  A.o_pos = o_pos;
  A.o_start = o_start;
  A.o_stop = o_stop;
  A.o_text = o_text;
  A.key_table = key_table;
  A.key_min = key_min;
  A.rm = rm;
  gather_fast_kernel<<<(unsigned)(2 * nb), ROW_T, 0, st>>>(A);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "gather_fast_kernel launch");
}

int launch_bucket_read(cudaStream_t st, const int32_t* start, const uint8_t* strand, int64_t n_max,
                       const uint64_t* totals, const uint32_t* key_table, int32_t key_min, uint32_t* bucket) {
  if (n_max <= 0) return HAWK_OK;
  bucket_read_kernel<<<blocks_for(n_max, 256), 256, 0, st>>>(start, strand, totals, key_table, key_min, bucket);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "bucket_read_kernel launch");
}

}  // namespace hawk

// device layer: bucket ids of a table that already sits in device memory
extern "C" int hawk_first_seen_dev(void* stream, const int32_t* d_start, const uint8_t* d_strand, int64_t n,
                                   int32_t key_min, int64_t key_span, uint32_t* d_key_table, uint32_t* d_bucket) {
  if (n < 0 || key_span < 0 || n >= 0xFFFFFFFFll || key_span > 0x7FFFFFFFll)
    return hawk_fail(HAWK_EINVAL, "hawk_first_seen_dev: bad sizes");
  if (n == 0) return HAWK_OK;
  if (!d_start || !d_strand || !d_key_table || !d_bucket || key_span == 0)
    return hawk_fail(HAWK_EINVAL, "hawk_first_seen_dev: null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  // the word behind the table (entry 2 * key_span) is the out-of-range flag
  cudaError_t e = cudaMemsetAsync(d_key_table, 0xFF, (size_t)key_span * 2 * 4, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_key_table + 2 * key_span, 0, 4, st);
  if (e != cudaSuccess) return hawk_check_cuda(e, "key table memset");
  const unsigned blocks = (unsigned)((n + 255) / 256);
  hawk::first_seen_mark_kernel<<<blocks, 256, 0, st>>>(d_start, d_strand, n, key_min, key_span, d_key_table,
                                                       d_key_table + 2 * key_span);
  hawk::first_seen_read_kernel<<<blocks, 256, 0, st>>>(d_start, d_strand, n, key_min, key_span, d_key_table, d_bucket);
  hawk_note_launch(2);
  int rc = hawk_check_cuda(cudaGetLastError(), "first_seen kernels launch");
  if (rc != HAWK_OK) return rc;
  uint32_t bad = 0;
  e = cudaMemcpyAsync(&bad, d_key_table + 2 * key_span, 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return hawk_check_cuda(e, "first_seen flag read");
  if (bad)
    return hawk_fail(HAWK_EINVAL, "hawk_first_seen_dev: a row's start lies outside [key_min, key_min + key_span) "
                     "(or its strand is not 0 / 1)");
  return HAWK_OK;
}
