// fused_kernels.cu -- K1 + K2 in one pass over the haplotype texts (sm_100a).
//
// pack_kernel (scan_kernels.cu) writes 0.625 B of planes per base for every base, and the
// staged K2 then reads back only the few per cent of them that lie next to a variant base.
// Here one kernel reads the ASCII once, builds the plane words in registers and
//   * stores them only where a later stage can read them: chunks within `reach` chunks of a
//     chunk that holds a variant (lower-case) base, and every chunk of a REF haplotype
//     (what rows_fast / gather_fast / the REF partner comparison touch);
//   * finds the candidate chunks (a variant base within reach of a guide core, or REF) from
//     the case words of the neighbouring chunks, which the same warp holds in registers;
//   * matches the PAM on both strands for the candidates with the fused filters
//     (search_guides.py:32-131, :395-420, :468-471 -- scan_chunk_small of hawk_core.h, the same
//     code the staged match_kernel runs) and appends {haplotype, chunk, hit masks} entries to a
//     per-warp segment in chunk order.
// A second, small kernel turns the entries into the sorted (hap << 32 | pos) record streams at
// exact prefix offsets, like the staged expand_kernel.
//
// Work decomposition is FLAT over the slot space: warp w owns the chunks
// [w * sub, (w + 1) * sub), sub = fused_sub_size(n_chunks), whatever haplotypes they belong to (a per-warp cursor maps
// chunks to haplotypes), so 5,009 haplotypes of 1 Mb and 428,529 haplotypes of 200 bases cost the
// same per base. A warp streams its sub-range 32 chunks (1 KB of text) per iteration, software-
// pipelined by one iteration: iteration i is packed, then iteration i - 1 -- whose neighbours on
// both sides are now in registers -- is finalised (store decision, candidate decision). One
// extra iteration either side of the sub-range is packed as halo (1.6 % of the work).
// Candidates are queued per warp and matched 32 at a time, every lane busy, although only a few
// per cent of the chunks are candidates; the matcher reads the planes back through L2 (they
// were stored a moment ago by this warp).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "hawk_core.h"
#include "hawk_kernels.h"
#include "hawk_post.h"

namespace hawk {

// first chunk of haplotype h's territory (its leading zero gap included); T(n_hap) = n_chunks
__device__ __forceinline__ int64_t territory(const int64_t* __restrict__ slot_off, int32_t n_hap, int64_t n_chunks,
                                             int32_t h) {
  return h >= n_hap ? n_chunks : (__ldg(&slot_off[h]) >> 5) - (HAWK_SLOT_GAP / HAWK_CHUNK);
}

struct FusedArgs {
  const uint4* ascii;
  int64_t n_chunks;
  uint4* q;
  uint32_t* v;
  uint32_t* nz;
  const int64_t* slot_off;
  int32_t n_hap;
  const HapScan* hs;
  ScanConst K;
  int32_t reach;      // planes are kept for chunks within `reach` chunks of a variant chunk (2..3)
  int32_t sub;        // chunks per warp sub-range (fused_sub_size)
  const uint64_t* seg_base;  // per sub-range: first entry slot, capacity
  const uint32_t* seg_cap;
  uint4* entries;       // {haplotype, chunk (haplotype-relative), hits strand 0, hits strand 1}
  uint32_t* cnt_ent;    // per sub-range: entries, hits per strand
  uint32_t* cnt_hit0;
  uint32_t* cnt_hit1;
  uint32_t* overflow;   // set when a segment was too small (the caller falls back to the staged K2)
  unsigned long long* bad;
};

struct Planes5 {
  uint32_t a, c, g, t, v;
};

struct Raw8 {
  uint32_t w[8];
};

// haplotype whose territory holds the chunks being finalised, everything relative to the
// sub-range start s (32-bit): warp-uniform
struct HapCursor {
  int32_t h;
  int32_t t_next;      // first chunk of the next territory
  int32_t c_lo, c_hi;  // chunks that can hold a scanned position: [c_lo, c_hi)
  int32_t is_ref;
};

// WIDE (the texts are 32-byte aligned) chose a 256-bit load when the texts were prefetched into
// registers; the cp.async ring copies 16 bytes at a time either way, so both values now compile to
// the same code (kept so the kernel names in earlier profiles stay comparable).
// SUBC: the sub-range size as a compile-time constant (FUSED_SUB_MAX, the large-input case), or 0 =
// take it from the arguments (small slot spaces)
template <bool WIDE, int REACH, int SUBC>
__global__ void __launch_bounds__(FUSED_WARPS * 32, FUSED_CTAS_PER_SM / FUSED_WARPS) fused_scan_kernel(const __grid_constant__ FusedArgs A) {
  __shared__ uint16_t q_off[FUSED_WARPS][64];  // candidate queue (ring): chunk offset in the sub-range
  __shared__ int32_t q_hap[FUSED_WARPS][64];   // ... and its haplotype
  __shared__ uint4 ring[FUSED_RING][64];       // texts in flight: FUSED_RING iterations x 1 KB (one warp per CTA)
  static_assert(FUSED_WARPS == 1, "the text ring is per CTA");
  unsigned ring_base = (unsigned)__cvta_generic_to_shared(&ring[0][0]) + 32u * (threadIdx.x & 31);
  asm volatile("" : "+r"(ring_base));  // keep it in a register (it was re-derived every iteration)
  const int lane = threadIdx.x & 31, warp = FUSED_WARPS == 1 ? 0 : (int)(threadIdx.x >> 5);
  const int64_t sub = (int64_t)blockIdx.x * FUSED_WARPS + warp;
  const int32_t SUB = SUBC ? SUBC : A.sub;
  const int64_t s = sub * SUB;
  if (s >= A.n_chunks) return;  // uniform over the warp; no block-wide barrier below
  const int32_t n = (int32_t)(s + SUB < A.n_chunks ? SUB : A.n_chunks - s);  // chunks owned
  const int32_t n_avail = (int32_t)(A.n_chunks - s < SUB + 32 ? A.n_chunks - s : SUB + 32);  // + halo
  const int32_t n_iter = (n + 31) >> 5;
  const ScanConst& K = A.K;
  BatchView B{};
  B.q = reinterpret_cast<const Planes*>(A.q);
  B.v = A.v;
  const uint4* const asc = A.ascii + 2 * s + 2 * lane;  // this lane's chunk of iteration 0
  uint4* const qs = A.q + s + lane;
  uint32_t* const vs = A.v + s + lane;
  uint32_t* const nzs = A.nz + (s >> 5);
  const uint32_t lane_bit = 1u << lane, lanes_below = lane_bit - 1u;

  auto load_cursor = [&](int32_t h) {
    HapCursor cu;
    cu.h = h;
    const int64_t tn = territory(A.slot_off, A.n_hap, A.n_chunks, h + 1) - s;
    cu.t_next = tn > (int64_t)(SUB + 64) ? SUB + 64 : (int32_t)tn;
    const HapScan* H = A.hs + h;
    const int64_t c0 = H->chunk0 - s;  // may be far below zero; the clamps keep the interval test exact
    const int32_t a = H->a, b = H->b;
    int64_t lo = c0 + (a >> 5), hi = b > a ? c0 + ((b + 31) >> 5) : lo;
    if (lo < -64) lo = -64;
    if (hi < -64) hi = -64;
    if (lo > SUB + 64) lo = SUB + 64;
    if (hi > SUB + 64) hi = SUB + 64;
    cu.c_lo = (int32_t)lo;
    cu.c_hi = (int32_t)hi;
    cu.is_ref = H->is_ref;
    return cu;
  };
  HapCursor cur;
  {
    int32_t lo = 0, hi = A.n_hap;  // T(lo) <= s < T(hi)
    while (hi - lo > 1) {
      const int32_t mid = (lo + hi) >> 1;
      if (territory(A.slot_off, A.n_hap, A.n_chunks, mid) <= s) lo = mid; else hi = mid;
    }
    cur = load_cursor(lo);
  }

  uint4* const seg = A.entries + A.seg_base[sub];
  const uint32_t cap = A.seg_cap[sub];
  uint32_t n_ent = 0, n_h0 = 0, n_h1 = 0;
  uint32_t qhead = 0, qn = 0;

  // match the oldest `take` (<= 32) queued candidates, append the ones with hits to the segment
  auto flush = [&](uint32_t take) {
    __syncwarp();  // the planes / case words below were stored by other lanes of this warp
    uint32_t m0 = 0, m1 = 0;
    int32_t h = 0, c = 0;
    if ((uint32_t)lane < take) {
      const uint32_t slot = (qhead + lane) & 63u;
      h = q_hap[warp][slot];
      // the case words and planes sit at the candidate's place in the slot space: their loads
      // do not wait for the haplotype's scan geometry (one round trip through L2 instead of two)
      const int64_t g = s + q_off[warp][slot];
      const HapScan H = A.hs[h];
      const uint32_t* vp = A.v + g;
      const uint32_t v0 = vp[-1], v1 = vp[0], v2 = vp[1];  // g >= 4: the slot space starts with a zero gap
      const Planes cur = B.q[g], nxt = B.q[g + 1];
      c = (int32_t)(g - H.chunk0);
      const bool use_v = !H.is_ref;
      const uint32_t w0 = v0 & K.prev_mask, w1 = v1, w2 = v2 & K.next_mask;
      if (!use_v || (w0 | w1 | w2) != 0) {
        uint32_t out[2], raw[2];
        scan_chunk_small_pre(K, H, c, w0, w1, w2, use_v, cur, nxt, out, raw);
        m0 = out[0];
        m1 = out[1];
      }
    }
    const uint32_t has = __ballot_sync(0xFFFFFFFFu, (m0 | m1) != 0);
    if ((m0 | m1) != 0) {
      const uint32_t at = n_ent + __popc(has & lanes_below);
      if (at < cap) seg[at] = make_uint4((uint32_t)h, (uint32_t)c, m0, m1);
    }
    n_ent += __popc(has);
    uint32_t k0 = __popc(m0), k1 = __popc(m1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      k0 += __shfl_xor_sync(0xFFFFFFFFu, k0, o);
      k1 += __shfl_xor_sync(0xFFFFFFFFu, k1, o);
    }
    n_h0 += k0;
    n_h1 += k1;
    qhead = (qhead + take) & 63u;
    qn -= take;
  };

  // text of iteration `it` (chunk 32 it + lane of the sub-range; -1 and n_iter are the halo) on its
  // way into ring slot (it + 1) mod FUSED_RING; lanes beyond the slot space get zeros. One commit
  // group per call whatever the lane copies, so "all but the newest FUSED_RING - 1 groups have
  // landed" means iteration `it` has.
  // running state of the ring: the next iteration to request (its text pointer, the chunks left
  // from its first chunk to the end of what this warp may read, its slot) and the slot to take
  const uint4* src_next = asc;
  int32_t left_next = n_avail;
  unsigned slot_next = ring_base, slot_take = ring_base;
  auto issue = [&]() {
    const bool ok = lane < left_next;
    const uint4* src = ok ? src_next : asc;
    const int bytes = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(slot_next), "l"(src), "r"(bytes) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(slot_next + 16u), "l"(src + 1), "r"(bytes) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    src_next += 64;
    left_next -= 32;
    slot_next = ring_base + ((slot_next - ring_base + 1024u) & (unsigned)(FUSED_RING * 1024 - 1));
  };
  auto take = [&](Raw8& r) {
    asm volatile("cp.async.wait_group %0;" ::"n"(FUSED_RING - 1) : "memory");
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]) : "r"(slot_take) : "memory");
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]) : "r"(slot_take + 16u) : "memory");
    slot_take = ring_base + ((slot_take - ring_base + 1024u) & (unsigned)(FUSED_RING * 1024 - 1));
  };
  // ... -> planes; returns the "case word non-zero" ballot
  auto pack = [&](int32_t it, const Raw8& r, Planes5& o) -> uint32_t {
    const PackedChunk k = pack_chunk_v3(r.w);
    o.a = k.a, o.c = k.c, o.g = k.g, o.t = k.t, o.v = k.v;
    if (k.invalid) {
      const int32_t cl = 32 * it + lane;
      if (cl >= 0 && cl < n) atomicMin(A.bad, (unsigned long long)((s + cl) * 32 + (__ffs(k.invalid) - 1)));
    }
    return __ballot_sync(0xFFFFFFFFu, k.v != 0);
  };

  // iterations [fast_lo, fast_hi) lie entirely inside the current (non-REF) haplotype's scan range
  // and territory and inside the sub-range: no boundary of any kind -- the common case by far
  int32_t fast_lo = 0, fast_hi = 0;
  auto set_fast_window = [&]() {
    int32_t hi = cur.c_hi < cur.t_next ? cur.c_hi : cur.t_next;
    if (n < hi) hi = n;
    fast_lo = cur.c_lo <= 0 ? 0 : (cur.c_lo + 31) >> 5;
    fast_hi = cur.is_ref ? fast_lo : (hi >> 5);
  };
  set_fast_window();

  // finalise iteration j: P are its planes, nzp / nzc / nzn the ballots of iterations j - 1, j, j + 1
  auto finalise = [&](int32_t j, const Planes5& P, uint32_t nzp, uint32_t nzc, uint32_t nzn) {
    const int32_t cl0 = 32 * j;
    if (lane == 0) nzs[j] = nzc;
    // variant chunk within one chunk (candidate) / within REACH chunks (planes are kept)
    const uint32_t near1 = nzc | __funnelshift_l(nzp, nzc, 1) | __funnelshift_r(nzc, nzn, 1);
    uint32_t keep_bits = near1;
    if (REACH >= 2) keep_bits |= __funnelshift_l(nzp, nzc, 2) | __funnelshift_r(nzc, nzn, 2);
    if (REACH >= 3) keep_bits |= __funnelshift_l(nzp, nzc, 3) | __funnelshift_r(nzc, nzn, 3);
    uint32_t cand_bits = near1;
    int32_t h_l = cur.h;
    if (j < fast_lo || j >= fast_hi) {  // a boundary somewhere in these 32 chunks
      while (cl0 >= cur.t_next) {  // uniform
        cur = load_cursor(cur.h + 1);
        set_fast_window();
      }
      h_l = cur.h;
      if (cl0 + 31 < cur.t_next) {
        // one haplotype's territory: uniform masks
        if (cur.is_ref) cand_bits = keep_bits = 0xFFFFFFFFu;
        cand_bits &= interval_mask(cur.c_lo, cur.c_hi, cl0);
      } else {
        // short haplotypes: the 32 chunks straddle territories, every lane finds its own
        const int64_t cg = s + cl0 + lane;
        while (h_l + 1 < A.n_hap && cg >= territory(A.slot_off, A.n_hap, A.n_chunks, h_l + 1)) ++h_l;
        const HapScan* H = A.hs + h_l;
        const int64_t crel = cg - H->chunk0;
        const int32_t a = H->a, b = H->b;
        const bool ref = H->is_ref != 0;
        const bool c = b > a && crel >= (a >> 5) && crel < ((b + 31) >> 5) && (ref || (near1 & lane_bit));
        cand_bits = __ballot_sync(0xFFFFFFFFu, c);
        keep_bits |= __ballot_sync(0xFFFFFFFFu, ref);
      }
      if (cl0 + 32 > n) {  // the last iteration of the slot space
        const uint32_t valid = (1u << (n - cl0)) - 1u;
        keep_bits &= valid;
        cand_bits &= valid;
      }
    }
    if (keep_bits) {  // uniform
      if (keep_bits & lane_bit) {
        qs[cl0] = make_uint4(P.a, P.c, P.g, P.t);
        vs[cl0] = P.v;
      }
    }
    // matches of older candidates: all their neighbour chunks have been stored by now
    if (qn >= 32) flush(32);
    if (cand_bits) {  // uniform
      if (cand_bits & lane_bit) {
        const uint32_t slot = (qhead + qn + __popc(cand_bits & lanes_below)) & 63u;
        q_off[warp][slot] = (uint16_t)(cl0 + lane);
        q_hap[warp][slot] = h_l;
      }
      qn += __popc(cand_bits);
    }
  };

  // the neighbour chunks this warp's matcher reads but another warp owns
  auto store_edge = [&](int32_t it, const Planes5& P, int which_lane) {
    if (lane == which_lane && 32 * it + lane < n_avail) {
      qs[32 * it] = make_uint4(P.a, P.c, P.g, P.t);
      vs[32 * it] = P.v;
    }
  };

  Planes5 prev{0u, 0u, 0u, 0u, 0u}, now;
  Raw8 raw;
  uint32_t nz_pp = 0, nz_p = 0, nz_c;  // ballots of iterations i - 2, i - 1, i
  // Software pipeline: the text of the next iteration is on its way into the warp's shared-memory
  // ring (cp.async, 16 bytes x 2 per lane and iteration: no registers held while in flight),
  // iteration i is packed, then iteration i - 1 (neighbours on both sides now known) is finalised.
  // Iteration -1 is the leading, iteration n_iter the trailing halo. Deeper rings (4, 8 KB per
  // warp) and requests of 2 - 8 KB at a time measured no better: profiles/README.md.
  if (s > 0) {  // start one iteration early: the leading halo
    src_next -= 64;
    left_next += 32;
  }
  for (int k = 0; k < FUSED_RING - 1; ++k) issue();
  if (s > 0) {
    issue();
    take(raw);
    nz_p = pack(-1, raw, now);
    store_edge(-1, now, 31);
  }
  for (int32_t i = 0; i <= n_iter; ++i) {
    issue();
    take(raw);
    nz_c = pack(i, raw, now);
    if (i == n_iter) store_edge(i, now, 0);
    if (i > 0) finalise(i - 1, prev, nz_pp, nz_p, nz_c);
    prev = now;
    nz_pp = nz_p;
    nz_p = nz_c;
  }
  while (qn > 0) flush(qn < 32 ? qn : 32);
  if (lane == 0) {
    A.cnt_ent[sub] = n_ent;
    A.cnt_hit0[sub] = n_h0;
    A.cnt_hit1[sub] = n_h1;
    if (n_ent > cap) atomicOr(A.overflow, 1u);
  }
}

// capacity of every sub-range's entry segment: all of its chunks when it overlaps a REF haplotype
// (every chunk of REF is a candidate) or when every haplotype is dense (unphased cohorts), else a
// quarter of them
__global__ void fused_caps_kernel(const int64_t* __restrict__ slot_off, const uint8_t* __restrict__ is_ref, int32_t n_hap,
                                  int64_t n_chunks, int64_t n_sub, int32_t sub_size, int32_t all_dense,
                                  uint32_t* __restrict__ cap) {
  const int64_t sub = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (sub >= n_sub) return;
  const int64_t s = sub * sub_size, e = s + sub_size < n_chunks ? s + sub_size : n_chunks;
  bool dense = all_dense != 0;
  if (!dense) {
    int32_t lo = 0, hi = n_hap;
    while (hi - lo > 1) {
      const int32_t mid = (lo + hi) >> 1;
      if (territory(slot_off, n_hap, n_chunks, mid) <= s) lo = mid; else hi = mid;
    }
    for (int32_t h = lo; h < n_hap && territory(slot_off, n_hap, n_chunks, h) < e; ++h)
      if (is_ref[h]) dense = true;
  }
  cap[sub] = dense ? (uint32_t)sub_size : (uint32_t)(sub_size / 4);
}

// entries -> (hap << 32 | pos) records at exact offsets; one warp per sub-range
__global__ void __launch_bounds__(128) fused_expand_kernel(const uint4* __restrict__ entries,
                                                           const uint64_t* __restrict__ seg_base,
                                                           const uint32_t* __restrict__ cnt_ent,
                                                           const uint64_t* __restrict__ base0,
                                                           const uint64_t* __restrict__ base1, int64_t n_sub,
                                                           uint64_t* __restrict__ hits0, uint64_t* __restrict__ hits1) {
  const int lane = threadIdx.x & 31;
  const int64_t sub = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (sub >= n_sub) return;
  const uint4* seg = entries + seg_base[sub];
  const uint32_t n = cnt_ent[sub];
  uint64_t run0 = base0[sub], run1 = base1[sub];
  // 128 entries per pass: the four loads of a lane are in flight together (the kernel is a chain
  // of dependent loads: one pass per 32 entries took 0.17 ms for 9.6 M entries, bound by latency)
  for (uint32_t k0 = 0; k0 < n; k0 += 128) {
    uint4 en[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t k = k0 + 32u * u + lane;
      en[u] = k < n ? seg[k] : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t pk = (uint32_t)__popc(en[u].z) | ((uint32_t)__popc(en[u].w) << 16);
      uint32_t incl = pk;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += y;
      }
      const uint32_t tot = __shfl_sync(0xFFFFFFFFu, incl, 31), excl = incl - pk;
      const uint64_t p0 = ((uint64_t)en[u].x << 32) | ((uint64_t)en[u].y << 5);
      uint64_t a = run0 + (excl & 0xFFFFu), b = run1 + (excl >> 16);
      uint32_t bits = en[u].z;
      while (bits) {
        const int jb = __ffs(bits) - 1;
        bits &= bits - 1;
        hits0[a++] = p0 + (uint32_t)jb;
      }
      bits = en[u].w;
      while (bits) {
        const int jb = __ffs(bits) - 1;
        bits &= bits - 1;
        hits1[b++] = p0 + (uint32_t)jb;
      }
      run0 += tot & 0xFFFFu;
      run1 += tot >> 16;
    }
  }
}

int32_t fused_sub_size(int64_t n_chunks) {
  static const int waves = getenv("HAWK_FUSED_WAVES") ? atoi(getenv("HAWK_FUSED_WAVES")) : 4;
  const int64_t want = n_chunks / ((int64_t)waves * 148 * 24);  // one warp per CTA, ~24 resident CTAs per SM (the sizes below were tuned with this figure)
  int32_t sub = FUSED_SUB_MAX;
  while (sub > 256 && sub > want) sub >>= 1;
  return sub;
}

int64_t fused_sub_ranges(int64_t n_chunks) {
  const int32_t sub = fused_sub_size(n_chunks);
  return (n_chunks + sub - 1) / sub;
}

int launch_fused_caps(cudaStream_t st, const int64_t* slot_off, const uint8_t* is_ref, int32_t n_hap, int64_t n_chunks,
                      int32_t all_dense, uint32_t* cap) {
  const int64_t n_sub = fused_sub_ranges(n_chunks);
  if (n_sub <= 0) return HAWK_OK;
  fused_caps_kernel<<<(unsigned)((n_sub + 127) / 128), 128, 0, st>>>(slot_off, is_ref, n_hap, n_chunks, n_sub,
                                                                    fused_sub_size(n_chunks), all_dense, cap);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "fused_caps_kernel launch");
}

int launch_fused_scan(cudaStream_t st, const FusedLaunch& L) {
  const int64_t n_sub = fused_sub_ranges(L.n_chunks);
  if (n_sub <= 0) return HAWK_OK;
  FusedArgs A{};
  A.ascii = (const uint4*)L.ascii;
  A.n_chunks = L.n_chunks;
  A.q = (uint4*)L.q;
  A.v = L.v;
  A.nz = L.nz;
  A.slot_off = L.slot_off;
  A.n_hap = L.n_hap;
  A.hs = L.hs;
  A.K = L.K;
  A.reach = L.reach;
  A.sub = fused_sub_size(L.n_chunks);
  A.seg_base = L.seg_base;
  A.seg_cap = L.seg_cap;
  A.entries = (uint4*)L.entries;
  A.cnt_ent = L.cnt_ent;
  A.cnt_hit0 = L.cnt_hit0;
  A.cnt_hit1 = L.cnt_hit1;
  A.overflow = L.overflow;
  A.bad = (unsigned long long*)L.bad;
  const unsigned blocks = (unsigned)((n_sub + FUSED_WARPS - 1) / FUSED_WARPS);
  const bool wide = ((uintptr_t)L.ascii & 31) == 0;
  // >= 2: a candidate chunk's matcher reads the case words two chunks from the variant chunk
  if (L.reach < 2 || L.reach > 3) return hawk_fail(HAWK_EINVAL, "fused scan: reach %d out of range", L.reach);
#define HAWK_FUSED_LAUNCH(W, R)                                                                   \
  do {                                                                                            \
    if (A.sub == FUSED_SUB_MAX) fused_scan_kernel<W, R, FUSED_SUB_MAX><<<blocks, FUSED_WARPS * 32, 0, st>>>(A); \
    else fused_scan_kernel<W, R, 0><<<blocks, FUSED_WARPS * 32, 0, st>>>(A);                      \
  } while (0)
  if (wide) {
    if (L.reach == 2) HAWK_FUSED_LAUNCH(true, 2); else HAWK_FUSED_LAUNCH(true, 3);
  } else {
    if (L.reach == 2) HAWK_FUSED_LAUNCH(false, 2); else HAWK_FUSED_LAUNCH(false, 3);
  }
#undef HAWK_FUSED_LAUNCH
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "fused_scan_kernel launch");
}

int launch_fused_expand(cudaStream_t st, const void* entries, const uint64_t* seg_base, const uint32_t* cnt_ent,
                        const uint64_t* base0, const uint64_t* base1, int64_t n_sub, uint64_t* hits0, uint64_t* hits1) {
  if (n_sub <= 0) return HAWK_OK;
  fused_expand_kernel<<<(unsigned)((n_sub + 3) / 4), 128, 0, st>>>((const uint4*)entries, seg_base, cnt_ent, base0, base1,
                                                                  n_sub, hits0, hits1);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "fused_expand_kernel launch");
}

}  // namespace hawk
