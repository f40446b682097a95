// scan2_kernels.cu -- K2, the PAM scan (search_guides.py:32-131 with the fused filters of
// :395-420 and :468-471), as a short pipeline of embarrassingly parallel kernels. Nothing
// here waits on another thread block, nothing is persistent, every output offset is an exact
// prefix sum, so the (hap << 32 | pos) record streams come out sorted by (haplotype,
// position) -- the reference's emission order -- without a sort and without capacity guesses.
//
//   hapscan      per haplotype: scan bounds, window bounds, first chunk (HapScan)
//   cand_count   one thread per 32-chunk slice (1,024 bp): candidate chunks of the slice.
//                Non-REF haplotypes: a chunk is a candidate iff it or a neighbouring chunk
//                holds a variant base (one bit per chunk in the nz summary plane; a guide
//                core reaches at most one chunk either side when G <= 32) -- this is the
//                "scan non-reference haplotypes only in windows overlapping their variants"
//                rule, at 1 bit per 32 bp. REF haplotypes / pam_search mode: every chunk.
//   match        one CTA per block of 256 slices: its candidates are compacted in shared memory
//                (chunk order) and matched 256 at a time, so every lane is busy although only a
//                few per cent of the chunks are candidates. Per candidate: the chunk's three
//                case words (masked to the guide's
//                reach), log-doubling sliding OR over that 96-bit window -> per-position "core
//                holds a variant" masks for both strands; two 128-bit plane loads; branch-free
//                AND-mask PAM test on both strands over shared funnel-shifted planes; interval
//                masks for the scan bounds and is_pamhit_in_range -> two 32-bit hit masks
//   expand       one CTA per slice block again: hit masks -> records at their exact offsets
// The per-block counts of cand_count / match are prefix-summed by the tile scan of
// post_kernels.cu; the host reads two totals (candidates, hits) to size the next stage.
#include <cuda_runtime.h>
#include <stdint.h>

#include "hawk_core.h"
#include "hawk_kernels.h"
#include "hawk_post.h"

#define CK_RET(expr)                \
  do {                              \
    int _rc = (expr);               \
    if (_rc != HAWK_OK) return _rc; \
  } while (0)

namespace hawk {

constexpr int S2_T = 256;  // threads per block everywhere in this file

// ---------------------------------------------------------------- per-haplotype scan geometry
__global__ void hapscan_kernel(BatchView B, ScanConst K, HapScan* __restrict__ hs) {
  const int32_t h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= B.n_hap) return;
  hs[h] = load_hap_scan(B, K, h);
}

// slice j of haplotype h covers chunks [cbase + 32 j, cbase + 32 j + 32), cbase = a >> 5.
// Block b handles S2_T consecutive slices of one haplotype; sblock_off[h] = first block of h;
// blk_tab[b] = {haplotype, first slice of the block} (one thread per block, binary search).
__global__ void sblock_table_kernel(const int64_t* __restrict__ sblock_off, int32_t n_hap, int64_t n_sblocks,
                                    int2* __restrict__ blk_tab) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_sblocks) return;
  int32_t lo = 0, hi = n_hap;  // sblock_off[lo] <= b < sblock_off[hi]
  while (hi - lo > 1) {
    const int32_t mid = (lo + hi) >> 1;
    if (sblock_off[mid] <= b) lo = mid; else hi = mid;
  }
  blk_tab[b] = make_int2(lo, (int32_t)(b - sblock_off[lo]) * S2_T);
}

int launch_hapscan(cudaStream_t st, const BatchView& B, const ScanConst& K, HapScan* hs) {
  if (B.n_hap <= 0) return HAWK_OK;
  hapscan_kernel<<<(B.n_hap + 127) / 128, 128, 0, st>>>(B, K, hs);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "hapscan_kernel launch");
}

struct SliceCtx {
  int32_t h;
  int32_t c32;   // first chunk of this thread's slice (haplotype-relative)
  bool live;     // slice intersects the scan interval
};

__device__ __forceinline__ SliceCtx slice_of_thread(const int2* __restrict__ blk_tab,
                                                    const HapScan* __restrict__ hs) {
  const int2 e = __ldg(&blk_tab[blockIdx.x]);
  SliceCtx x;
  x.h = e.x;
  const HapScan& H = hs[x.h];
  x.c32 = (H.a >> 5) + 32 * (e.y + (int32_t)threadIdx.x);
  x.live = H.b > H.a && x.c32 < ((H.b + 31) >> 5);
  return x;
}

// candidate chunks of one slice as a 32-bit mask
__device__ __forceinline__ uint32_t slice_candidates(const BatchView& B, const ScanConst& K, const HapScan& H,
                                                     int32_t c32) {
  const int32_t c_lo = H.a >> 5, c_end = (H.b + 31) >> 5;
  uint32_t cand;
  if (K.raw || H.is_ref) {
    cand = 0xFFFFFFFFu;
  } else {
    const int64_t bit = H.chunk0 + c32 - 1;  // >= 3: the slot space starts with a zero gap
    const uint32_t* wp = B.nz + (bit >> 5);
    const uint32_t sh = (uint32_t)(bit & 31);
    const uint32_t x0 = __ldg(wp), x1 = __ldg(wp + 1), x2 = __ldg(wp + 2);
    const uint32_t lo = funnel_r(x0, x1, sh), hi = funnel_r(x1, x2, sh);  // nz bits bit .. bit + 63
    const uint32_t mid = (lo >> 1) | (hi << 31);                          // chunks c32 .. c32 + 31
    if (K.small) {
      cand = mid | (mid << 1) | (lo & 1u) | (mid >> 1) | ((hi << 30) & 0x80000000u);
    } else {
      // long guides: a core reaches K.back chunks behind and K.ahead chunks ahead (<= 4 each);
      // 64-bit window of nz bits for chunks c32 - 4 .. c32 + 59
      const int64_t b4 = H.chunk0 + c32 - 4;  // >= 0: chunk0 >= 4
      const uint32_t* wq = B.nz + (b4 >> 5);
      const uint32_t s4 = (uint32_t)(b4 & 31);
      const uint32_t y0 = __ldg(wq), y1 = __ldg(wq + 1), y2 = __ldg(wq + 2);
      const uint64_t ww = ((uint64_t)funnel_r(y1, y2, s4) << 32) | funnel_r(y0, y1, s4);
      uint64_t cw = ww;
      for (int k = 1; k <= K.ahead; ++k) cw |= ww >> k;  // chunk c sees a variant in chunk c + k
      for (int k = 1; k <= K.back; ++k) cw |= ww << k;   // ... and in chunk c - k
      cand = (uint32_t)(cw >> 4);
    }
  }
  return cand & interval_mask(c_lo, c_end, c32);
}

__global__ void __launch_bounds__(S2_T) cand_count_kernel(BatchView B, ScanConst K, const HapScan* __restrict__ hs,
                                                          const int2* __restrict__ blk_tab,
                                                          uint32_t* __restrict__ slice_mask,
                                                          uint32_t* __restrict__ blk_cnt) {
  __shared__ uint32_t red[S2_T / 32];
  const SliceCtx x = slice_of_thread(blk_tab, hs);
  const uint32_t cand = x.live ? slice_candidates(B, K, hs[x.h], x.c32) : 0u;
  slice_mask[(int64_t)blockIdx.x * S2_T + threadIdx.x] = cand;
  uint32_t n = __popc(cand);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xFFFFFFFFu, n, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int k = 0; k < S2_T / 32; ++k) t += red[k];
    blk_cnt[blockIdx.x] = t;
  }
}

// Candidates of one slice block (256 slices = 8,192 chunks), in chunk order, as 16-bit chunk
// offsets from the block's first chunk in shared memory. Returns their number.
__device__ __forceinline__ uint32_t compact_block_candidates(uint32_t cand, uint16_t* list, uint32_t* wsum) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t mine = __popc(cand);
  uint32_t incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= d) incl += y;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  uint32_t wbase = 0, total = 0;
#pragma unroll
  for (int k = 0; k < S2_T / 32; ++k) {
    const uint32_t t = wsum[k];
    if (k < warp) wbase += t;
    total += t;
  }
  uint32_t p = wbase + (incl - mine);
  const uint32_t off = 32u * threadIdx.x;
  while (cand) {
    const int b = __ffs(cand) - 1;
    cand &= cand - 1;
    list[p++] = (uint16_t)(off + b);
  }
  __syncthreads();
  return total;
}

// ---------------------------------------------------------------- match
// One CTA per slice block: its candidates are compacted in shared memory, then matched 256 at
// a time (every lane busy). Hit masks go to masks[cand_base[block] + k]; per-block hit counts.
__global__ void __launch_bounds__(S2_T, 6) match_kernel(BatchView B, ScanConst K, const HapScan* __restrict__ hs,
                                                        const int2* __restrict__ blk_tab,
                                                     const uint32_t* __restrict__ slice_mask,
                                                     const uint64_t* __restrict__ cand_base, uint2* __restrict__ masks,
                                                     uint32_t* __restrict__ blk_hits0, uint32_t* __restrict__ blk_hits1,
                                                     unsigned long long* __restrict__ raw_tot) {
  __shared__ uint16_t list[S2_T * 32];
  __shared__ uint32_t wsum[S2_T / 32];
  __shared__ uint32_t red[3][S2_T / 32];
  __shared__ HapScan sH;  // block-uniform: kept out of the registers
  const int2 e = __ldg(&blk_tab[blockIdx.x]);
  if (threadIdx.x == 0) sH = hs[e.x];
  const uint32_t total = compact_block_candidates(slice_mask[(int64_t)blockIdx.x * S2_T + threadIdx.x], list, wsum);
  const HapScan& H = sH;  // published by the barriers inside the compaction
  const int32_t c_block = (H.a >> 5) + 32 * e.y;  // first chunk of the block's first slice
  uint2* const out_masks = masks + cand_base[blockIdx.x];
  const bool use_v = !(K.raw || H.is_ref);
  uint32_t n0 = 0, n1 = 0, nr = 0;
  for (uint32_t k = threadIdx.x; k < total; k += S2_T) {
    const int32_t c = c_block + list[k];
    uint32_t out[2] = {0, 0}, raw[2] = {0, 0};
    if (K.small) {
      uint32_t w0 = 0, w1 = 0, w2 = 0;
      bool go = true;
      if (use_v) {
        // the slot layout's zero gap makes c - 1 / c + 1 safe at the haplotype's ends
        const uint32_t* vp = B.v + H.chunk0 + c;
        w0 = __ldg(vp - 1) & K.prev_mask;
        w1 = __ldg(vp);
        w2 = __ldg(vp + 1) & K.next_mask;
        go = (w0 | w1 | w2) != 0;
      }
      if (go) scan_chunk_small(B, K, H, c, w0, w1, w2, use_v, out, raw);
    } else {
      scan_chunk(B, K, H, (int64_t)c, out, raw);
    }
    out_masks[k] = make_uint2(out[0], out[1]);
    n0 += __popc(out[0]);
    n1 += __popc(out[1]);
    nr += (uint32_t)__popc(raw[0]) | ((uint32_t)__popc(raw[1]) << 16);
  }
  // raw counts: <= 32 per chunk and strand, <= 32 rounds of a dense block: fits 16 bits per lane
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n0 += __shfl_xor_sync(0xFFFFFFFFu, n0, o);
    n1 += __shfl_xor_sync(0xFFFFFFFFu, n1, o);
  }
  uint32_t r0 = nr & 0xFFFFu, r1 = nr >> 16;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    r0 += __shfl_xor_sync(0xFFFFFFFFu, r0, o);
    r1 += __shfl_xor_sync(0xFFFFFFFFu, r1, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = n0;
    red[1][threadIdx.x >> 5] = n1;
    red[2][threadIdx.x >> 5] = r0;
    wsum[threadIdx.x >> 5] = r1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t0 = 0, t1 = 0, s0 = 0, s1 = 0;
    for (int k = 0; k < S2_T / 32; ++k) {
      t0 += red[0][k];
      t1 += red[1][k];
      s0 += red[2][k];
      s1 += wsum[k];
    }
    blk_hits0[blockIdx.x] = t0;
    blk_hits1[blockIdx.x] = t1;
    if (K.raw && (s0 | s1)) {
      atomicAdd(&raw_tot[0], (unsigned long long)s0);
      atomicAdd(&raw_tot[1], (unsigned long long)s1);
    }
  }
}

// ---------------------------------------------------------------- expand
// One CTA per slice block again: same compaction, then the block's hit masks become records at
// hit_base[block] + running offset, 256 candidates per pass (warp scan + carry across passes).
__global__ void __launch_bounds__(S2_T) expand_kernel(const HapScan* __restrict__ hs, const int2* __restrict__ blk_tab,
                                                      const uint32_t* __restrict__ slice_mask,
                                                      const uint64_t* __restrict__ cand_base,
                                                      const uint2* __restrict__ masks, const uint64_t* __restrict__ base0,
                                                      const uint64_t* __restrict__ base1, uint64_t* __restrict__ hits0,
                                                      uint64_t* __restrict__ hits1) {
  __shared__ uint16_t list[S2_T * 32];
  __shared__ uint32_t wsum[S2_T / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int2 e = __ldg(&blk_tab[blockIdx.x]);
  const int32_t c_block = (hs[e.x].a >> 5) + 32 * e.y;
  const uint32_t total = compact_block_candidates(slice_mask[(int64_t)blockIdx.x * S2_T + threadIdx.x], list, wsum);
  const uint2* const in_masks = masks + cand_base[blockIdx.x];
  const uint64_t hkey = (uint64_t)(uint32_t)e.x << 32;
  uint64_t run0 = base0[blockIdx.x], run1 = base1[blockIdx.x];  // uniform over the CTA
  for (uint32_t k0 = 0; k0 < total; k0 += S2_T) {
    const uint32_t k = k0 + threadIdx.x;
    uint2 m = make_uint2(0u, 0u);
    uint64_t p0 = 0;
    if (k < total) {
      m = in_masks[k];
      p0 = hkey | ((uint64_t)(uint32_t)(c_block + list[k]) << 5);
    }
    const uint32_t pk = (uint32_t)__popc(m.x) | ((uint32_t)__popc(m.y) << 16);  // <= 32 each
    uint32_t incl = pk;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= d) incl += y;
    }
    __syncthreads();  // wsum is reused from the previous pass
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t wb = 0, tot = 0;
#pragma unroll
    for (int j = 0; j < S2_T / 32; ++j) {
      const uint32_t t = wsum[j];
      if (j < warp) wb += t;
      tot += t;
    }
    const uint32_t excl = incl - pk;
    uint64_t a = run0 + (wb & 0xFFFFu) + (excl & 0xFFFFu), b = run1 + (wb >> 16) + (excl >> 16);
    uint32_t bits = m.x;
    while (bits) {
      const int j = __ffs(bits) - 1;
      bits &= bits - 1;
      hits0[a++] = p0 + (uint32_t)j;
    }
    bits = m.y;
    while (bits) {
      const int j = __ffs(bits) - 1;
      bits &= bits - 1;
      hits1[b++] = p0 + (uint32_t)j;
    }
    run0 += tot & 0xFFFFu;
    run1 += tot >> 16;
  }
}

}  // namespace hawk

// ------------------------------------------------------------------ device-layer entry points
using namespace hawk;

static BatchView view_of(const void* d_q, const uint32_t* d_v, const uint32_t* d_nz, const int64_t* d_slot_off,
                         const int32_t* d_len, const int32_t* d_scan_start, const int32_t* d_scan_stop,
                         const uint8_t* d_is_ref, int32_t n_hap) {
  BatchView B{};
  B.q = (const Planes*)d_q;
  B.v = d_v;
  B.nz = d_nz;
  B.slot_off = d_slot_off;
  B.len = d_len;
  B.scan_start = d_scan_start;
  B.scan_stop = d_scan_stop;
  B.is_ref = d_is_ref;
  B.n_hap = n_hap;
  return B;
}

static int check_params(const hawk_params* params) {
  if (!params || params->pam_len < 1 || params->pam_len > HAWK_MAX_PAM || params->guide_len < 1)
    return hawk_fail(HAWK_EINVAL, "scan: bad PAM / guide length");
  if (params->pam_len + params->guide_len + 2 * HAWK_GUIDESEQPAD > HAWK_MAX_WINDOW)
    return hawk_fail(HAWK_EINVAL, "scan: guide + PAM window exceeds HAWK_MAX_WINDOW");
  return HAWK_OK;
}

extern "C" int64_t hawk_scan_plan(const int32_t* scan_start, const int32_t* scan_stop, int32_t n_hap,
                                  int64_t* sblock_off) {
  int64_t total = 0;
  for (int32_t h = 0; h < n_hap; ++h) {
    if (sblock_off) sblock_off[h] = total;
    const int64_t a = scan_start[h] < 0 ? 0 : scan_start[h], b = scan_stop[h];
    if (b > a) {
      const int64_t slices = ((((b + 31) >> 5) - (a >> 5)) + 31) >> 5;
      total += (slices + S2_T - 1) / S2_T;
    }
  }
  if (sblock_off) sblock_off[n_hap] = total;
  return total;
}

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// stage-1 workspace: totals | HapScan per haplotype | candidate block bases + counts | tile sums
struct Scan2Ws {
  uint64_t* totals;  // [0] candidates, [1..2] hits per strand, [3..4] raw hits
  HapScan* hs;
  uint64_t* cand_base;
  uint32_t* cand_cnt;
  int2* blk_tab;
  uint32_t* slice_mask;
  uint64_t* tile_sums;
  size_t bytes;
};

static Scan2Ws ws_layout(void* base, int32_t n_hap, int64_t n_sblocks) {
  Scan2Ws w;
  char* p = (char*)base;
  size_t off = 0;
  w.totals = (uint64_t*)(p + off);
  off += 256;
  w.hs = (HapScan*)(p + off);
  off += al256((size_t)(n_hap > 0 ? n_hap : 1) * sizeof(HapScan));
  w.cand_base = (uint64_t*)(p + off);
  off += al256((size_t)(n_sblocks + 1) * 8);
  w.cand_cnt = (uint32_t*)(p + off);
  off += al256((size_t)(n_sblocks + 1) * 4);
  w.blk_tab = (int2*)(p + off);
  off += al256((size_t)(n_sblocks + 1) * 8);
  w.slice_mask = (uint32_t*)(p + off);
  off += al256((size_t)(n_sblocks + 1) * S2_T * 4);
  w.tile_sums = (uint64_t*)(p + off);
  off += al256(((size_t)scan_tiles(n_sblocks) + 2) * 8);
  w.bytes = off;
  return w;
}

// stage-2/3 workspace: hit block bases + counts per strand | tile sums
struct MatchWs {
  uint64_t* hit_base[2];
  uint32_t* hit_cnt[2];
  uint64_t* tile_sums;
  size_t bytes;
};

static MatchWs match_ws_layout(void* base, int64_t n_sblocks) {
  const int64_t nb = n_sblocks + 1;
  MatchWs w;
  char* p = (char*)base;
  size_t off = 0;
  for (int s = 0; s < 2; ++s) {
    w.hit_base[s] = (uint64_t*)(p + off);
    off += al256((size_t)nb * 8);
  }
  for (int s = 0; s < 2; ++s) {
    w.hit_cnt[s] = (uint32_t*)(p + off);
    off += al256((size_t)nb * 4);
  }
  w.tile_sums = (uint64_t*)(p + off);
  off += al256(((size_t)scan_tiles(nb) + 2) * 8);
  w.bytes = off;
  return w;
}

extern "C" size_t hawk_scan_workspace_bytes(int32_t n_hap, int64_t n_sblocks) {
  return ws_layout(nullptr, n_hap, n_sblocks).bytes;
}
extern "C" size_t hawk_scan_match_workspace_bytes(int64_t n_sblocks) { return match_ws_layout(nullptr, n_sblocks).bytes; }

// stage 1: candidate chunks. d_totals-style results land in the workspace head: after a stream
// sync the caller reads uint64 totals[0] = number of candidates (hawk_scan_totals).
extern "C" int hawk_scan_count_dev(void* stream, const void* d_q, const uint32_t* d_v, const uint32_t* d_nz,
                                   const int64_t* d_slot_off, const int32_t* d_len, const int32_t* d_scan_start,
                                   const int32_t* d_scan_stop, const uint8_t* d_is_ref, const int64_t* d_sblock_off,
                                   int32_t n_hap, int64_t n_sblocks, const hawk_params* params, int32_t raw_hits,
                                   void* d_workspace) {
  CK_RET(check_params(params));
  cudaStream_t st = (cudaStream_t)stream;
  const Scan2Ws W = ws_layout(d_workspace, n_hap, n_sblocks);
  cudaError_t e = cudaMemsetAsync(W.totals, 0, 256, st);
  if (e != cudaSuccess) return hawk_check_cuda(e, "totals memset");
  if (n_hap <= 0 || n_sblocks <= 0) return HAWK_OK;
  const BatchView B = view_of(d_q, d_v, d_nz, d_slot_off, d_len, d_scan_start, d_scan_stop, d_is_ref, n_hap);
  const ScanConst K = make_scan_const(*params, raw_hits);
  hawk_prof_begin(st, 1);
  hapscan_kernel<<<(n_hap + 127) / 128, 128, 0, st>>>(B, K, W.hs);
  hawk_note_launch(1);
  sblock_table_kernel<<<(unsigned)((n_sblocks + 255) / 256), 256, 0, st>>>(d_sblock_off, n_hap, n_sblocks, W.blk_tab);
  hawk_note_launch(1);
  cand_count_kernel<<<(unsigned)n_sblocks, S2_T, 0, st>>>(B, K, W.hs, W.blk_tab, W.slice_mask, W.cand_cnt);
  hawk_note_launch(1);
  CK_RET(exclusive_scan_u32(st, W.cand_cnt, n_sblocks, W.cand_base, W.tile_sums, W.totals + 0));
  hawk_prof_end(st);
  return hawk_check_cuda(cudaGetLastError(), "scan stage 1 launch");
}

// stage 2: PAM match of the candidates (d_masks: n_cand x 8 bytes). totals[1..2] = hits per
// strand, [3..4] raw hits.
extern "C" int hawk_scan_match_dev(void* stream, const void* d_q, const uint32_t* d_v, const uint32_t* d_nz,
                                   const int64_t* d_slot_off, const int32_t* d_len, const int32_t* d_scan_start,
                                   const int32_t* d_scan_stop, const uint8_t* d_is_ref, int32_t n_hap,
                                   int64_t n_sblocks, const hawk_params* params, int32_t raw_hits, int64_t n_cand,
                                   uint64_t* d_masks, void* d_workspace, void* d_match_workspace) {
  CK_RET(check_params(params));
  if (n_cand <= 0 || n_hap <= 0 || n_sblocks <= 0) return HAWK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const Scan2Ws W = ws_layout(d_workspace, n_hap, n_sblocks);
  const MatchWs M = match_ws_layout(d_match_workspace, n_sblocks);
  const BatchView B = view_of(d_q, d_v, d_nz, d_slot_off, d_len, d_scan_start, d_scan_stop, d_is_ref, n_hap);
  const ScanConst K = make_scan_const(*params, raw_hits);
  hawk_prof_begin(st, 4);
  match_kernel<<<(unsigned)n_sblocks, S2_T, 0, st>>>(B, K, W.hs, W.blk_tab, W.slice_mask, W.cand_base, (uint2*)d_masks,
                                                    M.hit_cnt[0], M.hit_cnt[1], (unsigned long long*)(W.totals + 3));
  hawk_note_launch(1);
  hawk_prof_end(st);
  hawk_prof_begin(st, 1);
  CK_RET(exclusive_scan_u32(st, M.hit_cnt[0], n_sblocks, M.hit_base[0], M.tile_sums, W.totals + 1));
  CK_RET(exclusive_scan_u32(st, M.hit_cnt[1], n_sblocks, M.hit_base[1], M.tile_sums, W.totals + 2));
  hawk_prof_end(st);
  return hawk_check_cuda(cudaGetLastError(), "scan stage 2 launch");
}

// stage 3: records. d_hits_* must hold totals[1] / totals[2] records.
extern "C" int hawk_scan_expand_dev(void* stream, int32_t n_hap, int64_t n_sblocks, int64_t n_cand,
                                    const uint64_t* d_masks, void* d_workspace, void* d_match_workspace,
                                    uint64_t* d_hits_fwd, uint64_t* d_hits_rev) {
  if (n_cand <= 0 || n_sblocks <= 0) return HAWK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const Scan2Ws W = ws_layout(d_workspace, n_hap, n_sblocks);
  const MatchWs M = match_ws_layout(d_match_workspace, n_sblocks);
  hawk_prof_begin(st, 3);
  expand_kernel<<<(unsigned)n_sblocks, S2_T, 0, st>>>(W.hs, W.blk_tab, W.slice_mask, W.cand_base, (const uint2*)d_masks,
                                                     M.hit_base[0], M.hit_base[1], d_hits_fwd, d_hits_rev);
  hawk_note_launch(1);
  hawk_prof_end(st);
  return hawk_check_cuda(cudaGetLastError(), "expand_kernel launch");
}

// device address of the uint64 totals[8] inside a workspace (copy them out after a sync)
extern "C" const uint64_t* hawk_scan_totals(const void* d_workspace) { return (const uint64_t*)d_workspace; }
