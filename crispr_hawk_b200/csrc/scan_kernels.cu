// scan_kernels.cu -- K1 (pack) and K2 (PAM scan + fused filters + ordered
// compaction) for sm_100a. Integer/bitwise, HBM-bound: no tensor cores.
//
// K2 structure (one CTA per *span* of SPAN_CHUNKS chunks of one haplotype,
// spans handed out by an atomic ticket so that a span's predecessors are always
// resident or finished):
//   phase 1  every thread evaluates chunks (32 positions each): REF haplotypes
//            load the planes unconditionally, non-REF haplotypes first look at
//            the case plane and touch the planes only where a variant base is in
//            reach of a guide core (search_guides.py:468-471) -- the "scan only the
//            windows overlapping variants" rule, driven by a 0.125 B/bp stream;
//   phase 2  per-strand hit bitmaps live in shared memory (fixed size, cannot
//            overflow); popcounts are block-scanned;
//   phase 3  decoupled look-back over a per-span status word gives the span's
//            global output offset, so the record stream comes out sorted by
//            (haplotype, position) without a sort pass;
//   phase 4  bitmaps are expanded to (hap << 32 | pos) records.
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "hawk_core.h"
#include "hawk_kernels.h"

namespace hawk {

// ------------------------------------------------------------------ K1: pack
// One thread per chunk: 32 ASCII bytes -> {A,C,G,T} plane words + case word.
__global__ void __launch_bounds__(256) pack_kernel(const uint4* __restrict__ ascii,
                                                   int64_t n_chunks, uint4* __restrict__ q,
                                                   uint32_t* __restrict__ v,
                                                   unsigned long long* __restrict__ bad) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += stride) {
    uint4 w[2];
    w[0] = __ldg(&ascii[2 * c]);
    w[1] = __ldg(&ascii[2 * c + 1]);
    const PackedChunk o = pack_chunk(reinterpret_cast<const uint32_t*>(w));
    q[c] = make_uint4(o.a, o.c, o.g, o.t);
    v[c] = o.v;
    if (o.invalid) atomicMin(bad, (unsigned long long)(c * 32 + (__ffs(o.invalid) - 1)));
  }
}

// ------------------------------------------------------------------ K2: scan
// Warp-autonomous, persistent. A *span* is HAWK_SPAN_CHUNKS = 256 chunks (8,192 base
// slots, 1 KB of case words) of one haplotype; spans are numbered in (haplotype, position)
// order and every warp ("unit") owns a contiguous span range [unit_span[u], unit_span[u+1])
// balanced on the host (hawk_scan_plan). No warp ever waits for another one:
//   tiles     each warp keeps a three-stage ring of case-word tiles in shared memory (1 KB +
//             a 4-word halo either side -- the slot layout keeps the halo zero), filled by
//             bulk-async copies (TMA) it issues itself three spans ahead, completion through
//             an mbarrier per stage;
//   phase A   128-bit reads of the tile find the chunks with a variant base in reach of a
//             guide core (search_guides.py:468-471; REF haplotypes / pam_search mode take every
//             chunk); candidates are compacted across the warp;
//   phase B   32 candidates at a time load the {A,C,G,T} planes and run the branch-free
//             AND-mask PAM test on both strands + the fused filters; a warp prefix sum of the
//             hit counts places the (hap << 32 | pos) records directly into the warp's private
//             segment of a staging buffer.
// seg_prefix_kernel + compact_kernel then concatenate the segments (exclusive prefix over the
// per-unit totals), which yields the stream sorted by (haplotype, position) without a sort.
constexpr int SCAN_WARPS = 8;
constexpr int SCAN_THREADS = SCAN_WARPS * 32;
constexpr int SCAN_CTAS_PER_SM = 4;
constexpr int SPAN = HAWK_SPAN_CHUNKS;
constexpr int SPAN_ROUNDS = SPAN / 128;
constexpr int HALO = 4;  // case words either side of a span (HAWK_SLOT_GAP / 32)
constexpr int VBUF_WORDS = SPAN + 2 * HALO;
constexpr int STAGES = 2;
constexpr int QCAP = 160;  // candidate ring entries per warp (<= 31 carried + 128 per round)
__device__ __forceinline__ uint32_t qwrap(uint32_t x) { return x >= QCAP ? x - QCAP : x; }  // x < 2 QCAP
static_assert(SPAN == 256, "candidate indices are stored as uint8");
static_assert(HALO * 32 == HAWK_SLOT_GAP, "halo must be covered by the layout's zero gap");
static_assert((VBUF_WORDS * 4) % 16 == 0, "tiles must keep 16-byte alignment");

__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
  return x;
}
__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
  return x;
}

// ---- mbarrier / bulk-copy wrappers (PTX ISA 8.x, sm_90+) ----
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}

struct ScanArgs {
  BatchView B;
  ScanConst K;
  const int2* span_tab;         // per span: {haplotype, first chunk}
  const int64_t* unit_span;     // n_units + 1: span range of every warp
  const double* unit_frac;      // n_units + 1: cumulative share of the output capacity
  const uint64_t* seg_prev;     // exact retry: per-unit totals of the previous launch [2][n_units], else null
  const uint64_t* seg_prev_off; // exact retry: their exclusive prefix
  uint64_t* seg_count;          // [2][n_units] per-unit totals of this launch
  uint64_t* stage[2];           // staging buffers, cap[s] records each
  int64_t cap[2];
  int32_t n_units;
  uint64_t* counts;             // [2..3] raw totals (atomicAdd), [4] overflow flag
};

// span -> {haplotype, first chunk}; one thread per span. Thread 0 also clears the counters.
__global__ void span_table_kernel(const int64_t* __restrict__ span_off, const int32_t* __restrict__ scan_start,
                                  int32_t n_hap, int64_t n_spans, int2* __restrict__ tab,
                                  uint64_t* __restrict__ counts) {
  int64_t sp = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (sp == 0)
    for (int k = 0; k < 8; ++k) counts[k] = 0;
  if (sp >= n_spans) return;
  int32_t lo = 0, hi = n_hap;  // span_off[lo] <= sp < span_off[hi]
  while (hi - lo > 1) {
    int32_t mid = (lo + hi) >> 1;
    if (span_off[mid] <= sp) lo = mid; else hi = mid;
  }
  int32_t a = scan_start[lo] < 0 ? 0 : scan_start[lo];
  tab[sp] = make_int2(lo, ((a >> 5) & ~3) + (int32_t)(sp - span_off[lo]) * SPAN);
}

__global__ void __launch_bounds__(SCAN_THREADS, SCAN_CTAS_PER_SM) scan_kernel(const __grid_constant__ ScanArgs A) {
  __shared__ __align__(16) uint32_t vbuf[SCAN_WARPS][STAGES][VBUF_WORDS];
  __shared__ __align__(16) uint4 queue[SCAN_WARPS][QCAP];  // candidate ring: {w(c-1), w(c), w(c+1), c}
  __shared__ __align__(8) uint64_t full_bar[SCAN_WARPS][STAGES];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int unit = blockIdx.x * SCAN_WARPS + warp;
  if (unit >= A.n_units) return;
  const int64_t sp0 = A.unit_span[unit], sp1 = A.unit_span[unit + 1];

  if (lane == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[warp][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  // this warp's segment of the staging buffers
  uint64_t* seg_dst[2];
  uint32_t seg_cap[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    uint64_t off, cap;
    if (A.seg_prev) {  // exact retry: segments sized by the previous launch's per-unit totals
      off = A.seg_prev_off[s * A.n_units + unit];
      cap = A.seg_prev[s * A.n_units + unit];
    } else {
      off = (uint64_t)(A.unit_frac[unit] * (double)A.cap[s]);
      const uint64_t hi = unit + 1 == A.n_units ? (uint64_t)A.cap[s] : (uint64_t)(A.unit_frac[unit + 1] * (double)A.cap[s]);
      cap = hi - off;
    }
    seg_dst[s] = A.stage[s] + off;
    seg_cap[s] = cap > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)cap;
  }

  // tile producer state (lane 0): haplotype of the last issued span
  int32_t iss_hap = -1, iss_nch4 = 0;
  int64_t iss_chunk0 = 0;
  auto issue = [&](int64_t span, int st) {
    const int2 e = __ldg(&A.span_tab[span]);
    if (e.x != iss_hap) {
      iss_hap = e.x;
      iss_chunk0 = A.B.slot_off[e.x] >> 5;
      iss_nch4 = ((A.B.len[e.x] + 127) >> 7) << 2;
    }
    // case words [c_first - HALO, min(c_first + SPAN + HALO, padded end + HALO))
    int32_t w_end = e.y + SPAN + HALO;
    if (w_end > iss_nch4 + HALO) w_end = iss_nch4 + HALO;
    const uint32_t bytes = (uint32_t)(w_end - (e.y - HALO)) * 4u;
    mbar_arrive_expect_tx(&full_bar[warp][st], bytes);
    bulk_g2s(&vbuf[warp][st][0], A.B.v + iss_chunk0 + (e.y - HALO), bytes, &full_bar[warp][st]);
  };
  if (lane == 0)
    for (int k = 0; k < STAGES && sp0 + k < sp1; ++k) issue(sp0 + k, k);

  uint32_t raw_acc[2] = {0, 0};
  uint32_t run[2] = {0, 0};  // records this warp has produced so far
  int32_t cur_hap = -1;
  HapScan H;
  H.is_ref = 0;
  uint32_t q_head = 0, q_n = 0;  // ring state (warp-uniform)

  // hit bits of one chunk per lane -> records in the warp's segment, chunk order = lane order
  auto emit = [&](const uint32_t out[2], int32_t c) {
    const uint32_t pk = (uint32_t)__popc(out[0]) | ((uint32_t)__popc(out[1]) << 16);
    if (!__any_sync(0xFFFFFFFFu, pk != 0)) return;
    uint32_t incl = pk;  // both strands' counts in one word (<= 1024 each)
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= d) incl += y;
    }
    const uint32_t tot = __shfl_sync(0xFFFFFFFFu, incl, 31), excl = incl - pk;
    const uint64_t p0 = ((uint64_t)(uint32_t)cur_hap << 32) | ((uint64_t)(uint32_t)c << 5);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      uint32_t bits = out[s];
      uint32_t p = run[s] + ((excl >> (16 * s)) & 0xFFFFu);
      const uint32_t t = (tot >> (16 * s)) & 0xFFFFu;
      if (run[s] + t <= seg_cap[s]) {  // warp-uniform: the whole batch fits
        while (bits) {
          const int b = __ffs(bits) - 1;
          bits &= bits - 1;
          seg_dst[s][p++] = p0 + (uint32_t)b;
        }
      } else {
        while (bits) {
          const int b = __ffs(bits) - 1;
          bits &= bits - 1;
          if (p < seg_cap[s]) seg_dst[s][p] = p0 + (uint32_t)b;
          ++p;
        }
      }
      run[s] += t;
    }
  };

  // match + filters for up to 32 queued candidates (n <= 32 taken from the ring head)
  auto drain = [&](uint32_t n, const uint32_t* vs, int32_t c_first) {
    uint32_t out[2] = {0, 0}, raw[2] = {0, 0};
    int32_t c = 0;
    if ((uint32_t)lane < n) {
      const uint4 e = queue[warp][qwrap(q_head + lane)];
      c = (int32_t)e.w;
      if (A.K.small) {
        scan_chunk_small(A.B, A.K, H, c, e.x, e.y, e.z, true, out, raw);
      } else {  // long guides: the candidate's span is still resident (queue is flushed per span)
        auto vword = [&](int64_t w) -> uint32_t {
          const int32_t j = (int32_t)w - c_first + HALO;
          return (w < 0 || w >= H.nchunks || j < 0 || j >= VBUF_WORDS) ? 0u : vs[j];
        };
        scan_chunk(A.B, A.K, H, (int64_t)c, vword, out, raw);
      }
    }
    q_head = qwrap(q_head + n);
    q_n -= n;
    emit(out, c);
  };

  int it = 0;
  for (int64_t span = sp0; span < sp1; ++span, ++it) {
    const int st = it % STAGES;
    const int2 e = __ldg(&A.span_tab[span]);
    const uint32_t* vs = vbuf[warp][st];
    if (e.x != cur_hap) {
      if (q_n) drain(q_n, vs, 0);  // K.small only: long-guide queues are empty between spans
      cur_hap = e.x;
      H = load_hap_scan(A.B, A.K, e.x);
    }
    mbar_wait(&full_bar[warp][st], (it / STAGES) & 1);
    const int32_t c_first = e.y;
    const int32_t c_lo = H.a >> 5, c_end = (H.b + 31) >> 5;

    if (A.K.raw || H.is_ref) {
      // ---- dense: every chunk of the scan interval is matched, 32 chunks per pass
      for (int32_t base = c_first; base < c_first + SPAN && base < c_end; base += 32) {
        const int32_t c = base + lane;
        uint32_t out[2] = {0, 0}, raw[2] = {0, 0};
        if (c >= c_lo && c < c_end) scan_chunk_small(A.B, A.K, H, c, 0u, 0u, 0u, false, out, raw);
        raw_acc[0] += __popc(raw[0]);
        raw_acc[1] += __popc(raw[1]);
        emit(out, c);
      }
    } else {
      // ---- sparse: chunks with a variant base in reach of a guide core go to the ring
#pragma unroll
      for (int r = 0; r < SPAN_ROUNDS; ++r) {
        const int32_t base = c_first + r * 128;
        if (base >= c_end) break;  // warp-uniform
        const int32_t c4 = base + 4 * lane;  // first of this lane's 4 chunks
        const uint4* t4 = reinterpret_cast<const uint4*>(vs + (c4 - c_first));  // words c4 - HALO ..
        const uint4 L = t4[0], M = t4[1], R = t4[2];
        uint32_t cm;
        if (A.K.small) {
          const uint32_t pm = A.K.prev_mask, nm = A.K.next_mask;
          cm = (((L.w & pm) | M.x | (M.y & nm)) ? 1u : 0u) | (((M.x & pm) | M.y | (M.z & nm)) ? 2u : 0u) |
               (((M.y & pm) | M.z | (M.w & nm)) ? 4u : 0u) | (((M.z & pm) | M.w | (R.x & nm)) ? 8u : 0u);
        } else {
          const uint32_t w[12] = {L.x, L.y, L.z, L.w, M.x, M.y, M.z, M.w, R.x, R.y, R.z, R.w};
          cm = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint32_t any = 0;
#pragma unroll
            for (int d = -4; d <= 4; ++d)
              if (d >= -A.K.back && d <= A.K.ahead) any |= w[4 + k + d];
            cm |= any ? (1u << k) : 0u;
          }
        }
        if (base < c_lo || base + 128 > c_end) {  // warp-uniform: round straddles the scan interval
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (c4 + k < c_lo || c4 + k >= c_end) cm &= ~(1u << k);
        }
        if (!__any_sync(0xFFFFFFFFu, cm != 0)) continue;
        // append this lane's candidates behind the earlier lanes' (chunk order)
        const uint32_t mine = __popc(cm);
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
          if (lane >= d) incl += y;
        }
        uint32_t idx = qwrap(qwrap(q_head + q_n) + (incl - mine));
        const uint32_t wv[6] = {L.w, M.x, M.y, M.z, M.w, R.x};
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (cm & (1u << k)) {
            queue[warp][idx] = make_uint4(wv[k], wv[k + 1], wv[k + 2], (uint32_t)(c4 + k));
            idx = qwrap(idx + 1);
          }
        q_n += __shfl_sync(0xFFFFFFFFu, incl, 31);
        __syncwarp();
        while (q_n >= 32) drain(32, vs, c_first);
      }
      if (!A.K.small && q_n) drain(q_n, vs, c_first);  // <= 31 left: the tile goes away with the span
    }
    __syncwarp();
    // refill this stage with the span STAGES ahead (generic reads above are ordered before the
    // async-proxy write by the fence)
    if (lane == 0 && span + STAGES < sp1) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(span + STAGES, st);
    }
  }
  if (q_n) drain(q_n, nullptr, 0);

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      A.seg_count[s * A.n_units + unit] = run[s];
      if (run[s] > seg_cap[s]) A.counts[4] = 1;  // segment overflow: retry with exact sizes
    }
  }
  // raw PAM-hit totals (pam_search semantics when K.raw)
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const uint32_t r = warp_sum_u32(raw_acc[s]);
    if (lane == 0 && r) atomicAdd((unsigned long long*)&A.counts[2 + s], (unsigned long long)r);
  }
}

// exclusive prefix of n values per strand (single CTA); optionally publishes the totals
__global__ void __launch_bounds__(1024) seg_prefix_kernel(const uint64_t* __restrict__ in, int32_t n,
                                                          uint64_t* __restrict__ out, uint64_t* totals) {
  __shared__ uint64_t part[1024];
  const int tid = threadIdx.x;
  const int per = (n + 1023) / 1024;
  for (int s = 0; s < 2; ++s) {
    const uint64_t* src = in + (size_t)s * n;
    uint64_t* dst = out + (size_t)s * n;
    const int lo = tid * per, hi = lo + per < n ? lo + per : n;
    uint64_t sum = 0;
    for (int j = lo; j < hi; ++j) sum += src[j];
    part[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      uint64_t y = tid >= o ? part[tid - o] : 0;
      __syncthreads();
      part[tid] += y;
      __syncthreads();
    }
    uint64_t run = part[tid] - sum;
    for (int j = lo; j < hi; ++j) {
      dst[j] = run;
      run += src[j];
    }
    if (totals && tid == 1023) totals[s] = part[1023];
    __syncthreads();
  }
}

// Concatenate the per-unit segments: one warp per (unit, strand)
struct CompactArgs {
  const uint64_t* seg_count;     // [2][n_units]
  const uint64_t* seg_base;      // their exclusive prefix
  const uint64_t* seg_prev;      // exact retry: segment sizes, else null
  const uint64_t* seg_prev_off;
  const double* unit_frac;
  const uint64_t* stage[2];
  uint64_t* hits[2];
  int64_t cap[2];      // staging capacity (segment geometry)
  int64_t out_cap[2];  // capacity of hits[]
  int32_t n_units;
};

__global__ void __launch_bounds__(128) compact_kernel(const __grid_constant__ CompactArgs A) {
  // one CTA per (unit, strand): segments are a few thousand records, so 128 threads with
  // 4 independent 8-byte copies in flight each cover one in a couple of passes
  const int unit = blockIdx.x, s = blockIdx.y, tid = threadIdx.x;
  uint64_t src_off, src_cap;
  if (A.seg_prev) {
    src_off = A.seg_prev_off[s * A.n_units + unit];
    src_cap = A.seg_prev[s * A.n_units + unit];
  } else {
    src_off = (uint64_t)(A.unit_frac[unit] * (double)A.cap[s]);
    const uint64_t hi = unit + 1 == A.n_units ? (uint64_t)A.cap[s] : (uint64_t)(A.unit_frac[unit + 1] * (double)A.cap[s]);
    src_cap = hi - src_off;
  }
  uint64_t n = A.seg_count[s * A.n_units + unit];
  if (n > src_cap) n = src_cap;  // overflowed segment: the launch is retried anyway
  const uint64_t base = A.seg_base[s * A.n_units + unit], cap = (uint64_t)A.out_cap[s];
  if (base >= cap) return;
  if (base + n > cap) n = cap - base;
  const uint64_t* __restrict__ src = A.stage[s] + src_off;
  uint64_t* __restrict__ dst = A.hits[s] + base;
  uint64_t k = tid;
  for (; k + 3 * 128 < n; k += 4 * 128) {
    const uint64_t a = src[k], b = src[k + 128], c = src[k + 256], d = src[k + 384];
    dst[k] = a;
    dst[k + 128] = b;
    dst[k + 256] = c;
    dst[k + 384] = d;
  }
  for (; k < n; k += 128) dst[k] = src[k];
}

}  // namespace hawk

// ------------------------------------------------------------------ launchers
using namespace hawk;

extern "C" int hawk_pack_dev(void* stream, const uint8_t* d_ascii, int64_t total_slots, void* d_q,
                             uint32_t* d_v, int64_t* d_bad) {
  if (total_slots < 0 || (total_slots % HAWK_CHUNK) != 0)
    return hawk_fail(HAWK_EINVAL, "hawk_pack_dev: total_slots must be a multiple of 32");
  if (((uintptr_t)d_ascii & 15) || ((uintptr_t)d_q & 15))
    return hawk_fail(HAWK_EINVAL, "hawk_pack_dev: buffers must be 16-byte aligned");
  int64_t n_chunks = total_slots / HAWK_CHUNK;
  if (n_chunks == 0) return HAWK_OK;
  int64_t blocks = (n_chunks + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;  // grid-stride: 16 CTAs per SM
  pack_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)d_ascii, n_chunks, (uint4*)d_q, d_v, (unsigned long long*)d_bad);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "pack_kernel launch");
}

extern "C" int32_t hawk_scan_units(int32_t sm_count, int64_t n_spans) {
  if (sm_count <= 0) sm_count = 148;
  int64_t n = (int64_t)sm_count * SCAN_CTAS_PER_SM * SCAN_WARPS;
  if (n > n_spans) n = n_spans;
  return (int32_t)(n < 0 ? 0 : n);
}

extern "C" int64_t hawk_scan_plan(const int32_t* scan_start, const int32_t* scan_stop,
                                  const uint8_t* is_ref, int32_t n_hap, int32_t raw_hits,
                                  int32_t n_units, int64_t* span_off, int64_t* unit_span,
                                  double* unit_frac) {
  // spans per haplotype (a span starts on a 4-chunk boundary: 128-bit reads of the case tile)
  auto spans_of = [&](int32_t h) -> int64_t {
    int64_t a = scan_start[h] < 0 ? 0 : scan_start[h], b = scan_stop[h];
    if (b <= a) return 0;
    int64_t chunks = ((b + 31) >> 5) - ((a >> 5) & ~(int64_t)3);
    return (chunks + HAWK_SPAN_CHUNKS - 1) / HAWK_SPAN_CHUNKS;
  };
  // a dense span (REF / pam_search mode: every chunk is matched) costs about 4 sparse ones and
  // emits about 16 times the records
  auto dense = [&](int32_t h) { return raw_hits || (is_ref && is_ref[h]); };
  int64_t total = 0;
  std::vector<int64_t> sp(n_hap + 1);
  std::vector<double> cw(n_hap + 1), ce(n_hap + 1);
  cw[0] = ce[0] = 0.0;
  for (int32_t h = 0; h < n_hap; ++h) {
    sp[h] = total;
    const int64_t n = spans_of(h);
    total += n;
    cw[h + 1] = cw[h] + (double)n * (dense(h) ? 4.0 : 1.0);
    ce[h + 1] = ce[h] + (double)n * (dense(h) ? 16.0 : 1.0);
  }
  sp[n_hap] = total;
  if (span_off)
    for (int32_t h = 0; h <= n_hap; ++h) span_off[h] = sp[h];
  if (!unit_span || n_units <= 0) return total;
  const double W = n_hap ? cw[n_hap] : 0.0, E = n_hap ? ce[n_hap] : 0.0;
  unit_span[0] = 0;
  if (unit_frac) unit_frac[0] = 0.0;
  int32_t h = 0;
  for (int32_t u = 1; u < n_units; ++u) {
    const double target = W * (double)u / (double)n_units;
    while (h + 1 < n_hap && cw[h + 1] <= target) ++h;  // haplotype holding the target weight
    const double wh = dense(h) ? 4.0 : 1.0, eh = dense(h) ? 16.0 : 1.0;
    int64_t k = (int64_t)((target - cw[h]) / wh);
    const int64_t nh = sp[h + 1] - sp[h];
    if (k > nh) k = nh;
    if (k < 0) k = 0;
    unit_span[u] = sp[h] + k;
    if (unit_span[u] < unit_span[u - 1]) unit_span[u] = unit_span[u - 1];
    if (unit_frac) unit_frac[u] = E > 0 ? (ce[h] + (double)k * eh) / E : 0.0;
  }
  unit_span[n_units] = total;
  if (unit_frac) unit_frac[n_units] = 1.0;
  return total;
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct ScanWs {
  int2* span_tab;
  uint64_t *seg_count, *seg_prev, *seg_prev_off, *seg_base, *stage[2];
  size_t bytes;
};

static ScanWs scan_ws_layout(void* base, int64_t n_spans, int32_t n_units, int64_t cap_fwd, int64_t cap_rev) {
  ScanWs w;
  char* p = (char*)base;
  size_t off = 0;
  w.span_tab = (int2*)(p + off);
  off = align256(off + (size_t)(n_spans > 0 ? n_spans : 1) * 8);
  const size_t seg = align256((size_t)(n_units > 0 ? n_units : 1) * 16);
  w.seg_count = (uint64_t*)(p + off); off += seg;
  w.seg_prev = (uint64_t*)(p + off); off += seg;
  w.seg_prev_off = (uint64_t*)(p + off); off += seg;
  w.seg_base = (uint64_t*)(p + off); off += seg;
  w.stage[0] = (uint64_t*)(p + off);
  off = align256(off + (size_t)(cap_fwd > 0 ? cap_fwd : 0) * 8);
  w.stage[1] = (uint64_t*)(p + off);
  off = align256(off + (size_t)(cap_rev > 0 ? cap_rev : 0) * 8);
  w.bytes = off;
  return w;
}

extern "C" size_t hawk_scan_workspace_bytes(int64_t n_spans, int32_t n_units, int64_t cap_fwd, int64_t cap_rev) {
  return scan_ws_layout(nullptr, n_spans, n_units, cap_fwd, cap_rev).bytes;
}

extern "C" int hawk_scan_dev(void* stream, const void* d_q, const uint32_t* d_v,
                             const int64_t* d_slot_off, const int32_t* d_len,
                             const int32_t* d_scan_start, const int32_t* d_scan_stop,
                             const uint8_t* d_is_ref, const int64_t* d_span_off,
                             const int64_t* d_unit_span, const double* d_unit_frac, int32_t n_hap,
                             int64_t n_spans, int32_t n_units, const hawk_params* params,
                             int32_t raw_hits, int32_t exact_retry, int64_t cap_fwd, int64_t cap_rev,
                             uint64_t* d_counts, void* d_workspace) {
  if (!params || params->pam_len < 1 || params->pam_len > HAWK_MAX_PAM || params->guide_len < 1)
    return hawk_fail(HAWK_EINVAL, "hawk_scan_dev: bad PAM / guide length");
  if (params->pam_len + params->guide_len + 2 * HAWK_GUIDESEQPAD > HAWK_MAX_WINDOW)
    return hawk_fail(HAWK_EINVAL, "hawk_scan_dev: guide + PAM window exceeds HAWK_MAX_WINDOW");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_spans <= 0 || n_hap <= 0 || n_units <= 0)
    return hawk_check_cuda(cudaMemsetAsync(d_counts, 0, 64, st), "counts memset");
  const ScanWs W = scan_ws_layout(d_workspace, n_spans, n_units, cap_fwd, cap_rev);
  if (exact_retry) {
    cudaError_t e = cudaMemcpyAsync(W.seg_prev, W.seg_count, (size_t)n_units * 16, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return hawk_check_cuda(e, "segment sizes copy");
    seg_prefix_kernel<<<1, 1024, 0, st>>>(W.seg_prev, n_units, W.seg_prev_off, nullptr);
    hawk_note_launch(1);
  }
  span_table_kernel<<<(unsigned)((n_spans + 255) / 256), 256, 0, st>>>(d_span_off, d_scan_start, n_hap,
                                                                      n_spans, W.span_tab, d_counts);
  hawk_note_launch(1);
  ScanArgs A;
  A.B = BatchView{};
  A.B.q = (const Planes*)d_q;
  A.B.v = d_v;
  A.B.slot_off = d_slot_off;
  A.B.len = d_len;
  A.B.scan_start = d_scan_start;
  A.B.scan_stop = d_scan_stop;
  A.B.is_ref = d_is_ref;
  A.B.n_hap = n_hap;
  A.K = make_scan_const(*params, raw_hits);
  A.span_tab = W.span_tab;
  A.unit_span = d_unit_span;
  A.unit_frac = d_unit_frac;
  A.seg_prev = exact_retry ? W.seg_prev : nullptr;
  A.seg_prev_off = exact_retry ? W.seg_prev_off : nullptr;
  A.seg_count = W.seg_count;
  A.stage[0] = W.stage[0];
  A.stage[1] = W.stage[1];
  A.cap[0] = cap_fwd;
  A.cap[1] = cap_rev;
  A.n_units = n_units;
  A.counts = d_counts;
  cudaFuncSetAttribute(scan_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  scan_kernel<<<(unsigned)((n_units + SCAN_WARPS - 1) / SCAN_WARPS), SCAN_THREADS, 0, st>>>(A);
  hawk_note_launch(1);
  // per-unit exclusive prefix + totals (counts[0..1])
  seg_prefix_kernel<<<1, 1024, 0, st>>>(W.seg_count, n_units, W.seg_base, d_counts);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "scan kernels launch");
}

extern "C" int hawk_scan_compact_dev(void* stream, const double* d_unit_frac, int32_t n_units,
                                     int64_t n_spans, int32_t exact_retry, int64_t cap_fwd,
                                     int64_t cap_rev, void* d_workspace, uint64_t* d_hits_fwd,
                                     uint64_t* d_hits_rev, int64_t out_cap_fwd, int64_t out_cap_rev) {
  if (n_spans <= 0 || n_units <= 0) return HAWK_OK;
  const ScanWs W = scan_ws_layout(d_workspace, n_spans, n_units, cap_fwd, cap_rev);
  CompactArgs C;
  C.seg_count = W.seg_count;
  C.seg_base = W.seg_base;
  C.seg_prev = exact_retry ? W.seg_prev : nullptr;
  C.seg_prev_off = exact_retry ? W.seg_prev_off : nullptr;
  C.unit_frac = d_unit_frac;
  C.stage[0] = W.stage[0];
  C.stage[1] = W.stage[1];
  C.hits[0] = d_hits_fwd;
  C.hits[1] = d_hits_rev;
  C.cap[0] = cap_fwd;
  C.cap[1] = cap_rev;
  C.out_cap[0] = out_cap_fwd;
  C.out_cap[1] = out_cap_rev;
  C.n_units = n_units;
  compact_kernel<<<dim3((unsigned)n_units, 2), 128, 0, (cudaStream_t)stream>>>(C);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "compact_kernel launch");
}
