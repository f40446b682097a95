// scan_kernels.cu -- K1 (pack), K2 (PAM scan + fused filters) and K3 (segment
// concatenation) for sm_100a, plus their host-side plan and launchers. Integer / bitwise
// work on HBM-resident bit planes: no tensor cores. See DESIGN.md section 3; each kernel is
// described where it is defined.
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
                                                   uint32_t* __restrict__ v, uint32_t* __restrict__ nz,
                                                   unsigned long long* __restrict__ bad) {
  // a warp packs 32 consecutive chunks per pass (chunk index = multiple of 32 + lane), so one
  // ballot gives the 32 "chunk holds a variant base" bits of a whole nz word
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t c0 = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); c0 < n_chunks; c0 += stride) {
    const int64_t c = c0 + lane;
    uint32_t vw = 0;
    if (c < n_chunks) {
      uint4 w[2];
      w[0] = __ldg(&ascii[2 * c]);
      w[1] = __ldg(&ascii[2 * c + 1]);
      const PackedChunk o = pack_chunk(reinterpret_cast<const uint32_t*>(w));
      q[c] = make_uint4(o.a, o.c, o.g, o.t);
      v[c] = o.v;
      vw = o.v;
      if (o.invalid) atomicMin(bad, (unsigned long long)(c * 32 + (__ffs(o.invalid) - 1)));
    }
    const uint32_t word = __ballot_sync(0xFFFFFFFFu, vw != 0);
    if (lane == 0) nz[c0 >> 5] = word;
  }
}

// ------------------------------------------------------------------ K2: scan
// Warp-autonomous, persistent. A *span* is HAWK_SPAN_CHUNKS = 1,024 chunks (32,768 base
// slots) of one haplotype; spans are numbered in (haplotype, position) order and every warp
// ("unit") owns a contiguous span range [unit_span[u], unit_span[u+1]) balanced on the host
// (hawk_scan_plan). No warp ever waits for another one:
//   phase A   (non-REF haplotypes) one 32-bit slice of the nz summary per lane = 32 chunks;
//             a chunk is a candidate iff it or a neighbour holds a variant base (a guide core
//             reaches at most one chunk either side when G <= 32; search_guides.py:468-471);
//             the candidates of the span are appended in order to a per-warp queue in shared
//             memory (warp prefix sum). Variant-free stretches cost 1 bit per 1,024 bp.
//   phase B   whenever 32 candidates are queued, one lane each: the chunk's three case words,
//             a log-doubling sliding OR over that 96-bit window -> per-position "core holds a
//             variant" masks for both strands; two 128-bit plane loads; the branch-free
//             AND-mask PAM test on both strands over shared funnel-shifted planes
//             (match_fixed<P>); interval masks for the scan bounds and is_pamhit_in_range;
//             a packed warp prefix sum of the hit counts places the (hap << 32 | pos) records
//             straight into the warp's private segment of a staging buffer.
//   dense     REF haplotypes / pam_search mode take every chunk, 32 per pass, no queue.
// seg_prefix_kernel + compact_kernel then concatenate the segments (exclusive prefix over the
// per-unit totals), which yields the stream sorted by (haplotype, position) without a sort.
constexpr int SCAN_WARPS = 8;
constexpr int SCAN_THREADS = SCAN_WARPS * 32;
constexpr int SCAN_CTAS_PER_SM = 4;
constexpr int SPAN = HAWK_SPAN_CHUNKS;
constexpr int QCAP = SPAN + 32;  // <= 31 carried candidates + a whole span
static_assert(SPAN == 1024, "one nz bit per chunk, 32 chunks per lane");

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
  __shared__ uint32_t queue_all[SCAN_WARPS][QCAP];  // candidate chunks of the current haplotype
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int unit = blockIdx.x * SCAN_WARPS + warp;
  if (unit >= A.n_units) return;
  const int64_t sp0 = A.unit_span[unit], sp1 = A.unit_span[unit + 1];

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

  uint32_t raw_acc[2] = {0, 0};
  uint32_t run[2] = {0, 0};  // records this warp has produced so far
  int32_t cur_hap = -1;
  HapScan H;
  H.is_ref = 0;
  uint32_t* const queue = queue_all[warp];
  uint32_t q_n = 0;  // queued candidates, at queue[0 .. q_n) (warp-uniform)

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

  // match + filters for the n <= 32 queued candidates at queue[off ..)
  auto drain = [&](uint32_t off, uint32_t n) {
    uint32_t out[2] = {0, 0}, raw[2] = {0, 0};
    int32_t c = 0;
    if ((uint32_t)lane < n) {
      c = (int32_t)queue[off + lane];
      if (A.K.small) {
        // the slot layout's zero gap makes c - 1 / c + 1 safe at the haplotype's ends
        const uint32_t* vp = A.B.v + H.chunk0 + c;
        const uint32_t w0 = __ldg(vp - 1) & A.K.prev_mask, w1 = __ldg(vp), w2 = __ldg(vp + 1) & A.K.next_mask;
        if (w0 | w1 | w2) scan_chunk_small(A.B, A.K, H, c, w0, w1, w2, true, out, raw);
      } else {
        scan_chunk(A.B, A.K, H, (int64_t)c, out, raw);
      }
    }
    emit(out, c);
  };
  // drain every full batch, keep the remainder (< 32, or nothing when `all`) at the front
  auto drain_queue = [&](bool all) {
    uint32_t off = 0;
    while (q_n - off >= 32) {
      drain(off, 32);
      off += 32;
    }
    if (all && q_n > off) {
      drain(off, q_n - off);
      off = q_n;
    }
    if (off) {
      const uint32_t rem = q_n - off;
      uint32_t t = 0;
      if ((uint32_t)lane < rem) t = queue[off + lane];
      __syncwarp();
      if ((uint32_t)lane < rem) queue[lane] = t;
      __syncwarp();
      q_n = rem;
    }
  };

  for (int64_t span = sp0; span < sp1; ++span) {
    const int2 e = __ldg(&A.span_tab[span]);
    if (e.x != cur_hap) {
      if (q_n) drain_queue(true);
      cur_hap = e.x;
      H = load_hap_scan(A.B, A.K, e.x);
    }
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
      continue;
    }
    // ---- sparse: this lane's 32 chunks [c32, c32 + 32) and one nz bit either side
    const int32_t c32 = c_first + 32 * lane;
    uint32_t cand = 0, any = 0;
    if (c32 < c_end) {
      const int64_t bit = H.chunk0 + c32 - 1;  // >= 3: the slot space starts with a zero gap
      const uint32_t* wp = A.B.nz + (bit >> 5);
      const uint32_t sh = (uint32_t)(bit & 31);
      const uint32_t x0 = __ldg(wp), x1 = __ldg(wp + 1), x2 = __ldg(wp + 2);
      const uint32_t lo = funnel_r(x0, x1, sh), hi = funnel_r(x1, x2, sh);  // bits bit .. bit + 63
      const uint32_t mid = (lo >> 1) | (hi << 31);                          // chunks c32 .. c32 + 31
      cand = mid | (mid << 1) | (lo & 1u) | (mid >> 1) | ((hi << 30) & 0x80000000u);
      any = lo | hi;
    }
    if (!A.K.small) {
      // long guides reach up to 4 chunks either side: take every chunk of a 32-chunk slice
      // whose neighbourhood (this slice, the one before, the one after) holds a variant;
      // the slices at the ends of the span are always taken
      const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, any, 1), next = __shfl_down_sync(0xFFFFFFFFu, any, 1);
      const bool edge = lane == 0 || lane == 31 || c32 + 32 >= c_end;
      cand = (c32 < c_end && ((any | prev | next) != 0 || edge)) ? 0xFFFFFFFFu : 0u;
    }
    if (c32 < c_end) cand &= interval_mask(c_lo, c_end, c32);
    if (!__any_sync(0xFFFFFFFFu, cand != 0)) continue;
    // append this lane's candidates behind the earlier lanes' (chunk order)
    const uint32_t mine = __popc(cand);
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= d) incl += y;
    }
    uint32_t* qp = queue + q_n + (incl - mine);
    while (cand) {
      const int b = __ffs(cand) - 1;
      cand &= cand - 1;
      *qp++ = (uint32_t)(c32 + b);
#ifdef HAWK_PREFETCH
      // the candidate's case words and planes are read when its batch is drained: start both
      // DRAM fetches now so they overlap instead of following one another
      asm volatile("prefetch.global.L2 [%0];" ::"l"(A.B.v + H.chunk0 + c32 + b));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(&A.B.q[H.chunk0 + c32 + b]));
#endif
    }
    q_n += __shfl_sync(0xFFFFFFFFu, incl, 31);
    __syncwarp();
    if (q_n >= 32) drain_queue(false);
  }
  if (q_n) drain_queue(true);

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
                             uint32_t* d_v, uint32_t* d_nz, int64_t* d_bad) {
  if (total_slots < 0 || (total_slots % HAWK_CHUNK) != 0)
    return hawk_fail(HAWK_EINVAL, "hawk_pack_dev: total_slots must be a multiple of 32");
  if (((uintptr_t)d_ascii & 15) || ((uintptr_t)d_q & 15))
    return hawk_fail(HAWK_EINVAL, "hawk_pack_dev: buffers must be 16-byte aligned");
  int64_t n_chunks = total_slots / HAWK_CHUNK;
  if (n_chunks == 0) return HAWK_OK;
  int64_t blocks = (n_chunks + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;  // grid-stride: 16 CTAs per SM
  pack_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)d_ascii, n_chunks, (uint4*)d_q, d_v, d_nz, (unsigned long long*)d_bad);
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

extern "C" int hawk_scan_dev(void* stream, const void* d_q, const uint32_t* d_v, const uint32_t* d_nz,
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
  A.B.nz = d_nz;
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
