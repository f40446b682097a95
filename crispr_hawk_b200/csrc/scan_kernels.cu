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

#include "hawk_core.h"
#include "hawk_kernels.h"

namespace hawk {

// ------------------------------------------------------------------ K1: pack
// One thread per chunk: 32 ASCII bytes -> {A,C,G,T} plane words + case word.
__global__ void __launch_bounds__(256) pack_kernel(const uint4* __restrict__ ascii,
                                                   int64_t n_chunks, uint4* __restrict__ q,
                                                   uint32_t* __restrict__ v,
                                                   unsigned long long* __restrict__ bad) {
  __shared__ uint8_t lut[256];
  lut[threadIdx.x] = iupac_entry((uint8_t)threadIdx.x);
  __syncthreads();
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += stride) {
    uint4 w[2];
    w[0] = __ldg(&ascii[2 * c]);
    w[1] = __ldg(&ascii[2 * c + 1]);
    const PackedChunk o = pack_chunk(reinterpret_cast<const uint32_t*>(w),
                                     [&](uint32_t byte) { return lut[byte]; });
    q[c] = make_uint4(o.a, o.c, o.g, o.t);
    v[c] = o.v;
    if (o.invalid) atomicMin(bad, (unsigned long long)(c * 32 + (__ffs(o.invalid) - 1)));
  }
}

// ------------------------------------------------------------------ K2: scan
constexpr int SCAN_THREADS = 256;
constexpr int SPAN_ITERS = HAWK_SPAN_CHUNKS / SCAN_THREADS;  // chunks per thread per span
static_assert(HAWK_SPAN_CHUNKS % SCAN_THREADS == 0, "span must be a multiple of the CTA size");

constexpr uint64_t ST_AGG = 1ull << 62, ST_INC = 2ull << 62, ST_MASK = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
  return x;
}

// Decoupled look-back (one warp): publishes this span's total and returns the
// sum of all predecessors' totals.
__device__ uint64_t span_lookback(volatile uint64_t* status, int64_t span, uint64_t total,
                                  int lane) {
  if (span == 0) {
    if (lane == 0) status[0] = ST_INC | total;
    return 0;
  }
  if (lane == 0) status[span] = ST_AGG | total;
  uint64_t excl = 0;
  int64_t j = span - 1;
  for (;;) {
    int64_t idx = j - lane;
    uint64_t val = ST_INC;  // virtual predecessor before span 0: inclusive prefix 0
    if (idx >= 0) {
      val = status[idx];
      while ((val >> 62) == 0) {
        __nanosleep(40);
        val = status[idx];
      }
    }
    unsigned inc = __ballot_sync(0xFFFFFFFFu, (val >> 62) == 2);
    uint64_t contrib = val & ST_MASK;
    if (inc) {
      int first = __ffs(inc) - 1;  // nearest predecessor holding an inclusive prefix
      excl += warp_sum_u64(lane <= first ? contrib : 0);
      break;
    }
    excl += warp_sum_u64(contrib);
    j -= 32;
  }
  if (lane == 0) status[span] = ST_INC | (excl + total);
  return excl;
}

struct ScanArgs {
  BatchView B;
  ScanConst K;
  const int64_t* span_off;  // n_hap + 1
  int64_t n_spans;
  uint64_t* hits[2];
  int64_t cap[2];
  uint64_t* counts;          // [0..1] filtered totals, [2..3] raw totals
  unsigned long long* ticket;
  uint64_t* status[2];       // n_spans each, zero-initialised
};

__global__ void __launch_bounds__(SCAN_THREADS) scan_kernel(const __grid_constant__ ScanArgs A) {
  // bitmap word w lives at w + w/32: phase 2/4 read with stride SPAN_ITERS, the skew keeps
  // those reads bank-conflict free
  __shared__ uint32_t bm[2][HAWK_SPAN_CHUNKS + HAWK_SPAN_CHUNKS / 32];
  __shared__ uint32_t warp_tot[2][SCAN_THREADS / 32];
  __shared__ uint64_t span_base[2];
  __shared__ long long s_span;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint64_t raw_acc[2] = {0, 0};

  for (;;) {
    if (tid == 0) s_span = (long long)atomicAdd(A.ticket, 1ull);
    __syncthreads();
    const int64_t span = s_span;
    if (span >= A.n_spans) break;

    // haplotype of this span: last h with span_off[h] <= span
    int32_t h;
    {
      int32_t lo = 0, hi = A.B.n_hap;  // span_off[lo] <= span < span_off[hi]
      while (hi - lo > 1) {
        int32_t mid = (lo + hi) >> 1;
        if (__ldg(&A.span_off[mid]) <= span) lo = mid; else hi = mid;
      }
      h = lo;
    }
    const HapScan H = load_hap_scan(A.B, A.K, h);
    const int64_t c_first = ((int64_t)H.a >> 5) + (span - __ldg(&A.span_off[h])) * HAWK_SPAN_CHUNKS;
    const int64_t c_end = ((int64_t)H.b + 31) >> 5;

    // phase 1: hit bitmaps
#pragma unroll 2
    for (int it = 0; it < SPAN_ITERS; ++it) {
      int w = it * SCAN_THREADS + tid;
      int64_t c = c_first + w;
      uint32_t out[2] = {0, 0}, raw[2] = {0, 0};
      if (c < c_end) scan_chunk(A.B, A.K, H, c, out, raw);
      bm[0][w + (w >> 5)] = out[0];
      bm[1][w + (w >> 5)] = out[1];
      raw_acc[0] += __popc(raw[0]);
      raw_acc[1] += __popc(raw[1]);
    }
    __syncthreads();

    // phase 2: block exclusive scan of popcounts; thread t owns words [t*ITERS, (t+1)*ITERS)
    uint32_t mine[2] = {0, 0};
#pragma unroll
    for (int k = 0; k < SPAN_ITERS; ++k) {
      int w = tid * SPAN_ITERS + k;
      mine[0] += __popc(bm[0][w + (w >> 5)]);
      mine[1] += __popc(bm[1][w + (w >> 5)]);
    }
    uint32_t incl[2] = {mine[0], mine[1]};
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t a = __shfl_up_sync(0xFFFFFFFFu, incl[0], o);
      uint32_t b = __shfl_up_sync(0xFFFFFFFFu, incl[1], o);
      if (lane >= o) {
        incl[0] += a;
        incl[1] += b;
      }
    }
    if (lane == 31) {
      warp_tot[0][warp] = incl[0];
      warp_tot[1][warp] = incl[1];
    }
    __syncthreads();
    uint32_t wbase[2] = {0, 0}, total[2] = {0, 0};
#pragma unroll
    for (int k = 0; k < SCAN_THREADS / 32; ++k) {
      uint32_t t0 = warp_tot[0][k], t1 = warp_tot[1][k];
      if (k < warp) {
        wbase[0] += t0;
        wbase[1] += t1;
      }
      total[0] += t0;
      total[1] += t1;
    }

    // phase 3: global offset of this span (warp 0 -> strand 0, warp 1 -> strand 1)
    if (warp < 2) {
      uint64_t excl = span_lookback(A.status[warp], span, total[warp], lane);
      if (lane == 0) {
        span_base[warp] = excl;
        if (span == A.n_spans - 1) A.counts[warp] = excl + total[warp];
      }
    }
    __syncthreads();

    // phase 4: expand bitmaps to records
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (mine[s] == 0) continue;
      uint64_t o = span_base[s] + wbase[s] + (incl[s] - mine[s]);
      uint64_t* dst = A.hits[s];
      const uint64_t cap = (uint64_t)A.cap[s];
      const uint64_t hkey = (uint64_t)(uint32_t)h << 32;
      for (int k = 0; k < SPAN_ITERS; ++k) {
        int w = tid * SPAN_ITERS + k;
        uint32_t bits = bm[s][w + (w >> 5)];
        uint64_t p0 = (uint64_t)(c_first + tid * SPAN_ITERS + k) << 5;
        while (bits) {
          int b = __ffs(bits) - 1;
          bits &= bits - 1;
          if (o < cap) dst[o] = hkey | (p0 + b);
          ++o;
        }
      }
    }
    __syncthreads();  // bitmaps and s_span are reused by the next span
  }

  // raw PAM-hit totals (pam_search semantics when K.raw)
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    uint64_t r = warp_sum_u64(raw_acc[s]);
    if (lane == 0 && r) atomicAdd((unsigned long long*)&A.counts[2 + s], (unsigned long long)r);
  }
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

extern "C" int64_t hawk_scan_plan(const int32_t* scan_start, const int32_t* scan_stop,
                                  int32_t n_hap, int64_t* span_off) {
  int64_t total = 0;
  for (int32_t h = 0; h < n_hap; ++h) {
    if (span_off) span_off[h] = total;
    int64_t a = scan_start[h] < 0 ? 0 : scan_start[h], b = scan_stop[h];
    if (b > a) {
      int64_t chunks = ((b + 31) >> 5) - (a >> 5);
      total += (chunks + HAWK_SPAN_CHUNKS - 1) / HAWK_SPAN_CHUNKS;
    }
  }
  if (span_off) span_off[n_hap] = total;
  return total;
}

extern "C" size_t hawk_scan_workspace_bytes(int64_t n_spans) {
  return 256 + (size_t)(n_spans > 0 ? n_spans : 1) * 16;
}

extern "C" int hawk_scan_dev(void* stream, int32_t sm_count, const void* d_q, const uint32_t* d_v,
                             const int64_t* d_slot_off, const int32_t* d_len,
                             const int32_t* d_scan_start, const int32_t* d_scan_stop,
                             const uint8_t* d_is_ref, const int64_t* d_span_off, int32_t n_hap,
                             int64_t n_spans, const hawk_params* params, int32_t raw_hits,
                             uint64_t* d_hits_fwd, uint64_t* d_hits_rev, int64_t cap_fwd,
                             int64_t cap_rev, uint64_t* d_counts, void* d_workspace) {
  if (!params || params->pam_len < 1 || params->pam_len > HAWK_MAX_PAM || params->guide_len < 1)
    return hawk_fail(HAWK_EINVAL, "hawk_scan_dev: bad PAM / guide length");
  if (params->pam_len + params->guide_len + 2 * HAWK_GUIDESEQPAD > HAWK_MAX_WINDOW)
    return hawk_fail(HAWK_EINVAL, "hawk_scan_dev: guide + PAM window exceeds HAWK_MAX_WINDOW");
  if (n_spans <= 0 || n_hap <= 0) return HAWK_OK;
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
  A.span_off = d_span_off;
  A.n_spans = n_spans;
  A.hits[0] = d_hits_fwd;
  A.hits[1] = d_hits_rev;
  A.cap[0] = cap_fwd;
  A.cap[1] = cap_rev;
  A.counts = d_counts;
  A.ticket = (unsigned long long*)d_workspace;
  A.status[0] = (uint64_t*)((char*)d_workspace + 256);
  A.status[1] = A.status[0] + n_spans;
  if (sm_count <= 0) sm_count = 148;
  int64_t blocks = (int64_t)sm_count * 8;  // persistent CTAs, 8 per SM (16 KB smem each)
  if (blocks > n_spans) blocks = n_spans;
  scan_kernel<<<(unsigned)blocks, SCAN_THREADS, 0, (cudaStream_t)stream>>>(A);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "scan_kernel launch");
}
