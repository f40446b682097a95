// api.cu -- C-ABI host layer of libhawkscan: contexts, packed batches, the
// search pipeline (K2 scan -> rows -> [resolve] -> compaction -> gather ->
// buckets) and result download. See include/hawkscan.h for the contract.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <new>
#include <vector>
#include <stdlib.h>

#include "hawk_core.h"
#include "hawk_kernels.h"
#include "hawk_post.h"
#include "hawk_host.h"

using namespace hawk;

// ------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";

int hawk_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int hawk_check_cuda(cudaError_t err, const char* what) {
  if (err == cudaSuccess) return HAWK_OK;
  return hawk_fail(err == cudaErrorMemoryAllocation ? HAWK_ENOMEM : HAWK_ECUDA, "%s: %s (%s)", what,
                   cudaGetErrorString(err), cudaGetErrorName(err));
}

#include <atomic>
static std::atomic<long long> g_launches{0};
void hawk_note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" int64_t hawk_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int hawk_abi_version(void) { return HAWK_ABI_VERSION; }
extern "C" const char* hawk_last_error(void) { return g_err; }
extern "C" const char* hawk_strerror(int code) {
  switch (code) {
    case HAWK_OK: return "ok";
    case HAWK_EINVAL: return "invalid argument";
    case HAWK_ECUDA: return "CUDA failure";
    case HAWK_ENOMEM: return "out of device memory";
    case HAWK_EIUPAC: return "non-IUPAC character";
    case HAWK_ECAPACITY: return "capacity exceeded";
    case HAWK_EALLELES: return "ambiguity code without variant alleles";
    case HAWK_EDUPREF: return "duplicate REF guide";
    case HAWK_EASSERT: return "the reference asserts on this input";
    case HAWK_ECFD: return "CFD score tables hold no entry for this guide";
    case HAWK_EFEATURE: return "scorer input holds a letter other than A, C, G, T";
    default: return "unknown error";
  }
}

// ------------------------------------------------------------------ small transfers
namespace hawk {
// 16-byte words when both pointers and the size allow it, bytes otherwise
__global__ void small_copy_kernel(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, size_t n, int wide) {
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
  if (wide) {
    for (size_t i = i0; i < n / 16; i += step) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
  } else {
    for (size_t i = i0; i < n; i += step) dst[i] = src[i];
  }
}
}  // namespace hawk

static const size_t SMALL_MAX = 8u << 20;  // larger transfers use the copy engines

static int launch_small_copy(cudaStream_t st, void* dst, const void* src, size_t n) {
  const int wide = (((uintptr_t)dst | (uintptr_t)src | n) & 15) == 0;
  const size_t items = wide ? n / 16 : n;
  size_t blocks = (items + 255) / 256;
  if (blocks > 592) blocks = 592;
  if (blocks < 1) blocks = 1;
  small_copy_kernel<<<(unsigned)blocks, 256, 0, st>>>((uint8_t*)dst, (const uint8_t*)src, n, wide);
  hawk_note_launch(1);
  return hawk_check_cuda(cudaGetLastError(), "small_copy_kernel launch");
}

static int arena_take(hawk_ctx* c, size_t n, uint8_t** out) {
  const size_t need = (n + 15) & ~(size_t)15;
  if (c->arena_used + need > c->arena_bytes) {
    // everything staged so far has been consumed once the stream is idle
    CKCUDA(cudaStreamSynchronize(c->stream));
    c->arena_used = 0;
    if (need > c->arena_bytes) {
      if (c->arena) cudaFreeHost(c->arena);
      c->arena = nullptr;
      c->arena_bytes = 0;
      size_t want = 4u << 20;
      while (want < 2 * need) want <<= 1;
      void* p = nullptr;
      if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return hawk_fail(HAWK_ENOMEM, "pinned staging allocation failed (%zu bytes)", want);
      }
      c->arena = (uint8_t*)p;
      c->arena_bytes = want;
    }
  }
  *out = c->arena + c->arena_used;
  c->arena_used += need;
  return HAWK_OK;
}

int hawk_ctx::small_h2d(void* dst_dev, const void* src_host, size_t n) {
  if (n == 0) return HAWK_OK;
  h2d_bytes += (int64_t)n;
  // the copy engine is the better mover for anything sizeable unless a streamed search keeps it
  // busy with its bulk input copies (a transfer would queue behind them, whatever its stream)
  if (n > SMALL_MAX || (!bulk_h2d && n >= (64u << 10)))
    return hawk_check_cuda(cudaMemcpyAsync(dst_dev, src_host, n, cudaMemcpyHostToDevice, stream), "H2D copy");
  // page-locked source (the caller pinned it): the SMs read it in place
  cudaPointerAttributes attr;
  if (n >= 4096 && cudaPointerGetAttributes(&attr, src_host) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
      attr.devicePointer != nullptr)
    return launch_small_copy(stream, dst_dev, attr.devicePointer, n);
  cudaGetLastError();
  uint8_t* stage = nullptr;
  CK(arena_take(this, n, &stage));
  memcpy(stage, src_host, n);
  return launch_small_copy(stream, dst_dev, stage, n);
}

int hawk_ctx::small_d2h_sync(void* dst_host, const void* src_dev, size_t n) {
  d2h_bytes += (int64_t)n;
  if (n > SMALL_MAX) {
    CKCUDA(cudaMemcpyAsync(dst_host, src_dev, n, cudaMemcpyDeviceToHost, stream));
    CKCUDA(cudaStreamSynchronize(stream));
    arena_used = 0;
    return HAWK_OK;
  }
  uint8_t* stage = nullptr;
  if (n) {
    CK(arena_take(this, n, &stage));
    CK(launch_small_copy(stream, stage, src_dev, n));
  }
  CKCUDA(cudaStreamSynchronize(stream));
  if (n) memcpy(dst_host, stage, n);
  arena_used = 0;
  return HAWK_OK;
}

// profiling hooks for the device layer: the context the calling thread is running a search on
static thread_local hawk_ctx* g_prof_ctx = nullptr;
void hawk_prof_begin(cudaStream_t st, int kind) {
  if (g_prof_ctx && g_prof_ctx->stream == st) g_prof_ctx->mark(kind, nullptr);
}
void hawk_prof_end(cudaStream_t st) {
  if (g_prof_ctx && g_prof_ctx->stream == st) g_prof_ctx->close_mark();
}
struct ProfScope {
  explicit ProfScope(hawk_ctx* c) { g_prof_ctx = c->profiling ? c : nullptr; }
  ~ProfScope() { g_prof_ctx = nullptr; }
};

// ------------------------------------------------------------------ context
extern "C" int hawk_ctx_create(int device, hawk_ctx** out) {
  if (!out) return hawk_fail(HAWK_EINVAL, "hawk_ctx_create: null output");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return hawk_fail(HAWK_ECUDA, "hawk_ctx_create: no CUDA device available (%s); libhawkscan has no CPU path",
                     e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return hawk_fail(HAWK_EINVAL, "hawk_ctx_create: device %d of %d", device, n);
  CKCUDA(cudaSetDevice(device));
  hawk_ctx* c = new (std::nothrow) hawk_ctx();
  if (!c) return hawk_fail(HAWK_ENOMEM, "hawk_ctx_create: host allocation");
  c->device = device;
  cudaDeviceProp prop;
  CKCUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  CKCUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  *out = c;
  return HAWK_OK;
}

extern "C" int hawk_ctx_destroy(hawk_ctx* c) {
  if (!c) return HAWK_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  c->trim();
  if (c->pinned) cudaFreeHost(c->pinned);
  if (c->arena) cudaFreeHost(c->arena);
  if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
  if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
  cudaStreamDestroy(c->stream);
  delete c;
  return HAWK_OK;
}

extern "C" int hawk_ctx_traffic(hawk_ctx* c, int64_t* h2d_bytes, int64_t* d2h_bytes) {
  if (!c) return hawk_fail(HAWK_EINVAL, "hawk_ctx_traffic: null context");
  if (h2d_bytes) *h2d_bytes = c->h2d_bytes;
  if (d2h_bytes) *d2h_bytes = c->d2h_bytes;
  return HAWK_OK;
}

extern "C" int hawk_ctx_info(hawk_ctx* c, int32_t* sm_count, int64_t* total_mem, int64_t* free_mem) {
  if (!c) return hawk_fail(HAWK_EINVAL, "hawk_ctx_info: null context");
  CKCUDA(cudaSetDevice(c->device));
  size_t f = 0, t = 0;
  CKCUDA(cudaMemGetInfo(&f, &t));
  if (sm_count) *sm_count = c->sm_count;
  if (total_mem) *total_mem = (int64_t)t;
  if (free_mem) *free_mem = (int64_t)f;
  return HAWK_OK;
}

// ------------------------------------------------------------------ layout
extern "C" int hawk_layout(const int32_t* len, int32_t n_hap, int64_t* slot_off, int64_t* total_slots) {
  if (n_hap < 0 || (n_hap > 0 && (!len || !slot_off)))
    return hawk_fail(HAWK_EINVAL, "hawk_layout: bad arguments");
  int64_t off = HAWK_SLOT_GAP;
  for (int32_t h = 0; h < n_hap; ++h) {
    if (len[h] < 0) return hawk_fail(HAWK_EINVAL, "hawk_layout: negative length at %d", h);
    slot_off[h] = off;
    off += ((int64_t)len[h] + HAWK_SLOT_ALIGN - 1) / HAWK_SLOT_ALIGN * HAWK_SLOT_ALIGN + HAWK_SLOT_GAP;
  }
  if (slot_off) slot_off[n_hap] = off;
  if (total_slots) *total_slots = off;
  return HAWK_OK;
}

// ------------------------------------------------------------------ batch
int batch_create_impl(hawk_ctx* c, const uint8_t* ascii, bool ascii_on_device,
                             const int64_t* slot_off, const int32_t* len, int32_t n_hap,
                             hawk_batch** out, int64_t* bad_slot, bool defer_planes) {
  if (!c || !out || n_hap < 0 || (n_hap > 0 && ((!ascii && !defer_planes) || !slot_off || !len)))
    return hawk_fail(HAWK_EINVAL, "hawk_batch_create: bad arguments");
  CKCUDA(cudaSetDevice(c->device));
  if (bad_slot) *bad_slot = -1;
  std::vector<int64_t> expect(n_hap + 1);
  int64_t total = 0;
  CK(hawk_layout(len, n_hap, expect.data(), &total));
  for (int32_t h = 0; h <= n_hap; ++h)
    if (n_hap && slot_off[h] != expect[h])
      return hawk_fail(HAWK_EINVAL, "hawk_batch_create: slot_off[%d] does not follow hawk_layout", h);
  hawk_batch* b = new (std::nothrow) hawk_batch();
  if (!b) return hawk_fail(HAWK_ENOMEM, "hawk_batch_create: host allocation");
  b->ctx = c;
  b->n_hap = n_hap;
  b->total_slots = total;
  b->slot_off.assign(expect.begin(), expect.end());
  b->len.assign(len, len + n_hap);
  cudaStream_t st = c->stream;
  int rc = HAWK_OK;
  DevBuf d_ascii, d_bad;
  const size_t n_chunks = (size_t)total / HAWK_CHUNK + HAWK_SLACK_CHUNKS;
  // K1 writes every chunk of the slot space; only the readable slack behind it is zeroed
  const size_t used = (size_t)total / HAWK_CHUNK;
  do {
    if ((rc = b->q.alloc(c, n_chunks * 16))) break;
    if ((rc = b->v.alloc(c, n_chunks * 4))) break;
    const size_t nz_words = (used + 31) / 32 + HAWK_SLACK_CHUNKS;
    if ((rc = b->nz.alloc(c, nz_words * 4, true))) break;
    if ((rc = hawk_check_cuda(cudaMemsetAsync(b->q.as<uint8_t>() + used * 16, 0, (n_chunks - used) * 16, st), "slack memset"))) break;
    if ((rc = hawk_check_cuda(cudaMemsetAsync(b->v.as<uint8_t>() + used * 4, 0, (n_chunks - used) * 4, st), "slack memset"))) break;
    if ((rc = upload(c, b->d_slot_off, b->slot_off.data(), (size_t)(n_hap + 1) * 8))) break;
    if ((rc = upload(c, b->d_len, b->len.data(), (size_t)n_hap * 4))) break;
    if (defer_planes) {
      // the caller fills the planes later (hawk_edits_ensure)
    } else if (!ascii) {  // no haplotypes: the slot space is the leading gap only
      if ((rc = hawk_check_cuda(cudaMemsetAsync(b->q.p, 0, n_chunks * 16, st), "gap memset"))) break;
      if ((rc = hawk_check_cuda(cudaMemsetAsync(b->v.p, 0, n_chunks * 4, st), "gap memset"))) break;
    } else if (total > 0) {
      const uint8_t* src = ascii;
      if (!ascii_on_device) {
        if ((rc = upload(c, d_ascii, ascii, (size_t)total))) break;
        src = d_ascii.as<uint8_t>();
      }
      int64_t init = INT64_MAX;
      if ((rc = upload(c, d_bad, &init, 8))) break;
      cudaEvent_t ev;
      c->mark(0, &ev);
      rc = hawk_pack_dev(st, src, total, b->q.p, b->v.as<uint32_t>(), b->nz.as<uint32_t>(), d_bad.as<int64_t>());
      c->close_mark();
      if (rc) break;
      int64_t bad = INT64_MAX;
      if ((rc = c->small_d2h_sync(&bad, d_bad.p, 8))) break;
      if (bad != INT64_MAX) {
        if (bad_slot) *bad_slot = bad;
        rc = hawk_fail(HAWK_EIUPAC, "non-IUPAC character at slot %lld", (long long)bad);
        break;
      }
    }
  } while (0);
  if (rc != HAWK_OK) {
    delete b;
    return rc;
  }
  *out = b;
  return HAWK_OK;
}

extern "C" int hawk_batch_create(hawk_ctx* c, const uint8_t* ascii, const int64_t* slot_off,
                                 const int32_t* len, int32_t n_hap, hawk_batch** out,
                                 int64_t* bad_slot) {
  return batch_create_impl(c, ascii, false, slot_off, len, n_hap, out, bad_slot);
}

extern "C" int hawk_batch_create_dev(hawk_ctx* c, const uint8_t* d_ascii, const int64_t* slot_off,
                                     const int32_t* len, int32_t n_hap, hawk_batch** out,
                                     int64_t* bad_slot) {
  if ((uintptr_t)d_ascii & 15) return hawk_fail(HAWK_EINVAL, "hawk_batch_create_dev: d_ascii must be 16-byte aligned");
  return batch_create_impl(c, d_ascii, true, slot_off, len, n_hap, out, bad_slot);
}

// ---- N1: haplotypes as edit lists against the reference text --------------------------------
extern "C" int hawk_batch_create_from_edits(hawk_ctx* c, const uint8_t* ref_ascii, int64_t ref_len,
                                            int32_t region_start, int32_t n_hap, const int64_t* edit_off,
                                            const int32_t* edit_pos, const int32_t* edit_reflen,
                                            const int32_t* edit_altlen, const int64_t* edit_altoff,
                                            const uint8_t* alt_pool, int64_t alt_pool_len, hawk_batch** out,
                                            int64_t* bad_slot) {
  if (!c || !out || !ref_ascii || ref_len <= 0 || n_hap < 0 || (n_hap > 0 && !edit_off))
    return hawk_fail(HAWK_EINVAL, "hawk_batch_create_from_edits: bad arguments");
  if (ref_len > INT32_MAX - 4096) return hawk_fail(HAWK_EINVAL, "hawk_batch_create_from_edits: reference too long");
  CKCUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  Trace tr;
  const int64_t n_edits = n_hap ? edit_off[n_hap] : 0;
  const size_t ne = (size_t)(n_edits > 0 ? n_edits : 1), nh = (size_t)(n_hap > 0 ? n_hap : 1);
  DevBuf d_ref, d_eoff, d_pos, d_rl, d_al, d_ao, d_op, d_eh, d_pool, d_so, d_len, d_segcnt, d_bad, d_ascii;
  // the reference text with readable slack behind it (block copies read whole aligned words)
  CK(d_ref.alloc(c, (size_t)ref_len + 32));
  CKCUDA(cudaMemsetAsync(d_ref.as<uint8_t>() + ref_len, 0, 32, st));
  CK(c->small_h2d(d_ref.p, ref_ascii, (size_t)ref_len));
  const int64_t zero64 = 0;
  const int32_t zero32 = 0;
  const uint8_t zero8 = 0;
  CK(upload(c, d_eoff, n_hap ? (const void*)edit_off : (const void*)&zero64, (size_t)(n_hap + 1) * 8));
  CK(upload(c, d_pos, n_edits ? (const void*)edit_pos : (const void*)&zero32, ne * 4));
  CK(upload(c, d_rl, n_edits ? (const void*)edit_reflen : (const void*)&zero32, ne * 4));
  CK(upload(c, d_al, n_edits ? (const void*)edit_altlen : (const void*)&zero32, ne * 4));
  CK(upload(c, d_ao, n_edits ? (const void*)edit_altoff : (const void*)&zero64, ne * 8));
  CK(upload(c, d_pool, alt_pool_len > 0 ? alt_pool : &zero8, (size_t)(alt_pool_len > 0 ? alt_pool_len : 1)));
  CK(d_op.alloc(c, ne * 4));
  CK(d_eh.alloc(c, ne * 4));
  CK(d_len.alloc(c, nh * 4));
  CK(d_segcnt.alloc(c, nh * 4));
  const int32_t int_max = INT32_MAX;
  CK(upload(c, d_bad, &int_max, 4));
  // pass 0 on the device: validation, output positions, lengths, segment counts
  CK(launch_derive(st, n_hap, d_eoff.as<int64_t>(), d_pos.as<int32_t>(), d_rl.as<int32_t>(), d_al.as<int32_t>(),
                   d_ao.as<int64_t>(), ref_len, alt_pool_len, region_start, d_op.as<int32_t>(), d_len.as<int32_t>(),
                   d_segcnt.as<int32_t>(), d_bad.as<int32_t>(), nullptr, nullptr, nullptr, nullptr, 0, d_eh.as<int32_t>()));
  // the reference's own planes (the windows around the edits are cut from them) and the ALT pool
  // check; anything unusual -- a non-IUPAC character, lower-case (soft-masked) reference bases,
  // which every haplotype would inherit as variant bases -- goes the long way round (texts + K1)
  DevBuf d_refq, d_refv, d_refnz, d_flags;
  const int64_t ref_chunks = (ref_len + HAWK_CHUNK - 1) / HAWK_CHUNK;
  bool windows_ok = c->edit_planes == 1 && n_hap > 0 && getenv("HAWK_DENSE_EDITS") == nullptr;
  std::vector<uint32_t> ref_nz;
  unsigned long long flags[2] = {~0ull, ~0ull};  // first bad reference slot, first bad pool offset
  if (windows_ok) {
    // text padded with zeros to whole chunks (+ one chunk the funnel shifts may read)
    DevBuf d_reftext;
    const size_t padded = (size_t)(ref_chunks + 1) * HAWK_CHUNK;
    CK(d_reftext.alloc(c, padded));
    CKCUDA(cudaMemsetAsync(d_reftext.as<uint8_t>() + ref_len, 0, padded - (size_t)ref_len, st));
    CKCUDA(cudaMemcpyAsync(d_reftext.p, d_ref.p, (size_t)ref_len, cudaMemcpyDeviceToDevice, st));
    CK(d_refq.alloc(c, (size_t)(ref_chunks + 1) * 16));
    CK(d_refv.alloc(c, (size_t)(ref_chunks + 1) * 4));
    const size_t nzw = (size_t)(ref_chunks + 1 + 31) / 32;
    CK(d_refnz.alloc(c, nzw * 4, true));
    CK(upload(c, d_flags, flags, 16));
    CK(hawk_pack_dev(st, d_reftext.as<uint8_t>(), (int64_t)padded, d_refq.p, d_refv.as<uint32_t>(), d_refnz.as<uint32_t>(),
                     d_flags.as<int64_t>()));
    CK(launch_pool_check(st, d_pool.as<uint8_t>(), alt_pool_len, d_flags.as<unsigned long long>() + 1));
    ref_nz.resize(nzw);
    CK(c->small_d2h_sync(ref_nz.data(), d_refnz.p, nzw * 4));
    CK(c->small_d2h_sync(flags, d_flags.p, 16));
    for (uint32_t w : ref_nz) windows_ok = windows_ok && w == 0;
    windows_ok = windows_ok && flags[0] == ~0ull && flags[1] == ~0ull;
  }
  std::vector<int32_t> len(n_hap), seg_count(n_hap);
  int32_t bad_hap = INT32_MAX;
  if (n_hap) {
    CK(c->small_d2h_sync(len.data(), d_len.p, (size_t)n_hap * 4));
    CK(c->small_d2h_sync(seg_count.data(), d_segcnt.p, (size_t)n_hap * 4));
  }
  CK(c->small_d2h_sync(&bad_hap, d_bad.p, 4));
  tr.tick("edits: upload + derive");
  if (bad_hap != INT32_MAX)
    return hawk_fail(HAWK_EINVAL,
                     "hawk_batch_create_from_edits: the edits of haplotype %d are not sorted, non-overlapping "
                     "SNVs / anchored insertions / anchored deletions inside the reference", bad_hap);
  std::vector<int64_t> slot_off(n_hap + 1), seg_off(n_hap + 1);
  int64_t total = 0;
  CK(hawk_layout(len.data(), n_hap, slot_off.data(), &total));
  seg_off[0] = 0;
  for (int32_t h = 0; h < n_hap; ++h) seg_off[h + 1] = seg_off[h] + seg_count[h];
  CK(upload(c, d_so, slot_off.data(), (size_t)(n_hap + 1) * 8));
  hawk_batch* b = nullptr;
  if (windows_ok) {
    // planes only where a search reads them, built at the first search (hawk_edits_ensure)
    CK(batch_create_impl(c, nullptr, true, slot_off.data(), len.data(), n_hap, &b, bad_slot, true));
    b->edits_lazy = true;
    b->ref_len = ref_len;
    b->ref_chunks = ref_chunks;
    b->n_edits = n_edits;
    b->has_edits.resize(n_hap);
    b->h_edit_off.assign(edit_off, edit_off + n_hap + 1);
    std::vector<int32_t> plain;
    for (int32_t h = 0; h < n_hap; ++h) {
      b->has_edits[h] = edit_off[h + 1] > edit_off[h];
      if (!b->has_edits[h]) plain.push_back(h);
    }
    b->n_plain = (int32_t)plain.size();
    int rc = b->n_plain ? upload(c, b->d_plain, plain.data(), plain.size() * 4) : HAWK_OK;
    if (rc) {
      hawk_batch_destroy(b);
      return rc;
    }
    b->ref_text.move_from(d_ref);
    b->ref_q.move_from(d_refq);
    b->ref_v.move_from(d_refv);
    b->edit_outpos.move_from(d_op);
    b->edit_hap.move_from(d_eh);
  } else {
    CK(d_ascii.alloc(c, (size_t)total));
    CK(hawk_materialize_dev(st, d_ref.as<uint8_t>(), ref_len, d_eoff.as<int64_t>(), d_pos.as<int32_t>(),
                            d_rl.as<int32_t>(), d_al.as<int32_t>(), d_ao.as<int64_t>(), d_op.as<int32_t>(),
                            d_pool.as<uint8_t>(), d_so.as<int64_t>(), d_len.as<int32_t>(), n_hap, total, n_edits,
                            n_hap ? *std::max_element(len.begin(), len.end()) : 0, d_ascii.as<uint8_t>()));
    CK(batch_create_impl(c, d_ascii.as<uint8_t>(), true, slot_off.data(), len.data(), n_hap, &b, bad_slot));
  }
  tr.tick("edits: materialise + pack");
  // pass 1: the run-length coordinate maps, straight into the batch
  const size_t n_seg = (size_t)seg_off[n_hap];
  int rc = upload(c, b->seg_off, seg_off.data(), (size_t)(n_hap + 1) * 8);
  if (!rc) rc = b->seg_rel.alloc(c, (n_seg ? n_seg : 1) * 4);
  if (!rc) rc = b->seg_gen.alloc(c, (n_seg ? n_seg : 1) * 4);
  if (!rc) rc = b->seg_step.alloc(c, n_seg ? n_seg : 1);
  if (!rc)
    rc = launch_derive(st, n_hap, d_eoff.as<int64_t>(), d_pos.as<int32_t>(), d_rl.as<int32_t>(), d_al.as<int32_t>(),
                       d_ao.as<int64_t>(), ref_len, alt_pool_len, region_start, nullptr, nullptr, nullptr, nullptr,
                       b->seg_off.as<int64_t>(), b->seg_rel.as<int32_t>(), b->seg_gen.as<int32_t>(),
                       b->seg_step.as<uint8_t>(), 1);
  if (!rc) rc = hawk_check_cuda(cudaStreamSynchronize(st), "segments sync");
  if (rc != HAWK_OK) {
    hawk_batch_destroy(b);
    return rc;
  }
  // N2: anchored edits are the reference's normalised variants (variant.py:456-486) already
  b->var_off.move_from(d_eoff);
  b->var_pos.move_from(d_pos);
  b->var_rl.move_from(d_rl);
  b->var_al.move_from(d_al);
  b->var_ao.move_from(d_ao);
  b->var_pool.move_from(d_pool);
  b->var_pos_base = region_start;
  b->has_variants = true;
  b->h_seg_off = seg_off;
  b->first_gen.assign(n_hap, region_start);
  b->linear.resize(n_hap);
  for (int32_t h = 0; h < n_hap; ++h) b->linear[h] = seg_count[h] == 1 ? 1 : 0;
  b->gmin = region_start;
  b->gmax = (int32_t)(region_start + ref_len - 1);  // edits never create coordinates outside the reference
  b->has_posmap = true;
  tr.tick("edits: segments");
  *out = b;
  return HAWK_OK;
}

int hawk_edits_ensure(hawk_batch* b, int need) {
  if (!b || !b->edits_lazy) return HAWK_OK;
  if (need > HAWK_SLOT_GAP / HAWK_CHUNK) need = HAWK_EDITS_DENSE;  // a window may not cross into the next haplotype
  if (need != HAWK_EDITS_DENSE && b->edits_reach >= need) return HAWK_OK;
  hawk_ctx* c = b->ctx;
  CKCUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  if (b->total_slots == 0) return HAWK_OK;
  Trace tr;
  if (need == HAWK_EDITS_DENSE) {
    // every text, then K1 (synth_kernels.cu, scan_kernels.cu)
    DevBuf d_ascii, d_bad;
    CK(d_ascii.alloc(c, (size_t)b->total_slots));
    int32_t max_len = 0;
    for (int32_t l : b->len) max_len = l > max_len ? l : max_len;
    CK(hawk_materialize_dev(st, b->ref_text.as<uint8_t>(), b->ref_len, b->var_off.as<int64_t>(), b->var_pos.as<int32_t>(),
                            b->var_rl.as<int32_t>(), b->var_al.as<int32_t>(), b->var_ao.as<int64_t>(),
                            b->edit_outpos.as<int32_t>(), b->var_pool.as<uint8_t>(), b->d_slot_off.as<int64_t>(),
                            b->d_len.as<int32_t>(), b->n_hap, b->total_slots, b->n_edits, max_len, d_ascii.as<uint8_t>()));
    const int64_t init = INT64_MAX;
    CK(upload(c, d_bad, &init, 8));
    CK(hawk_pack_dev(st, d_ascii.as<uint8_t>(), b->total_slots, b->q.p, b->v.as<uint32_t>(), b->nz.as<uint32_t>(),
                     d_bad.as<int64_t>()));
    CKCUDA(cudaStreamSynchronize(st));  // d_ascii is released on return
    b->edits_lazy = false;
    b->sparse = false;
    tr.tick("edits: all planes (texts + K1)");
    return HAWK_OK;
  }
  const size_t nz_words = ((size_t)b->total_slots / HAWK_CHUNK + 31) / 32 + HAWK_SLACK_CHUNKS;
  CKCUDA(cudaMemsetAsync(b->nz.p, 0, nz_words * 4, st));
  CK(launch_edits_plain(st, b->ref_q.p, b->ref_v.as<uint32_t>(), b->ref_chunks, b->d_plain.as<int32_t>(), b->n_plain,
                        b->d_slot_off.as<int64_t>(), b->q.p, b->v.as<uint32_t>()));
  CK(launch_edit_windows(st, b->ref_q.p, b->var_off.as<int64_t>(), b->var_pos.as<int32_t>(), b->var_rl.as<int32_t>(),
                         b->var_al.as<int32_t>(), b->var_ao.as<int64_t>(), b->edit_outpos.as<int32_t>(),
                         b->edit_hap.as<int32_t>(), b->var_pool.as<uint8_t>(), b->d_slot_off.as<int64_t>(),
                         b->d_len.as<int32_t>(), b->n_hap, b->n_edits, b->q.p, b->v.as<uint32_t>(), b->nz.as<uint32_t>(), need));
  b->edits_reach = need;
  b->sparse = true;
  b->sparse_reach = need;
  tr.tick("edits: window planes launched");
  return HAWK_OK;
}

extern "C" int hawk_ctx_set_edit_planes(hawk_ctx* c, int32_t mode) {
  if (!c || mode < 0 || mode > 1) return hawk_fail(HAWK_EINVAL, "hawk_ctx_set_edit_planes: mode must be 0 or 1");
  c->edit_planes = mode;
  return HAWK_OK;
}

extern "C" int hawk_batch_layout(hawk_batch* b, int64_t* slot_off, int32_t* len) {
  if (!b) return hawk_fail(HAWK_EINVAL, "hawk_batch_layout: null batch");
  if (slot_off) memcpy(slot_off, b->slot_off.data(), (size_t)(b->n_hap + 1) * 8);
  if (len && b->n_hap) memcpy(len, b->len.data(), (size_t)b->n_hap * 4);
  return HAWK_OK;
}

extern "C" int hawk_batch_repack_dev(hawk_batch* b, const uint8_t* d_ascii, int64_t* bad_slot) {
  if (!b || !d_ascii || ((uintptr_t)d_ascii & 15))
    return hawk_fail(HAWK_EINVAL, "hawk_batch_repack_dev: bad arguments");
  hawk_ctx* c = b->ctx;
  CKCUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  if (bad_slot) *bad_slot = -1;
  if (b->total_slots == 0) return HAWK_OK;
  b->edits_lazy = false;  // the planes now stand for the caller's texts
  DevBuf d_bad;
  int64_t init = INT64_MAX;
  CK(upload(c, d_bad, &init, 8));
  cudaEvent_t ev;
  c->mark(0, &ev);
  int rc = hawk_pack_dev(st, d_ascii, b->total_slots, b->q.p, b->v.as<uint32_t>(), b->nz.as<uint32_t>(),
                         d_bad.as<int64_t>());
  c->close_mark();
  CK(rc);
  int64_t bad = INT64_MAX;
  CK(c->small_d2h_sync(&bad, d_bad.p, 8));
  if (bad != INT64_MAX) {
    if (bad_slot) *bad_slot = bad;
    return hawk_fail(HAWK_EIUPAC, "non-IUPAC character at slot %lld", (long long)bad);
  }
  b->sparse = false;
  return HAWK_OK;
}

extern "C" int hawk_encode_search_dev(hawk_ctx* c, hawk_batch* b, const uint8_t* d_ascii, const hawk_params* params,
                                      const int32_t* scan_start, const int32_t* scan_stop, const uint8_t* is_ref,
                                      hawk_result** out, int64_t* bad_slot) {
  if (!b || !d_ascii || ((uintptr_t)d_ascii & 15))
    return hawk_fail(HAWK_EINVAL, "hawk_encode_search_dev: bad arguments (texts must be 16-byte aligned device memory)");
  return hawk_search_impl(c, b, params, scan_start, scan_stop, is_ref, nullptr, out, d_ascii, bad_slot);
}

extern "C" void* hawk_ctx_stream(hawk_ctx* c) { return c ? (void*)c->stream : nullptr; }

extern "C" int hawk_ctx_sync(hawk_ctx* c) {
  if (!c) return hawk_fail(HAWK_EINVAL, "hawk_ctx_sync: null context");
  CKCUDA(cudaSetDevice(c->device));
  CKCUDA(cudaStreamSynchronize(c->stream));
  return HAWK_OK;
}

extern "C" int hawk_ctx_set_fused(hawk_ctx* c, int32_t mode) {
  if (!c || mode < 0 || mode > 2) return hawk_fail(HAWK_EINVAL, "hawk_ctx_set_fused: mode must be 0, 1 or 2");
  c->fused_mode = mode;
  return HAWK_OK;
}

extern "C" int hawk_ctx_set_profiling(hawk_ctx* c, int32_t enabled) {
  if (!c) return hawk_fail(HAWK_EINVAL, "hawk_ctx_set_profiling: null context");
  c->profiling = enabled != 0;
  return HAWK_OK;
}

extern "C" int hawk_ctx_profile(hawk_ctx* c, double* ms, int64_t* n) {
  if (!c || !ms || !n) return hawk_fail(HAWK_EINVAL, "hawk_ctx_profile: bad arguments");
  CKCUDA(cudaSetDevice(c->device));
  CKCUDA(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < 5; ++i) { ms[i] = 0; n[i] = 0; }
  for (auto& s : c->spans) {
    float t = 0;
    if (cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess && s.kind >= 0 && s.kind < 5) {
      ms[s.kind] += t;
      n[s.kind] += 1;
    }
    cudaEventDestroy(s.a);
    cudaEventDestroy(s.b);
  }
  c->spans.clear();
  return HAWK_OK;
}

extern "C" int hawk_batch_destroy(hawk_batch* b) {
  if (!b) return HAWK_OK;
  cudaSetDevice(b->ctx->device);
  delete b;
  return HAWK_OK;
}

extern "C" int hawk_batch_export_nibbles(hawk_batch* b, int32_t hap, uint8_t* nibbles, uint8_t* lower) {
  if (!b || hap < 0 || hap >= b->n_hap || !nibbles)
    return hawk_fail(HAWK_EINVAL, "hawk_batch_export_nibbles: bad arguments");
  hawk_ctx* c = b->ctx;
  CKCUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  int32_t L = b->len[hap];
  if (L == 0) return HAWK_OK;
  DevBuf dn, dl, t_ascii, t_q, t_v, t_nz, t_so, t_bad;
  const void* q = b->q.p;
  const uint32_t* v = b->v.as<uint32_t>();
  int64_t chunk0 = b->slot_off[hap] >> 5;
  if (b->edits_lazy) {
    // an edit-list batch holds planes around the edits only: this ONE haplotype is materialised and
    // packed into scratch (synth_kernels.cu + K1, ~len bytes), the batch itself stays as it is
    const int32_t one_len = L;
    int64_t so[2], total = 0;
    CK(hawk_layout(&one_len, 1, so, &total));
    CK(t_ascii.alloc(c, (size_t)total));
    CK(upload(c, t_so, so, 16));
    CK(hawk_materialize_range(st, b->ref_text.as<uint8_t>(), b->ref_len, b->var_off.as<int64_t>() + hap,
                              b->var_pos.as<int32_t>(), b->var_rl.as<int32_t>(), b->var_al.as<int32_t>(), b->var_ao.as<int64_t>(),
                              b->edit_outpos.as<int32_t>(), b->var_pool.as<uint8_t>(), t_so.as<int64_t>(),
                              b->d_len.as<int32_t>() + hap, 1, total, b->h_edit_off[hap], b->h_edit_off[hap + 1], L,
                              t_ascii.as<uint8_t>()));
    const size_t n_chunks = (size_t)total / HAWK_CHUNK;
    CK(t_q.alloc(c, n_chunks * 16));
    CK(t_v.alloc(c, n_chunks * 4));
    CK(t_nz.alloc(c, (n_chunks + 31) / 32 * 4 + 16));
    const int64_t init = INT64_MAX;
    CK(upload(c, t_bad, &init, 8));
    CK(hawk_pack_dev(st, t_ascii.as<uint8_t>(), total, t_q.p, t_v.as<uint32_t>(), t_nz.as<uint32_t>(), t_bad.as<int64_t>()));
    q = t_q.p;
    v = t_v.as<uint32_t>();
    chunk0 = so[0] >> 5;
  } else if (b->sparse) {
    return hawk_fail(HAWK_EINVAL, "hawk_batch_export_nibbles: the batch keeps planes only around variant bases");
  }
  CK(dn.alloc(c, L));
  if (lower) CK(dl.alloc(c, L));
  CK(launch_export_nibbles(st, q, v, chunk0, L, dn.as<uint8_t>(), lower ? dl.as<uint8_t>() : nullptr));
  CKCUDA(cudaMemcpyAsync(nibbles, dn.p, L, cudaMemcpyDeviceToHost, st));
  if (lower) CKCUDA(cudaMemcpyAsync(lower, dl.p, L, cudaMemcpyDeviceToHost, st));
  CKCUDA(cudaStreamSynchronize(st));
  return HAWK_OK;
}

extern "C" int hawk_batch_set_posmap(hawk_batch* b, const int64_t* seg_off, const int32_t* seg_rel,
                                     const int32_t* seg_gen, const uint8_t* seg_step) {
  if (!b || !seg_off) return hawk_fail(HAWK_EINVAL, "hawk_batch_set_posmap: bad arguments");
  CKCUDA(cudaSetDevice(b->ctx->device));
  for (int32_t h = 0; h < b->n_hap; ++h) {
    if (seg_off[h + 1] <= seg_off[h])
      return hawk_fail(HAWK_EINVAL, "hawk_batch_set_posmap: haplotype %d has no segment", h);
    if (seg_rel[seg_off[h]] != 0)
      return hawk_fail(HAWK_EINVAL, "hawk_batch_set_posmap: first segment of haplotype %d must start at 0", h);
  }
  hawk_ctx* c = b->ctx;
  cudaStream_t st = c->stream;
  size_t n = (size_t)seg_off[b->n_hap];
  b->h_seg_off.assign(seg_off, seg_off + b->n_hap + 1);
  b->first_gen.resize(b->n_hap);
  b->linear.resize(b->n_hap);
  int64_t gmin = INT64_MAX, gmax = INT64_MIN;
  for (int32_t h = 0; h < b->n_hap; ++h) {
    const int64_t s0 = seg_off[h], s1 = seg_off[h + 1];
    b->first_gen[h] = seg_gen[s0];
    b->linear[h] = (s1 - s0 == 1 && seg_step[s0] == 1) ? 1 : 0;
    for (int64_t k = s0; k < s1; ++k) {
      const int64_t rel_end = k + 1 < s1 ? seg_rel[k + 1] : b->len[h];
      const int64_t lo = seg_gen[k], hi = seg_gen[k] + (seg_step[k] && rel_end > seg_rel[k] ? rel_end - seg_rel[k] - 1 : 0);
      if (lo < gmin) gmin = lo;
      if (hi > gmax) gmax = hi;
    }
  }
  if (gmin > gmax) gmin = 0, gmax = -1;
  if (gmin < INT32_MIN || gmax > INT32_MAX) return hawk_fail(HAWK_EINVAL, "hawk_batch_set_posmap: coordinates exceed 32 bits");
  b->gmin = (int32_t)gmin;
  b->gmax = (int32_t)gmax;
  b->seg_idx.release();  // the coarse index follows the segments: rebuilt at the next phased search
  b->seg_idx_stride = 0;
  CK(upload(c, b->seg_off, seg_off, (size_t)(b->n_hap + 1) * 8));
  CK(upload(c, b->seg_rel, seg_rel, n * 4));
  CK(upload(c, b->seg_gen, seg_gen, n * 4));
  CK(upload(c, b->seg_step, seg_step, n));
  CKCUDA(cudaStreamSynchronize(st));
  b->has_posmap = true;
  return HAWK_OK;
}

extern "C" int hawk_batch_set_alleles(hawk_batch* b, const int64_t* va_off, const int32_t* va_idx,
                                      const int64_t* va_ent_off, const uint8_t* va_ref) {
  if (!b || !va_off || !va_ent_off) return hawk_fail(HAWK_EINVAL, "hawk_batch_set_alleles: bad arguments");
  CKCUDA(cudaSetDevice(b->ctx->device));
  hawk_ctx* c = b->ctx;
  cudaStream_t st = c->stream;
  size_t n_sites = (size_t)va_off[b->n_hap];
  size_t n_ent = (size_t)va_ent_off[n_sites];
  CK(upload(c, b->va_off, va_off, (size_t)(b->n_hap + 1) * 8));
  CK(upload(c, b->va_idx, va_idx, n_sites * 4));
  CK(upload(c, b->va_ent_off, va_ent_off, (n_sites + 1) * 8));
  CK(upload(c, b->va_ref, va_ref, n_ent));
  CKCUDA(cudaStreamSynchronize(st));
  b->has_alleles = true;
  return HAWK_OK;
}

// ------------------------------------------------------------------ search
static BatchView batch_view(const hawk_batch* b, const int32_t* d_a, const int32_t* d_b,
                            const uint8_t* d_isref) {
  BatchView B{};
  B.q = b->q.as<Planes>();
  B.v = b->v.as<uint32_t>();
  B.nz = b->nz.as<uint32_t>();
  B.slot_off = b->d_slot_off.as<int64_t>();
  B.len = b->d_len.as<int32_t>();
  B.scan_start = d_a;
  B.scan_stop = d_b;
  B.is_ref = d_isref;
  B.n_hap = b->n_hap;
  B.seg_off = b->seg_off.as<int64_t>();
  B.seg_rel = b->seg_rel.as<int32_t>();
  B.seg_gen = b->seg_gen.as<int32_t>();
  B.seg_step = b->seg_step.as<uint8_t>();
  B.va_off = b->va_off.as<int64_t>();
  B.va_idx = b->va_idx.as<int32_t>();
  B.va_ent_off = b->va_ent_off.as<int64_t>();
  B.va_ref = b->va_ref.as<uint8_t>();
  return B;
}

struct ScanOut {
  DevBuf hits[2];
  int64_t n[2] = {0, 0};
  int64_t scanned_bp = 0;
};

// K2: candidates -> match -> records; two host round trips read the totals that size the
// next stage (candidates, then hits), everything else is stream-ordered
struct ScanInputs {  // per-search arrays on the device, carved out of one upload
  DevBuf buf;
  bool cached = false;  // the bounds are the batch's own (hawk_batch_set_scan): nothing to upload
  int32_t *a = nullptr, *b = nullptr;
  uint8_t* is_ref = nullptr;
  int64_t* sblock_off = nullptr;
};

static inline size_t al8z(size_t x) { return (x + 7) & ~(size_t)7; }
struct ScanLayout {
  size_t o_sb, o_a, o_b, o_ref, bytes;
  explicit ScanLayout(int32_t n_hap) {
    o_sb = 0;
    o_a = o_sb + (size_t)(n_hap + 1) * 8;
    o_b = o_a + al8z((size_t)n_hap * 4);
    o_ref = o_b + al8z((size_t)n_hap * 4);
    bytes = o_ref + al8z((size_t)n_hap);
  }
};

// sub-ranges of the fused kernel that overlap a REF haplotype's territory (full-size entry segments)
static int64_t count_dense_subs(const hawk_batch* b, const uint8_t* is_ref) {
  const int64_t n_chunks = b->total_slots / HAWK_CHUNK;
  int64_t n_dense = 0, last = -1;
  for (int32_t h = 0; h < b->n_hap; ++h) {
    if (!is_ref[h]) continue;
    const int64_t t0 = (b->slot_off[h] >> 5) - HAWK_SLOT_GAP / HAWK_CHUNK;
    const int64_t t1 = h + 1 < b->n_hap ? (b->slot_off[h + 1] >> 5) - HAWK_SLOT_GAP / HAWK_CHUNK : n_chunks;
    const int32_t sub = fused_sub_size(n_chunks);
    int64_t s0 = t0 / sub, s1 = (t1 - 1) / sub;
    if (s0 <= last) s0 = last + 1;
    if (s1 >= s0) {
      n_dense += s1 - s0 + 1;
      last = s1;
    }
  }
  return n_dense;
}

extern "C" int hawk_batch_set_scan(hawk_batch* b, const int32_t* scan_start, const int32_t* scan_stop,
                                   const uint8_t* is_ref) {
  if (!b || (b->n_hap > 0 && (!scan_start || !scan_stop || !is_ref)))
    return hawk_fail(HAWK_EINVAL, "hawk_batch_set_scan: bad arguments");
  hawk_ctx* c = b->ctx;
  CKCUDA(cudaSetDevice(c->device));
  const int32_t n = b->n_hap;
  b->has_scan = false;
  b->h_scan_a.assign(scan_start, scan_start + n);
  b->h_scan_b.assign(scan_stop, scan_stop + n);
  b->h_scan_ref.assign(is_ref, is_ref + n);
  const ScanLayout L(n);
  std::vector<char> host(L.bytes, 0);
  b->scan_sblocks = hawk_scan_plan(scan_start, scan_stop, n, (int64_t*)(host.data() + L.o_sb));
  b->scan_bp = 0;
  b->scan_ref_h = -1;
  b->scan_n_ref = 0;
  for (int32_t h = 0; h < n; ++h) {
    const int64_t a = scan_start[h] < 0 ? 0 : scan_start[h], e = scan_stop[h];
    if (e > a) b->scan_bp += e - a;
    if (is_ref[h]) {
      if (b->scan_ref_h < 0) b->scan_ref_h = h;
      ++b->scan_n_ref;
    }
  }
  if (n > 0) {
    memcpy(host.data() + L.o_a, scan_start, (size_t)n * 4);
    memcpy(host.data() + L.o_b, scan_stop, (size_t)n * 4);
    memcpy(host.data() + L.o_ref, is_ref, (size_t)n);
  }
  b->scan_dense_subs = count_dense_subs(b, is_ref);
  CK(b->d_scan.alloc(c, L.bytes));
  CKCUDA(cudaMemcpyAsync(b->d_scan.p, host.data(), L.bytes, cudaMemcpyHostToDevice, c->stream));
  CKCUDA(cudaStreamSynchronize(c->stream));
  c->h2d_bytes += (int64_t)L.bytes;
  b->has_scan = true;
  return HAWK_OK;
}

static int run_scan(hawk_ctx* c, hawk_batch* b, const hawk_params* params, const int32_t* scan_start,
                    const int32_t* scan_stop, const uint8_t* is_ref, int raw, ScanInputs& in, ScanOut& out) {
  cudaStream_t st = c->stream;
  Trace tr;
  const int32_t n_hap = b->n_hap;
  const ScanLayout L(n_hap);
  int64_t n_sblocks;
  char* dp;
  if (in.cached) {
    n_sblocks = b->scan_sblocks;
    out.scanned_bp = b->scan_bp;
    dp = (char*)b->d_scan.p;
  } else {
    char* hp = (char*)c->pinned_get(L.bytes);
    if (!hp) return hawk_fail(HAWK_ENOMEM, "pinned staging allocation failed");
    n_sblocks = hawk_scan_plan(scan_start, scan_stop, n_hap, (int64_t*)(hp + L.o_sb));
    out.scanned_bp = 0;
    for (int32_t h = 0; h < n_hap; ++h) {
      int64_t a = scan_start[h] < 0 ? 0 : scan_start[h], e = scan_stop[h];
      if (e > a) out.scanned_bp += e - a;
    }
    if (n_hap > 0) {
      memcpy(hp + L.o_a, scan_start, (size_t)n_hap * 4);
      memcpy(hp + L.o_b, scan_stop, (size_t)n_hap * 4);
      memcpy(hp + L.o_ref, is_ref, (size_t)n_hap);
    }
    CK(in.buf.alloc(c, L.bytes));
    CK(c->small_h2d(in.buf.p, hp, L.bytes));
    dp = (char*)in.buf.p;
  }
  in.sblock_off = (int64_t*)(dp + L.o_sb);
  in.a = (int32_t*)(dp + L.o_a);
  in.b = (int32_t*)(dp + L.o_b);
  in.is_ref = (uint8_t*)(dp + L.o_ref);
  for (int s = 0; s < 2; ++s) CK(out.hits[s].alloc(c, 16));
  if (n_sblocks == 0) return HAWK_OK;
  DevBuf d_ws, d_mws, d_masks;
  CK(d_ws.alloc(c, hawk_scan_workspace_bytes(n_hap, n_sblocks)));
  tr.tick("scan: plan + upload");
  ProfScope prof_scope(c);
  int rc = hawk_scan_count_dev(st, b->q.p, b->v.as<uint32_t>(), b->nz.as<uint32_t>(), b->d_slot_off.as<int64_t>(),
                               b->d_len.as<int32_t>(), in.a, in.b, in.is_ref, in.sblock_off, n_hap, n_sblocks, params,
                               raw, d_ws.p);
  CK(rc);
  uint64_t totals[8];
  CK(c->small_d2h_sync(totals, hawk_scan_totals(d_ws.p), 64));
  const int64_t n_cand = (int64_t)totals[0];
  tr.tick("scan: candidates");
  if (n_cand == 0) return HAWK_OK;
  CK(d_mws.alloc(c, hawk_scan_match_workspace_bytes(n_sblocks)));
  CK(d_masks.alloc(c, (size_t)n_cand * 8));
  rc = hawk_scan_match_dev(st, b->q.p, b->v.as<uint32_t>(), b->nz.as<uint32_t>(), b->d_slot_off.as<int64_t>(),
                           b->d_len.as<int32_t>(), in.a, in.b, in.is_ref, n_hap, n_sblocks, params, raw, n_cand,
                           d_masks.as<uint64_t>(), d_ws.p, d_mws.p);
  CK(rc);
  CK(c->small_d2h_sync(totals, hawk_scan_totals(d_ws.p), 64));
  out.n[0] = (int64_t)totals[1];
  out.n[1] = (int64_t)totals[2];
  tr.tick("scan: match");
  for (int s = 0; s < 2; ++s) CK(out.hits[s].alloc(c, (size_t)(out.n[s] > 0 ? out.n[s] : 1) * 8));
  rc = hawk_scan_expand_dev(st, n_hap, n_sblocks, n_cand, d_masks.as<uint64_t>(), d_ws.p, d_mws.p,
                            out.hits[0].as<uint64_t>(), out.hits[1].as<uint64_t>());
  CK(rc);
  tr.tick("scan: expand launched");
  return HAWK_OK;
}

// chunks either side of a variant chunk whose planes the stages after the scan can read: a hit
// needs a variant base inside its core, and its padded window reaches G + PAD before / C + PAD - 1
// behind the PAM position, so no base further than G + C + PAD - 1 from a variant base is read
// -- and never fewer than 2 chunks: a candidate chunk lies within one chunk of a variant chunk and
// its matcher reads the case words of both of ITS neighbours
static int fused_reach(const ScanConst& K) {
  const int r = (31 + K.G + K.C + HAWK_GUIDESEQPAD - 1) >> 5;
  return r < 2 ? 2 : r;
}

// K1 + K2 fused (fused_kernels.cu): the texts are read once, planes are stored only where later
// stages read them, hit entries come out per warp sub-range; one host round trip (hit totals,
// the first non-IUPAC slot, the overflow flag), then the entries become the two record streams.
static int run_scan_fused(hawk_ctx* c, hawk_batch* b, const uint8_t* d_ascii, const hawk_params* params,
                          const int32_t* scan_start, const int32_t* scan_stop, const uint8_t* is_ref, ScanInputs& in,
                          ScanOut& out, int64_t* bad_slot, bool* fell_back) {
  cudaStream_t st = c->stream;
  Trace tr;
  const int32_t n_hap = b->n_hap;
  const ScanConst K = make_scan_const(*params, 0);
  const int64_t n_chunks = b->total_slots / HAWK_CHUNK;
  *fell_back = false;
  // [1..2] hit totals, [3] overflow flag, [4] first bad slot
  uint64_t tot0[8] = {0, 0, 0, 0, (uint64_t)INT64_MAX, 0, 0, 0};
  const int64_t n_sub = fused_sub_ranges(n_chunks);
  const int all_dense = K.unphased;
  int64_t n_dense;
  uint64_t* d_tot;
  if (in.cached) {  // the batch's own bounds (hawk_batch_set_scan): only the totals block goes up
    const ScanLayout L(n_hap);
    char* dp = (char*)b->d_scan.p;
    in.a = (int32_t*)(dp + L.o_a);
    in.b = (int32_t*)(dp + L.o_b);
    in.is_ref = (uint8_t*)(dp + L.o_ref);
    out.scanned_bp = b->scan_bp;
    n_dense = all_dense ? n_sub : b->scan_dense_subs;
    CK(in.buf.alloc(c, 64));
    CK(c->small_h2d(in.buf.p, tot0, 64));
    d_tot = (uint64_t*)in.buf.p;
  } else {
    const size_t o_a = 0, o_b = o_a + al8z((size_t)n_hap * 4), o_ref = o_b + al8z((size_t)n_hap * 4),
                 o_tot = o_ref + al8z((size_t)n_hap), total_bytes = o_tot + 64;
    char* hp = (char*)c->pinned_get(total_bytes);
    if (!hp) return hawk_fail(HAWK_ENOMEM, "pinned staging allocation failed");
    out.scanned_bp = 0;
    for (int32_t h = 0; h < n_hap; ++h) {
      int64_t a = scan_start[h] < 0 ? 0 : scan_start[h], e = scan_stop[h];
      if (e > a) out.scanned_bp += e - a;
    }
    memcpy(hp + o_a, scan_start, (size_t)n_hap * 4);
    memcpy(hp + o_b, scan_stop, (size_t)n_hap * 4);
    memcpy(hp + o_ref, is_ref, (size_t)n_hap);
    memcpy(hp + o_tot, tot0, 64);
    CK(in.buf.alloc(c, total_bytes));
    CK(c->small_h2d(in.buf.p, hp, total_bytes));
    char* dp = (char*)in.buf.p;
    in.a = (int32_t*)(dp + o_a);
    in.b = (int32_t*)(dp + o_b);
    in.is_ref = (uint8_t*)(dp + o_ref);
    d_tot = (uint64_t*)(dp + o_tot);
    n_dense = all_dense ? n_sub : count_dense_subs(b, is_ref);
  }
  for (int s = 0; s < 2; ++s) CK(out.hits[s].alloc(c, 16));
  const size_t sub_size = (size_t)fused_sub_size(n_chunks);
  const size_t total_cap = (size_t)(n_sub - n_dense) * (sub_size / 4) + (size_t)n_dense * sub_size;
  DevBuf d_hs, d_cap, d_segbase, d_tiles, d_entries, d_cnt, d_base;
  CK(d_hs.alloc(c, (size_t)(n_hap > 0 ? n_hap : 1) * sizeof(HapScan)));
  CK(d_cap.alloc(c, (size_t)(n_sub + 1) * 4));
  CK(d_segbase.alloc(c, (size_t)(n_sub + 1) * 8));
  CK(d_tiles.alloc(c, ((size_t)scan_tiles(n_sub) + 2) * 8));
  CK(d_entries.alloc(c, (total_cap ? total_cap : 1) * 16));
  CK(d_cnt.alloc(c, (size_t)(n_sub + 1) * 4 * 3));
  CK(d_base.alloc(c, (size_t)(n_sub + 1) * 8 * 2));
  uint32_t* cnt_ent = d_cnt.as<uint32_t>();
  uint32_t* cnt_h0 = cnt_ent + (n_sub + 1);
  uint32_t* cnt_h1 = cnt_h0 + (n_sub + 1);
  uint64_t* base0 = d_base.as<uint64_t>();
  uint64_t* base1 = base0 + (n_sub + 1);
  tr.tick("fused: upload");
  ProfScope prof_scope(c);
  BatchView B = batch_view(b, in.a, in.b, in.is_ref);
  hawk_prof_begin(st, 1);
  CK(launch_hapscan(st, B, K, d_hs.as<HapScan>()));
  CK(launch_fused_caps(st, b->d_slot_off.as<int64_t>(), in.is_ref, n_hap, n_chunks, all_dense, d_cap.as<uint32_t>()));
  CK(exclusive_scan_u32(st, d_cap.as<uint32_t>(), n_sub, d_segbase.as<uint64_t>(), d_tiles.as<uint64_t>(), nullptr));
  hawk_prof_end(st);
  FusedLaunch L{};
  L.ascii = d_ascii;
  L.n_chunks = n_chunks;
  L.q = b->q.p;
  L.v = b->v.as<uint32_t>();
  L.nz = b->nz.as<uint32_t>();
  L.slot_off = b->d_slot_off.as<int64_t>();
  L.n_hap = n_hap;
  L.hs = d_hs.as<HapScan>();
  L.K = K;
  L.reach = fused_reach(K);
  L.store_all = 0;
  L.seg_base = d_segbase.as<uint64_t>();
  L.seg_cap = d_cap.as<uint32_t>();
  L.entries = d_entries.p;
  L.cnt_ent = cnt_ent;
  L.cnt_hit0 = cnt_h0;
  L.cnt_hit1 = cnt_h1;
  L.overflow = (uint32_t*)(d_tot + 3);
  L.bad = (int64_t*)(d_tot + 4);
  cudaEvent_t ev;
  c->mark(0, &ev);
  int rc = launch_fused_scan(st, L);
  c->close_mark();
  CK(rc);
  b->sparse = true;
  b->sparse_reach = L.reach;
  hawk_prof_begin(st, 1);
  CK(exclusive_scan_u32(st, cnt_h0, n_sub, base0, d_tiles.as<uint64_t>(), d_tot + 1));
  CK(exclusive_scan_u32(st, cnt_h1, n_sub, base1, d_tiles.as<uint64_t>(), d_tot + 2));
  hawk_prof_end(st);
  uint64_t totals[8];
  CK(c->small_d2h_sync(totals, d_tot, 64));
  tr.tick("fused: scan + sync");
  if ((int64_t)totals[4] != INT64_MAX) {
    if (bad_slot) *bad_slot = (int64_t)totals[4];
    return hawk_fail(HAWK_EIUPAC, "non-IUPAC character at slot %lld", (long long)totals[4]);
  }
  if (totals[3]) {  // a segment overflowed (variants denser than one chunk in four): staged K2 on the kept planes
    *fell_back = true;
    return HAWK_OK;
  }
  out.n[0] = (int64_t)totals[1];
  out.n[1] = (int64_t)totals[2];
  for (int s = 0; s < 2; ++s) CK(out.hits[s].alloc(c, (size_t)(out.n[s] > 0 ? out.n[s] : 1) * 8));
  hawk_prof_begin(st, 3);
  CK(launch_fused_expand(st, d_entries.p, d_segbase.as<uint64_t>(), cnt_ent, base0, base1, n_sub,
                         out.hits[0].as<uint64_t>(), out.hits[1].as<uint64_t>()));
  hawk_prof_end(st);
  tr.tick("fused: expand launched");
  return HAWK_OK;
}

static int check_scan_args(hawk_ctx* c, hawk_batch* b, const hawk_params* p, const int32_t* a,
                           const int32_t* e, hawk_result** out) {
  if (!c || !b || !p || !out || (b->n_hap > 0 && (!a || !e)))
    return hawk_fail(HAWK_EINVAL, "search: bad arguments");
  if (b->ctx != c) return hawk_fail(HAWK_EINVAL, "search: batch belongs to another context");
  if (p->pam_len < 1 || p->pam_len > HAWK_MAX_PAM || p->guide_len < 1)
    return hawk_fail(HAWK_EINVAL, "search: PAM length must be 1..%d and guide length >= 1", HAWK_MAX_PAM);
  if (p->pam_len + p->guide_len + 2 * HAWK_GUIDESEQPAD > HAWK_MAX_WINDOW)
    return hawk_fail(HAWK_EINVAL, "search: guide + PAM window exceeds %d", HAWK_MAX_WINDOW);
  for (int i = 0; i < p->pam_len; ++i)
    if (p->pam_fwd[i] < 1 || p->pam_fwd[i] > 15 || p->pam_rc[i] < 1 || p->pam_rc[i] > 15)
      return hawk_fail(HAWK_EINVAL, "search: PAM nibble %d out of range", i);
  return HAWK_OK;
}

extern "C" int hawk_pam_search(hawk_ctx* c, hawk_batch* b, const hawk_params* params,
                               const int32_t* scan_start, const int32_t* scan_stop,
                               hawk_result** out) {
  CK(check_scan_args(c, b, params, scan_start, scan_stop, out));
  CK(hawk_edits_ensure(b, HAWK_EDITS_DENSE));  // the raw PAM scan reads every chunk
  if (b->sparse) return hawk_fail(HAWK_EINVAL, "hawk_pam_search: the batch keeps planes only around variant bases; re-encode it");
  CKCUDA(cudaSetDevice(c->device));
  hawk_result* r = new (std::nothrow) hawk_result();
  if (!r) return hawk_fail(HAWK_ENOMEM, "hawk_pam_search: host allocation");
  r->ctx = c;
  r->window = params->pam_len + params->guide_len + 2 * HAWK_GUIDESEQPAD;
  std::vector<uint8_t> isref(b->n_hap, 1);
  ScanInputs in;
  ScanOut so;
  int rc = run_scan(c, b, params, scan_start, scan_stop, isref.data(), 1, in, so);
  if (rc != HAWK_OK) {
    delete r;
    return rc;
  }
  for (int s = 0; s < 2; ++s) {
    r->n_hits[s] = so.n[s];
    r->hits[s].move_from(so.hits[s]);
  }
  r->scanned_bp = so.scanned_bp;
  *out = r;
  return HAWK_OK;
}

// phased / variant-free pipeline downstream of the scan (table_kernels.cu)
// Coarse index of the batch's posmap segments for rows_fast (one entry per haplotype and 4 kb of
// its text), built once per batch; skipped where it would be large (haplotypes of very unequal
// length: the stride follows the longest) -- row_coords then searches all segments as before.
static int ensure_seg_index(hawk_ctx* c, hawk_batch* b) {
  if (b->seg_idx_stride != 0 || b->n_hap <= 0) return HAWK_OK;
  int32_t max_len = 0;
  for (int32_t h = 0; h < b->n_hap; ++h) max_len = b->len[h] > max_len ? b->len[h] : max_len;
  const int32_t stride = (max_len >> HAWK_SEG_IDX_SHIFT) + 2;
  if ((int64_t)stride * b->n_hap > (64ll << 20)) {
    b->seg_idx_stride = -1;
    return HAWK_OK;
  }
  CK(b->seg_idx.alloc(c, (size_t)stride * b->n_hap * 4));
  CK(launch_seg_index(c->stream, b->seg_off.as<int64_t>(), b->seg_rel.as<int32_t>(), b->n_hap, stride,
                      b->seg_idx.as<int32_t>()));
  b->seg_idx_stride = stride;
  return HAWK_OK;
}

static int search_fast(hawk_ctx* c, hawk_batch* b, const ScanConst& K, const BatchView& B_in, ScanOut& so,
                       int32_t ref_h, const StreamLink* link, hawk_result* r) {
  cudaStream_t st = c->stream;
  Trace tr;
  BatchView B = B_in;
  CK(ensure_seg_index(c, b));
  if (b->seg_idx_stride > 0) {
    B.seg_idx = b->seg_idx.as<int32_t>();
    B.seg_idx_stride = b->seg_idx_stride;
  }
  const int64_t n_hits[2] = {so.n[0], so.n[1]};
  const uint64_t* recs[2] = {so.hits[0].as<uint64_t>(), so.hits[1].as<uint64_t>()};
  RefInfo ref{ref_h, 0, 0, 0};
  DevBuf d_refrange, d_refbm[2], start[2], stop[2], keep[2], blk_cnt[2], blk_base[2], d_tot, d_kb;
  CK(d_refrange.alloc(c, 32, true));
  CK(launch_ref_range(st, recs[0], n_hits[0], recs[1], n_hits[1], ref_h, d_refrange.as<int64_t>()));
  if (ref_h >= 0) {
    ref.linear = b->linear[ref_h];
    ref.g0 = b->first_gen[ref_h];
    ref.len = b->len[ref_h];
    if (ref.linear) {
      const size_t words = (size_t)(ref.len + 31) / 32 + 1;
      for (int s = 0; s < 2; ++s) CK(d_refbm[s].alloc(c, words * 4, true));
      CK(launch_ref_bitmap(st, recs[0], recs[1], d_refrange.as<int64_t>(), d_refbm[0].as<uint32_t>(),
                           d_refbm[1].as<uint32_t>()));
    }
  }
  int64_t n_blk[2];
  CK(d_tot.alloc(c, 16, true));
  for (int s = 0; s < 2; ++s) {
    n_blk[s] = row_blocks(n_hits[s]);
    CK(start[s].alloc(c, (size_t)n_hits[s] * 4));
    CK(stop[s].alloc(c, (size_t)n_hits[s] * 4));
    CK(keep[s].alloc(c, (size_t)n_hits[s]));
    CK(blk_cnt[s].alloc(c, (size_t)(n_blk[s] + 1) * 4));
    CK(blk_base[s].alloc(c, (size_t)(n_blk[s] + 1) * 8));
  }
  {
    const uint32_t* bm[2] = {d_refbm[0].as<uint32_t>(), d_refbm[1].as<uint32_t>()};
    int32_t* st_[2] = {start[0].as<int32_t>(), start[1].as<int32_t>()};
    int32_t* sp_[2] = {stop[0].as<int32_t>(), stop[1].as<int32_t>()};
    uint8_t* kp_[2] = {keep[0].as<uint8_t>(), keep[1].as<uint8_t>()};
    uint32_t* bc_[2] = {blk_cnt[0].as<uint32_t>(), blk_cnt[1].as<uint32_t>()};
    CK(launch_rows_fast(st, B, K, recs, n_hits, ref, bm, d_refrange.as<int64_t>(), link ? link->drop_ref : 0, st_, sp_,
                        kp_, bc_));
  }
  {
    const uint32_t* bc_[2] = {blk_cnt[0].as<uint32_t>(), blk_cnt[1].as<uint32_t>()};
    uint64_t* bb_[2] = {blk_base[0].as<uint64_t>(), blk_base[1].as<uint64_t>()};
    CK(launch_blk_prefix(st, bc_, n_blk, bb_, d_tot.as<uint64_t>()));
  }
  // no host round trip here: the table is allocated for the upper bound (every hit kept)
  // and the surviving-row totals are read once, after the last kernel
  const int64_t n_max = n_hits[0] + n_hits[1];
  const int W = K.C + 2 * HAWK_GUIDESEQPAD;
  r->text_stride = (W + 15) / 16 * 16;
  CK(r->hap.alloc(c, (size_t)n_max * 4));
  CK(r->strand.alloc(c, (size_t)n_max));
  CK(r->pos.alloc(c, (size_t)n_max * 4));
  CK(r->start.alloc(c, (size_t)n_max * 4));
  CK(r->stop.alloc(c, (size_t)n_max * 4));
  CK(r->bucket.alloc(c, (size_t)n_max * 4));
  CK(r->text.alloc(c, (size_t)n_max * r->text_stride));
  uint64_t kept[2] = {0, 0};
  if (n_max > 0) {
    CK(d_kb.alloc(c, (size_t)(b->n_hap + 1) * 16));
    CK(launch_hap_offsets(st, recs[0], recs[1], n_hits[0], n_hits[1], keep[0].as<uint8_t>(), keep[1].as<uint8_t>(),
                          blk_base[0].as<uint64_t>(), blk_base[1].as<uint64_t>(), d_tot.as<uint64_t>(), b->n_hap,
                          d_kb.as<uint64_t>()));
    // first-seen bucket ids: direct-address table over (start, strand) when the coordinate
    // range allows it, else the hash table of post_kernels.cu
    const int64_t key_span = b->gmax >= b->gmin ? ((int64_t)b->gmax - b->gmin + 1) * 2 : 0;
    const bool direct = link || (key_span > 0 && key_span <= HAWK_DIRECT_KEY_SPAN && n_max < 0xFFFFFFFFll);
    DevBuf key_table_own;
    uint32_t* key_table = nullptr;
    int32_t key_min = b->gmin;
    RowMap rm{0, -1, 0, 0};
    if (link) {  // one group of a streamed search: the key table and the row numbering are global
      if (link->row_base + n_max >= 0xFFFFFFFFll)
        return hawk_fail(HAWK_ECAPACITY, "streamed search: more than 2^32 guide rows");
      key_table = link->key_table;
      key_min = link->key_min;
      rm = RowMap{link->row_base, link->ref_local, link->ref_global, link->hap_add};
    } else if (direct) {
      CK(key_table_own.alloc(c, (size_t)key_span * 4));
      CKCUDA(cudaMemsetAsync(key_table_own.p, 0xFF, (size_t)key_span * 4, st));
      key_table = key_table_own.as<uint32_t>();
    }
    {
      const uint8_t* kp_[2] = {keep[0].as<uint8_t>(), keep[1].as<uint8_t>()};
      const uint64_t* bb_[2] = {blk_base[0].as<uint64_t>(), blk_base[1].as<uint64_t>()};
      const int32_t* st_[2] = {start[0].as<int32_t>(), start[1].as<int32_t>()};
      const int32_t* sp_[2] = {stop[0].as<int32_t>(), stop[1].as<int32_t>()};
      const uint64_t* kb_[2] = {d_kb.as<uint64_t>() + (size_t)(b->n_hap + 1), d_kb.as<uint64_t>()};
      CK(launch_gather_fast(st, B, K, recs, kp_, bb_, st_, sp_, kb_, n_hits, r->text_stride, r->hap.as<int32_t>(),
                            r->strand.as<uint8_t>(), r->pos.as<int32_t>(), r->start.as<int32_t>(), r->stop.as<int32_t>(),
                            r->text.as<uint8_t>(), direct ? key_table : nullptr, key_min, rm));
    }
    if (direct)
      CK(launch_bucket_read(st, r->start.as<int32_t>(), r->strand.as<uint8_t>(), n_max, d_tot.as<uint64_t>(),
                            key_table, key_min, r->bucket.as<uint32_t>()));
    c->close_mark();
    if (link) {
      int64_t rr[4];
      CK(c->small_d2h_sync(rr, d_refrange.p, 32));
      r->ref_hits[0] = rr[1] - rr[0];
      r->ref_hits[1] = rr[3] - rr[2];
    }
    CK(c->small_d2h_sync(kept, d_tot.p, 16));
    tr.tick("table: pipeline + sync");
    const int64_t n = (int64_t)(kept[0] + kept[1]);
    r->n_guides = n;
    if (!direct && n > 0) {
      uint64_t tsize = 1024;
      while (tsize < (uint64_t)n * 2) tsize <<= 1;
      DevBuf keys, vals;
      CK(keys.alloc(c, tsize * 8));
      CK(vals.alloc(c, tsize * 8));
      CKCUDA(cudaMemsetAsync(keys.p, 0xFF, tsize * 8, st));
      CKCUDA(cudaMemsetAsync(vals.p, 0xFF, tsize * 8, st));
      CK(launch_buckets(st, r->start.as<int32_t>(), r->strand.as<uint8_t>(), n, keys.as<unsigned long long>(),
                        vals.as<unsigned long long>(), tsize, r->bucket.as<uint32_t>()));
      CKCUDA(cudaStreamSynchronize(st));  // keys / vals are released on return
    }
  } else {
    c->close_mark();
  }
  return HAWK_OK;
}

// unphased pipeline downstream of the scan (resolve_kernels.cu): two passes over the hits, one
// host round trip (row total + error flag) between them
static int search_unphased(hawk_ctx* c, hawk_batch* b, const ScanConst& K, const BatchView& B, ScanOut& so, int32_t ref_h,
                           hawk_result* r) {
  cudaStream_t st = c->stream;
  Trace tr;
  const int64_t n_hits[2] = {so.n[0], so.n[1]};
  const uint64_t* recs[2] = {so.hits[0].as<uint64_t>(), so.hits[1].as<uint64_t>()};
  RefInfo ref{ref_h, 1, 0, 0};
  DevBuf d_refrange, d_refbm[2], start[2], stop[2], cnt[2], rpivot[2], blk_sum[2], blk_base[2], d_tot, d_kb;
  CK(d_refrange.alloc(c, 32, true));
  CK(launch_ref_range(st, recs[0], n_hits[0], recs[1], n_hits[1], ref_h, d_refrange.as<int64_t>()));
  size_t bm_words = 1;
  if (ref_h >= 0) {
    ref.g0 = b->first_gen[ref_h];
    ref.len = b->len[ref_h];
    bm_words = (size_t)(ref.len + 31) / 32 + 1;
  }
  for (int s = 0; s < 2; ++s) CK(d_refbm[s].alloc(c, bm_words * 4, true));
  if (ref_h >= 0)
    CK(launch_ref_bitmap(st, recs[0], recs[1], d_refrange.as<int64_t>(), d_refbm[0].as<uint32_t>(), d_refbm[1].as<uint32_t>()));
  int64_t n_blk[2];
  CK(d_tot.alloc(c, 32, true));  // [0..1] kept rows per strand, [2] error code
  for (int s = 0; s < 2; ++s) {
    n_blk[s] = resolve_blocks(n_hits[s]);
    CK(start[s].alloc(c, (size_t)n_hits[s] * 4));
    CK(stop[s].alloc(c, (size_t)n_hits[s] * 4));
    CK(cnt[s].alloc(c, (size_t)n_hits[s] * 4));
    CK(rpivot[s].alloc(c, (size_t)n_hits[s] * 8));  // two words per hit (resolve_kernels.cu: ResStrand)
    CK(blk_sum[s].alloc(c, (size_t)(n_blk[s] + 1) * 8));
    CK(blk_base[s].alloc(c, (size_t)(n_blk[s] + 1) * 8));
  }
  const uint32_t* bm[2] = {d_refbm[0].as<uint32_t>(), d_refbm[1].as<uint32_t>()};
  int32_t* st_[2] = {start[0].as<int32_t>(), start[1].as<int32_t>()};
  int32_t* sp_[2] = {stop[0].as<int32_t>(), stop[1].as<int32_t>()};
  uint32_t* cn_[2] = {cnt[0].as<uint32_t>(), cnt[1].as<uint32_t>()};
  int32_t* rp_[2] = {rpivot[0].as<int32_t>(), rpivot[1].as<int32_t>()};
  uint64_t* bs_[2] = {blk_sum[0].as<uint64_t>(), blk_sum[1].as<uint64_t>()};
  uint64_t* bb_[2] = {blk_base[0].as<uint64_t>(), blk_base[1].as<uint64_t>()};
  CK(launch_resolve_count(st, B, K, recs, n_hits, ref, bm, st_, sp_, cn_, rp_, bs_, (int*)(d_tot.as<uint64_t>() + 2)));
  {
    const uint64_t* in_[2] = {bs_[0], bs_[1]};
    CK(launch_blk_prefix64(st, in_, n_blk, bb_, d_tot.as<uint64_t>()));
  }
  uint64_t tot[4] = {0, 0, 0, 0};
  CK(c->small_d2h_sync(tot, d_tot.p, 32));
  tr.tick("resolve: count + sync");
  const int err = (int)(int32_t)(uint32_t)tot[2];
  if (err == HAWK_EALLELES)
    return hawk_fail(HAWK_EALLELES, "ambiguity code inside a guide window has no variant_alleles entry");
  if (err == HAWK_ECAPACITY)
    return hawk_fail(HAWK_ECAPACITY, "resolve_guide expansion of one hit exceeds %llu strings", (unsigned long long)HAWK_MAX_EXPANSION);
  const int64_t n = (int64_t)(tot[0] + tot[1]);
  if (n >= 0xFFFFFFFFll) return hawk_fail(HAWK_ECAPACITY, "unphased search: more than 2^32 guide rows");
  const int W = K.C + 2 * HAWK_GUIDESEQPAD;
  r->n_guides = n;
  r->text_stride = (W + 15) / 16 * 16;
  CK(r->hap.alloc(c, (size_t)n * 4));
  CK(r->strand.alloc(c, (size_t)n));
  CK(r->pos.alloc(c, (size_t)n * 4));
  CK(r->start.alloc(c, (size_t)n * 4));
  CK(r->stop.alloc(c, (size_t)n * 4));
  CK(r->bucket.alloc(c, (size_t)n * 4));
  CK(r->text.alloc(c, (size_t)n * r->text_stride));
  if (n > 0) {
    CK(d_kb.alloc(c, (size_t)(b->n_hap + 1) * 16));
    const uint32_t* cc_[2] = {cn_[0], cn_[1]};
    const uint64_t* cb_[2] = {bb_[0], bb_[1]};
    CK(launch_hap_offsets_cnt(st, recs, n_hits, cc_, cb_, d_tot.as<uint64_t>(), b->n_hap, d_kb.as<uint64_t>()));
    const int64_t key_span = b->gmax >= b->gmin ? ((int64_t)b->gmax - b->gmin + 1) * 2 : 0;
    const bool direct = key_span > 0 && key_span <= HAWK_DIRECT_KEY_SPAN;
    DevBuf key_table;
    if (direct) {
      CK(key_table.alloc(c, (size_t)key_span * 4));
      CKCUDA(cudaMemsetAsync(key_table.p, 0xFF, (size_t)key_span * 4, st));
    }
    const int32_t* cs_[2] = {st_[0], st_[1]};
    const int32_t* ce_[2] = {sp_[0], sp_[1]};
    const int32_t* cr_[2] = {rp_[0], rp_[1]};
    const uint64_t* kb_[2] = {d_kb.as<uint64_t>() + (size_t)(b->n_hap + 1), d_kb.as<uint64_t>()};
    CK(launch_resolve_write(st, B, K, recs, n_hits, cs_, ce_, cc_, cr_, cb_, kb_, ref_h, r->text_stride,
                            r->hap.as<int32_t>(), r->strand.as<uint8_t>(), r->pos.as<int32_t>(), r->start.as<int32_t>(),
                            r->stop.as<int32_t>(), r->text.as<uint8_t>(), direct ? key_table.as<uint32_t>() : nullptr, b->gmin));
    if (direct) {
      CK(launch_bucket_read(st, r->start.as<int32_t>(), r->strand.as<uint8_t>(), n, d_tot.as<uint64_t>(),
                            key_table.as<uint32_t>(), b->gmin, r->bucket.as<uint32_t>()));
    } else {
      uint64_t tsize = 1024;
      while (tsize < (uint64_t)n * 2) tsize <<= 1;
      DevBuf keys, vals;
      CK(keys.alloc(c, tsize * 8));
      CK(vals.alloc(c, tsize * 8));
      CKCUDA(cudaMemsetAsync(keys.p, 0xFF, tsize * 8, st));
      CKCUDA(cudaMemsetAsync(vals.p, 0xFF, tsize * 8, st));
      CK(launch_buckets(st, r->start.as<int32_t>(), r->strand.as<uint8_t>(), n, keys.as<unsigned long long>(),
                        vals.as<unsigned long long>(), tsize, r->bucket.as<uint32_t>()));
    }
    c->close_mark();
    CKCUDA(cudaStreamSynchronize(st));  // the temporaries above are released on return
  } else {
    c->close_mark();
  }
  tr.tick("resolve: write + sync");
  return HAWK_OK;
}

extern "C" int hawk_search(hawk_ctx* c, hawk_batch* b, const hawk_params* params,
                           const int32_t* scan_start, const int32_t* scan_stop,
                           const uint8_t* is_ref, hawk_result** out) {
  return hawk_search_impl(c, b, params, scan_start, scan_stop, is_ref, nullptr, out);
}

int hawk_search_impl(hawk_ctx* c, hawk_batch* b, const hawk_params* params, const int32_t* scan_start,
                     const int32_t* scan_stop, const uint8_t* is_ref, const StreamLink* link, hawk_result** out,
                     const uint8_t* fused_text, int64_t* bad_slot) {
  // NULL bounds: the ones attached to the batch (hawk_batch_set_scan), already on the device
  const bool cached = b && !scan_start && !scan_stop && !is_ref && b->has_scan;
  if (cached) {
    scan_start = b->h_scan_a.data();
    scan_stop = b->h_scan_b.data();
    is_ref = b->h_scan_ref.data();
  }
  CK(check_scan_args(c, b, params, scan_start, scan_stop, out));
  if (bad_slot) *bad_slot = -1;
  if (b->n_hap > 0 && !is_ref) return hawk_fail(HAWK_EINVAL, "hawk_search: is_ref missing");
  if (!b->has_posmap) return hawk_fail(HAWK_EINVAL, "hawk_search: call hawk_batch_set_posmap first");
  const bool unphased = (params->flags & HAWK_F_UNPHASED) != 0;
  if (unphased && link) return hawk_fail(HAWK_EINVAL, "hawk_search: the streamed search is phased / variant-free only");
  if (unphased && !b->has_alleles)
    return hawk_fail(HAWK_EINVAL, "hawk_search: unphased search needs hawk_batch_set_alleles");
  CKCUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  int32_t ref_h = -1, n_ref = 0;
  if (cached) {
    ref_h = b->scan_ref_h;
    n_ref = b->scan_n_ref;
  } else {
    for (int32_t h = 0; h < b->n_hap; ++h)
      if (is_ref[h]) {
        if (ref_h < 0) ref_h = h;
        ++n_ref;
      }
  }
  if (n_ref > 1)
    return hawk_fail(HAWK_EDUPREF, "hawk_search: %d haplotypes are labelled REF; the reference aborts on "
                     "the duplicate REF guides this produces (search_guides.py:328-334)", n_ref);

  hawk_result* r = new (std::nothrow) hawk_result();
  if (!r) return hawk_fail(HAWK_ENOMEM, "hawk_search: host allocation");
  r->ctx = c;
  const ScanConst K = make_scan_const(*params, 0);
  const int W = K.C + 2 * HAWK_GUIDESEQPAD;
  r->window = W;

  int rc = HAWK_OK;
  ScanInputs in;
  in.cached = cached;
  DevBuf d_refrange, d_err;
  DevBuf start[2], stop[2], keep[2], kept_excl[2], tile_sums, cnt[2], off[2], text_pre[2], row_hit[2];
  ScanOut so;
  do {
    Trace tr;
    bool staged = true;
    if (!fused_text && b->edits_lazy) {
      // an edit-list batch builds its planes now: windows around the edits as far as this guide /
      // PAM geometry reaches -- or all of them when a haplotype with edits is scanned like REF
      // (every chunk), the search is unphased, or the windows are too wide for the fast form
      int need = K.small && !unphased ? fused_reach(K) : HAWK_EDITS_DENSE;
      for (int32_t h = 0; h < b->n_hap; ++h)
        if (is_ref[h] && b->has_edits[h]) need = HAWK_EDITS_DENSE;
      if ((rc = hawk_edits_ensure(b, need))) break;
    }
    if (fused_text) {
      b->edits_lazy = false;  // re-encoded from the caller's texts
      // The fused kernel reads the texts once, keeps planes only where a later stage reads them and
      // matches in the same pass; it is flat over the slot space, so short haplotypes (an unphased
      // cohort's ~200-base indel windows: the staged K2 spends a thread block per haplotype) cost
      // the same per base as long ones. Since its K1 became pack_chunk_v3 with the texts arriving
      // through a cp.async ring it also wins on long haplotypes (config 2: 1.36 ms against
      // 1.27 + 0.09 + 0.24 ms staged), so mode 2 takes it whenever the geometry has the fast form;
      // mode 0 keeps the staged kernels (all planes stored) for callers who search the batch again
      // with another geometry.
      const bool use_fused = c->fused_mode != 0;
      if (K.small && use_fused && b->total_slots > 0) {
        bool fell_back = false;
        if ((rc = run_scan_fused(c, b, fused_text, params, scan_start, scan_stop, is_ref, in, so, bad_slot, &fell_back))) break;
        staged = fell_back;
      } else {
        if ((rc = hawk_batch_repack_dev(b, fused_text, bad_slot))) break;
      }
    } else if (b->sparse && (!K.small || fused_reach(K) > b->sparse_reach)) {
      rc = hawk_fail(HAWK_EINVAL, "hawk_search: this batch keeps planes only around variant bases "
                     "(hawk_encode_search_dev) and this guide / PAM geometry reaches further; re-encode it");
      break;
    }
    if (staged && (rc = run_scan(c, b, params, scan_start, scan_stop, is_ref, 0, in, so))) break;
    tr.tick("run_scan total");
    r->scanned_bp = so.scanned_bp;
    cudaEvent_t ev_post;
    c->mark(2, &ev_post);
    const BatchView B = batch_view(b, in.a, in.b, in.is_ref);
    const int64_t n_hits[2] = {so.n[0], so.n[1]};
    if (!unphased) {
      if ((rc = search_fast(c, b, K, B, so, ref_h, link, r))) break;
      r->params = *params;
      r->is_table = true;
      for (int s = 0; s < 2; ++s) {
        r->n_hits[s] = n_hits[s];
        r->hits[s].move_from(so.hits[s]);
      }
      break;
    }
    // windows of the scan's fast form (<= 80 characters) and a REF whose coordinates are linear
    // (always, for an unphased cohort) take the two-pass pipeline over the hit stream
    static const bool old_unphased = getenv("HAWK_OLD_UNPHASED") != nullptr;
    if (!old_unphased && K.small && W <= 80 && (ref_h < 0 || b->linear[ref_h])) {
      if ((rc = search_unphased(c, b, K, B, so, ref_h, r))) break;
      for (int s = 0; s < 2; ++s) {
        r->n_hits[s] = n_hits[s];
        r->hits[s].move_from(so.hits[s]);
      }
      break;
    }
    if ((rc = d_refrange.alloc(c, 32, true))) break;
    if ((rc = d_err.alloc(c, 4, true))) break;
    if ((rc = launch_ref_range(st, so.hits[0].as<uint64_t>(), n_hits[0], so.hits[1].as<uint64_t>(),
                               n_hits[1], ref_h, d_refrange.as<int64_t>())))
      break;
    int64_t n_rows[2] = {n_hits[0], n_hits[1]};
    for (int s = 0; s < 2 && rc == HAWK_OK; ++s) {
      if ((rc = start[s].alloc(c, (size_t)n_hits[s] * 4))) break;
      if ((rc = stop[s].alloc(c, (size_t)n_hits[s] * 4))) break;
      if ((rc = keep[s].alloc(c, (size_t)n_hits[s]))) break;
      rc = launch_rows(st, B, K, so.hits[s].as<uint64_t>(), n_hits[s], s, ref_h,
                       d_refrange.as<int64_t>(), unphased ? 0 : 1, start[s].as<int32_t>(),
                       stop[s].as<int32_t>(), keep[s].as<uint8_t>());
    }
    if (rc) break;
    tr.tick("post: rows launched");
    int64_t max_tiles = scan_tiles(n_hits[0] > n_hits[1] ? n_hits[0] : n_hits[1]);
    if (unphased) {
      // resolve_guide: count, scan, write
      uint64_t totals[2] = {0, 0};
      if ((rc = tile_sums.alloc(c, (size_t)(max_tiles + 1) * 8))) break;
      for (int s = 0; s < 2 && rc == HAWK_OK; ++s) {
        if (n_hits[s] == 0) continue;
        if ((rc = cnt[s].alloc(c, (size_t)n_hits[s] * 8))) break;
        if ((rc = off[s].alloc(c, (size_t)n_hits[s] * 8))) break;
        if ((rc = launch_expand_count(st, B, K, so.hits[s].as<uint64_t>(), n_hits[s], s,
                                      cnt[s].as<uint64_t>(), d_err.as<int>())))
          break;
        if ((rc = exclusive_scan_u64(st, cnt[s].as<uint64_t>(), n_hits[s], off[s].as<uint64_t>(),
                                     tile_sums.as<uint64_t>())))
          break;
        rc = c->small_d2h_sync(&totals[s], tile_sums.as<uint64_t>() + scan_tiles(n_hits[s]), 8);
      }
      if (rc) break;
      int err = 0;
      if ((rc = c->small_d2h_sync(&err, d_err.p, 4))) break;
      if (err == HAWK_EALLELES) {
        rc = hawk_fail(HAWK_EALLELES, "ambiguity code inside a guide window has no variant_alleles entry");
        break;
      }
      if (err == HAWK_ECAPACITY) {
        rc = hawk_fail(HAWK_ECAPACITY, "resolve_guide expansion of one hit exceeds %llu strings",
                       (unsigned long long)HAWK_MAX_EXPANSION);
        break;
      }
      for (int s = 0; s < 2 && rc == HAWK_OK; ++s) {
        n_rows[s] = (int64_t)totals[s];
        if (n_rows[s] == 0) continue;
        if ((rc = text_pre[s].alloc(c, (size_t)n_rows[s] * W))) break;
        if ((rc = row_hit[s].alloc(c, (size_t)n_rows[s] * 8))) break;
        if ((rc = keep[s].alloc(c, (size_t)n_rows[s]))) break;
        rc = launch_expand_write(st, B, K, so.hits[s].as<uint64_t>(), n_hits[s], s,
                                 off[s].as<uint64_t>(), n_rows[s], ref_h, d_refrange.as<int64_t>(),
                                 start[s].as<int32_t>(), text_pre[s].as<uint8_t>(),
                                 row_hit[s].as<int64_t>(), keep[s].as<uint8_t>());
      }
      if (rc) break;
      int64_t t2 = scan_tiles(n_rows[0] > n_rows[1] ? n_rows[0] : n_rows[1]);
      if (t2 > max_tiles) {
        max_tiles = t2;
        if ((rc = tile_sums.alloc(c, (size_t)(max_tiles + 1) * 8))) break;
      }
    } else {
      if ((rc = tile_sums.alloc(c, (size_t)(max_tiles + 1) * 8))) break;
    }
    // stable compaction offsets of the surviving rows
    uint64_t kept_total[2] = {0, 0};
    for (int s = 0; s < 2 && rc == HAWK_OK; ++s) {
      if (n_rows[s] == 0) continue;
      if ((rc = kept_excl[s].alloc(c, (size_t)n_rows[s] * 8))) break;
      if ((rc = exclusive_scan_u8(st, keep[s].as<uint8_t>(), n_rows[s], kept_excl[s].as<uint64_t>(),
                                  tile_sums.as<uint64_t>())))
        break;
      rc = c->small_d2h_sync(&kept_total[s], tile_sums.as<uint64_t>() + scan_tiles(n_rows[s]), 8);
    }
    if (rc) break;
    const int64_t n = (int64_t)(kept_total[0] + kept_total[1]);
    tr.tick("post: keep scans + sync");
    r->n_guides = n;
    r->text_stride = (W + 15) / 16 * 16;
    if ((rc = r->hap.alloc(c, (size_t)n * 4))) break;
    if ((rc = r->strand.alloc(c, (size_t)n))) break;
    if ((rc = r->pos.alloc(c, (size_t)n * 4))) break;
    if ((rc = r->start.alloc(c, (size_t)n * 4))) break;
    if ((rc = r->stop.alloc(c, (size_t)n * 4))) break;
    if (n >= 0xFFFFFFFFll) {
      rc = hawk_fail(HAWK_ECAPACITY, "search: more than 2^32 guide rows");
      break;
    }
    if ((rc = r->bucket.alloc(c, (size_t)n * 4))) break;
    if ((rc = r->text.alloc(c, (size_t)n * r->text_stride))) break;
    if (n > 0) {
      GatherLaunch g;
      g.B = B;
      g.K = K;
      for (int s = 0; s < 2; ++s) {
        g.recs[s] = so.hits[s].as<uint64_t>();
        g.row_hit[s] = unphased ? row_hit[s].as<int64_t>() : nullptr;
        g.keep[s] = keep[s].as<uint8_t>();
        g.kept_excl[s] = kept_excl[s].as<uint64_t>();
        g.text_pre[s] = unphased ? text_pre[s].as<uint8_t>() : nullptr;
        g.start[s] = start[s].as<int32_t>();
        g.stop[s] = stop[s].as<int32_t>();
        g.n_rows[s] = n_rows[s];
        g.kept_total[s] = kept_total[s];
      }
      g.o_hap = r->hap.as<int32_t>();
      g.o_strand = r->strand.as<uint8_t>();
      g.o_pos = r->pos.as<int32_t>();
      g.o_start = r->start.as<int32_t>();
      g.o_stop = r->stop.as<int32_t>();
      g.o_text = r->text.as<uint8_t>();
      g.text_stride = r->text_stride;
      if ((rc = launch_gather(st, g))) break;
      // first-seen bucket ids
      uint64_t tsize = 1024;
      while (tsize < (uint64_t)n * 2) tsize <<= 1;
      DevBuf keys, vals;
      if ((rc = keys.alloc(c, tsize * 8))) break;
      if ((rc = vals.alloc(c, tsize * 8))) break;
      if ((rc = hawk_check_cuda(cudaMemsetAsync(keys.p, 0xFF, tsize * 8, st), "bucket keys memset"))) break;
      if ((rc = hawk_check_cuda(cudaMemsetAsync(vals.p, 0xFF, tsize * 8, st), "bucket vals memset"))) break;
      if ((rc = launch_buckets(st, r->start.as<int32_t>(), r->strand.as<uint8_t>(), n,
                               keys.as<unsigned long long>(), vals.as<unsigned long long>(), tsize,
                               r->bucket.as<uint32_t>())))
        break;
    }
    c->close_mark();
    tr.tick("post: gather+buckets launched");
    if ((rc = hawk_check_cuda(cudaStreamSynchronize(st), "search sync"))) break;
    tr.tick("post: final sync");
    for (int s = 0; s < 2; ++s) {
      r->n_hits[s] = n_hits[s];
      r->hits[s].move_from(so.hits[s]);
    }
  } while (0);
  if (rc != HAWK_OK) {
    cudaStreamSynchronize(st);
    delete r;
    return rc;
  }
  *out = r;
  return HAWK_OK;
}

extern "C" int hawk_result_destroy(hawk_result* r) {
  if (!r) return HAWK_OK;
  cudaSetDevice(r->ctx->device);
  delete r;
  return HAWK_OK;
}

extern "C" int hawk_result_info(hawk_result* r, int64_t* n_guides, int64_t* n_hits, int32_t* window,
                                int32_t* text_stride, int64_t* scanned_bp) {
  if (!r) return hawk_fail(HAWK_EINVAL, "hawk_result_info: null result");
  if (n_guides) *n_guides = r->n_guides;
  if (n_hits) {
    n_hits[0] = r->n_hits[0];
    n_hits[1] = r->n_hits[1];
  }
  if (window) *window = r->window;
  if (text_stride) *text_stride = r->text_stride;
  if (scanned_bp) *scanned_bp = r->scanned_bp;
  return HAWK_OK;
}

extern "C" int hawk_result_fetch(hawk_result* r, int32_t* hap, uint8_t* strand, int32_t* pos,
                                 int32_t* start, int32_t* stop, uint32_t* bucket, uint8_t* text) {
  if (!r) return hawk_fail(HAWK_EINVAL, "hawk_result_fetch: null result");
  CKCUDA(cudaSetDevice(r->ctx->device));
  cudaStream_t st = r->ctx->stream;
  size_t n = (size_t)r->n_guides;
  if (n == 0) return HAWK_OK;
  r->ctx->d2h_bytes += (int64_t)(n * ((hap ? 4 : 0) + (strand ? 1 : 0) + (pos ? 4 : 0) + (start ? 4 : 0) + (stop ? 4 : 0) +
                                     (bucket ? 4 : 0) + (text ? (size_t)r->text_stride : 0)));
  if (hap) CKCUDA(cudaMemcpyAsync(hap, r->hap.p, n * 4, cudaMemcpyDeviceToHost, st));
  if (strand) CKCUDA(cudaMemcpyAsync(strand, r->strand.p, n, cudaMemcpyDeviceToHost, st));
  if (pos) CKCUDA(cudaMemcpyAsync(pos, r->pos.p, n * 4, cudaMemcpyDeviceToHost, st));
  if (start) CKCUDA(cudaMemcpyAsync(start, r->start.p, n * 4, cudaMemcpyDeviceToHost, st));
  if (stop) CKCUDA(cudaMemcpyAsync(stop, r->stop.p, n * 4, cudaMemcpyDeviceToHost, st));
  if (bucket) CKCUDA(cudaMemcpyAsync(bucket, r->bucket.p, n * 4, cudaMemcpyDeviceToHost, st));
  if (text) CKCUDA(cudaMemcpyAsync(text, r->text.p, n * (size_t)r->text_stride, cudaMemcpyDeviceToHost, st));
  CKCUDA(cudaStreamSynchronize(st));
  return HAWK_OK;
}

extern "C" int hawk_result_device_columns(hawk_result* r, void** cols) {
  if (!r || !cols) return hawk_fail(HAWK_EINVAL, "hawk_result_device_columns: bad arguments");
  cols[0] = r->hap.p;
  cols[1] = r->strand.p;
  cols[2] = r->pos.p;
  cols[3] = r->start.p;
  cols[4] = r->stop.p;
  cols[5] = r->bucket.p;
  cols[6] = r->text.p;
  return HAWK_OK;
}

extern "C" int hawk_result_fetch_hits(hawk_result* r, int32_t strand, uint64_t* hits) {
  if (!r || strand < 0 || strand > 1) return hawk_fail(HAWK_EINVAL, "hawk_result_fetch_hits: bad arguments");
  CKCUDA(cudaSetDevice(r->ctx->device));
  size_t n = (size_t)r->n_hits[strand];
  if (n == 0) return HAWK_OK;
  if (!hits) return hawk_fail(HAWK_EINVAL, "hawk_result_fetch_hits: null output");
  r->ctx->d2h_bytes += (int64_t)(n * 8);
  CKCUDA(cudaMemcpyAsync(hits, r->hits[strand].p, n * 8, cudaMemcpyDeviceToHost, r->ctx->stream));
  CKCUDA(cudaStreamSynchronize(r->ctx->stream));
  return HAWK_OK;
}
