// hawk_host.h -- objects of the C-ABI host layer (contexts, device buffers, batches, results),
// shared by api.cu and stream_api.cu (internal)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <vector>

#include "hawk_core.h"
#include "hawk_kernels.h"

#define CK(expr)                                   \
  do {                                             \
    int _rc = (expr);                              \
    if (_rc != HAWK_OK) return _rc;                \
  } while (0)
#define CKCUDA(expr) CK(hawk_check_cuda((expr), #expr))


// HAWK_TRACE=1: host-side wall-clock trace of the search pipeline on stderr (debugging aid)
struct Trace {
  bool on;
  std::chrono::steady_clock::time_point t0;
  Trace() : on(getenv("HAWK_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
  void tick(const char* what) {
    if (!on) return;
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[hawk] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

// ------------------------------------------------------------------ objects
struct hawk_ctx {
  int device;
  cudaStream_t stream;
  int sm_count;
  // Device-memory cache: every buffer of the library lives on this one stream, so a freed
  // block can be handed to the next request without synchronising (stream order protects
  // it). Avoids the per-search cost of the driver allocator for multi-GB temporaries.
  struct Block { void* p; size_t bytes; };
  std::vector<Block> free_blocks;
  void* take(size_t n, size_t* got) {
    size_t best = (size_t)-1;
    for (size_t i = 0; i < free_blocks.size(); ++i)
      if (free_blocks[i].bytes >= n && free_blocks[i].bytes <= 2 * n + (4u << 20) &&
          (best == (size_t)-1 || free_blocks[i].bytes < free_blocks[best].bytes))
        best = i;
    if (best != (size_t)-1) {
      Block b = free_blocks[best];
      free_blocks.erase(free_blocks.begin() + best);
      *got = b.bytes;
      return b.p;
    }
    size_t want = n < (1u << 20) ? ((n + 511) & ~(size_t)511) : ((n + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1));
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {  // give the cached blocks back to the driver and try once more
      cudaGetLastError();
      cudaStreamSynchronize(stream);
      trim();
      e = cudaMalloc(&p, want);
      if (e != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
      }
    }
    *got = want;
    return p;
  }
  void give(void* p, size_t bytes) { free_blocks.push_back(Block{p, bytes}); }
  // pinned host staging for the small per-search uploads (one H2D copy instead of six)
  void* pinned = nullptr;
  size_t pinned_bytes = 0;
  void* pinned_get(size_t n) {
    if (n > pinned_bytes) {
      if (pinned) cudaFreeHost(pinned);
      pinned = nullptr;
      pinned_bytes = 0;
      const size_t want = (n * 2 + 4095) & ~(size_t)4095;
      if (cudaHostAlloc(&pinned, want, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        pinned = nullptr;
        return nullptr;
      }
      pinned_bytes = want;
    }
    return pinned;
  }
  void trim() {
    for (auto& b : free_blocks) cudaFree(b.p);
    free_blocks.clear();
  }
  // Small transfers (metadata up, totals down) are moved by the SMs through pinned, mapped
  // host memory instead of the copy engines: a copy-engine transfer queues behind every bulk
  // copy of the same direction, whatever its stream, which would stall the streamed search
  // (hawk_search_stream) behind its own PCIe traffic. Implemented in api.cu.
  uint8_t* arena = nullptr;  // pinned bump arena; reset whenever the stream has been synchronised
  size_t arena_bytes = 0, arena_used = 0;
  int small_h2d(void* dst_dev, const void* src_host, size_t n);        // asynchronous on `stream`
  int small_d2h_sync(void* dst_host, const void* src_dev, size_t n);   // returns after a stream sync
  // bulk-copy streams of the streamed search (created on first use)
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  bool bulk_h2d = false;  // a streamed search owns the H2D copy engine: uploads go through the SMs
  // bytes the host layer moved across PCIe since the context was created (hawk_ctx_traffic)
  int64_t h2d_bytes = 0, d2h_bytes = 0;
  // hawk_encode_search_dev: 0 = K1 then the staged K2, 1 = the fused kernel whenever the guide
  // geometry allows, 2 = choose by haplotype shape (hawk_ctx_set_fused)
  int fused_mode = 2;
  // hawk_batch_create_from_edits: 1 = planes only around the edits, built at the first search
  // (edits_kernels.cu); 0 = materialise every text and run K1 (hawk_ctx_set_edit_planes)
  int edit_planes = 1;
  // optional per-kernel timing (hawk_ctx_set_profiling)
  bool profiling = false;
  struct Span { cudaEvent_t a, b; int kind; };
  std::vector<Span> spans;
  void mark(int kind, cudaEvent_t* a) {
    if (!profiling) return;
    Span s; s.kind = kind;
    cudaEventCreate(&s.a); cudaEventCreate(&s.b);
    cudaEventRecord(s.a, stream);
    spans.push_back(s);
    if (a) *a = s.a;
  }
  void close_mark() {
    if (!profiling || spans.empty()) return;
    cudaEventRecord(spans.back().b, stream);
  }
};

// device buffer owned through the context's block cache
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  hawk_ctx* ctx = nullptr;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  int alloc(hawk_ctx* c, size_t n, bool zero = false) {
    release();
    ctx = c;
    p = c->take(n ? n : 16, &bytes);
    if (!p) {
      bytes = 0;
      return hawk_fail(HAWK_ENOMEM, "out of device memory (%zu bytes requested)", n);
    }
    if (zero) return hawk_check_cuda(cudaMemsetAsync(p, 0, n ? n : 16, c->stream), "cudaMemsetAsync");
    return HAWK_OK;
  }
  void release() {
    if (p) ctx->give(p, bytes);
    p = nullptr;
    bytes = 0;
  }
  void move_from(DevBuf& o) {
    release();
    p = o.p;
    bytes = o.bytes;
    ctx = o.ctx;
    o.p = nullptr;
    o.bytes = 0;
  }
  template <class T>
  T* as() const { return (T*)p; }
};

static inline int upload(hawk_ctx* c, DevBuf& b, const void* src, size_t bytes) {
  CK(b.alloc(c, bytes));
  if (bytes) CK(c->small_h2d(b.p, src, bytes));
  return HAWK_OK;
}

struct hawk_batch {
  hawk_ctx* ctx;
  int32_t n_hap;
  int64_t total_slots;
  std::vector<int64_t> slot_off;
  std::vector<int32_t> len;
  DevBuf q, v, nz, d_slot_off, d_len;
  DevBuf seg_off, seg_rel, seg_gen, seg_step;
  DevBuf seg_idx;              // coarse index of the segments (table_kernels.cu), built at the first phased search
  int32_t seg_idx_stride = 0;  // 0: not built; -1: not worth building for this batch
  DevBuf va_off, va_idx, va_ent_off, va_ref;
  bool has_posmap = false, has_alleles = false;
  // planes are kept only around variant bases and for REF haplotypes (hawk_encode_search_dev):
  // valid for searches whose windows reach at most `sparse_reach` chunks from a variant chunk
  bool sparse = false;
  int32_t sparse_reach = 0;
  // N2: the haplotypes' variant tables (hawk_batch_set_variants, or kept from the edit lists)
  DevBuf var_off, var_pos, var_rl, var_al, var_ao, var_pool;
  int32_t var_pos_base = 0;
  bool has_variants = false;
  // N1: a batch created from edit lists keeps the reference's planes and the edits, and builds
  // its own planes on demand (hawk_edits_ensure): windows of `edits_reach` chunks around the
  // edits for a search, or all of them (materialise + K1) for whoever reads whole haplotypes
  bool edits_lazy = false;
  int32_t edits_reach = 0;
  DevBuf ref_text, ref_q, ref_v, edit_outpos, edit_hap, d_plain;
  int64_t ref_len = 0, ref_chunks = 0, n_edits = 0;
  int32_t n_plain = 0;
  std::vector<uint8_t> has_edits;  // per haplotype
  std::vector<int64_t> h_edit_off; // n_hap + 1
  // scan bounds attached to the batch (hawk_batch_set_scan): searches that pass NULL bounds use
  // these device-resident copies instead of uploading n_hap-sized arrays per call
  bool has_scan = false;
  std::vector<int32_t> h_scan_a, h_scan_b;
  std::vector<uint8_t> h_scan_ref;
  DevBuf d_scan;  // [sblock_off (n_hap + 1) x 8 | scan_start | scan_stop | is_ref]
  int64_t scan_bp = 0, scan_sblocks = 0, scan_dense_subs = 0;
  int32_t scan_ref_h = -1, scan_n_ref = 0;
  // host-side facts about the coordinate maps (hawk_batch_set_posmap)
  std::vector<int64_t> h_seg_off;
  std::vector<int32_t> first_gen;   // posmap(0) per haplotype
  std::vector<uint8_t> linear;      // one step-1 segment
  int32_t gmin = 0, gmax = -1;      // genomic coordinate range over all haplotypes
};

struct hawk_result {
  hawk_ctx* ctx;
  int64_t n_guides = 0;
  int64_t n_hits[2] = {0, 0};
  int32_t window = 0, text_stride = 0;
  int64_t scanned_bp = 0;
  int64_t ref_hits[2] = {0, 0};  // REF records per strand (filled for the groups of a streamed search)
  DevBuf hits[2];
  DevBuf hap, strand, pos, start, stop, bucket, text;
  hawk_params params;     // of the search that produced the table
  bool is_table = false;  // a phased / variant-free hawk_search result (hawk_result_annotate)
  DevBuf gv_idx;          // N2: variant indices of the last hawk_result_annotate
  int64_t gv_total = 0;
};


// direct-address first-seen table: at most this many (start, strand) keys
#define HAWK_DIRECT_KEY_SPAN (1ll << 28)

// One group of a streamed search (hawk_search_stream): REF + a block of the other haplotypes
// searched as a batch of its own, its rows appended to a table shared by all groups.
struct StreamLink {
  int64_t row_base;     // global index of the group's first row
  int32_t drop_ref;     // REF rows were emitted by an earlier group: only keep them as partners
  int32_t ref_local;    // REF haplotype's index inside the group's batch, -1: none
  int32_t ref_global;   // ... and in the caller's batch
  int32_t hap_add;      // local index + hap_add = caller's index (non-REF haplotypes)
  uint32_t* key_table;  // shared direct-address (start, strand) -> first row table
  int32_t key_min;
};

// planes of an edit-list batch: `need` = chunks a search reaches from a variant chunk, or
// HAWK_EDITS_DENSE for every chunk (no-op for other batches and when already satisfied)
#define HAWK_EDITS_DENSE (1 << 20)
int hawk_edits_ensure(hawk_batch* b, int need);
int hawk_materialize_range(cudaStream_t stream, const uint8_t* d_ref, int64_t ref_len, const int64_t* d_edit_off,
                           const int32_t* d_edit_pos, const int32_t* d_edit_reflen, const int32_t* d_edit_altlen,
                           const int64_t* d_edit_altoff, const int32_t* d_edit_outpos, const uint8_t* d_alt_pool,
                           const int64_t* d_slot_off, const int32_t* d_len, int32_t n_hap, int64_t total_slots,
                           int64_t e_lo, int64_t e_hi, int32_t max_len, uint8_t* d_ascii_out);
int batch_create_impl(hawk_ctx* c, const uint8_t* ascii, bool ascii_on_device, const int64_t* slot_off,
                      const int32_t* len, int32_t n_hap, hawk_batch** out, int64_t* bad_slot, bool defer_planes = false);
// `fused_text`: device-resident texts of the batch's layout; the batch is re-encoded from them
// inside the search (fused K1 + K2 when the guide geometry allows, else K1 then the staged K2)
int hawk_search_impl(hawk_ctx* c, hawk_batch* b, const hawk_params* params, const int32_t* scan_start,
                     const int32_t* scan_stop, const uint8_t* is_ref, const StreamLink* link, hawk_result** out,
                     const uint8_t* fused_text = nullptr, int64_t* bad_slot = nullptr);
