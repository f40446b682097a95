// hawk_post.h -- launch wrappers of post_kernels.cu (internal)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "hawk_core.h"

// resolve_guide products above this are refused (the reference would need that
// many Python strings per hit)
#define HAWK_MAX_EXPANSION (1ull << 24)

namespace hawk {

struct GatherLaunch {
  BatchView B;
  ScanConst K;
  const uint64_t* recs[2];
  const int64_t* row_hit[2];
  const uint8_t* keep[2];
  const uint64_t* kept_excl[2];
  const uint8_t* text_pre[2];
  const int32_t* start[2];
  const int32_t* stop[2];
  int64_t n_rows[2];
  uint64_t kept_total[2];
  int32_t* o_hap;
  uint8_t* o_strand;
  int32_t* o_pos;
  int32_t* o_start;
  int32_t* o_stop;
  uint8_t* o_text;
  int32_t text_stride;
};

// REF haplotype as seen by the redundancy filter
struct RefInfo {
  int32_t h;       // index, -1: none
  int32_t linear;  // its posmap is a single step-1 segment
  int32_t g0;      // posmap(0)
  int32_t len;
};

// row numbering of one group of a streamed search (all zero / -1 otherwise)
struct RowMap {
  int64_t row_base;    // added to the row index before it enters the shared first-seen table
  int32_t ref_local;   // haplotype index that maps to ref_global, -1: none
  int32_t ref_global;
  int32_t hap_add;     // other haplotypes: index + hap_add
};

// annot_kernels.cu (N2: post-search pure functions on the guide table)
int launch_annot_variants(cudaStream_t st, const BatchView& B, const ScanConst& K, const VariantView& V,
                          const int32_t* hap, const uint8_t* strand, const int32_t* pos, const int32_t* stop,
                          const uint8_t* text, int32_t text_stride, int64_t n, uint32_t* cnt, const uint64_t* off,
                          int32_t* idx, int32_t* flags, int pass);
int launch_annot_text(cudaStream_t st, const ScanConst& K, const uint8_t* strand, const uint8_t* text,
                      int32_t text_stride, int64_t n, uint8_t* rc_text, int32_t* gc_num, int32_t* gc_den);

// table_kernels.cu (phased / variant-free pipeline)
int64_t row_blocks(int64_t n);
int launch_ref_bitmap(cudaStream_t st, const uint64_t* r0, const uint64_t* r1, const int64_t* ref_range,
                      uint32_t* bm0, uint32_t* bm1);
int launch_rows_fast(cudaStream_t st, const BatchView& B, const ScanConst& K, const uint64_t* const recs[2],
                     const int64_t n[2], const RefInfo& ref, const uint32_t* const ref_bm[2], const int64_t* ref_range,
                     int32_t drop_ref, int32_t* const start[2], int32_t* const stop[2], uint8_t* const keep[2],
                     uint32_t* const blk_cnt[2]);
int launch_seg_index(cudaStream_t st, const int64_t* seg_off, const int32_t* seg_rel, int32_t n_hap, int32_t stride,
                     int32_t* idx);
int launch_blk_prefix(cudaStream_t st, const uint32_t* const cnt[2], const int64_t n_blk[2], uint64_t* const base[2],
                      uint64_t* totals);
int launch_hap_offsets(cudaStream_t st, const uint64_t* r0, const uint64_t* r1, int64_t n0, int64_t n1,
                       const uint8_t* k0, const uint8_t* k1, const uint64_t* b0, const uint64_t* b1,
                       const uint64_t* totals, int32_t n_hap, uint64_t* kb);
int launch_gather_fast(cudaStream_t st, const BatchView& B, const ScanConst& K, const uint64_t* const recs[2],
                       const uint8_t* const keep[2], const uint64_t* const blk_base[2], const int32_t* const start[2],
                       const int32_t* const stop[2], const uint64_t* const kb_other[2], const int64_t n[2],
                       int32_t text_stride, int32_t* o_hap, uint8_t* o_strand, int32_t* o_pos, int32_t* o_start,
                       int32_t* o_stop, uint8_t* o_text, uint32_t* key_table, int32_t key_min, const RowMap& rm);
int launch_bucket_read(cudaStream_t st, const int32_t* start, const uint8_t* strand, int64_t n_max,
                       const uint64_t* totals, const uint32_t* key_table, int32_t key_min, uint32_t* bucket);

// fused_kernels.cu: K1 + K2 in one pass over the texts (flat over the slot space)
constexpr int FUSED_RING = 2;          // slots (1 KB of text each) of a warp's cp.async ring, power of two
constexpr int FUSED_CTAS_PER_SM = 28;  // launch bound (ptxas settles on 63 registers; a bound of 32 CTAs ran 7 % slower, 24 the same)
constexpr int FUSED_WARPS = 1;    // warps per CTA (one: the warp index is then provably uniform -> uniform datapath)
constexpr int FUSED_SUB_MAX = 4096;  // chunks per warp sub-range (128 KB of text) for large slot spaces
// ... and fewer (a power of two >= 256) for small ones, so that there are a few waves of warps to
// run: 275 MB of text in 4,096-chunk sub-ranges would be 2,100 warps for 3,552 warp slots
int32_t fused_sub_size(int64_t n_chunks);
struct FusedLaunch {
  const uint8_t* ascii;
  int64_t n_chunks;
  void* q;
  uint32_t* v;
  uint32_t* nz;
  const int64_t* slot_off;
  int32_t n_hap;
  const HapScan* hs;
  ScanConst K;
  int32_t reach, store_all;
  int32_t sub;  // fused_sub_size(n_chunks)
  const uint64_t* seg_base;
  const uint32_t* seg_cap;
  void* entries;
  uint32_t *cnt_ent, *cnt_hit0, *cnt_hit1, *overflow;
  int64_t* bad;
};
int64_t fused_sub_ranges(int64_t n_chunks);
int launch_fused_caps(cudaStream_t st, const int64_t* slot_off, const uint8_t* is_ref, int32_t n_hap, int64_t n_chunks,
                      int32_t all_dense, uint32_t* cap);
int launch_fused_scan(cudaStream_t st, const FusedLaunch& L);
int launch_fused_expand(cudaStream_t st, const void* entries, const uint64_t* seg_base, const uint32_t* cnt_ent,
                        const uint64_t* base0, const uint64_t* base1, int64_t n_sub, uint64_t* hits0, uint64_t* hits1);
// scan2_kernels.cu: per-haplotype scan geometry
int launch_hapscan(cudaStream_t st, const BatchView& B, const ScanConst& K, HapScan* hs);

// resolve_kernels.cu: unphased guide-table pipeline over the hit stream
int64_t resolve_blocks(int64_t n);
int launch_resolve_count(cudaStream_t st, const BatchView& B, const ScanConst& K, const uint64_t* const recs[2],
                         const int64_t n[2], const RefInfo& ref, const uint32_t* const ref_bm[2], int32_t* const start[2],
                         int32_t* const stop[2], uint32_t* const cnt[2], int32_t* const rpivot[2],
                         uint64_t* const blk_sum[2], int* err);
int launch_blk_prefix64(cudaStream_t st, const uint64_t* const sum[2], const int64_t n_blk[2], uint64_t* const base[2],
                        uint64_t* totals);
int launch_hap_offsets_cnt(cudaStream_t st, const uint64_t* const recs[2], const int64_t n[2], const uint32_t* const cnt[2],
                           const uint64_t* const base[2], const uint64_t* totals, int32_t n_hap, uint64_t* kb);
int launch_resolve_write(cudaStream_t st, const BatchView& B, const ScanConst& K, const uint64_t* const recs[2],
                         const int64_t n[2], const int32_t* const start[2], const int32_t* const stop[2],
                         const uint32_t* const cnt[2], const int32_t* const rpivot[2], const uint64_t* const blk_base[2],
                         const uint64_t* const kb_other[2], int32_t ref_h, int32_t text_stride, int32_t* o_hap,
                         uint8_t* o_strand, int32_t* o_pos, int32_t* o_start, int32_t* o_stop, uint8_t* o_text,
                         uint32_t* key_table, int32_t key_min);

// synth_kernels.cu: edit lists -> lengths, output positions, posmap segments (two passes)
int launch_derive(cudaStream_t st, int32_t n_hap, const int64_t* edit_off, const int32_t* pos, const int32_t* reflen,
                  const int32_t* altlen, const int64_t* altoff, int64_t ref_len, int64_t alt_pool_len,
                  int32_t region_start, int32_t* outpos, int32_t* len, int32_t* seg_count, int32_t* bad,
                  const int64_t* seg_off, int32_t* seg_rel, int32_t* seg_gen, uint8_t* seg_step, int pass,
                  int32_t* edit_hap = nullptr);

// edits_kernels.cu: planes of edit-list haplotypes, only where a search reads them
int launch_pool_check(cudaStream_t st, const uint8_t* pool, int64_t n, unsigned long long* bad);
int launch_edits_plain(cudaStream_t st, const void* ref_q, const uint32_t* ref_v, int64_t ref_chunks, const int32_t* plain,
                       int32_t n_plain, const int64_t* slot_off, void* q, uint32_t* v);
int launch_edit_windows(cudaStream_t st, const void* ref_q, const int64_t* edit_off, const int32_t* pos, const int32_t* reflen,
                        const int32_t* altlen, const int64_t* altoff, const int32_t* outpos, const int32_t* edit_hap,
                        const uint8_t* pool, const int64_t* slot_off, const int32_t* len, int32_t n_hap, int64_t n_edits,
                        void* q, uint32_t* v, uint32_t* nz, int32_t reach);

int64_t scan_tiles(int64_t n);
int exclusive_scan_u8(cudaStream_t st, const uint8_t* in, int64_t n, uint64_t* out, uint64_t* tile_sums);
int exclusive_scan_u64(cudaStream_t st, const uint64_t* in, int64_t n, uint64_t* out, uint64_t* tile_sums);
// also stores the grand total at *total_out (device pointer, may be null)
int exclusive_scan_u32(cudaStream_t st, const uint32_t* in, int64_t n, uint64_t* out, uint64_t* tile_sums,
                       uint64_t* total_out);
int launch_ref_range(cudaStream_t st, const uint64_t* r0, int64_t n0, const uint64_t* r1, int64_t n1,
                     int32_t ref_h, int64_t* out);
int launch_rows(cudaStream_t st, const BatchView& B, const ScanConst& K, const uint64_t* recs,
                int64_t n, int s, int32_t ref_h, const int64_t* ref_range, int dedup,
                int32_t* start, int32_t* stop, uint8_t* keep);
int launch_expand_count(cudaStream_t st, const BatchView& B, const ScanConst& K,
                        const uint64_t* recs, int64_t n, int s, uint64_t* cnt, int* err);
int launch_expand_write(cudaStream_t st, const BatchView& B, const ScanConst& K,
                        const uint64_t* recs, int64_t n_hits, int s, const uint64_t* off,
                        int64_t n_rows, int32_t ref_h, const int64_t* ref_range,
                        const int32_t* start, uint8_t* text, int64_t* row_hit, uint8_t* keep);
int launch_gather(cudaStream_t st, const GatherLaunch& g);
int launch_buckets(cudaStream_t st, const int32_t* start, const uint8_t* strand, int64_t n,
                   unsigned long long* keys, unsigned long long* vals, uint64_t table_size,
                   uint32_t* bucket);
int launch_export_nibbles(cudaStream_t st, const void* q, const uint32_t* v, int64_t chunk0,
                          int32_t len, uint8_t* nib, uint8_t* lower);

}  // namespace hawk
