/*
 * hawkscan.h -- C-ABI of the B200-native CRISPR-HAWK guide-discovery scan.
 *
 * The reference (pinellolab/CRISPR-HAWK v0.2.2) has no FFI; its narrowest seam
 * for this path is three Python callables invoked from crisprhawk.py
 * (SURVEY.md 8b). Each entry point below cites the reference interface it
 * replaces (paths relative to /root/reference/src/crisprhawk). The Python
 * mirror in crispr_hawk_b200/ binds these with ctypes; INTEGRATION.md shows the
 * stub a reference maintainer would add.
 *
 * Conventions: plain pointers and sizes, no torch/C++ types; every function
 * returns 0 (HAWK_OK) or a negative HAWK_E* code; hawk_last_error() gives the
 * message of the calling thread's last failure. Two layers:
 *   - host layer  (hawk_batch_*, hawk_search*): caller-owned HOST buffers in and
 *     out, library-owned device buffers behind opaque handles, synchronous.
 *   - device layer (hawk_*_dev): caller-owned DEVICE buffers and stream (e.g.
 *     torch tensors), asynchronous, no allocation inside.
 *
 * Slot layout (both layers). Haplotype h owns base slots
 * [slot_off[h], slot_off[h] + len[h]) of one flat slot space; slot_off[h] is a
 * multiple of HAWK_SLOT_ALIGN and every haplotype is preceded and followed by at
 * least HAWK_SLOT_GAP unused slots (the halo the scan's tiles read); unused slots
 * hold 0. Device planes:
 *   q  : one uint4 {A,C,G,T} per 32-slot chunk -- the 4-bit IUPAC mask of
 *        encoder.py:18-34, bit-sliced: bit i of .x/.y/.z/.w is bit 0/1/2/3 of
 *        the nibble of slot 32*chunk + i (0.5 B per base).
 *   v  : one uint32 per chunk, bit i = base was lower-case (a variant base,
 *        haplotype.py:106-121) (0.125 B per base).
 *   nz : one bit per chunk (chunk c -> bit c & 31 of word c >> 5): the chunk's v word is
 *        non-zero. Lets the scan skip the case plane of variant-free stretches.
 */
#ifndef HAWKSCAN_H
#define HAWKSCAN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HAWK_ABI_VERSION 3
#define HAWK_SLOT_ALIGN 128 /* bases; haplotypes start on a 64-byte plane boundary */
#define HAWK_SLOT_GAP 128   /* unused (zero) slots before the first and after every haplotype */
#define HAWK_CHUNK 32       /* bases per chunk (one uint4 of planes, one uint32 of case bits) */
#define HAWK_SLACK_CHUNKS 8 /* readable zero chunks after the last slot (nz: 8 zero words) */
#define HAWK_MAX_PAM 16
#define HAWK_GUIDESEQPAD 10 /* guide.py:21 */
#define HAWK_MAX_WINDOW 148 /* G + P + 2*PAD upper bound supported on device (G + P <= 128) */

enum {
  HAWK_OK = 0,
  HAWK_EINVAL = -1,   /* bad argument */
  HAWK_ECUDA = -2,    /* CUDA runtime failure */
  HAWK_ENOMEM = -3,   /* allocation failure */
  HAWK_EIUPAC = -4,   /* non-IUPAC character (encoder.py:38-44) */
  HAWK_ECAPACITY = -5,/* caller-provided output capacity too small (dev layer) */
  HAWK_EALLELES = -6, /* ambiguity code without variant_alleles entry (KeyError, search_guides.py:207-213) */
  HAWK_EDUPREF = -7,  /* two REF guides at one (start, strand) (search_guides.py:328-334) */
  HAWK_EASSERT = -8,  /* input on which the reference itself fails an assert (annotation.py:191) */
  HAWK_ECFD = -9,     /* CFDon: a letter or key the score tables do not hold (KeyError, cfdscore.py:89-94) */
  HAWK_EFEATURE = -10 /* one-hot input: a letter other than A, C, G, T (KeyError, scores/deepCpf1/seqdeepcpf1.py:19, 91) */
};

/* mode flags of hawk_params.flags */
#define HAWK_F_UNPHASED 1u /* variants_present and not phased: is_pamhit_valid + resolve_guide (search_guides.py:473-479) */

typedef struct hawk_ctx hawk_ctx;       /* one device, one stream, reusable workspace */
typedef struct hawk_batch hawk_batch;   /* packed haplotypes of one region, device resident */
typedef struct hawk_result hawk_result; /* guide table of one search, device resident */

/* Replaces pam.py:127-142 (PAM.encode): caller passes the nibble lists
 * (`bits_list` of the forward PAM and of its reverse complement). */
typedef struct hawk_params {
  int32_t pam_len;                 /* P, 1..HAWK_MAX_PAM */
  int32_t guide_len;               /* G >= 1 */
  int32_t right;                   /* --right: guide downstream of the PAM on strand 0 */
  uint32_t flags;                  /* HAWK_F_* */
  uint8_t pam_fwd[HAWK_MAX_PAM];   /* IUPAC nibbles of the PAM, 5'->3' */
  uint8_t pam_rc[HAWK_MAX_PAM];    /* nibbles of its reverse complement (pam.py:63) */
} hawk_params;

int hawk_abi_version(void);
const char *hawk_last_error(void);
const char *hawk_strerror(int code);
/* CUDA kernels launched by this library in this process so far (bench.py's gpu_launches) */
int64_t hawk_launch_count(void);

/* ---- context ------------------------------------------------------------ */
int hawk_ctx_create(int device, hawk_ctx **ctx);
int hawk_ctx_destroy(hawk_ctx *ctx);
/* SM count etc. for callers sizing grids/benchmarks */
int hawk_ctx_info(hawk_ctx *ctx, int32_t *sm_count, int64_t *total_mem, int64_t *free_mem);
/* bytes the host layer has moved across PCIe through this context so far, per direction
 * (bench.py's e2e h2d_bytes_per_step / d2h_bytes_per_step are differences of these) */
int hawk_ctx_traffic(hawk_ctx *ctx, int64_t *h2d_bytes, int64_t *d2h_bytes);

/* ---- layout helper -------------------------------------------------------
 * slot_off[0..n_hap]: slot_off[0] = HAWK_SLOT_GAP, slot_off[h+1] = slot_off[h] +
 * len[h] rounded up to HAWK_SLOT_ALIGN + HAWK_SLOT_GAP; returns total slots
 * (= slot_off[n_hap], a multiple of HAWK_SLOT_ALIGN) in *total_slots. */
int hawk_layout(const int32_t *len, int32_t n_hap, int64_t *slot_off, int64_t *total_slots);

/* ---- host layer ----------------------------------------------------------
 * hawk_batch_create replaces crisprhawk.py:64-81 (encode_haplotypes) /
 * encoder.py:48-57 (encode) for a whole region at once: `ascii` is the slot
 * space filled with the haplotype texts (case preserved, unused slots 0),
 * total_slots bytes of HOST memory. On a non-IUPAC byte returns HAWK_EIUPAC
 * and stores its slot index in *bad_slot (the Python mirror turns it into the
 * reference's CrisprHawkIupacTableError). */
int hawk_batch_create(hawk_ctx *ctx, const uint8_t *ascii, const int64_t *slot_off,
                      const int32_t *len, int32_t n_hap, hawk_batch **batch, int64_t *bad_slot);
int hawk_batch_destroy(hawk_batch *batch);
/* encoder.encode's return value for one haplotype: one nibble (1..15) per base */
int hawk_batch_export_nibbles(hawk_batch *batch, int32_t hap, uint8_t *nibbles /* len[hap] */,
                              uint8_t *lower /* len[hap], may be NULL */);

/* Coordinate maps (haplotype.py:90-104,138-159 `posmap`), run-length encoded:
 * haplotype h owns segments [seg_off[h], seg_off[h+1]); inside segment k
 * posmap(i) = seg_gen[k] + seg_step[k] * (i - seg_rel[k]), seg_step in {0,1}. */
int hawk_batch_set_posmap(hawk_batch *batch, const int64_t *seg_off, const int32_t *seg_rel,
                          const int32_t *seg_gen, const uint8_t *seg_step);

/* Unphased only (haplotype.py:287-291 `variant_alleles`): haplotype h owns
 * sites [va_off[h], va_off[h+1]) sorted by va_idx (relative index); site j owns
 * entries [va_ent_off[j], va_ent_off[j+1]); va_ref[e] = nibble of the entry's
 * REF allele when it is a single base, else 0. */
int hawk_batch_set_alleles(hawk_batch *batch, const int64_t *va_off, const int32_t *va_idx,
                           const int64_t *va_ent_off, const uint8_t *va_ref);

/* Scan bounds as batch metadata (like the coordinate maps): scan_start / scan_stop are the
 * haplotype-relative bounds of compute_scan_start_stop (search_guides.py:49-84) for the PAM length
 * that will be searched, is_ref[h] = (samples == "REF"). Once attached, hawk_search /
 * hawk_encode_search_dev may be called with scan_start = scan_stop = is_ref = NULL and use these
 * device-resident copies: a cohort of 4.3e5 indel-window haplotypes otherwise spends a
 * millisecond per search re-staging and re-uploading three n_hap-sized arrays. */
int hawk_batch_set_scan(hawk_batch *batch, const int32_t *scan_start, const int32_t *scan_stop,
                        const uint8_t *is_ref);

/* Replaces search_guides.py:510-548 (search) up to, not including, the
 * construction of Python Guide objects. scan_start/scan_stop are the
 * haplotype-relative bounds of compute_scan_start_stop (:49-84); is_ref[h] =
 * (samples == "REF"). The result is the guide table in the reference's
 * emission order (haplotype, strand, position, expansion), already filtered by
 * is_pamhit_in_range (:395-420), the REF-core filter (:468-471),
 * is_pamhit_valid + resolve_guide (unphased, :216-257, :372-392) and
 * remove_redundant_guides (:340-369).
 * REF: at most one haplotype may be flagged is_ref (more: HAWK_EDUPREF up front -- the reference
 * aborts as soon as two REF guides share a (start, strand), which two REF haplotypes always
 * produce), and it should be haplotype 0, where every haplotype list the reference builds has
 * it (haplotypes.py reconstructs REF first). With REF elsewhere the rows are still exact, but
 * bucket ids are taken over the KEPT rows: a key whose first guide in emission order is a
 * redundant ALT guide that remove_redundant_guides drops sorts by its first kept guide instead
 * of that dropped one (group_guides_position :306-337 orders keys before filtering). */
int hawk_search(hawk_ctx *ctx, hawk_batch *batch, const hawk_params *params,
                const int32_t *scan_start, const int32_t *scan_stop, const uint8_t *is_ref,
                hawk_result **result);
int hawk_result_destroy(hawk_result *result);
/* Replaces search_guides.py:102-131 (pam_search): raw PAM occurrences inside
 * the scan bounds on both strands, no window / REF-core filter. The result
 * holds hit lists only (n_guides == 0). */
int hawk_pam_search(hawk_ctx *ctx, hawk_batch *batch, const hawk_params *params,
                    const int32_t *scan_start, const int32_t *scan_stop, hawk_result **result);
/* n_guides: rows of the table; n_hits[2]: records per strand in the result's
 * hit lists (hawk_pam_search: raw PAM hits; hawk_search: hits that survived
 * the in-range and REF-core filters, before resolution / redundancy removal);
 * window: characters per row (G + P + 20); text_stride: bytes per row of the text
 * column (window rounded up to 16, zero padded); scanned_bp: sum of scan_stop -
 * scan_start, the unit of the haplotype-bp/s metric. */
int hawk_result_info(hawk_result *result, int64_t *n_guides, int64_t *n_hits, int32_t *window,
                     int32_t *text_stride, int64_t *scanned_bp);
/* Copy the table to HOST arrays of n_guides rows: haplotype index, strand,
 * PAM position (haplotype-relative), genomic start/stop
 * (adjust_guide_position, :260-280), bucket = smallest row index among the
 * rows sharing the row's (start, strand) key, i.e. buckets ordered by bucket id
 * are in first-seen order (group_guides_position, :306-337), and the padded
 * window text (extract_guide_sequence / resolved string), `text_stride` bytes per
 * row of which the first `window` are the text. Any pointer may be NULL. */
int hawk_result_fetch(hawk_result *result, int32_t *hap, uint8_t *strand, int32_t *pos,
                      int32_t *start, int32_t *stop, uint32_t *bucket, uint8_t *text);
/* Borrowed DEVICE addresses of the table's columns (hap, strand, pos, start, stop, bucket,
 * text; n_guides rows each, text_stride bytes per text row), valid until the result is
 * destroyed and ordered on the context's stream: lets a multi-GPU caller gather the per-rank
 * tables over NCCL without a detour through host memory (crispr_hawk_b200/shard.py). */
int hawk_result_device_columns(hawk_result *result, void **cols /* [7] */);
/* hit list of one strand: packed (hap << 32 | pos), ascending */
int hawk_result_fetch_hits(hawk_result *result, int32_t strand, uint64_t *hits /* n_hits[strand] */);

/* Same as hawk_batch_create for haplotype texts that already live in DEVICE memory
 * (slot layout, total_slots bytes, 16-byte aligned): no host copy is made. */
int hawk_batch_create_dev(hawk_ctx *ctx, const uint8_t *d_ascii, const int64_t *slot_off,
                          const int32_t *len, int32_t n_hap, hawk_batch **batch, int64_t *bad_slot);

/* N1 (next row): haplotypes given as edit lists against the reference text of
 * the padded region instead of as texts -- what Haplotype._update_sequence / _update_posmap
 * (haplotype.py:106-159, 185-252) do per variant on the host. Haplotype h applies edits
 * [edit_off[h], edit_off[h+1]), sorted by edit_pos (0-based index into ref_ascii of the anchor
 * base) and non-overlapping; each is an SNV (reflen = altlen = 1), an anchored insertion
 * (reflen 1, ALT = anchor + inserted bases) or an anchored deletion (altlen 1, REF = anchor +
 * deleted bases); ALT text = alt_pool[edit_altoff .. + edit_altlen). The texts are
 * materialised on the device (ALT characters lower-case, as the reference writes them), packed
 * by K1, and the run-length coordinate maps are attached (posmap(0) = region_start), so the
 * batch is ready for hawk_search. Only the edits cross PCIe. hawk_batch_layout returns the
 * resulting lengths / slot offsets (n_hap + 1 and n_hap entries). */
int hawk_batch_create_from_edits(hawk_ctx *ctx, const uint8_t *ref_ascii, int64_t ref_len,
                                 int32_t region_start, int32_t n_hap, const int64_t *edit_off,
                                 const int32_t *edit_pos, const int32_t *edit_reflen,
                                 const int32_t *edit_altlen, const int64_t *edit_altoff,
                                 const uint8_t *alt_pool, int64_t alt_pool_len, hawk_batch **batch,
                                 int64_t *bad_slot);
int hawk_batch_layout(hawk_batch *batch, int64_t *slot_off, int32_t *len);

/* ---- streamed search: host buffers in, host table out, PCIe overlapped --------------------
 * One call for encode_haplotypes + search (crisprhawk.py:64-115) when the haplotype texts and
 * the guide table both live in HOST memory: the haplotypes other than REF are cut into groups;
 * while group g is packed and searched (as REF + the group's block, so that
 * remove_redundant_guides has its REF partners), the texts of group g + 1 are on their way to
 * the device and the guide rows of group g - 1 on their way back. The table is the same as
 * hawk_batch_create + hawk_batch_set_posmap + hawk_search + hawk_result_fetch give: rows in the
 * reference's emission order, REF rows once, bucket = smallest row index of the row's
 * (start, strand) key over the whole table. Phased / variant-free searches only
 * (HAWK_F_UNPHASED is refused); REF, if present, should be haplotype 0 (any other position is
 * served by a single group). `ascii`, and the columns of `out`, should be pinned host memory
 * (cudaHostAlloc / cudaHostRegister / torch pin_memory): pageable memory works but the copies
 * then do not overlap. n_groups = 0 lets the library choose (about 192 MB of text per group).
 * Rows beyond out->capacity are counted but not copied and the call returns HAWK_ECAPACITY with
 * the needed row count in *n_guides; out = NULL counts only. Columns of `out` may be NULL. */
typedef struct hawk_table_out {
  int32_t *hap;
  uint8_t *strand;
  int32_t *pos;
  int32_t *start;
  int32_t *stop;
  uint32_t *bucket;   /* row indices: a table holds fewer than 2^32 rows */
  uint8_t *text;       /* capacity * text_stride bytes */
  int64_t capacity;    /* rows every non-NULL column can hold */
  int32_t text_stride; /* must equal hawk_table_text_stride(pam_len, guide_len) */
} hawk_table_out;
/* bytes per row of the text column: G + P + 20 rounded up to 16 */
int32_t hawk_table_text_stride(int32_t pam_len, int32_t guide_len);
/* The haplotype groups hawk_search_stream would use (host only, no device needed): group g
 * covers haplotypes [lo[g], hi[g]); REF, when it is haplotype 0, is added to every group and is
 * not listed. Returns the number of groups (<= capacity, at most 256) or a negative HAWK_E*. */
int32_t hawk_stream_plan(const int64_t *slot_off, int32_t n_hap, const uint8_t *is_ref, int32_t n_groups,
                         int32_t *lo, int32_t *hi, int32_t capacity);
int hawk_search_stream(hawk_ctx *ctx, const uint8_t *ascii, const int64_t *slot_off, const int32_t *len,
                       int32_t n_hap, const int64_t *seg_off, const int32_t *seg_rel,
                       const int32_t *seg_gen, const uint8_t *seg_step, const hawk_params *params,
                       const int32_t *scan_start, const int32_t *scan_stop, const uint8_t *is_ref,
                       int32_t n_groups, const hawk_table_out *out, int64_t *n_guides,
                       int64_t *n_hits /* [2] */, int64_t *scanned_bp, int64_t *bad_slot);
/* The same from edit lists (see hawk_batch_create_from_edits): only the reference text and the
 * edits cross PCIe, group by group; the guide rows of one group leave while the next group's
 * texts are materialised, packed and searched. */
int hawk_search_stream_edits(hawk_ctx *ctx, const uint8_t *ref_ascii, int64_t ref_len,
                             int32_t region_start, int32_t n_hap, const int64_t *edit_off,
                             const int32_t *edit_pos, const int32_t *edit_reflen,
                             const int32_t *edit_altlen, const int64_t *edit_altoff,
                             const uint8_t *alt_pool, int64_t alt_pool_len, const hawk_params *params,
                             const int32_t *scan_start, const int32_t *scan_stop, const uint8_t *is_ref,
                             int32_t n_groups, const hawk_table_out *out, int64_t *n_guides,
                             int64_t *n_hits /* [2] */, int64_t *scanned_bp);

/* ---- N2 (next row): post-search pure functions on the guide table -----------
 * What annotation.annotate_guides does to every Guide right after search() (annotation.py:
 * 563-572), for the rows of a phased / variant-free hawk_search result, in row order:
 *   _annotate_variants / polish_guide_variants (:246-315): the variants of the row's haplotype
 *       that are visible in the guide, as a CSR list (gv_off, n_guides + 1 entries; the indices
 *       count inside the haplotype's own variant list and come back through
 *       hawk_result_fetch_variants, *gv_total of them). The host joins the sorted variant ids
 *       and looks the allele frequencies up (annotate_variants_afs, :334-365: string work).
 *   reverse_guides (:27-51): rc_text = the window text, reverse-complemented (IUPAC-aware,
 *       case kept, utils.py:46-79) for rows of strand 1; text_stride bytes per row.
 *   gc_content (:513-541): gc_num / gc_den = G+C+S and A+C+G+T+S+W counts of the guide without
 *       its PAM (Bio.SeqUtils.gc_fraction, ambiguous="remove"); gc = gc_num / gc_den.
 * The variant table: haplotype h carries variants [var_off[h], var_off[h+1]) sorted by
 * position, each in the reference's normalised form (variant.py:456-486, adjust_multiallelic):
 * var_pos = genomic coordinate, allele lengths, ALT text (upper-case, as in the variant id) =
 * alt_pool[var_altoff .. + var_altlen). Batches made by hawk_batch_create_from_edits keep their
 * edit lists as this table (anchored edits are normalised already). At most one variant per
 * position and haplotype is what a phased VCF yields; with more the reference's own answer
 * depends on Python's set order, so hawk_batch_set_variants refuses such a table (HAWK_EINVAL)
 * and the Python seam hands the list to the reference's own annotation functions.
 * Any output pointer may be NULL (gv_off = NULL skips the variant pass). Returns HAWK_EASSERT on
 * the input the reference's _find_insertion_stop asserts on. */
int hawk_batch_set_variants(hawk_batch *batch, const int64_t *var_off, const int32_t *var_pos,
                            const int32_t *var_reflen, const int32_t *var_altlen,
                            const int64_t *var_altoff, const uint8_t *alt_pool, int64_t alt_pool_len);
int hawk_result_annotate(hawk_result *result, hawk_batch *batch, uint8_t *rc_text, int32_t *gc_num,
                         int32_t *gc_den, int64_t *gv_off, int64_t *gv_total);
int hawk_result_fetch_variants(hawk_result *result, int32_t *gv_idx /* gv_total */);

/* N4 (next row): CFDon of every guide row against the REF guide of its (start, strand) key
 * (scoring.py:303-387 cfdon_score + group_guides_position, scores/crisprhawk_scores.py:65-87,
 * scores/cfdscore/cfdscore.py:53-95 compute_cfd) on the table of a phased / variant-free
 * hawk_search whose REF, if any, is haplotype 0. mm[(i * 4 + w) * 4 + g], i < 20, w / g in
 * A, C, G, T(U): the factor of the reference's key "r<w>:d<revcomp(g)>,<i + 1>"; pam2[a * 4 + b]:
 * the factor of the PAM's last two letters; NaN = key absent from the reference's dict. The
 * product runs in the reference's order in double precision (bit-identical floats). scores[row] =
 * NaN where the key has no REF guide. HAWK_ECFD (+ *bad_row) where the reference raises KeyError. */
int hawk_result_cfdon(hawk_result *result, const uint8_t *is_ref, int32_t n_hap, const double *mm /* 320 */,
                      const double *pam2 /* 16 */, double *scores /* n_guides */, int64_t *bad_row);

/* N4 (next row), second half: the learned scorers' inputs of every guide row, batched -- what the
 * reference builds guide by guide in Python before it calls Azimuth / RS3 / DeepCpf1 / CRISPRon /
 * sgDesigner (the models themselves stay its host code).
 *   kmers  (host, n_guides x L bytes, row-major, may be NULL): scoring.py:50-84
 *          _extract_guide_sequences (lead = 4) / _extract_guide_sequences_sgdesigner (lead = 0):
 *          sequence[(PAD - lead) : (-PAD + 3)].upper() of the text annotation.reverse_guides leaves
 *          (strand 1: IUPAC-aware reverse complement); L = G + P + lead + 3 (30 for SpCas9 NGG / 20,
 *          34 for Cpf1 TTTV / 23).
 *   onehot (n_guides x 4 x L float32, may be NULL; HOST memory, or DEVICE memory of the context's
 *          device if onehot_on_device): scores/deepCpf1/seqdeepcpf1.py:71-92 preprocess, channels
 *          A, C, G, T. A row with any other letter: HAWK_EFEATURE and *bad_row = the smallest such
 *          row (the reference's NTENCODING lookup raises KeyError); checked only when onehot is
 *          requested.
 * Rows in emission order (row i of hawk_result_fetch). Works on any result that holds the window
 * text column (phased, variant-free, unphased). */
int hawk_result_featurize(hawk_result *result, int32_t lead, uint8_t *kmers, float *onehot,
                          int32_t onehot_on_device, int64_t *bad_row);

/* N2, the row collapse of the report (reports._collapse_report_entries, reports.py:958-1008, for
 * the score-free column set): rows of a phased / variant-free hawk_search result that agree in
 * (start, stop, strand, origin = is_ref of the haplotype, guide + PAM text incl. case) form one
 * report row. perm[k] = table row at position k of the order (start, stop, group), ties in
 * emission order (what pandas' "first" aggregations see); head[k] = 1 where a group starts.
 * *collision = 1 if two different keys shared a 64-bit hash (never merged: heads are confirmed
 * on the bytes; the groups of such a key may then be split in two runs). The host orders the
 * groups like pandas sorts the groupby keys and joins the strings (crispr_hawk_b200/report_rows.py). */
int hawk_result_collapse(hawk_result *result, const uint8_t *is_ref, int32_t n_hap, uint32_t *perm /* n_guides */,
                         uint8_t *head /* n_guides */, int32_t *collision);

/* Re-run K1 into an existing batch from device-resident texts of the same layout (the
 * coordinate maps / allele tables attached to the batch are kept). */
int hawk_batch_repack_dev(hawk_batch *batch, const uint8_t *d_ascii, int64_t *bad_slot);

/* encode_haplotypes + search (crisprhawk.py:64-115) in ONE pass over device-resident texts: K1
 * and K2 fused. The batch is re-encoded from `d_ascii` (its own layout, 16-byte aligned, device
 * memory) and searched at once: the texts are read a single time, the PAM is matched while the
 * plane words are still in registers, and planes are stored only where the rest of the pipeline
 * reads them (around variant bases and for REF haplotypes). The result equals
 * hawk_batch_repack_dev + hawk_search. Afterwards the batch is *sparse*: it serves this result
 * (hawk_result_annotate) and further hawk_search calls whose guide + PAM are not longer; anything
 * else (hawk_pam_search, hawk_batch_export_nibbles, longer guides) needs hawk_batch_repack_dev
 * first and says so. Guides longer than 32 nt (or guide + PAM > 33) take K1 then the staged K2. */
int hawk_encode_search_dev(hawk_ctx *ctx, hawk_batch *batch, const uint8_t *d_ascii, const hawk_params *params,
                           const int32_t *scan_start, const int32_t *scan_stop, const uint8_t *is_ref,
                           hawk_result **result, int64_t *bad_slot);
/* Which kernels hawk_encode_search_dev runs: 0 = K1, then the staged K2 (the batch stays dense);
 * 1 = the fused kernel whenever the guide geometry allows; 2 (default) = the library's choice,
 * which is the same as 1 since the fused kernel became the faster one on every haplotype shape
 * (it used to be: fused for unphased cohorts and short haplotypes only). Results are identical. */
int hawk_ctx_set_fused(hawk_ctx *ctx, int32_t mode);
/* How hawk_batch_create_from_edits (and hawk_search_stream_edits) obtain the planes:
 * 1 (default) = only where a search reads them -- every chunk of a haplotype without edits, and
 * the chunks within a search's reach of an edit, cut from the reference's own planes at the
 * first hawk_search (O(edits) work; the haplotype texts are never written). Whole-haplotype
 * readers (hawk_pam_search, hawk_batch_export_nibbles, a search that marks a haplotype with
 * edits as REF, unphased or long-guide searches) make the batch build every plane first, on
 * their own. 0 = always materialise every text and run K1 (O(haplotype bases)). A reference with
 * lower-case or non-IUPAC characters takes the second way whatever the mode. Same results. */
int hawk_ctx_set_edit_planes(hawk_ctx *ctx, int32_t mode);

/* The context's cudaStream_t (so callers can time with events on the stream the kernels
 * run on) and optional per-kernel timing: with profiling on, every K1 / K2 launch of the
 * host layer is bracketed by CUDA events; hawk_ctx_profile returns and resets the sums.
 * ms[0] = K1 pack_kernel, ms[1] = K2 candidate kernels + prefix sums, ms[2] = guide-table
 * pipeline, ms[3] = K2 expand_kernel, ms[4] = K2 match_kernel; n[i] = bracketed regions. */
void *hawk_ctx_stream(hawk_ctx *ctx);
/* wait for everything queued on the context's stream (the asynchronous calls of this header) */
int hawk_ctx_sync(hawk_ctx *ctx);
int hawk_ctx_set_profiling(hawk_ctx *ctx, int32_t enabled);
int hawk_ctx_profile(hawk_ctx *ctx, double *ms /* [5] */, int64_t *n /* [5] */);

/* ---- device layer (asynchronous on `stream`, a cudaStream_t) -------------- */
/* K1: ASCII slot space -> planes. d_bad: one int64, must hold INT64_MAX on
 * entry; receives the smallest slot index with a non-IUPAC byte. */
int hawk_pack_dev(void *stream, const uint8_t *d_ascii, int64_t total_slots, void *d_q,
                  uint32_t *d_v, uint32_t *d_nz /* total_slots / 1024 rounded up, words */,
                  int64_t *d_bad);

/* K2, the PAM scan (search_guides.py:32-131 + the filters of :395-420, :468-471), as three
 * stream-ordered stages; the caller synchronises after stages 1 and 2 to read the totals
 * that size the next stage. Haplotypes are cut into slices of 32 chunks (1,024 base slots);
 * hawk_scan_plan (host) fills sblock_off (n_hap + 1): first thread block (256 slices) of
 * every haplotype, and returns the number of blocks. All d_* are device pointers;
 * per-haplotype arrays have n_hap entries; d_sblock_off is the uploaded plan.
 *   stage 1  hawk_scan_count_dev: candidate chunks (non-REF haplotypes: chunks with a variant
 *            base in reach of a guide core, found in the nz summary plane; REF haplotypes
 *            and raw_hits = 1 (pam_search semantics): every chunk). totals[0] = candidates.
 *   stage 2  hawk_scan_match_dev: PAM match on both strands + filters for the candidates ->
 *            hit masks (d_masks, n_cand x 8 bytes, in candidate = (haplotype, chunk) order).
 *            totals[1..2] = hits per strand, totals[3..4] = raw PAM hits (raw_hits = 1).
 *   stage 3  hawk_scan_expand_dev: (hap << 32 | pos) records, ascending, totals[1] / totals[2]
 *            of them in d_hits_fwd / d_hits_rev.
 * hawk_scan_totals(workspace) is the device address of uint64 totals[8]. The two workspaces
 * need no initialisation and must stay untouched between the stages. */
int64_t hawk_scan_plan(const int32_t *scan_start, const int32_t *scan_stop, int32_t n_hap,
                       int64_t *sblock_off);
size_t hawk_scan_workspace_bytes(int32_t n_hap, int64_t n_sblocks);
size_t hawk_scan_match_workspace_bytes(int64_t n_sblocks);
const uint64_t *hawk_scan_totals(const void *d_workspace);
int hawk_scan_count_dev(void *stream, const void *d_q, const uint32_t *d_v, const uint32_t *d_nz,
                        const int64_t *d_slot_off, const int32_t *d_len, const int32_t *d_scan_start,
                        const int32_t *d_scan_stop, const uint8_t *d_is_ref, const int64_t *d_sblock_off,
                        int32_t n_hap, int64_t n_sblocks, const hawk_params *params, int32_t raw_hits,
                        void *d_workspace);
int hawk_scan_match_dev(void *stream, const void *d_q, const uint32_t *d_v, const uint32_t *d_nz,
                        const int64_t *d_slot_off, const int32_t *d_len, const int32_t *d_scan_start,
                        const int32_t *d_scan_stop, const uint8_t *d_is_ref, int32_t n_hap,
                        int64_t n_sblocks, const hawk_params *params, int32_t raw_hits, int64_t n_cand,
                        uint64_t *d_masks, void *d_workspace, void *d_match_workspace);
int hawk_scan_expand_dev(void *stream, int32_t n_hap, int64_t n_sblocks, int64_t n_cand,
                         const uint64_t *d_masks, void *d_workspace, void *d_match_workspace,
                         uint64_t *d_hits_fwd, uint64_t *d_hits_rev);

/* First-seen bucket ids of a guide table in device memory (group_guides_position,
 * search_guides.py:306-337): d_bucket[i] = smallest row index sharing row i's (start, strand)
 * key. The final step of the multi-GPU merge, after the per-rank tables were concatenated in
 * rank order. Keys are direct addresses: key_span = largest start - key_min + 1;
 * d_key_table holds 2 * key_span + 1 uint32 (any content); n < 2^32. Synchronises the stream
 * before returning; a start outside [key_min, key_min + key_span) returns HAWK_EINVAL (the
 * bucket column is then undefined) instead of touching memory outside the table. */
int hawk_first_seen_dev(void *stream, const int32_t *d_start, const uint8_t *d_strand, int64_t n,
                        int32_t key_min, int64_t key_span, uint32_t *d_key_table, uint32_t *d_bucket);

/* ---- multi-GPU final merge: one-sided push over NVLink ---------------------------------------
 * One process per GPU, rank blocks = contiguous haplotype ranges of one region, every rank with
 * REF as its local haplotype 0 (remove_redundant_guides needs it). The merged table lives in ONE
 * buffer owned by the gathering rank: hawk_merge_layout gives the byte offset of every column
 * (hap i32, strand u8, pos i32, start i32, stop i32, bucket u32, text u8 x text_stride; each
 * column padded to 256 bytes; text is optional) and the size to allocate; hawk_peer_alloc
 * allocates it and exports a 64-byte CUDA IPC handle, which the caller ships to the other ranks
 * by any means (bench.py: one torch.distributed broadcast); they map it with hawk_peer_open and
 * write their rows into their slice with hawk_result_push -- rows [row_lo, n_guides) of a
 * result (row_lo = number of leading REF rows, which the gathering rank owns) to row offset
 * `at`, haplotype indices shifted by hap_add on the way (index hap_keep excepted) -- all peers
 * at the same time, stores going over NVLink straight into the owner's HBM. After a barrier the
 * owner runs hawk_first_seen_dev on its columns. hawk_result_push is asynchronous on the
 * context's stream; hawk_peer_close synchronises it first. */
int hawk_merge_layout(int64_t total_rows, int32_t text_stride, int32_t with_text, int64_t *off /* [7] */,
                      int64_t *bytes);
int hawk_peer_alloc(hawk_ctx *ctx, int64_t bytes, void **d_ptr, uint8_t *handle /* [64] */);
int hawk_peer_free(hawk_ctx *ctx, void *d_ptr);
int hawk_peer_open(hawk_ctx *ctx, const uint8_t *handle /* [64] */, void **d_ptr);
int hawk_peer_close(hawk_ctx *ctx, void *d_ptr);
int hawk_result_push(hawk_result *result, int64_t row_lo, int32_t hap_add, int32_t hap_keep, void *d_buffer,
                     int64_t total_rows, int64_t at, int32_t with_text, int64_t *pushed_bytes);

/* N1 (next row): materialise haplotype texts on the device from the reference text and
 * per-haplotype sorted, non-overlapping edit lists (haplotype.py:106-121,185-252
 * conventions: ALT allele characters lower-case; SNV 1 base, insertion anchor + inserted
 * bases, deletion the 1-base anchor). edit_outpos[e] = haplotype index where edit e's ALT
 * text starts. Output: the ASCII slot space consumed by hawk_pack_dev. d_ref must be readable
 * for 8 bytes past ref_len (block copies read whole aligned words). */
int hawk_materialize_dev(void *stream, const uint8_t *d_ref, int64_t ref_len,
                         const int64_t *d_edit_off, const int32_t *d_edit_pos,
                         const int32_t *d_edit_reflen, const int32_t *d_edit_altlen,
                         const int64_t *d_edit_altoff, const int32_t *d_edit_outpos,
                         const uint8_t *d_alt_pool, const int64_t *d_slot_off,
                         const int32_t *d_len, int32_t n_hap, int64_t total_slots,
                         int64_t n_edits, int32_t max_len /* longest haplotype */,
                         uint8_t *d_ascii_out);

#ifdef __cplusplus
}
#endif
#endif /* HAWKSCAN_H */
