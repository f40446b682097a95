/* Plain-C client of include/hawkscan.h: proves the boundary is a C ABI (no C++ or torch types)
 * and exercises the host-only entry points. With a GPU (argv[1] = "gpu") it also runs one
 * tiny search end to end: REF-only, NGG / 20 nt, the KAT1 region of SURVEY.md Appendix A. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hawkscan.h"

static const char *KAT_REF =
    "AGACTTTCAAAGATATGCTGGGTAGAGGTCGAGGTTATTATTTGTTACCAATTCTCATTGTGTTTCGGAA"
    "CTTGCGTTTTAGGTATGTCTTAGTGACTCTAAATACCAAGGCAGTCCTCGATCCGTTCCTAATAAGGAAT"
    "GGTGATTCCCTGTCATACCAATCTACCCCCTGTTATGCGCGTTTGTCGTTAGACCAATGTCAGCGCAGCG"
    "GCAGATCAAGCAGGAGGCGGAATGTAAACAGAAGGTATGCTTAGGTGGATAGGGAGTGAGCAACAAACGG";

#define CHECK(expr)                                                                   \
  do {                                                                                \
    int rc_ = (expr);                                                                 \
    if (rc_ != HAWK_OK) {                                                             \
      fprintf(stderr, "%s -> %d (%s): %s\n", #expr, rc_, hawk_strerror(rc_), hawk_last_error()); \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

int main(int argc, char **argv) {
  if (hawk_abi_version() != HAWK_ABI_VERSION) return 2;
  int32_t len[1] = {(int32_t)strlen(KAT_REF)};
  int64_t slot_off[2], total = 0;
  CHECK(hawk_layout(len, 1, slot_off, &total));
  if (slot_off[0] != HAWK_SLOT_GAP || total % HAWK_SLOT_ALIGN != 0 || total < slot_off[0] + len[0] + HAWK_SLOT_GAP) return 3;
  int32_t a[1] = {100}, b[1] = {177};
  int64_t sblock_off[2];
  if (hawk_scan_plan(a, b, 1, sblock_off) != 1 || sblock_off[1] != 1) return 4;
  printf("abi %d layout ok (total %lld slots)\n", hawk_abi_version(), (long long)total);
  if (argc < 2 || strcmp(argv[1], "gpu") != 0) return 0;

  hawk_ctx *ctx = NULL;
  CHECK(hawk_ctx_create(0, &ctx));
  uint8_t *ascii = (uint8_t *)calloc((size_t)total, 1);
  memcpy(ascii + slot_off[0], KAT_REF, (size_t)len[0]);
  hawk_batch *batch = NULL;
  int64_t bad = -1;
  CHECK(hawk_batch_create(ctx, ascii, slot_off, len, 1, &batch, &bad));
  int64_t seg_off[2] = {0, 1};
  int32_t seg_rel[1] = {0}, seg_gen[1] = {901};
  uint8_t seg_step[1] = {1};
  CHECK(hawk_batch_set_posmap(batch, seg_off, seg_rel, seg_gen, seg_step));
  hawk_params prm;
  memset(&prm, 0, sizeof prm);
  prm.pam_len = 3;
  prm.guide_len = 20;
  const uint8_t fwd[3] = {15, 4, 4}, rc[3] = {2, 2, 15}; /* NGG, CCN */
  memcpy(prm.pam_fwd, fwd, 3);
  memcpy(prm.pam_rc, rc, 3);
  uint8_t is_ref[1] = {1};
  hawk_result *res = NULL;
  CHECK(hawk_search(ctx, batch, &prm, a, b, is_ref, &res));
  int64_t n = 0, hits[2];
  int32_t window = 0, stride = 0;
  int64_t bp = 0;
  CHECK(hawk_result_info(res, &n, hits, &window, &stride, &bp));
  int32_t *start = (int32_t *)malloc((size_t)n * 4);
  uint8_t *strand = (uint8_t *)malloc((size_t)n);
  uint8_t *text = (uint8_t *)malloc((size_t)n * (size_t)stride);
  CHECK(hawk_result_fetch(res, NULL, strand, NULL, start, NULL, NULL, text));
  printf("guides %lld hits %lld/%lld window %d scanned %lld\n", (long long)n, (long long)hits[0], (long long)hits[1],
         window, (long long)bp);
  for (int64_t i = 0; i < n; ++i) printf("%d %c %.*s\n", start[i], strand[i] ? '-' : '+', window, text + i * stride);
  /* 14 guides, 3 on '+', first at 989 */
  int ok = n == 14 && hits[0] == 3 && hits[1] == 11 && start[0] == 989 && bp == 77 &&
           memcmp(text, "TTAGGTATGTCTTAGTGACTCTAAATACCAAGGCAGTCCTCGA", 43) == 0;
  /* N2 from plain C: reverse complements, GC counts; no variants on a REF-only batch */
  uint8_t *rc_text = (uint8_t *)malloc((size_t)n * (size_t)stride);
  int32_t *gc_num = (int32_t *)malloc((size_t)n * 4), *gc_den = (int32_t *)malloc((size_t)n * 4);
  int64_t *gv_off = (int64_t *)malloc((size_t)(n + 1) * 8), gv_total = -1;
  int64_t var_off[2] = {0, 0};
  CHECK(hawk_batch_set_variants(batch, var_off, NULL, NULL, NULL, NULL, NULL, 0));
  CHECK(hawk_result_annotate(res, batch, rc_text, gc_num, gc_den, gv_off, &gv_total));
  int ok2 = gv_total == 0 && gv_off[n] == 0 && memcmp(rc_text, text, 43) == 0 /* strand 0: unchanged */ &&
            gc_den[0] == 20 && gc_num[0] == 7 /* CTTAGTGACTCTAAATACCA */;
  /* the streamed call gives the same table in caller-owned host columns */
  hawk_table_out out;
  memset(&out, 0, sizeof out);
  out.start = (int32_t *)malloc((size_t)n * 4);
  out.strand = (uint8_t *)malloc((size_t)n);
  out.text = (uint8_t *)malloc((size_t)n * (size_t)stride);
  out.capacity = n;
  out.text_stride = hawk_table_text_stride(prm.pam_len, prm.guide_len);
  int64_t n2 = 0, hits2[2] = {0, 0}, bp2 = 0;
  CHECK(hawk_search_stream(ctx, ascii, slot_off, len, 1, seg_off, seg_rel, seg_gen, seg_step, &prm, a, b, is_ref, 0, &out,
                           &n2, hits2, &bp2, &bad));
  int ok3 = n2 == n && hits2[0] == hits[0] && hits2[1] == hits[1] && bp2 == bp && out.text_stride == stride &&
            memcmp(out.start, start, (size_t)n * 4) == 0 && memcmp(out.strand, strand, (size_t)n) == 0 &&
            memcmp(out.text, text, (size_t)n * (size_t)stride) == 0;
  out.capacity = 3; /* too small: the needed row count comes back with HAWK_ECAPACITY */
  int rc_small = hawk_search_stream(ctx, ascii, slot_off, len, 1, seg_off, seg_rel, seg_gen, seg_step, &prm, a, b, is_ref, 0,
                                    &out, &n2, hits2, &bp2, &bad);
  int ok4 = rc_small == HAWK_ECAPACITY && n2 == n;
  int64_t h2d = 0, d2h = 0;
  CHECK(hawk_ctx_traffic(ctx, &h2d, &d2h));
  printf("annotate %d stream %d capacity %d traffic %d\n", ok2, ok3, ok4, h2d > 0 && d2h > 0);
  hawk_result_destroy(res);
  hawk_batch_destroy(batch);
  hawk_ctx_destroy(ctx);
  return ok && ok2 && ok3 && ok4 ? 0 : 5;
}
