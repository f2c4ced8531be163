/*
 * oracle.h — CPU restatement of Microcket's post-alignment hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or run it, and only as the checker.
 * The product (microcket_b200/) never links or calls this code and has no
 * CPU fallback.
 *
 * Parity status
 *   sam2pairs, krmdup : PINNED — checked against the reference's own sources
 *                       compiled into oracle/_ref/ (see Makefile) on the
 *                       SURVEY Appendix-B vectors and on seeded synthetic
 *                       inputs (tests/test_oracle_vs_ref.py,
 *                       tests/golden/*).
 *   coordinate dedup, binning : PARITY UNPINNED — the reference has no
 *                       implementation (binning lives in the absent
 *                       juicer_tools.jar, microcket:525-529; krmdup never
 *                       sees coordinates).  Restated from the .pairs column
 *                       semantics (anno/4DN.DCIC.header:2) and the in-house
 *                       bin = int(pos/res) convention
 *                       (util/analyze.EBV/calc.loop2EBV.pl:28).
 */
#ifndef MICROCKET_ORACLE_H
#define MICROCKET_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- sam2pairs (src/sam2pairs/{sam2pairs.cpp,pairutil.h,flash2pairs.h,unc2pairs.h}) ---- */
typedef struct {
    uint32_t lowMap, manyHits, unpaired, selfCircle, trans, cis10K, cis1K, cis0; /* sam2pairs.cpp:211-218 order */
    uint64_t selfCircle_true;   /* all self-circles, not only thread 0's share */
    uint64_t groups;            /* read groups processed (last group excluded, pairutil.h:176) */
    uint64_t cigar_errors;      /* groups dropped because cigar2segment would have returned false (UB in the reference) */
} orc_s2p_stats;

/* mode: 0 = flash, 1 = unc.  Pair lines come out in input (group) order.
 * *pairs_out / *sam_out are malloc'ed; caller frees with orc_free. */
int orc_sam2pairs(const char *sam, size_t n, int mode, float min_mapped_ratio, int min_mapq,
                  int emu_threads, int write_sam,
                  char **pairs_out, size_t *pairs_len, char **sam_out, size_t *sam_len,
                  orc_s2p_stats *st);

/* One CIGAR → segment summary (pairutil.h:63-126).  Returns 1 ok / 0 error. */
typedef struct { int segCnt, leftClip, rightClip, mappable; int left[4], right[4]; } orc_segment;
int orc_cigar2segment(const char *cigar, size_t len, int start, orc_segment *s);

/* ---- krmdup (src/preprocess/krmdup.cpp) ---- */
typedef struct { uint32_t uniq, dup, discard; } orc_dd_stats;
typedef struct orc_dedup orc_dedup;           /* persistent key sets = one krmdup process */
orc_dedup *orc_dedup_new(int hskip1, int klen1, int hskip2, int klen2);
void       orc_dedup_free(orc_dedup *);
/* Whole interleaved FASTQ in, read1/read2 FASTQ out, in krmdup's file order
 * (per 65 536-pair batch: A, C, G, T buckets; krmdup.cpp:216-226). */
int orc_krmdup(orc_dedup *d, const char *fq, size_t n,
               char **r1_out, size_t *r1_len, char **r2_out, size_t *r2_len, orc_dd_stats *st);
/* key only: returns 0 = valid (key written), 1 = discarded by loader, 2 = discarded by worker */
int orc_krmdup_key(const char *seq1, size_t l1, const char *seq2, size_t l2,
                   int hskip1, int klen1, int hskip2, int klen2, uint64_t *key, int *bucket);

/* ---- packed pairs: coordinate dedup + binning (parity unpinned) ---- */
typedef struct {
    uint32_t pos1, pos2;
    uint16_t chr1, chr2;       /* index into the chromosome table */
    uint8_t  strands;          /* bit0: strand1 is '-', bit1: strand2 is '-' */
    uint8_t  cls;              /* 0 trans, 1 cis10K, 2 cis1K, 3 cis0 */
    uint16_t lane;
} orc_pair;                    /* 16 bytes, same layout as mk_pair in include/microcket_b200.h */

/* parse .pairs text (lines starting with '#' skipped); chromosome names looked up in names[0..n_chrom) */
long orc_pairs_parse(const char *text, size_t n, const char *const *names, int n_chrom,
                     orc_pair *out, size_t cap);
/* keep[i] = 1 iff pair i is the first in input order with its (lane,chr1,pos1,s1,chr2,pos2,s2) key */
size_t orc_coord_dedup(const orc_pair *p, size_t n, uint8_t *keep);
/* COO counts for one resolution, sorted by (bin1,bin2), bin = offset[chr] + pos/res,
 * offset[] = running sum of (len/res + 1).  Returns nnz (or -1 if cap too small). */
long orc_bin_coo(const orc_pair *p, size_t n, const uint8_t *keep,
                 const uint32_t *chrom_len, int n_chrom, uint32_t res,
                 uint32_t *bin1, uint32_t *bin2, uint32_t *cnt, size_t cap);

void orc_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
