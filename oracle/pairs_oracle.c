/*
 * pairs_oracle.c — CPU restatement of coordinate-keyed duplicate removal and
 * of contact binning over packed pairs.  TEST INFRASTRUCTURE ONLY (oracle.h).
 *
 * PARITY UNPINNED: the reference has no implementation of either step
 * (krmdup keys on read bases, krmdup.cpp:168-193; binning is delegated to the
 * absent juicer_tools.jar, microcket:525-529).  Semantics restated from
 *   .pairs columns ............ anno/4DN.DCIC.header:2
 *   bin = int(pos / res) ...... util/analyze.EBV/calc.loop2EBV.pl:28,
 *                               util/analyze.EBV/calc.inter.EBV.matrix.and.circos.pl:34,39
 *   chromosome set and order .. anno/<genome>.info
 * tests/ cross-check this file against an independent numpy version.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

long orc_pairs_parse(const char *text, size_t n, const char *const *names, int n_chrom,
                     orc_pair *out, size_t cap) {
    size_t cnt = 0;
    for (size_t i = 0; i < n;) {
        const char *e = (const char *)memchr(text + i, '\n', n - i);
        size_t len = e ? (size_t)(e - (text + i)) : n - i;
        const char *p = text + i, *end = p + len;
        i += len + 1;
        if (len == 0 || *p == '#') continue;
        const char *f[8]; size_t fl[8]; int nf = 0;
        while (nf < 7 && p <= end) {
            const char *t = (const char *)memchr(p, '\t', (size_t)(end - p));
            if (!t) t = end;
            f[nf] = p; fl[nf] = (size_t)(t - p); ++nf; p = t + 1;
        }
        if (nf < 7) return -2;
        if (cnt >= cap) return -1;
        orc_pair r; memset(&r, 0, sizeof r);
        int c[2] = {-1, -1};
        for (int k = 0; k < 2; ++k)
            for (int j = 0; j < n_chrom; ++j)
                if (strlen(names[j]) == fl[1 + 2 * k] && memcmp(names[j], f[1 + 2 * k], fl[1 + 2 * k]) == 0) { c[k] = j; break; }
        if (c[0] < 0 || c[1] < 0) return -3;
        r.chr1 = (uint16_t)c[0]; r.chr2 = (uint16_t)c[1];
        r.pos1 = (uint32_t)strtoul(f[2], NULL, 10); r.pos2 = (uint32_t)strtoul(f[4], NULL, 10);
        r.strands = (uint8_t)((f[5][0] == '-' ? 1 : 0) | (f[6][0] == '-' ? 2 : 0));
        if (r.chr1 != r.chr2) r.cls = 0;
        else { uint32_t d = r.pos2 - r.pos1; r.cls = d >= 10000u ? 1 : (d >= 1000u ? 2 : 3); }
        out[cnt++] = r;
    }
    return (long)cnt;
}

typedef struct { orc_pair p; size_t idx; } keyed;

static int key_cmp(const orc_pair *a, const orc_pair *b) {
    if (a->lane != b->lane) return a->lane < b->lane ? -1 : 1;
    if (a->chr1 != b->chr1) return a->chr1 < b->chr1 ? -1 : 1;
    if (a->pos1 != b->pos1) return a->pos1 < b->pos1 ? -1 : 1;
    if (a->chr2 != b->chr2) return a->chr2 < b->chr2 ? -1 : 1;
    if (a->pos2 != b->pos2) return a->pos2 < b->pos2 ? -1 : 1;
    if (a->strands != b->strands) return a->strands < b->strands ? -1 : 1;
    return 0;
}

static int keyed_cmp(const void *x, const void *y) {
    const keyed *a = (const keyed *)x, *b = (const keyed *)y;
    int c = key_cmp(&a->p, &b->p);
    if (c) return c;
    return a->idx < b->idx ? -1 : (a->idx > b->idx ? 1 : 0);
}

size_t orc_coord_dedup(const orc_pair *p, size_t n, uint8_t *keep) {
    keyed *k = (keyed *)malloc((n ? n : 1) * sizeof(keyed));
    for (size_t i = 0; i < n; ++i) { k[i].p = p[i]; k[i].idx = i; }
    qsort(k, n, sizeof(keyed), keyed_cmp);
    size_t kept = 0;
    for (size_t i = 0; i < n; ++i) {
        int first = i == 0 || key_cmp(&k[i].p, &k[i - 1].p) != 0;
        keep[k[i].idx] = (uint8_t)first;
        kept += (size_t)first;
    }
    free(k);
    return kept;
}

static int u64_cmp(const void *x, const void *y) {
    uint64_t a = *(const uint64_t *)x, b = *(const uint64_t *)y;
    return a < b ? -1 : (a > b ? 1 : 0);
}

long orc_bin_coo(const orc_pair *p, size_t n, const uint8_t *keep,
                 const uint32_t *chrom_len, int n_chrom, uint32_t res,
                 uint32_t *bin1, uint32_t *bin2, uint32_t *cnt, size_t cap) {
    uint64_t *off = (uint64_t *)malloc((size_t)(n_chrom + 1) * sizeof(uint64_t));
    off[0] = 0;
    for (int c = 0; c < n_chrom; ++c) off[c + 1] = off[c] + chrom_len[c] / res + 1;
    uint64_t *key = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
    size_t m = 0;
    for (size_t i = 0; i < n; ++i) {
        if (keep && !keep[i]) continue;
        uint64_t a = off[p[i].chr1] + p[i].pos1 / res, b = off[p[i].chr2] + p[i].pos2 / res;
        if (a > b) { uint64_t t = a; a = b; b = t; }      /* upper triangle */
        key[m++] = (a << 32) | b;
    }
    qsort(key, m, sizeof(uint64_t), u64_cmp);
    size_t nnz = 0;
    for (size_t i = 0; i < m;) {
        size_t j = i; while (j < m && key[j] == key[i]) ++j;
        if (nnz >= cap) { free(off); free(key); return -1; }
        bin1[nnz] = (uint32_t)(key[i] >> 32); bin2[nnz] = (uint32_t)key[i]; cnt[nnz] = (uint32_t)(j - i);
        ++nnz; i = j;
    }
    free(off); free(key);
    return (long)nnz;
}
