/*
 * krmdup_oracle.c — plain-C restatement of the reference's krmdup.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Parity: PINNED against
 * oracle/_ref/krmdup (the reference's own source) by tests/.
 *
 * Follows /root/reference/src/preprocess/krmdup.cpp:
 *   loader, bucket by first key base, loader discards ... :88-149
 *   length test, 2-bit key packing, worker discards ..... :156-198
 *   first-occurrence-wins set probe ..................... :201-212
 *   per-batch output order A,C,G,T ...................... :216-226, :328-364
 *   log ................................................. :383-389
 *
 * The four unordered_sets of :322 persist for the whole process, so an
 * orc_dedup object == one krmdup process (one lane when the driver runs
 * with -b, microcket:428-451).
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

#define DD_BATCH (1u << 16)   /* krmdup.cpp:19 */

typedef struct { uint64_t *slot; uint8_t *used; size_t cap, cnt; } kset;

struct orc_dedup {
    int hskip1, klen1, hskip2, klen2;
    kset set[4];
};

static uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

static void kset_init(kset *s, size_t cap) {
    s->cap = cap; s->cnt = 0;
    s->slot = (uint64_t *)malloc(cap * sizeof(uint64_t));
    s->used = (uint8_t *)calloc(cap, 1);
}

static int kset_insert(kset *s, uint64_t k);   /* 1 = newly inserted, 0 = already present */

static void kset_grow(kset *s) {
    kset n; kset_init(&n, s->cap * 2);
    for (size_t i = 0; i < s->cap; ++i) if (s->used[i]) kset_insert(&n, s->slot[i]);
    free(s->slot); free(s->used); *s = n;
}

static int kset_insert(kset *s, uint64_t k) {
    if ((s->cnt + 1) * 10 > s->cap * 7) kset_grow(s);
    size_t m = s->cap - 1, i = (size_t)mix64(k) & m;
    while (s->used[i]) { if (s->slot[i] == k) return 0; i = (i + 1) & m; }
    s->used[i] = 1; s->slot[i] = k; ++s->cnt;
    return 1;
}

orc_dedup *orc_dedup_new(int hskip1, int klen1, int hskip2, int klen2) {
    if (klen1 + klen2 > 32 || klen1 + klen2 < 16) return NULL;    /* krmdup.cpp:259-262 */
    orc_dedup *d = (orc_dedup *)calloc(1, sizeof *d);
    d->hskip1 = hskip1; d->klen1 = klen1; d->hskip2 = hskip2; d->klen2 = klen2;
    for (int i = 0; i < 4; ++i) kset_init(&d->set[i], 1 << 12);
    return d;
}

void orc_dedup_free(orc_dedup *d) {
    if (!d) return;
    for (int i = 0; i < 4; ++i) { free(d->set[i].slot); free(d->set[i].used); }
    free(d);
}

static int base_code(char c, uint64_t *code) {      /* krmdup.cpp:171-174 */
    switch (c) {
    case 'A': case 'a': *code = 1; return 1;
    case 'T': case 't': *code = 2; return 1;
    case 'C': case 'c': *code = 0; return 1;
    case 'G': case 'g': *code = 3; return 1;
    default: return 0;
    }
}

int orc_krmdup_key(const char *seq1, size_t l1, const char *seq2, size_t l2,
                   int hskip1, int klen1, int hskip2, int klen2, uint64_t *key, int *bucket) {
    size_t e1 = (size_t)(hskip1 + klen1), e2 = (size_t)(hskip2 + klen2);
    char first = l1 >= e1 ? seq1[hskip1] : 'N';                    /* :103-106 */
    if (first == 'N') return 1;                                     /* :108-111 */
    *bucket = first == 'A' ? 0 : first == 'C' ? 1 : first == 'G' ? 2 : 3;   /* :113-141 */
    if (l1 < e1 || l2 < e2) return 2;                               /* :158-161 */
    uint64_t k = 0, c;
    for (size_t i = (size_t)hskip1; i < e1; ++i) { if (!base_code(seq1[i], &c)) return 2; k = (k << 2) | c; }
    for (size_t i = (size_t)hskip2; i < e2; ++i) { if (!base_code(seq2[i], &c)) return 2; k = (k << 2) | c; }
    *key = k;
    return 0;
}

typedef struct { const char *p; size_t n; } sl;
typedef struct { char *p; size_t n, cap; } obuf;
static void oput(obuf *b, const char *s, size_t n) {
    if (b->n + n + 1 > b->cap) {
        size_t nc = b->cap ? b->cap * 2 : 1 << 16;
        while (nc < b->n + n + 1) nc *= 2;
        b->p = (char *)realloc(b->p, nc); b->cap = nc;
    }
    memcpy(b->p + b->n, s, n); b->n += n;
}
static void orec(obuf *b, sl id, sl seq, sl qual) {   /* "%s\n%s\n+\n%s\n", krmdup.cpp:206-209 */
    oput(b, id.p, id.n); oput(b, "\n", 1); oput(b, seq.p, seq.n); oput(b, "\n+\n", 3);
    oput(b, qual.p, qual.n); oput(b, "\n", 1);
}

int orc_krmdup(orc_dedup *d, const char *fq, size_t n,
               char **r1_out, size_t *r1_len, char **r2_out, size_t *r2_len, orc_dd_stats *st) {
    st->uniq = st->dup = st->discard = 0;
    size_t nl = 0, cap = 1024;
    sl *lines = (sl *)malloc(cap * sizeof(sl));
    for (size_t i = 0; i < n;) {
        const char *e = (const char *)memchr(fq + i, '\n', n - i);
        size_t len = e ? (size_t)(e - (fq + i)) : n - i;
        if (nl == cap) { cap *= 2; lines = (sl *)realloc(lines, cap * sizeof(sl)); }
        lines[nl].p = fq + i; lines[nl].n = len; ++nl;
        i += len + 1;
    }
    size_t npair = nl / 8;       /* a trailing partial record is outside the restated domain */
    obuf o1 = {0, 0, 0}, o2 = {0, 0, 0};
    uint8_t *keep = (uint8_t *)malloc(DD_BATCH), *bkt = (uint8_t *)malloc(DD_BATCH);
    for (size_t b0 = 0; b0 < npair; b0 += DD_BATCH) {
        size_t bn = npair - b0 < DD_BATCH ? npair - b0 : DD_BATCH;
        /* the set probe order inside a bucket is input order; buckets are independent sets */
        for (size_t j = 0; j < bn; ++j) {
            const sl *L = lines + (b0 + j) * 8;
            uint64_t key; int bucket = 3;
            int rc = orc_krmdup_key(L[1].p, L[1].n, L[5].p, L[5].n, d->hskip1, d->klen1, d->hskip2, d->klen2, &key, &bucket);
            keep[j] = 0; bkt[j] = (uint8_t)bucket;
            if (rc) { ++st->discard; continue; }
            if (kset_insert(&d->set[bucket], key)) { keep[j] = 1; ++st->uniq; } else ++st->dup;
        }
        for (int bucket = 0; bucket < 4; ++bucket)
            for (size_t j = 0; j < bn; ++j) {
                if (!keep[j] || bkt[j] != bucket) continue;
                const sl *L = lines + (b0 + j) * 8;
                orec(&o1, L[0], L[1], L[3]);
                orec(&o2, L[4], L[5], L[7]);
            }
    }
    free(keep); free(bkt); free(lines);
    if (r1_out) { *r1_out = o1.p; *r1_len = o1.n; } else free(o1.p);
    if (r2_out) { *r2_out = o2.p; *r2_len = o2.n; } else free(o2.p);
    return 0;
}
