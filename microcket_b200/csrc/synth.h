/*
 * synth.h — counter-based synthetic SAM / FASTQ generator shared by host (g++)
 * and device (nvcc) code.  Every read group is a pure function of
 * (seed, group index), so any shard can be produced independently on the CPU or
 * on the GPU and the bytes are identical.
 *
 * Read IDs carry the truth the way the reference's simulation utilities expect
 * (`<n>#chrA:posA-chrB:posB`, util/simulation/split.sim3C.pl:41-47,
 * check.accuracy.pl:24-32).  Nothing here emits the inputs on which the
 * reference itself is undefined: no `= X P` ops, no clip after
 * an `N`, no `*` CIGAR with MAPQ >= 1.
 */
#ifndef MICROCKET_SYNTH_H
#define MICROCKET_SYNTH_H
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define MK_HD __host__ __device__ __forceinline__
#else
#define MK_HD static inline
#endif

#define MK_SYN_MAXCHR 32

typedef struct {
    uint64_t seed;
    int mode;                 /* 0 = flash (stitched, single-end records), 1 = unc (paired records) */
    int n_chrom;
    uint32_t chrom_len[MK_SYN_MAXCHR];
    uint64_t chrom_off[MK_SYN_MAXCHR + 1];   /* running sum of lengths */
    char chrom_name[MK_SYN_MAXCHR][8];
    uint8_t chrom_name_len[MK_SYN_MAXCHR];
    int read_len;             /* cycles per mate (unc) */
    int min_stitch, max_stitch; /* stitched read length range (flash) */
    /* mixture, per 1024 */
    int w_chimeric;           /* groups with a split alignment */
    int w_trans, w_far, w_mid; /* pair geometry: trans / cis >= 10 kb / cis 1-10 kb; rest cis < 1 kb */
    int w_noise;              /* low MAPQ, unmapped, secondary, missing mates, odd clips, introns */
    int w_selfcircle;
    int dup_per_1024;         /* FASTQ generator: fraction of pairs that re-use an earlier fragment */
    int n_lanes;
    /* SAM generator: a group re-uses the fragment (loci, strands, CIGARs, bases) of group `hash % sam_dup_universe` with
       probability sam_dup_per_1024/1024 — PCR duplicates with their own read ids.  The source is anywhere in the universe
       (the whole job, all shards), so duplicates cross shard boundaries in both directions.  0 = off. */
    int sam_dup_per_1024;
    uint64_t sam_dup_universe;
} mk_synth_cfg;

MK_HD uint64_t mk_mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

typedef struct { uint64_t s; } mk_rng;
MK_HD uint64_t mk_next(mk_rng *r) { r->s += 0x9E3779B97F4A7C15ULL; return mk_mix64(r->s); }
MK_HD uint32_t mk_below(mk_rng *r, uint32_t n) { return (uint32_t)((mk_next(r) >> 32) * (uint64_t)n >> 32); }

/* byte sink: counts when p == NULL */
typedef struct { char *p; size_t n; } mk_sink;
MK_HD void mk_putc(mk_sink *w, char c) { if (w->p) w->p[w->n] = c; ++w->n; }
MK_HD void mk_puts(mk_sink *w, const char *s) { while (*s) mk_putc(w, *s++); }
MK_HD void mk_putn(mk_sink *w, const char *s, int n) { for (int i = 0; i < n; ++i) mk_putc(w, s[i]); }
MK_HD void mk_putu(mk_sink *w, uint64_t v) {
    char t[20]; int n = 0;
    do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) mk_putc(w, t[--n]);
}

MK_HD void mk_put_bases(mk_sink *w, uint64_t key, int n) {
    const char B[4] = {'A', 'C', 'G', 'T'};
    mk_rng r; r.s = key;
    uint64_t word = 0;
    for (int i = 0; i < n; ++i) {
        if ((i & 31) == 0) word = mk_next(&r);
        mk_putc(w, B[word & 3]); word >>= 2;
    }
}
MK_HD void mk_put_qual(mk_sink *w, uint64_t key, int n) {
    mk_rng r; r.s = key ^ 0x5151515151515151ULL;
    uint64_t word = 0;
    for (int i = 0; i < n; ++i) {
        if ((i & 15) == 0) word = mk_next(&r);
        int q = (int)(word & 15); word >>= 4;
        mk_putc(w, q == 0 ? ',' : (q < 3 ? ':' : 'F'));
    }
}

typedef struct { int chr; uint32_t pos; } mk_locus;

MK_HD mk_locus mk_pick_locus(const mk_synth_cfg *c, mk_rng *r, uint32_t margin) {
    uint64_t total = c->chrom_off[c->n_chrom];
    uint64_t g = mk_next(r) % total;
    int lo = 0, hi = c->n_chrom - 1;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (c->chrom_off[mid] <= g) lo = mid; else hi = mid - 1; }
    mk_locus L; L.chr = lo;
    uint32_t len = c->chrom_len[lo];
    uint32_t p = (uint32_t)(g - c->chrom_off[lo]) + 1;
    uint32_t maxp = len > margin + 1 ? len - margin : 1;
    if (p > maxp) p = maxp;
    L.pos = p;
    return L;
}

/* one alignment record of a SAM line */
typedef struct {
    int flag, chr; uint32_t pos; int mapq;
    int op_n[6]; char op_c[6]; int n_op;       /* CIGAR */
    int mate_chr; uint32_t mate_pos;
} mk_aln;

MK_HD void mk_cigar_set(mk_aln *a, int n0, char c0, int n1, char c1, int n2, char c2) {
    a->n_op = 0;
    if (n0 > 0) { a->op_n[a->n_op] = n0; a->op_c[a->n_op++] = c0; }
    if (n1 > 0) { a->op_n[a->n_op] = n1; a->op_c[a->n_op++] = c1; }
    if (n2 > 0) { a->op_n[a->n_op] = n2; a->op_c[a->n_op++] = c2; }
}
MK_HD int mk_cigar_qlen(const mk_aln *a) {       /* bases present in SEQ */
    int q = 0;
    for (int i = 0; i < a->n_op; ++i) if (a->op_c[i] == 'M' || a->op_c[i] == 'I' || a->op_c[i] == 'S') q += a->op_n[i];
    return q;
}
MK_HD uint32_t mk_cigar_rlen(const mk_aln *a) {  /* reference span */
    uint32_t q = 0;
    for (int i = 0; i < a->n_op; ++i) if (a->op_c[i] == 'M' || a->op_c[i] == 'D' || a->op_c[i] == 'N') q += (uint32_t)a->op_n[i];
    return q;
}

MK_HD void mk_put_qname(mk_sink *w, const mk_synth_cfg *c, uint64_t idx, mk_locus A, mk_locus B) {
    mk_puts(w, "SIM:"); mk_putu(w, idx); mk_putc(w, '#');
    mk_putn(w, c->chrom_name[A.chr], c->chrom_name_len[A.chr]); mk_putc(w, ':'); mk_putu(w, A.pos); mk_putc(w, '-');
    mk_putn(w, c->chrom_name[B.chr], c->chrom_name_len[B.chr]); mk_putc(w, ':'); mk_putu(w, B.pos);
}

MK_HD void mk_put_aln(mk_sink *w, const mk_synth_cfg *c, uint64_t idx, mk_locus A, mk_locus B, const mk_aln *a, uint64_t seqkey) {
    mk_put_qname(w, c, idx, A, B); mk_putc(w, '\t');
    mk_putu(w, (uint64_t)a->flag); mk_putc(w, '\t');
    if (a->chr < 0) mk_putc(w, '*'); else mk_putn(w, c->chrom_name[a->chr], c->chrom_name_len[a->chr]);
    mk_putc(w, '\t'); mk_putu(w, a->pos); mk_putc(w, '\t'); mk_putu(w, (uint64_t)a->mapq); mk_putc(w, '\t');
    int qlen;
    if (a->n_op == 0) { mk_putc(w, '*'); qlen = c->mode ? c->read_len : c->min_stitch; }
    else { for (int i = 0; i < a->n_op; ++i) { mk_putu(w, (uint64_t)a->op_n[i]); mk_putc(w, a->op_c[i]); } qlen = mk_cigar_qlen(a); }
    mk_putc(w, '\t');
    if (a->mate_chr < 0) { mk_puts(w, "*\t0\t0\t"); }
    else {
        if (a->mate_chr == a->chr) mk_putc(w, '='); else mk_putn(w, c->chrom_name[a->mate_chr], c->chrom_name_len[a->mate_chr]);
        mk_putc(w, '\t'); mk_putu(w, a->mate_pos); mk_puts(w, "\t0\t");
    }
    mk_put_bases(w, seqkey, qlen); mk_putc(w, '\t');
    mk_put_qual(w, seqkey, qlen);
    mk_puts(w, "\tNM:i:"); mk_putu(w, seqkey & 3);
    mk_puts(w, "\tAS:i:"); mk_putu(w, (uint64_t)qlen);
    mk_puts(w, "\tXS:i:"); mk_putu(w, (seqkey >> 8) & 31);
    mk_putc(w, '\n');
}

/* second locus according to the geometry mixture */
MK_HD mk_locus mk_pick_partner(const mk_synth_cfg *c, mk_rng *r, mk_locus A, int *self_circle) {
    uint32_t t = mk_below(r, 1024);
    *self_circle = 0;
    if (t < (uint32_t)c->w_trans) {
        mk_locus B = mk_pick_locus(c, r, 2000);
        return B;
    }
    t -= (uint32_t)c->w_trans;
    uint32_t d;
    if (t < (uint32_t)c->w_far) d = 10000u + mk_below(r, 2000000u);
    else if (t < (uint32_t)(c->w_far + c->w_mid)) d = 1000u + mk_below(r, 9000u);
    else if (t < (uint32_t)(c->w_far + c->w_mid + c->w_selfcircle)) { d = mk_below(r, 11u); *self_circle = 1; }
    else d = 11u + mk_below(r, 989u);
    mk_locus B; B.chr = A.chr;
    uint32_t len = c->chrom_len[A.chr];
    if (A.pos + d + 2000u < len) B.pos = A.pos + d;
    else B.pos = A.pos > d ? A.pos - d : 1;
    return B;
}

/* a clipped / indel-bearing single alignment of `len` query bases starting at pos */
MK_HD void mk_plain_cigar(mk_aln *a, mk_rng *r, int len, int noisy) {
    uint32_t v = noisy ? mk_below(r, 16) : 15;
    const int clips[8] = {5, 15, 20, 21, 30, 60, 80, 100};
    int k = clips[mk_below(r, 8)]; if (k >= len - 10) k = len / 3;
    switch (v) {
    case 0: mk_cigar_set(a, k, 'S', len - k, 'M', 0, 'M'); break;
    case 1: mk_cigar_set(a, len - k, 'M', k, 'S', 0, 'M'); break;
    case 2: { int i = 1 + (int)mk_below(r, 4); int m = (len - i) / 2; mk_cigar_set(a, m, 'M', i, 'I', len - i - m, 'M'); break; }
    case 3: { int d = 1 + (int)mk_below(r, 6); int m = len / 2; mk_cigar_set(a, m, 'M', d, 'D', len - m, 'M'); break; }
    case 4: mk_cigar_set(a, k, 'H', len - k, 'M', 0, 'M'); break;
    default: mk_cigar_set(a, len, 'M', 0, 'M', 0, 'M'); break;
    }
}

/*
 * Emit read group `idx`.  Returns the number of SAM lines written.
 * unc: R1/R2 records (flags 64/128), chimeric R1 or R2 as primary `aMbS` +
 *      supplementary `aHbM` (bwa mem -5 without -Y), noise as configured.
 * flash: single-end records (flags 0/16/2048/2064).
 */
MK_HD uint64_t mk_sam_fragment(const mk_synth_cfg *c, uint64_t idx) {
    if (c->sam_dup_per_1024 <= 0 || c->sam_dup_universe == 0) return idx;
    mk_rng d; d.s = mk_mix64(c->seed ^ 0xD0B1E5ULL ^ (idx * 0xA24BAED4963EE407ULL));
    if (mk_below(&d, 1024) < (uint32_t)c->sam_dup_per_1024) return mk_next(&d) % c->sam_dup_universe;
    return idx;
}

MK_HD int mk_gen_group(const mk_synth_cfg *c, uint64_t idx, mk_sink *w) {
    mk_rng r; r.s = mk_mix64(c->seed ^ (mk_sam_fragment(c, idx) * 0xD1342543DE82EF95ULL));
    int sc;
    mk_locus A = mk_pick_locus(c, &r, 4000);
    mk_locus B = mk_pick_partner(c, &r, A, &sc);
    int sA = (int)(mk_next(&r) & 1), sB = (int)(mk_next(&r) & 1);   /* 1 = '-' */
    uint32_t kind = mk_below(&r, 1024);
    int noisy = kind >= (uint32_t)(1024 - c->w_noise);
    int chim = !noisy && kind < (uint32_t)c->w_chimeric;
    uint64_t sk = mk_next(&r);
    int lines = 0;
    mk_aln a[4];
    for (int i = 0; i < 4; ++i) { a[i].mate_chr = -1; a[i].mate_pos = 0; a[i].mapq = 60; }

    if (c->mode == 0) {
        int L = c->min_stitch + (int)mk_below(&r, (uint32_t)(c->max_stitch - c->min_stitch + 1));
        if (noisy) {
            uint32_t v = mk_below(&r, 8);
            if (v == 0) {          /* unmapped */
                a[0].flag = 4; a[0].chr = -1; a[0].pos = 0; a[0].mapq = 0; a[0].n_op = 0;
                mk_put_aln(w, c, idx, A, B, &a[0], sk); return 1;
            } else if (v == 1) {   /* three hits */
                int p = L / 3;
                a[0].flag = sA ? 16 : 0; a[0].chr = A.chr; a[0].pos = A.pos; mk_cigar_set(&a[0], p, 'M', L - p, 'S', 0, 'M');
                a[1].flag = 2048 | (sB ? 16 : 0); a[1].chr = B.chr; a[1].pos = B.pos; mk_cigar_set(&a[1], p, 'H', p, 'M', L - 2 * p, 'H');
                mk_locus C = mk_pick_locus(c, &r, 4000);
                a[2].flag = 2048; a[2].chr = C.chr; a[2].pos = C.pos; mk_cigar_set(&a[2], 2 * p, 'H', L - 2 * p, 'M', 0, 'M');
                for (int i = 0; i < 3; ++i) mk_put_aln(w, c, idx, A, B, &a[i], sk + (uint64_t)i);
                return 3;
            } else if (v == 2) {   /* spliced (STAR style) */
                int p = L / 2; int gap = 200 + (int)mk_below(&r, 3000);
                a[0].flag = sA ? 16 : 0; a[0].chr = A.chr; a[0].pos = A.pos; mk_cigar_set(&a[0], p, 'M', gap, 'N', L - p, 'M');
                if (mk_below(&r, 4) == 0) { a[0].n_op = 5; a[0].op_n[3] = 300; a[0].op_c[3] = 'N'; a[0].op_n[4] = 20; a[0].op_c[4] = 'M'; a[0].op_n[2] = L - p - 20; }
                mk_put_aln(w, c, idx, A, B, &a[0], sk); return 1;
            } else if (v == 3) {   /* low MAPQ */
                a[0].flag = sA ? 16 : 0; a[0].chr = A.chr; a[0].pos = A.pos; a[0].mapq = (int)mk_below(&r, 10);
                mk_cigar_set(&a[0], L, 'M', 0, 'M', 0, 'M');
                mk_put_aln(w, c, idx, A, B, &a[0], sk); return 1;
            } else if (v == 4) {   /* secondary alignment beside a good record */
                a[0].flag = sA ? 16 : 0; a[0].chr = A.chr; a[0].pos = A.pos; mk_cigar_set(&a[0], L, 'M', 0, 'M', 0, 'M');
                a[1].flag = 256 | (sB ? 16 : 0); a[1].chr = B.chr; a[1].pos = B.pos; a[1].mapq = 60; mk_cigar_set(&a[1], L, 'M', 0, 'M', 0, 'M');
                mk_put_aln(w, c, idx, A, B, &a[0], sk); mk_put_aln(w, c, idx, A, B, &a[1], sk + 1); return 2;
            } else {               /* clipped / indel single record (lowMap boundary cases) */
                a[0].flag = sA ? 16 : 0; a[0].chr = A.chr; a[0].pos = A.pos; mk_plain_cigar(&a[0], &r, L, 1);
                mk_put_aln(w, c, idx, A, B, &a[0], sk); return 1;
            }
        }
        if (!chim) {
            a[0].flag = sA ? 16 : 0; a[0].chr = A.chr; a[0].pos = A.pos; mk_cigar_set(&a[0], L, 'M', 0, 'M', 0, 'M');
            mk_put_aln(w, c, idx, A, B, &a[0], sk); return 1;
        }
        /* two records: the junction splits the stitched read at p */
        int p = 30 + (int)mk_below(&r, (uint32_t)(L - 60));
        a[0].flag = sA ? 16 : 0; a[0].chr = A.chr; a[0].pos = A.pos;
        a[1].flag = 2048 | (sB ? 16 : 0); a[1].chr = B.chr; a[1].pos = B.pos;
        if (!sA) mk_cigar_set(&a[0], p, 'M', L - p, 'S', 0, 'M'); else mk_cigar_set(&a[0], L - p, 'S', p, 'M', 0, 'M');
        if (!sB) mk_cigar_set(&a[1], p, 'H', L - p, 'M', 0, 'M'); else mk_cigar_set(&a[1], L - p, 'M', p, 'H', 0, 'M');
        int first = (int)(mk_next(&r) & 1);      /* supplementary may be the longer piece */
        mk_put_aln(w, c, idx, A, B, &a[first ? 1 : 0], sk); mk_put_aln(w, c, idx, A, B, &a[first ? 0 : 1], sk + 1);
        return 2;
    }

    /* ---- unc ---- */
    int L = c->read_len;
    if (noisy) {
        uint32_t v = mk_below(&r, 10);
        if (v == 0) {              /* both mates unmapped */
            a[0].flag = 77; a[0].chr = -1; a[0].pos = 0; a[0].mapq = 0; a[0].n_op = 0;
            a[1].flag = 141; a[1].chr = -1; a[1].pos = 0; a[1].mapq = 0; a[1].n_op = 0;
            mk_put_aln(w, c, idx, A, B, &a[0], sk); mk_put_aln(w, c, idx, A, B, &a[1], sk + 1); return 2;
        }
        a[0].flag = 65 | (sA ? 16 : 0) | (sB ? 32 : 0); a[0].chr = A.chr; a[0].pos = A.pos; a[0].mate_chr = B.chr; a[0].mate_pos = B.pos;
        a[1].flag = 129 | (sB ? 16 : 0) | (sA ? 32 : 0); a[1].chr = B.chr; a[1].pos = B.pos; a[1].mate_chr = A.chr; a[1].mate_pos = A.pos;
        mk_cigar_set(&a[0], L, 'M', 0, 'M', 0, 'M'); mk_cigar_set(&a[1], L, 'M', 0, 'M', 0, 'M');
        if (v == 1) { a[0].mapq = (int)mk_below(&r, 10); }                  /* R1 filtered → missing mate */
        else if (v == 2) { a[1].mapq = (int)mk_below(&r, 10); }
        else if (v == 3) { mk_plain_cigar(&a[0], &r, L, 1); }               /* clip boundary cases */
        else if (v == 4) { mk_plain_cigar(&a[1], &r, L, 1); }
        else if (v == 5) {         /* secondary record between the mates */
            a[2] = a[1]; a[2].flag |= 256;
            mk_put_aln(w, c, idx, A, B, &a[0], sk); mk_put_aln(w, c, idx, A, B, &a[2], sk + 2); mk_put_aln(w, c, idx, A, B, &a[1], sk + 1);
            return 3;
        } else if (v == 6) {       /* spliced R1 that must mate with R2 on the same chromosome */
            int gap = 100 + (int)mk_below(&r, 600);
            a[0].flag = 65 | (sA ? 16 : 0); a[1].flag = 129 | (sA ? 0 : 16);
            a[1].chr = A.chr;
            mk_cigar_set(&a[0], L / 2, 'M', gap, 'N', L - L / 2, 'M');
            uint32_t span = (uint32_t)(L + gap);
            if (!sA) { a[0].pos = A.pos; a[1].pos = A.pos + span + mk_below(&r, 700); }
            else { a[1].pos = A.pos; a[0].pos = A.pos + (uint32_t)L + mk_below(&r, 700); }
        } else if (v == 7) {       /* spliced R2 */
            int gap = 100 + (int)mk_below(&r, 600);
            a[0].flag = 65 | (sA ? 16 : 0); a[1].flag = 129 | (sA ? 0 : 16);
            a[1].chr = A.chr;
            mk_cigar_set(&a[1], L / 2, 'M', gap, 'N', L - L / 2, 'M');
            if (!sA) { a[0].pos = A.pos; a[1].pos = A.pos + 50 + mk_below(&r, 900); }
            else { a[1].pos = A.pos; a[0].pos = A.pos + (uint32_t)(L / 2 + gap) + 10 + mk_below(&r, 800); }
        } else if (v == 8) {       /* both mates split: 2 + 2 records, silently dropped */
            int p = L / 2;
            mk_cigar_set(&a[0], p, 'M', L - p, 'S', 0, 'M'); mk_cigar_set(&a[1], p, 'M', L - p, 'S', 0, 'M');
            a[2] = a[0]; a[2].flag |= 2048; a[2].chr = B.chr; a[2].pos = B.pos + 300; mk_cigar_set(&a[2], p, 'H', L - p, 'M', 0, 'M');
            a[3] = a[1]; a[3].flag |= 2048; a[3].chr = A.chr; a[3].pos = A.pos + 300; mk_cigar_set(&a[3], p, 'H', L - p, 'M', 0, 'M');
            mk_put_aln(w, c, idx, A, B, &a[0], sk); mk_put_aln(w, c, idx, A, B, &a[2], sk + 2);
            mk_put_aln(w, c, idx, A, B, &a[1], sk + 1); mk_put_aln(w, c, idx, A, B, &a[3], sk + 3);
            return 4;
        } else {                   /* a record that is neither first nor second in pair */
            a[2] = a[0]; a[2].flag = 1 | 2048; a[2].pos = A.pos + 77;
            mk_put_aln(w, c, idx, A, B, &a[0], sk); mk_put_aln(w, c, idx, A, B, &a[2], sk + 2); mk_put_aln(w, c, idx, A, B, &a[1], sk + 1);
            return 3;
        }
        mk_put_aln(w, c, idx, A, B, &a[0], sk); mk_put_aln(w, c, idx, A, B, &a[1], sk + 1);
        return 2;
    }
    if (!chim) {
        a[0].flag = 65 | (sA ? 16 : 0) | (sB ? 32 : 0); a[0].chr = A.chr; a[0].pos = A.pos; a[0].mate_chr = B.chr; a[0].mate_pos = B.pos;
        a[1].flag = 129 | (sB ? 16 : 0) | (sA ? 32 : 0); a[1].chr = B.chr; a[1].pos = B.pos; a[1].mate_chr = A.chr; a[1].mate_pos = A.pos;
        mk_plain_cigar(&a[0], &r, L, 0); mk_plain_cigar(&a[1], &r, L, 0);
        mk_put_aln(w, c, idx, A, B, &a[0], sk); mk_put_aln(w, c, idx, A, B, &a[1], sk + 1);
        return 2;
    }
    /* chimeric: the split mate has a piece at A (5' end) and a piece at B; the lone mate sits near the B piece */
    int split_r1 = (int)(mk_next(&r) & 1);
    int p = 25 + (int)mk_below(&r, (uint32_t)(L - 50));
    uint32_t v = mk_below(&r, 16);
    int bad_strand = v == 0, too_far = v == 1, wrong_chr = v == 2, supp_first = v >= 12;
    int fsplit = split_r1 ? 64 : 128, flone = split_r1 ? 128 : 64;
    /* primary piece at A */
    a[0].flag = 1 | fsplit | (sA ? 16 : 0); a[0].chr = A.chr; a[0].pos = A.pos;
    if (!sA) mk_cigar_set(&a[0], p, 'M', L - p, 'S', 0, 'M'); else mk_cigar_set(&a[0], L - p, 'S', p, 'M', 0, 'M');
    /* supplementary piece at B */
    a[1].flag = 1 | fsplit | 2048 | (sB ? 16 : 0); a[1].chr = B.chr; a[1].pos = B.pos;
    if (!sB) mk_cigar_set(&a[1], p, 'H', L - p, 'M', 0, 'M'); else mk_cigar_set(&a[1], L - p, 'M', p, 'H', 0, 'M');
    /* lone mate: opposite strand to the B piece, downstream of it when the piece is '+' */
    int sl = bad_strand ? sB : !sB;
    a[2].flag = 1 | flone | (sl ? 16 : 0); a[2].chr = wrong_chr ? A.chr : B.chr;
    uint32_t off = too_far ? 1500u + mk_below(&r, 3000u) : 20u + mk_below(&r, 800u);
    if (!sB) a[2].pos = B.pos + off; else a[2].pos = B.pos > off + (uint32_t)L ? B.pos - off : 1;
    mk_cigar_set(&a[2], L, 'M', 0, 'M', 0, 'M');
    a[0].mate_chr = a[1].mate_chr = a[2].chr; a[0].mate_pos = a[1].mate_pos = a[2].pos;
    a[2].mate_chr = a[0].chr; a[2].mate_pos = a[0].pos;
    int o0 = supp_first ? 1 : 0, o1 = supp_first ? 0 : 1;
    if (split_r1) {
        mk_put_aln(w, c, idx, A, B, &a[o0], sk); mk_put_aln(w, c, idx, A, B, &a[o1], sk + 1); mk_put_aln(w, c, idx, A, B, &a[2], sk + 2);
    } else {
        mk_put_aln(w, c, idx, A, B, &a[2], sk + 2); mk_put_aln(w, c, idx, A, B, &a[o0], sk); mk_put_aln(w, c, idx, A, B, &a[o1], sk + 1);
    }
    (void)lines;
    return 3;
}

/* ---------------- FASTQ (krmdup input) ---------------- */
/*
 * Pair `idx` of an interleaved FASTQ.  A pair re-uses the fragment of an
 * earlier pair with probability dup_per_1024/1024; the first 21+ bases of each
 * mate are a pure function of the fragment (its two 5' ends), so sequence-key
 * duplicates coincide with coordinate duplicates.
 */
MK_HD uint64_t mk_fastq_fragment(const mk_synth_cfg *c, uint64_t idx) {
    mk_rng r; r.s = mk_mix64(c->seed ^ 0xF00DULL ^ (idx * 0xA24BAED4963EE407ULL));
    uint32_t t = mk_below(&r, 1024);
    if (idx > 0 && t < (uint32_t)c->dup_per_1024) return mk_next(&r) % idx;
    return idx;
}

MK_HD void mk_gen_fastq_pair(const mk_synth_cfg *c, uint64_t idx, mk_sink *w) {
    uint64_t frag = mk_fastq_fragment(c, idx);
    mk_rng fr; fr.s = mk_mix64(c->seed ^ 0xBEEFULL ^ (frag * 0x9FB21C651E98DF25ULL));
    uint64_t k1 = mk_next(&fr), k2 = mk_next(&fr);
    mk_rng r; r.s = mk_mix64(c->seed ^ 0xC0FFEEULL ^ (idx * 0xD6E8FEB86659FD93ULL));
    uint32_t odd = mk_below(&r, 4096);
    int L1 = c->read_len - (int)mk_below(&r, 30), L2 = c->read_len - (int)mk_below(&r, 30);
    int n_at = -1, n_mate = 0;
    if (odd == 0) L1 = 12 + (int)mk_below(&r, 9);          /* too short: discarded by the loader */
    else if (odd == 1) L2 = 12 + (int)mk_below(&r, 9);     /* too short: discarded by the worker */
    else if (odd == 2) { n_at = 5; }                       /* N at the first key base */
    else if (odd == 3) { n_at = 6 + (int)mk_below(&r, 15); n_mate = (int)(mk_next(&r) & 1); }
    for (int m = 0; m < 2; ++m) {
        int L = m ? L2 : L1;
        mk_puts(w, "@SIM:"); mk_putu(w, idx); mk_putc(w, '#'); mk_putu(w, frag); mk_putc(w, '/'); mk_putc(w, m ? '2' : '1'); mk_putc(w, '\n');
        /* head (fragment-determined) then tail (pair-determined) */
        size_t start = w->n;
        int head = L < 32 ? L : 32;
        mk_put_bases(w, m ? k2 : k1, head);
        if (L > head) mk_put_bases(w, mk_next(&r), L - head);
        if (n_at >= 0 && n_mate == m && n_at < L && w->p) w->p[start + (size_t)n_at] = 'N';
        mk_puts(w, "\n+\n");
        mk_put_qual(w, mk_next(&r), L);
        mk_putc(w, '\n');
    }
}

/* ---------------- genomes (chromosome sizes as in anno/hg38.info, anno/mm10.info; lexical order) ---------------- */
static const char *const MK_HG38_NAMES[25] = {"chr1","chr10","chr11","chr12","chr13","chr14","chr15","chr16","chr17","chr18","chr19","chr2","chr20","chr21","chr22","chr3","chr4","chr5","chr6","chr7","chr8","chr9","chrM","chrX","chrY"};
static const uint32_t MK_HG38_LEN[25] = {248956422u,133797422u,135086622u,133275309u,114364328u,107043718u,101991189u,90338345u,83257441u,80373285u,58617616u,242193529u,64444167u,46709983u,50818468u,198295559u,190214555u,181538259u,170805979u,159345973u,145138636u,138394717u,16569u,156040895u,57227415u};
static const char *const MK_MM10_NAMES[22] = {"chr1","chr10","chr11","chr12","chr13","chr14","chr15","chr16","chr17","chr18","chr19","chr2","chr3","chr4","chr5","chr6","chr7","chr8","chr9","chrM","chrX","chrY"};
static const uint32_t MK_MM10_LEN[22] = {195471971u,130694993u,122082543u,120129022u,120421639u,124902244u,104043685u,98207768u,94987271u,90702639u,61431566u,182113224u,160039680u,156508116u,151834684u,149736546u,145441459u,129401213u,124595110u,16299u,171031299u,91744698u};

static inline void mk_synth_init(mk_synth_cfg *c, uint64_t seed, int mode, int mm10) {
    c->seed = seed; c->mode = mode;
    c->n_chrom = mm10 ? 22 : 25;
    c->chrom_off[0] = 0;
    for (int i = 0; i < c->n_chrom; ++i) {
        const char *nm = mm10 ? MK_MM10_NAMES[i] : MK_HG38_NAMES[i];
        int l = 0; while (nm[l]) { c->chrom_name[i][l] = nm[l]; ++l; }
        c->chrom_name_len[i] = (uint8_t)l;
        c->chrom_len[i] = mm10 ? MK_MM10_LEN[i] : MK_HG38_LEN[i];
        c->chrom_off[i + 1] = c->chrom_off[i] + c->chrom_len[i];
    }
    c->read_len = 150; c->min_stitch = 150; c->max_stitch = 290;
    c->w_chimeric = mode ? 256 : 358;       /* unc: 25 % split; flash: 35 % two-record */
    c->w_trans = 256; c->w_far = 236; c->w_mid = 20; c->w_selfcircle = 5;
    c->w_noise = 51;                        /* 5 % */
    c->dup_per_1024 = 205; c->n_lanes = 1;
    c->sam_dup_per_1024 = 0; c->sam_dup_universe = 0;
}

#endif
