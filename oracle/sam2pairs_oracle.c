/*
 * sam2pairs_oracle.c — plain-C restatement of the reference's sam2pairs.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Parity: PINNED against
 * oracle/_ref/sam2pairs (the reference's own sources) by tests/.
 *
 * Follows, by reference file:line (all under /root/reference/src/sam2pairs/):
 *   record filter + grouping ......... pairutil.h:136-177, sam2pairs.cpp:116-131
 *   CIGAR walk ....................... pairutil.h:63-126
 *   integrity (fp32) ................. pairutil.h:180-208
 *   stitched-read resolution ......... flash2pairs.h:17-155
 *   paired-read resolution ........... unc2pairs.h:16-358
 *   ordering / classes / emission .... unc2pairs.h:310-348 == flash2pairs.h:104-144
 *   batch split, log, selfCircle quirk sam2pairs.cpp:143-219
 *
 * Deliberate restatement choices (inputs outside these are undefined in the
 * reference itself):
 *   - a group whose CIGAR makes cigar2segment() return false is dropped and
 *     counted in cigar_errors (the reference reads right[-1]);
 *   - numeric tokens must be all digits; anything else parses as 0.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define S2P_BATCH (1u << 18)        /* pairutil.h:48 */
#define MIN_CLIP 20                 /* pairutil.h:54 */
#define MAX_SELF_CIRCLE 10u         /* pairutil.h:57 */
#define MAX_PAIR_DIST 1000          /* pairutil.h:58 */

typedef struct { const char *p; size_t n; } str_t;

typedef struct {
    str_t line, qname, chr, cigar;
    unsigned flag, pos, mapq;
} samrec;

enum { ST_SILENT = 0, ST_LOWMAP, ST_MANYHITS, ST_UNPAIRED, ST_SELFCIRCLE,
       ST_TRANS, ST_CIS10K, ST_CIS1K, ST_CIS0, ST_CIGARERR };

typedef struct { int status; str_t chrA, chrB; unsigned posA, posB; char sA, sB; } pairres;

static int is_ws(char c) { return c == ' ' || (c >= 9 && c <= 13); }

static str_t next_tok(const char **cur, const char *end) {
    const char *p = *cur;
    while (p < end && is_ws(*p)) ++p;
    str_t t; t.p = p;
    while (p < end && !is_ws(*p)) ++p;
    t.n = (size_t)(p - t.p);
    *cur = p;
    return t;
}

static unsigned tok_uint(str_t t) {
    unsigned v = 0;
    if (t.n == 0) return 0;
    for (size_t i = 0; i < t.n; ++i) {
        if (t.p[i] < '0' || t.p[i] > '9') return 0;
        v = v * 10u + (unsigned)(t.p[i] - '0');
    }
    return v;
}

static void parse_rec(str_t line, samrec *r) {
    const char *cur = line.p, *end = line.p + line.n;
    r->line = line;
    r->qname = next_tok(&cur, end);
    r->flag = tok_uint(next_tok(&cur, end));
    r->chr = next_tok(&cur, end);
    r->pos = tok_uint(next_tok(&cur, end));
    r->mapq = tok_uint(next_tok(&cur, end));
    r->cigar = next_tok(&cur, end);
}

/* pairutil.h:63-126 */
int orc_cigar2segment(const char *cigar, size_t len, int start, orc_segment *s) {
    memset(s, 0, sizeof(*s));
    int idx = 0, cur = start, val = 0;
    int last_right = 0;
    s->left[0] = start;
    for (size_t j = 0; j < len; ++j) {
        char c = cigar[j];
        if (c >= '0' && c <= '9') { val = val * 10 + (c - '0'); continue; }
        switch (c) {
        case 'H': case 'S':
            if (j == len - 1) s->rightClip = val;
            else if (idx == 0) s->leftClip = val;
            else return 0;
            break;
        case 'M':
            s->mappable += val; cur += val; last_right = cur - 1;
            if (idx < 4) s->right[idx] = last_right;
            break;
        case 'D':
            cur += val; last_right = cur - 1;
            if (idx < 4) s->right[idx] = last_right;
            break;
        case 'I':
            break;
        case 'N':
            cur += val; ++idx; last_right = 0;
            if (idx < 4) { s->left[idx] = cur; s->right[idx] = 0; }
            break;
        default:
            return 0;
        }
        val = 0;
    }
    if (last_right == 0) return 0;          /* pairutil.h:119 */
    s->segCnt = idx + 1;
    if (idx >= 4) s->right[3] = last_right; /* only the count matters beyond 2 segments */
    return 1;
}

static int seg_of(const samrec *r, orc_segment *s) {
    return orc_cigar2segment(r->cigar.p, r->cigar.n, (int)r->pos, s);
}

/* pairutil.h:180-188 — the comparison is carried out in fp32 */
static int integrity1(const orc_segment *s, float ratio) {
    int total = s->mappable;
    if (s->leftClip > MIN_CLIP) total += s->leftClip;
    if (s->rightClip > MIN_CLIP) total += s->rightClip;
    volatile float lim = (float)total * ratio;
    return (float)s->mappable >= lim;
}

/* pairutil.h:190-208 — including the s1.rightClip test at :200 */
static int integrity2(const orc_segment *a, const orc_segment *b, float ratio) {
    int ta = a->mappable, tb = b->mappable;
    if (a->leftClip > MIN_CLIP) ta += a->leftClip;
    if (a->rightClip > MIN_CLIP) ta += a->rightClip;
    if (b->leftClip > MIN_CLIP) tb += b->leftClip;
    if (a->rightClip > MIN_CLIP) tb += b->rightClip;
    int big = ta > tb ? ta : tb;
    volatile float lim = (float)big * ratio;
    return (float)(a->mappable + b->mappable) >= lim;
}

static int str_cmp(str_t a, str_t b) {
    size_t m = a.n < b.n ? a.n : b.n;
    int c = m ? memcmp(a.p, b.p, m) : 0;
    if (c) return c;
    return a.n < b.n ? -1 : (a.n > b.n ? 1 : 0);
}

/* unc2pairs.h:310-348 == flash2pairs.h:104-144 */
static void order_and_class(pairres *r, str_t c1, unsigned p1, char s1, str_t c2, unsigned p2, char s2) {
    int cc = str_cmp(c1, c2);
    if (cc < 0 || (cc == 0 && p1 < p2)) {
        r->chrA = c1; r->posA = p1; r->sA = s1; r->chrB = c2; r->posB = p2; r->sB = s2;
    } else {
        r->chrA = c2; r->posA = p2; r->sA = s2; r->chrB = c1; r->posB = p1; r->sB = s1;
    }
    if (cc == 0) {
        unsigned d = r->posB - r->posA;
        if (d <= MAX_SELF_CIRCLE) r->status = ST_SELFCIRCLE;
        else if (d >= 10000u) r->status = ST_CIS10K;
        else if (d >= 1000u) r->status = ST_CIS1K;
        else r->status = ST_CIS0;
    } else r->status = ST_TRANS;
}

static char strand_of(const samrec *r) { return (r->flag & 16u) ? '-' : '+'; }
static int same_chr(const samrec *a, const samrec *b) { return str_cmp(a->chr, b->chr) == 0; }

/* flash2pairs.h:25-154 */
static void resolve_flash(const samrec *g, size_t n, float ratio, pairres *r) {
    orc_segment a, b;
    if (n == 1) {
        if (!seg_of(&g[0], &a)) { r->status = ST_CIGARERR; return; }
        if (a.segCnt > 2) { r->status = ST_MANYHITS; return; }
        if (!integrity1(&a, ratio)) { r->status = ST_LOWMAP; return; }
        unsigned p1 = g[0].pos, p2 = (unsigned)a.right[a.segCnt - 1];
        unsigned d = p2 - p1;
        r->status = d >= 10000u ? ST_CIS10K : (d >= 1000u ? ST_CIS1K : ST_CIS0);
        r->chrA = r->chrB = g[0].chr; r->posA = p1; r->posB = p2; r->sA = '+'; r->sB = '-';
        return;
    }
    if (n == 2) {
        int ok = seg_of(&g[0], &a) & seg_of(&g[1], &b);
        if (!ok) { r->status = ST_CIGARERR; return; }
        if (a.segCnt != 1 || b.segCnt != 1) { r->status = ST_MANYHITS; return; }
        if (!integrity2(&a, &b, ratio)) { r->status = ST_LOWMAP; return; }
        unsigned p1 = a.leftClip > a.rightClip ? (unsigned)a.right[0] : g[0].pos;
        unsigned p2 = b.leftClip > b.rightClip ? (unsigned)b.right[0] : g[1].pos;
        order_and_class(r, g[0].chr, p1, strand_of(&g[0]), g[1].chr, p2, strand_of(&g[1]));
        return;
    }
    r->status = ST_MANYHITS;
}

/* the mate test shared by unc2pairs.h:191-308: does `lone` pair with candidate `c`? */
static int mates(const samrec *lone, const orc_segment *ls, const samrec *c, const orc_segment *cs) {
    if (strand_of(lone) == '+')
        return strand_of(c) == '-' && same_chr(lone, c) && ls->left[0] < cs->left[0] &&
               cs->right[0] - ls->left[0] <= MAX_PAIR_DIST;
    return strand_of(c) == '+' && same_chr(lone, c) && cs->left[0] < ls->left[0] &&
           ls->right[0] - cs->left[0] <= MAX_PAIR_DIST;
}

static unsigned distal_end(const orc_segment *s) {   /* unc2pairs.h:237-248 */
    return s->leftClip > s->rightClip ? (unsigned)s->right[0] : (unsigned)s->left[0];
}

/* unc2pairs.h:29-358 */
static void resolve_unc(const samrec *g, size_t n, float ratio, pairres *r) {
    const samrec *R1[4], *R2[4];
    size_t n1 = 0, n2 = 0;
    for (size_t i = 0; i < n; ++i) {
        if (g[i].flag & 64u) { if (n1 < 4) R1[n1] = &g[i]; ++n1; }
        else if (g[i].flag & 128u) { if (n2 < 4) R2[n2] = &g[i]; ++n2; }
    }
    r->status = ST_SILENT;
    if (n1 == 0 || n2 == 0) return;
    if (n1 + n2 > 3) return;

    orc_segment s1, s2, s3;
    str_t c1, c2; unsigned p1, p2; char t1, t2;
    if (n1 == 1 && n2 == 1) {
        if (!seg_of(R1[0], &s1)) { r->status = ST_CIGARERR; return; }
        if (!integrity1(&s1, ratio)) { r->status = ST_LOWMAP; return; }
        if (!seg_of(R2[0], &s2)) { r->status = ST_CIGARERR; return; }
        if (!integrity1(&s2, ratio)) { r->status = ST_LOWMAP; return; }
        if (s1.segCnt + s2.segCnt > 3) { r->status = ST_MANYHITS; return; }
        t1 = strand_of(R1[0]); t2 = strand_of(R2[0]); c1 = R1[0]->chr; c2 = R2[0]->chr;
        if (s1.segCnt == 1 && s2.segCnt == 1) {
            p1 = t1 == '+' ? (unsigned)s1.left[0] : (unsigned)s1.right[0];
            p2 = t2 == '+' ? (unsigned)s2.left[0] : (unsigned)s2.right[0];
        } else if (s1.segCnt == 2) {                     /* unc2pairs.h:146-167 */
            int same = same_chr(R1[0], R2[0]);
            if (t1 == '+') {
                if (t2 == '-' && same && s1.left[1] < s2.left[0] && s2.right[0] - s1.left[1] <= MAX_PAIR_DIST) {
                    p1 = (unsigned)s1.left[0]; p2 = (unsigned)s2.right[0];
                } else { r->status = ST_UNPAIRED; return; }
            } else {
                if (t2 == '+' && same && s2.left[0] < s1.left[0] && s1.right[0] - s2.left[0] <= MAX_PAIR_DIST) {
                    p1 = (unsigned)s1.right[1]; p2 = (unsigned)s2.left[0];
                } else { r->status = ST_UNPAIRED; return; }
            }
        } else {                                         /* unc2pairs.h:168-189 */
            int same = same_chr(R1[0], R2[0]);
            if (t1 == '+') {
                if (t2 == '-' && same && s1.left[0] < s2.left[0] && s2.right[0] - s1.left[0] <= MAX_PAIR_DIST) {
                    p1 = (unsigned)s1.left[0]; p2 = (unsigned)s2.right[1];
                } else { r->status = ST_UNPAIRED; return; }
            } else {
                if (t2 == '+' && same && s2.left[1] < s1.left[0] && s1.right[0] - s2.left[1] <= MAX_PAIR_DIST) {
                    p1 = (unsigned)s1.right[0]; p2 = (unsigned)s2.left[0];
                } else { r->status = ST_UNPAIRED; return; }
            }
        }
    } else if (n1 == 1) {                                /* 1+2, unc2pairs.h:84-98,191-249 */
        if (!seg_of(R1[0], &s1)) { r->status = ST_CIGARERR; return; }
        if (!integrity1(&s1, ratio)) { r->status = ST_LOWMAP; return; }
        if (!(seg_of(R2[0], &s2) & seg_of(R2[1], &s3))) { r->status = ST_CIGARERR; return; }
        if (!integrity2(&s2, &s3, ratio)) { r->status = ST_LOWMAP; return; }
        if (s1.segCnt != 1 || s2.segCnt != 1 || s3.segCnt != 1) { r->status = ST_MANYHITS; return; }
        t1 = strand_of(R1[0]); c1 = R1[0]->chr;
        p1 = t1 == '+' ? (unsigned)s1.left[0] : (unsigned)s1.right[0];
        const samrec *other; const orc_segment *os;
        if (mates(R1[0], &s1, R2[0], &s2)) { other = R2[1]; os = &s3; }
        else if (mates(R1[0], &s1, R2[1], &s3)) { other = R2[0]; os = &s2; }
        else { r->status = ST_UNPAIRED; return; }
        c2 = other->chr; t2 = strand_of(other); p2 = distal_end(os);
    } else {                                             /* 2+1, unc2pairs.h:99-121,250-308 */
        if (!(seg_of(R1[0], &s1) & seg_of(R1[1], &s2))) { r->status = ST_CIGARERR; return; }
        if (!integrity2(&s1, &s2, ratio)) { r->status = ST_LOWMAP; return; }
        if (!seg_of(R2[0], &s3)) { r->status = ST_CIGARERR; return; }
        if (!integrity1(&s3, ratio)) { r->status = ST_LOWMAP; return; }
        if (s1.segCnt != 1 || s2.segCnt != 1 || s3.segCnt != 1) { r->status = ST_MANYHITS; return; }
        t2 = strand_of(R2[0]); c2 = R2[0]->chr;
        p2 = t2 == '+' ? (unsigned)s3.left[0] : (unsigned)s3.right[0];
        const samrec *other; const orc_segment *os;
        if (mates(R2[0], &s3, R1[0], &s1)) { other = R1[1]; os = &s2; }
        else if (mates(R2[0], &s3, R1[1], &s2)) { other = R1[0]; os = &s1; }
        else { r->status = ST_UNPAIRED; return; }
        c1 = other->chr; t1 = strand_of(other); p1 = distal_end(os);
    }
    order_and_class(r, c1, p1, t1, c2, p2, t2);
}

typedef struct { char *p; size_t n, cap; } buf_t;
static void buf_put(buf_t *b, const char *s, size_t n) {
    if (b->n + n + 1 > b->cap) {
        size_t nc = b->cap ? b->cap * 2 : 1 << 16;
        while (nc < b->n + n + 1) nc *= 2;
        b->p = (char *)realloc(b->p, nc); b->cap = nc;
    }
    memcpy(b->p + b->n, s, n); b->n += n;
}
static void buf_uint(buf_t *b, unsigned v) {
    char t[16]; int n = snprintf(t, sizeof t, "%u", v); buf_put(b, t, (size_t)n);
}

int orc_sam2pairs(const char *sam, size_t n, int mode, float ratio, int min_mapq,
                  int emu_threads, int write_sam,
                  char **pairs_out, size_t *pairs_len, char **sam_out, size_t *sam_len,
                  orc_s2p_stats *st) {
    memset(st, 0, sizeof(*st));
    if (emu_threads < 2) return -1;             /* sam2pairs.cpp:36-39 */

    /* split into lines (getline semantics) */
    size_t nl = 0, cap = 1024;
    str_t *lines = (str_t *)malloc(cap * sizeof(str_t));
    for (size_t i = 0; i < n;) {
        const char *e = (const char *)memchr(sam + i, '\n', n - i);
        size_t len = e ? (size_t)(e - (sam + i)) : n - i;
        if (nl == cap) { cap *= 2; lines = (str_t *)realloc(lines, cap * sizeof(str_t)); }
        lines[nl].p = sam + i; lines[nl].n = len; ++nl;
        i += len + 1;
    }

    /* filter + group: kept records in order; group = maximal run of equal QNAME */
    samrec *recs = (samrec *)malloc((nl ? nl : 1) * sizeof(samrec));
    size_t nrec = 0;
    int seen_first = 0;
    for (size_t i = 0; i < nl; ++i) {
        if (!seen_first && lines[i].n && lines[i].p[0] == '@') continue;   /* sam2pairs.cpp:117-119 */
        samrec r; parse_rec(lines[i], &r);
        if (r.mapq < (unsigned)min_mapq) continue;                         /* pairutil.h:157 */
        if (r.flag & 0x700u) continue;                                     /* pairutil.h:160 */
        seen_first = 1;
        recs[nrec++] = r;
    }
    size_t *gstart = (size_t *)malloc((nrec + 2) * sizeof(size_t));
    size_t ng = 0;
    for (size_t i = 0; i < nrec; ++i)
        if (i == 0 || str_cmp(recs[i].qname, recs[i - 1].qname) != 0) gstart[ng++] = i;
    gstart[ng] = nrec;

    /* the stream's last group is never processed (pairutil.h:176 + sam2pairs.cpp:150-151) */
    size_t P = ng ? ng - 1 : 0;
    st->groups = P;
    size_t full = (P / S2P_BATCH) * S2P_BATCH, rem = P % S2P_BATCH;
    unsigned T = (unsigned)emu_threads;
    unsigned share_full = S2P_BATCH / (T - 1);          /* sam2pairs.cpp:172-173, tn = 0 */
    unsigned share_last = (unsigned)rem / T;            /* sam2pairs.cpp:150-151, tn = 0 */

    buf_t po = {0, 0, 0}, so = {0, 0, 0};
    for (size_t g = 0; g < P; ++g) {
        const samrec *gr = recs + gstart[g];
        size_t gn = gstart[g + 1] - gstart[g];
        pairres r; memset(&r, 0, sizeof r);
        if (mode == 0) resolve_flash(gr, gn, ratio, &r); else resolve_unc(gr, gn, ratio, &r);
        switch (r.status) {
        case ST_LOWMAP: ++st->lowMap; break;
        case ST_MANYHITS: ++st->manyHits; break;
        case ST_UNPAIRED: ++st->unpaired; break;
        case ST_CIGARERR: ++st->cigar_errors; break;
        case ST_SELFCIRCLE: {
            ++st->selfCircle_true;
            unsigned rr = (unsigned)(g % S2P_BATCH);     /* only thread 0's tally reaches the log (sam2pairs.cpp:202-210) */
            if (g < full ? rr < share_full : rr < share_last) ++st->selfCircle;
            break; }
        case ST_TRANS: ++st->trans; break;
        case ST_CIS10K: ++st->cis10K; break;
        case ST_CIS1K: ++st->cis1K; break;
        case ST_CIS0: ++st->cis0; break;
        default: break;
        }
        if (r.status >= ST_TRANS && r.status <= ST_CIS0) {
            str_t rid = gr[gn - 1].qname;
            buf_put(&po, rid.p, rid.n); buf_put(&po, "\t", 1);
            buf_put(&po, r.chrA.p, r.chrA.n); buf_put(&po, "\t", 1); buf_uint(&po, r.posA); buf_put(&po, "\t", 1);
            buf_put(&po, r.chrB.p, r.chrB.n); buf_put(&po, "\t", 1); buf_uint(&po, r.posB); buf_put(&po, "\t", 1);
            char tail[4] = { r.sA, '\t', r.sB, '\n' };
            buf_put(&po, tail, 4);
            if (write_sam)
                for (size_t i = 0; i < gn; ++i) { buf_put(&so, gr[i].line.p, gr[i].line.n); buf_put(&so, "\n", 1); }
        }
    }
    free(lines); free(recs); free(gstart);
    if (pairs_out) { *pairs_out = po.p; *pairs_len = po.n; } else free(po.p);
    if (sam_out) { *sam_out = so.p; *sam_len = so.n; } else free(so.p);
    return 0;
}

void orc_free(void *p) { free(p); }
