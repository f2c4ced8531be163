/*
 * oracle_cli.c — command-line front end for the CPU restatement, used by
 * tests/ and by bench.py's cpu_baseline leg when oracle/_ref is absent.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 *   oracle_cli sam2pairs <in.sam> <flash|unc> <prefix> [T=4] [ratio=0.5] [Q=10] [sam=1|0]
 *       pairs → stdout (input order), <prefix>.<mode>2pairs.log, <prefix>.<mode>.sam
 *       (argument order of sam2pairs.cpp:25-54)
 *   oracle_cli krmdup <in.fq> <prefix>
 *       appends <prefix>.read1.fq, <prefix>.read2.fq, <prefix>.log (krmdup.cpp:266-271,377-389)
 */
#include "oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static char *slurp(const char *path, size_t *n) {
    FILE *f = strcmp(path, "-") ? fopen(path, "rb") : stdin;
    if (!f) return NULL;
    size_t cap = 1 << 20, len = 0; char *b = (char *)malloc(cap);
    for (;;) {
        if (len == cap) { cap *= 2; b = (char *)realloc(b, cap); }
        size_t r = fread(b + len, 1, cap - len, f);
        if (!r) break;
        len += r;
    }
    if (f != stdin) fclose(f);
    *n = len; return b;
}

int main(int argc, char **argv) {
    if (argc >= 5 && !strcmp(argv[1], "sam2pairs")) {
        int mode = !strcmp(argv[3], "flash") ? 0 : (!strcmp(argv[3], "unc") ? 1 : -1);
        if (mode < 0) { fprintf(stderr, "Error: Unknown mode, must be 'flash' or 'unc'.\n"); return 6; }
        int T = argc > 5 ? atoi(argv[5]) : 4;
        if (T < 2) { fprintf(stderr, "Error: at least 2 threads are required.\n"); return 5; }
        float ratio = argc > 6 ? (float)atof(argv[6]) : 0.5f;
        int Q = argc > 7 ? atoi(argv[7]) : 10;
        int wsam = 1;
        if (argc > 8 && (argv[8][0] == 'N' || argv[8][0] == 'n' || argv[8][0] == '0')) wsam = 0;
        size_t n; char *in = slurp(argv[2], &n);
        if (!in) { fprintf(stderr, "Error: read input file failed!\n"); return 10; }
        char *po, *so; size_t pl, sl; orc_s2p_stats st;
        if (orc_sam2pairs(in, n, mode, ratio, Q, T, wsam, &po, &pl, &so, &sl, &st)) return 1;
        fwrite(po, 1, pl, stdout);
        char path[4096];
        if (wsam) {
            snprintf(path, sizeof path, "%s.%s.sam", argv[4], argv[3]);
            FILE *f = fopen(path, "wb"); if (!f) return 11;
            fwrite(so, 1, sl, f); fclose(f);
        }
        snprintf(path, sizeof path, "%s.%s2pairs.log", argv[4], argv[3]);
        FILE *f = fopen(path, "w"); if (!f) return 10;
        fprintf(f, "lowMap\t%u\nmanyHits\t%u\nunpaired\t%u\nselfCircle\t%u\ntrans\t%u\ncis10K\t%u\ncis1K\t%u\ncis0\t%u\n",
                st.lowMap, st.manyHits, st.unpaired, st.selfCircle, st.trans, st.cis10K, st.cis1K, st.cis0);
        fclose(f);
        return 0;
    }
    if (argc >= 4 && !strcmp(argv[1], "krmdup")) {
        size_t n; char *in = slurp(argv[2], &n);
        if (!in) { fprintf(stderr, "Error: read fastq failed!\n"); return 10; }
        orc_dedup *d = orc_dedup_new(5, 16, 5, 16);
        char *r1, *r2; size_t l1, l2; orc_dd_stats st;
        orc_krmdup(d, in, n, &r1, &l1, &r2, &l2, &st);
        char path[4096];
        snprintf(path, sizeof path, "%s.read1.fq", argv[3]); FILE *f = fopen(path, "a"); if (!f) return 1; fwrite(r1, 1, l1, f); fclose(f);
        snprintf(path, sizeof path, "%s.read2.fq", argv[3]); f = fopen(path, "a"); if (!f) return 1; fwrite(r2, 1, l2, f); fclose(f);
        snprintf(path, sizeof path, "%s.log", argv[3]); f = fopen(path, "a"); if (!f) return 10;
        fprintf(f, "Total\t%u\nUniq\t%u\nDup\t%u\nDiscard\t%u\n", st.uniq + st.dup + st.discard, st.uniq, st.dup, st.discard);
        fclose(f);
        orc_dedup_free(d);
        return 0;
    }
    fprintf(stderr, "usage: oracle_cli sam2pairs <in.sam> <flash|unc> <prefix> [T] [ratio] [Q] [sam]\n"
                    "       oracle_cli krmdup <in.fq> <prefix>\n");
    return 2;
}
