// cli_krmdup.cpp — drop-in `krmdup` and `krmdup.pipe` executables over libmicrocket_b200.so.
// Same options, files (append mode), log and exit codes as the reference (src/preprocess/krmdup.cpp:229-397,
// src/preprocess/krmdup.pipe.cpp).  Built twice: -DKRMDUP_PIPE writes interleaved FASTQ to stdout instead of
// <prefix>.read1.fq / <prefix>.read2.fq.
#include <getopt.h>
#include <time.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include "../../include/microcket_b200.h"
using namespace std;

static void usage(const char *prg) {
    cerr << "\nUsage: " << prg << " [options] -i <interleaved.paired-end.fq> -o <output.prefix>\n"
         << "\nOptions:\n"
         << "  -k <int>  Skip the heading cycles in read 1 (default: 5)\n"
         << "  -K <int>  Skip the heading cycles in read 2 (default: 5)\n"
         << "  -s <int>  Size of the KEY in read 1 (default: 16)\n"
         << "  -S <int>  Size of the KEY in read 2 (default: 16)\n"
         << "\nThis program is designed to remove the duplicate reads from the FASTQ data on the GPU."
         << "\n\nIMPORTANT NOTEs:"
         << "\nThe total KEY size in read1 and read2 must >=16 and <=32."
         << "\nWhen running this program on the adapter-and-quality trimmed data, please mind the read length,"
         << "\nreads that are shorter than Skip1+Key1 or Skip2+Key2 will be discarded."
#ifdef KRMDUP_PIPE
         << "\n\nLog file will be written, while reads will be output to STDOUT in interleaved-fastq format.\n\n";
#else
         << "\n\nLog and FASTQ files will be written.\n\n";
#endif
    exit(2);
}

#ifdef KRMDUP_PIPE
// the two record streams come out in the same pair order: interleave them 4 lines at a time
struct Interleaver {
    string a, b;
    void feed(const char *r1, size_t n1, const char *r2, size_t n2) {
        a.append(r1, n1); b.append(r2, n2);
        size_t pa = 0, pb = 0;
        while (true) {
            size_t ea = pa, eb = pb; int k;
            for (k = 0; k < 4; ++k) { size_t p = a.find('\n', ea); if (p == string::npos) break; ea = p + 1; }
            if (k < 4) break;
            for (k = 0; k < 4; ++k) { size_t p = b.find('\n', eb); if (p == string::npos) break; eb = p + 1; }
            if (k < 4) break;
            fwrite(a.data() + pa, 1, ea - pa, stdout); fwrite(b.data() + pb, 1, eb - pb, stdout);
            pa = ea; pb = eb;
        }
        a.erase(0, pa); b.erase(0, pb);
    }
};
#endif

int main(int argc, char *argv[]) {
    mk_dedup_cfg cfg; mk_dedup_default_cfg(&cfg);
    const char *readx = NULL, *prefix = NULL;
    int opt;
    while ((opt = getopt(argc, argv, "i:o:k:K:s:S:")) != -1) {
        switch (opt) {
        case 'i': readx = optarg; break;
        case 'o': prefix = optarg; break;
        case 'k': cfg.hskip1 = atoi(optarg); break;
        case 'K': cfg.hskip2 = atoi(optarg); break;
        case 's': cfg.klen1 = atoi(optarg); break;
        case 'S': cfg.klen2 = atoi(optarg); break;
        default: usage(argv[0]);
        }
    }
    if (!readx || !prefix) usage(argv[0]);
    if (cfg.klen1 + cfg.klen2 > 32 || cfg.klen1 + cfg.klen2 < 16) { cerr << "Error: invalid key sizes!\n"; exit(1); }
    if (const char *d = getenv("MICROCKET_DEVICE")) cfg.device = atoi(d);
    if (const char *w = getenv("MICROCKET_WINDOW_MB")) cfg.window_bytes = (size_t)atol(w) << 20;
#ifndef KRMDUP_PIPE
    FILE *f1 = fopen((string(prefix) + ".read1.fq").c_str(), "a"), *f2 = fopen((string(prefix) + ".read2.fq").c_str(), "a");
    if (!f1 || !f2) { cerr << "Error: open output files failed!\n"; exit(1); }
#else
    Interleaver il;
#endif
    FILE *fin = (readx[0] == '-' && readx[1] == '\0') ? fopen("/dev/stdin", "rb") : fopen(readx, "rb");
    if (!fin) { cerr << "Error: read fastq failed!\n"; return 10; }
    mk_ctx *ctx = NULL;
    const bool trace = getenv("MICROCKET_TRACE") != NULL;
    struct timespec ts0; clock_gettime(CLOCK_MONOTONIC, &ts0);
    auto since = [&]() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (t.tv_sec - ts0.tv_sec) + 1e-9 * (t.tv_nsec - ts0.tv_nsec); };
    if (mk_dedup_create(&cfg, &ctx) != MK_OK) { cerr << "Error: " << mk_last_error() << "\n"; return 20; }
    if (trace) fprintf(stderr, "[krmdup] context ready after %.3f s\n", since());
    const size_t IN = 64u << 20, OUT = 32u << 20;
    vector<char> in(IN), o1(OUT), o2(OUT);
    auto drain = [&]() -> int {
        while (true) {
            size_t a = 0, b = 0;
            if (mk_dedup_pull(ctx, o1.data(), OUT, &a, o2.data(), OUT, &b) != MK_OK) return -1;
            if (!a && !b) return 0;
#ifdef KRMDUP_PIPE
            il.feed(o1.data(), a, o2.data(), b);
#else
            if (a) fwrite(o1.data(), 1, a, f1);
            if (b) fwrite(o2.data(), 1, b, f2);
#endif
        }
    };
    while (true) {
        size_t n = fread(in.data(), 1, IN, fin);
        if (n == 0) break;
        if (mk_dedup_push(ctx, in.data(), n, 0) != MK_OK || drain()) { cerr << "Error: " << mk_last_error() << "\n"; return 20; }
    }
    mk_dedup_stats st;
    if (mk_dedup_push(ctx, NULL, 0, 1) != MK_OK || drain() || mk_dedup_finish(ctx, &st) != MK_OK) { cerr << "Error: " << mk_last_error() << "\n"; return 20; }
    fclose(fin);
#ifndef KRMDUP_PIPE
    fclose(f1); fclose(f2);
#else
    fflush(stdout);
#endif
    if (trace) fprintf(stderr, "[krmdup] streaming done after %.3f s\n", since());
    ofstream flog((string(prefix) + ".log").c_str(), ios::app);
    if (flog.fail()) { cerr << "Error: write log failed!\n"; return 10; }
    flog << "Total\t" << st.uniq + st.dup + st.discard << "\nUniq\t" << st.uniq << "\nDup\t" << st.dup << "\nDiscard\t" << st.discard << '\n';
    flog.close();
    if (trace) fprintf(stderr, "[krmdup] log written after %.3f s\n", since());
    mk_destroy(ctx);
    if (trace) fprintf(stderr, "[krmdup] context destroyed after %.3f s\n", since());
    return 0;
}
