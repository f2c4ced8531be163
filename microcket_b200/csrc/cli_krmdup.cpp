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
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "../../include/microcket_b200.h"
using namespace std;

// Three-stage host pipeline: a reader thread fills input blocks, the main thread pushes them through the GPU context and
// pulls kept records into output blocks, a writer thread appends those to the files (or stdout).  Reading, the GPU round
// trip and writing overlap instead of adding up (the reference overlaps them with its own loader / worker / writer threads,
// krmdup.cpp:300-372).
template <class T> struct Chan {                      // blocking queue
    deque<T> q; mutex m; condition_variable cv;
    void put(T v) { { lock_guard<mutex> l(m); q.push_back(std::move(v)); } cv.notify_one(); }
    T get() { unique_lock<mutex> l(m); cv.wait(l, [&] { return !q.empty(); }); T v = std::move(q.front()); q.pop_front(); return v; }
};
struct InBlock { vector<char> buf; size_t n = 0; };
struct OutBlock { char *a = NULL, *b = NULL; size_t na = 0, nb = 0; bool last = false; };   // pinned (mk_host_alloc): kept records are DMA'd straight into them

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
    cfg.window_bytes = (size_t)64 << 20;          // small windows: less pinned memory to set up, and the first output arrives early
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
    const size_t IN = 32u << 20, OUT = 32u << 20;
    Chan<InBlock *> in_free, in_full; Chan<OutBlock *> out_free, out_full;
    vector<InBlock> in_pool(12); vector<OutBlock> out_pool(3);      // 384 MiB of read-ahead: the reader runs while the CUDA context comes up
    for (auto &b : in_pool) { b.buf.resize(IN); in_free.put(&b); }

    thread reader([&] {
        while (true) {
            InBlock *b = in_free.get();
            b->n = fread(b->buf.data(), 1, IN, fin);
            in_full.put(b);
            if (b->n == 0) break;
        }
    });
    auto fail_early = [&]() {                                        // (the reader is already filling blocks: let it run to EOF)
        cerr << "Error: " << mk_last_error() << "\n";
        while (true) { InBlock *b = in_full.get(); if (b->n == 0) break; in_free.put(b); }
        reader.join();
        return 20;
    };
    if (mk_dedup_create(&cfg, &ctx) != MK_OK) return fail_early();
    for (auto &b : out_pool) {                                       // (pinned memory needs the CUDA context)
        if (mk_host_alloc(OUT, (void **)&b.a) != MK_OK || mk_host_alloc(OUT, (void **)&b.b) != MK_OK) return fail_early();
        out_free.put(&b);
    }
    if (trace) fprintf(stderr, "[krmdup] context ready after %.3f s\n", since());
    thread writer([&] {
        while (true) {
            OutBlock *b = out_full.get();
            if (b->last) break;
#ifdef KRMDUP_PIPE
            il.feed(b->a, b->na, b->b, b->nb);
#else
            if (b->na) fwrite(b->a, 1, b->na, f1);
            if (b->nb) fwrite(b->b, 1, b->nb, f2);
#endif
            out_free.put(b);
        }
    });
    auto drain = [&]() -> int {
        while (true) {
            OutBlock *b = out_free.get();
            b->na = b->nb = 0; b->last = false;
            if (mk_dedup_pull(ctx, b->a, OUT, &b->na, b->b, OUT, &b->nb) != MK_OK) { out_free.put(b); return -1; }
            if (!b->na && !b->nb) { out_free.put(b); return 0; }
            out_full.put(b);
        }
    };
    auto stop_threads = [&]() { OutBlock *b = out_free.get(); b->last = true; out_full.put(b); writer.join(); reader.join(); };
    int rc = 0;
    while (true) {
        InBlock *b = in_full.get();
        if (b->n == 0) break;
        if (!rc && (mk_dedup_push(ctx, b->buf.data(), b->n, 0) != MK_OK || drain())) rc = 20;
        in_free.put(b);                                  // keep the reader going to EOF even after an error
    }
    if (rc) { stop_threads(); cerr << "Error: " << mk_last_error() << "\n"; return 20; }
    mk_dedup_stats st;
    if (mk_dedup_push(ctx, NULL, 0, 1) != MK_OK || drain() || mk_dedup_finish(ctx, &st) != MK_OK) { stop_threads(); cerr << "Error: " << mk_last_error() << "\n"; return 20; }
    stop_threads();
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
