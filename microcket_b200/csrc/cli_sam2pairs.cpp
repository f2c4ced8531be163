// cli_sam2pairs.cpp — drop-in `sam2pairs` executable over libmicrocket_b200.so.
// Same argv, stdout/stderr text, files and exit codes as the reference's main (src/sam2pairs/sam2pairs.cpp:23-229):
//   sam2pairs <in.sam> <mode=flash|unc> <out.prefix> [thread=4] [min_mapped_ratio=0.5] [min.mapQ=10] [sam=1|0]
// `thread` only selects which self-circle share is logged (the reference logs thread 0's, sam2pairs.cpp:202-210);
// the work itself runs on the GPU.  MICROCKET_DEVICE selects the CUDA device (default 0).
// Extension (argv[8], or MICROCKET_OUTPUT): `sorted` writes the pair lines to stdout already in the order of the driver's
// `LANG=C sort -k2,2d -k4,4d -k3,3n -k5,5n` (microcket:480,484,502,506), so that sort becomes a pass-through;
// `sorted-dedup` also removes coordinate duplicates first (first occurrence wins).  Pairs, their text and the line offsets
// then stay in HBM until the end of the input (mk_s2p_run_device + mk_pairs_sort_text_device).
// Extension (MICROCKET_RMDUP=1 or =k,s,K,S): krmdup's duplicate removal (src/preprocess/krmdup.cpp) is taken on the SAM,
// before grouping, so a driver may skip its FASTQ krmdup step; krmdup's four log lines are appended to <out.prefix>.rmdup.log
// (the file `krmdup -o $sid.rmdup` writes, microcket:413,445).  MICROCKET_RMDUP_PAIRS sizes the key table (read pairs).
// Extension: <in.sam> may be a BAM file or stream; it is decoded on host threads to the text `samtools view` would pipe in
// (bam_input.hpp; microcket:478,500 `samtools view -@ $vthread x.bam | sam2pairs /dev/stdin ...`).  MICROCKET_BAM_THREADS.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>
#include "../../include/microcket_b200.h"
#include "bam_input.hpp"
using namespace std;

static int fail(const char *what) { cerr << "Error: " << what << ": " << mk_last_error() << "\n"; return 20; }
#define CK(call) do { if ((call) != MK_OK) return fail("sam2pairs"); } while (0)

// device array that grows by reallocation (contents kept)
struct Grow {
    void *p = NULL; size_t cap = 0; int dev = 0;
    int need(size_t bytes, size_t used) {
        if (bytes <= cap) return MK_OK;
        size_t n = cap ? cap : ((size_t)64 << 20); while (n < bytes) n *= 2;
        void *q = NULL;
        if (mk_dev_alloc(dev, n, &q) != MK_OK) return MK_ERR_NOMEM;
        if (used && p && mk_copy_device(q, p, used) != MK_OK) return MK_ERR_CUDA;
        mk_dev_free(p); p = q; cap = n;
        return MK_OK;
    }
    ~Grow() { mk_dev_free(p); }
};

// The whole input through the device-resident API in chunks; text, packed pairs and line offsets accumulate in HBM, then one
// sort (and optionally one dedup) and one pass over the lines.
static int run_sorted(mk_ctx *ctx, const mk_s2p_cfg &cfg, mkbam::SamSource &src, FILE *fsam, bool dedup, mk_s2p_stats *st) {
    const size_t CHUNK = (size_t)(getenv("MICROCKET_CHUNK_MB") ? atol(getenv("MICROCKET_CHUNK_MB")) : 1024) << 20;
    const int dev = cfg.device;
    char *h_in = NULL; void *d_in = NULL, *d_sam = NULL; char *h_sam = NULL;
    CK(mk_host_alloc(CHUNK + 64, (void **)&h_in));
    CK(mk_dev_alloc(dev, CHUNK + 64, &d_in));
    if (fsam) { CK(mk_dev_alloc(dev, CHUNK + 64, &d_sam)); CK(mk_host_alloc(CHUNK + 64, (void **)&h_sam)); }
    Grow text, pairs, off; text.dev = pairs.dev = off.dev = dev;
    size_t text_total = 0, n_total = 0, have = 0;
    while (true) {
        const size_t got = src.read(h_in + have, CHUNK - have);
        const bool eof = got == 0;
        size_t tot = have + got;
        if (tot == 0) break;
        size_t cut = tot;
        if (!eof) {
            while (cut > 0 && h_in[cut - 1] != '\n') --cut;
            if (cut == 0) { if (tot == CHUNK) { cerr << "Error: a line longer than the chunk (" << CHUNK << " bytes)\n"; return 20; } have = tot; continue; }
        } else if (h_in[tot - 1] != '\n') { h_in[tot++] = '\n'; cut = tot; }   // getline accepts a last line without '\n'
        const size_t max_pairs = cut / 32 + 1024;
        CK(text.need(text_total + cut + 4096, text_total));
        CK(pairs.need((n_total + max_pairs) * sizeof(mk_pair), n_total * sizeof(mk_pair)));
        CK(off.need((n_total + max_pairs + 1) * 8, (n_total + 1) * 8));
        CK(mk_copy_to_device(d_in, h_in, cut));
        mk_s2p_dev_io io; memset(&io, 0, sizeof io);
        io.d_pairs_text = (char *)text.p + text_total; io.pairs_text_cap = text.cap - text_total;
        io.d_pairs = (mk_pair *)pairs.p + n_total; io.pairs_cap = max_pairs;
        io.d_sam_text = (char *)d_sam; io.sam_text_cap = d_sam ? CHUNK + 64 : 0;
        io.d_line_off = (uint64_t *)off.p + n_total; io.line_off_cap = max_pairs + 1; io.line_off_base = text_total;
        CK(mk_s2p_run_device(ctx, (const char *)d_in, cut, eof ? 1 : 0, &io, NULL));
        text_total += io.pairs_text_len; n_total += io.n_pairs;
        if (fsam && io.sam_text_len) { CK(mk_copy_to_host(h_sam, d_sam, io.sam_text_len)); fwrite(h_sam, 1, io.sam_text_len, fsam); }
        if (eof) break;
        // the unprocessed trailing read group and the partial last line go in front of the next chunk
        const size_t keep_from = io.consumed;
        if (keep_from == 0 && cut == tot && tot == CHUNK) { cerr << "Error: a read group larger than the chunk\n"; return 20; }
        have = tot - keep_from;
        memmove(h_in, h_in + keep_from, have);
    }
    CK(mk_s2p_finish(ctx, st));
    mk_host_free(h_in); mk_dev_free(d_in); mk_dev_free(d_sam); mk_host_free(h_sam);
    if (n_total == 0) return 0;
    // chromosome ranks under sort -d, in id order
    const int n_ids = mk_s2p_chrom_count(ctx);
    vector<string> names(n_ids); vector<const char *> cn(n_ids); vector<uint16_t> rank(n_ids);
    for (int i = 0; i < n_ids; ++i) { char b[64]; CK(mk_s2p_chrom_name(ctx, i, b, sizeof b)); names[i] = b; cn[i] = names[i].c_str(); }
    CK(mk_pairs_chrom_ranks(cn.data(), n_ids, rank.data()));
    mk_pairs_ws *ws = NULL;
    CK(mk_pairs_ws_create(dev, n_total, &ws));
    void *d_keep = NULL, *d_out = NULL;
    if (dedup) {
        // the keep mask comes from the one-sort dedup on a scratch copy (it reorders its input); no genome table is needed for
        // duplicate removal alone: every chromosome gets the full 32-bit range and one coarse "resolution"
        void *work = NULL, *b1 = NULL, *b2 = NULL, *bc = NULL;
        CK(mk_dev_alloc(dev, n_total * sizeof(mk_pair), &work)); CK(mk_copy_device(work, pairs.p, n_total * sizeof(mk_pair)));
        CK(mk_dev_alloc(dev, n_total, &d_keep));
        CK(mk_dev_alloc(dev, n_total * 4, &b1)); CK(mk_dev_alloc(dev, n_total * 4, &b2)); CK(mk_dev_alloc(dev, n_total * 4, &bc));
        vector<uint32_t> len(n_ids, 0xFFFFFFFFu);
        size_t kept = 0, nnz = 0;
        CK(mk_pairs_dedup_bin_indexed_device(ws, (mk_pair *)work, n_total, len.data(), n_ids, NULL, 0, 1u << 30, cfg.lane,
                                             (uint32_t *)b1, (uint32_t *)b2, (uint32_t *)bc, n_total, (uint8_t *)d_keep, NULL, &kept, &nnz, NULL));
        mk_dev_free(work); mk_dev_free(b1); mk_dev_free(b2); mk_dev_free(bc);
    }
    CK(mk_dev_alloc(dev, text_total + 64, &d_out));
    size_t out_len = 0, n_lines = 0;
    CK(mk_pairs_sort_text_device(ws, (const mk_pair *)pairs.p, n_total, (const uint8_t *)d_keep, (const char *)text.p, (const uint64_t *)off.p,
                                 rank.data(), n_ids, 0, (char *)d_out, text_total + 64, &out_len, &n_lines, NULL));
    const size_t PIECE = (size_t)256 << 20;
    char *h_out = NULL;
    CK(mk_host_alloc(PIECE, (void **)&h_out));
    for (size_t o = 0; o < out_len; o += PIECE) {
        const size_t m = out_len - o < PIECE ? out_len - o : PIECE;
        CK(mk_copy_to_host(h_out, (const char *)d_out + o, m));
        fwrite(h_out, 1, m, stdout);
    }
    mk_host_free(h_out); mk_dev_free(d_out); mk_dev_free(d_keep); mk_pairs_ws_destroy(ws);
    return 0;
}

int main(int argc, char *argv[]) {
    if (argc < 4) {
        cerr << "\nUsage: " << argv[0] << " <in.sam> <mode=flash|unc> <out.prefix> [thread=4] [min_mapped_ratio=0.5] [min.mapQ=10] [sam=1|0]"
             << "\n\nTask: extract the pairs from the alignment result."
             << "\n2 files will be written: out.mode.stat and out.mode.sam."
             << "\nThe pairs (without header) will be output to stdout (to pipe with sort utility)."
             << "\n\nThis program is part of Microcket, and is NOT supposed to be called manually by the user.\n\n";
        exit(2);
    }
    mk_s2p_cfg cfg; mk_s2p_default_cfg(&cfg);
    cfg.emu_threads = 4; cfg.write_sam = 1; cfg.emit_text = 1;
    if (argc > 4) {
        cfg.emu_threads = atoi(argv[4]);
        if (cfg.emu_threads < 2) { cerr << "Error: at least 2 threads are required.\n"; return 5; }
        if (argc > 5) {
            cfg.min_mapped_ratio = atof(argv[5]);
            cerr << "INFO: min_mapped_ratio is set to " << cfg.min_mapped_ratio << ".\n";
            if (argc > 6) {
                cfg.min_mapq = atoi(argv[6]);
                cerr << "INFO: min_mapQ is set to " << cfg.min_mapq << ".\n";
                if (argc > 7 && (argv[7][0] == 'N' || argv[7][0] == 'n' || argv[7][0] == '0')) {
                    cfg.write_sam = 0;
                    cerr << "WARN: sam output is skipped.\n";
                }
            }
        }
    }
    string mode = argv[2];
    if (mode == "flash") cfg.mode = 0; else if (mode == "unc") cfg.mode = 1;
    else { cerr << "Error: Unknown mode, must be 'flash' or 'unc'.\n"; return 6; }
    if (const char *d = getenv("MICROCKET_DEVICE")) cfg.device = atoi(d);
    if (const char *w = getenv("MICROCKET_WINDOW_MB")) cfg.window_bytes = (size_t)atol(w) << 20;
    string outmode = argc > 8 ? argv[8] : (getenv("MICROCKET_OUTPUT") ? getenv("MICROCKET_OUTPUT") : "");
    if (!outmode.empty() && outmode != "sorted" && outmode != "sorted-dedup") { cerr << "Error: Unknown output mode, must be 'sorted' or 'sorted-dedup'.\n"; return 6; }
    if (!outmode.empty()) { cfg.emit_packed = 1; if (!cfg.window_bytes) cfg.window_bytes = (size_t)1020 << 20; }

    if (const char *r = getenv("MICROCKET_RMDUP")) {
        if (r[0] && r[0] != '0') {
            cfg.rmdup = 1;
            int a, b, c, d;
            if (sscanf(r, "%d,%d,%d,%d", &a, &b, &c, &d) == 4) { cfg.hskip1 = a; cfg.klen1 = b; cfg.hskip2 = c; cfg.klen2 = d; }
            if (cfg.klen1 + cfg.klen2 < 16 || cfg.klen1 + cfg.klen2 > 32) { cerr << "ERROR: key size must be larger than 16 and smaller than 32!\n"; return 1; }   // krmdup.cpp:256-262
            cfg.rmdup_capacity = getenv("MICROCKET_RMDUP_PAIRS") ? strtoull(getenv("MICROCKET_RMDUP_PAIRS"), NULL, 10) : (uint64_t)1 << 27;
        }
    }
    FILE *fin = fopen(argv[1], "rb");
    if (!fin) { cerr << "Error: read input file failed!\n"; return 10; }
    long in_size = 0;
    if (cfg.rmdup && !getenv("MICROCKET_RMDUP_PAIRS")) {                  // a regular file: no more read pairs than bytes / 128
        if (fseek(fin, 0, SEEK_END) == 0) { in_size = ftell(fin); if (in_size > 0) cfg.rmdup_capacity = (uint64_t)in_size / 128 + 1024; }
        fseek(fin, 0, SEEK_SET);
    }
    // SAM text as it is; BAM decoded to text.  Looking at the first bytes blocks on a pipe, so it waits until the GPU context
    // exists unless the key table has to be sized from a regular file first
    std::unique_ptr<mkbam::SamSource> srcp;
    if (in_size > 0) {
        srcp.reset(new mkbam::SamSource(fin));
        if (srcp->is_bam()) cfg.rmdup_capacity = (uint64_t)in_size / 16 + 1024;   // compressed records: far fewer bytes per read pair
    }
    string base = string(argv[3]) + "." + argv[2];
    FILE *fsam = NULL;
    if (cfg.write_sam) {
        fsam = fopen((base + ".sam").c_str(), "wb");
        if (!fsam) { cerr << "Error: write sam file failed!\n"; fclose(fin); return 11; }
    }
    mk_ctx *ctx = NULL;
    if (mk_s2p_create(&cfg, NULL, 0, &ctx) != MK_OK) return fail("cannot create the GPU context");

    if (!srcp) srcp.reset(new mkbam::SamSource(fin));
    mkbam::SamSource &src = *srcp;
    mk_s2p_stats st;
    if (!outmode.empty()) {
        if (int rc = run_sorted(ctx, cfg, src, fsam, outmode == "sorted-dedup", &st)) return rc;
    } else {
    const size_t IN = 64u << 20, OUT = 32u << 20;
    vector<char> in(IN), out(OUT), samo(OUT);
    auto drain = [&]() -> int {
        while (true) {
            size_t a = 0, b = 0;
            if (mk_s2p_pull(ctx, out.data(), OUT, &a, fsam ? samo.data() : NULL, OUT, &b) != MK_OK) return -1;
            if (!a && !b) return 0;
            if (a) fwrite(out.data(), 1, a, stdout);
            if (b && fsam) fwrite(samo.data(), 1, b, fsam);
        }
    };
    while (true) {
        size_t n = src.read(in.data(), IN);
        if (n == 0) break;
        if (mk_s2p_push(ctx, in.data(), n, 0) != MK_OK) return fail("sam2pairs");
        if (drain()) return fail("sam2pairs");
    }
    if (mk_s2p_push(ctx, NULL, 0, 1) != MK_OK) return fail("sam2pairs");
    if (drain()) return fail("sam2pairs");
    if (mk_s2p_finish(ctx, &st) != MK_OK) return fail("sam2pairs");
    if (drain()) return fail("sam2pairs");
    }
    if (src.failed()) { cerr << "Error: " << src.error() << "\n"; return 10; }
    fclose(fin);
    if (fsam) fclose(fsam);
    fflush(stdout);

    ofstream flog((base + "2pairs.log").c_str());
    if (flog.fail()) { cerr << "Error: write log file failed!\n"; return 10; }
    flog << "lowMap\t" << st.lowMap << "\nmanyHits\t" << st.manyHits << "\nunpaired\t" << st.unpaired << "\nselfCircle\t" << st.selfCircle
         << "\ntrans\t" << st.trans << "\ncis10K\t" << st.cis10K << "\ncis1K\t" << st.cis1K << "\ncis0\t" << st.cis0 << '\n';
    flog.close();
    if (cfg.rmdup) {
        mk_dedup_stats dd;
        if (mk_s2p_rmdup_stats(ctx, &dd) != MK_OK) return fail("sam2pairs");
        ofstream fr((string(argv[3]) + ".rmdup.log").c_str(), ios::app);                          // krmdup.cpp:375-390 appends
        if (fr.fail()) { cerr << "Error: write log file failed!\n"; return 10; }
        fr << "Total\t" << (dd.uniq + dd.dup + dd.discard) << "\nUniq\t" << dd.uniq << "\nDup\t" << dd.dup << "\nDiscard\t" << dd.discard << '\n';
    }
    mk_destroy(ctx);
    return 0;
}
