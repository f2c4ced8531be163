// cli_sam2pairs.cpp — drop-in `sam2pairs` executable over libmicrocket_b200.so.
// Same argv, stdout/stderr text, files and exit codes as the reference's main (src/sam2pairs/sam2pairs.cpp:23-229):
//   sam2pairs <in.sam> <mode=flash|unc> <out.prefix> [thread=4] [min_mapped_ratio=0.5] [min.mapQ=10] [sam=1|0]
// `thread` only selects which self-circle share is logged (the reference logs thread 0's, sam2pairs.cpp:202-210);
// the work itself runs on the GPU.  MICROCKET_DEVICE selects the CUDA device (default 0).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include "../../include/microcket_b200.h"
using namespace std;

static int fail(const char *what) { cerr << "Error: " << what << ": " << mk_last_error() << "\n"; return 20; }

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

    FILE *fin = fopen(argv[1], "rb");
    if (!fin) { cerr << "Error: read input file failed!\n"; return 10; }
    string base = string(argv[3]) + "." + argv[2];
    FILE *fsam = NULL;
    if (cfg.write_sam) {
        fsam = fopen((base + ".sam").c_str(), "wb");
        if (!fsam) { cerr << "Error: write sam file failed!\n"; fclose(fin); return 11; }
    }
    mk_ctx *ctx = NULL;
    if (mk_s2p_create(&cfg, NULL, 0, &ctx) != MK_OK) return fail("cannot create the GPU context");

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
        size_t n = fread(in.data(), 1, IN, fin);
        if (n == 0) break;
        if (mk_s2p_push(ctx, in.data(), n, 0) != MK_OK) return fail("sam2pairs");
        if (drain()) return fail("sam2pairs");
    }
    if (mk_s2p_push(ctx, NULL, 0, 1) != MK_OK) return fail("sam2pairs");
    if (drain()) return fail("sam2pairs");
    mk_s2p_stats st;
    if (mk_s2p_finish(ctx, &st) != MK_OK) return fail("sam2pairs");
    if (drain()) return fail("sam2pairs");
    fclose(fin);
    if (fsam) fclose(fsam);
    fflush(stdout);

    ofstream flog((base + "2pairs.log").c_str());
    if (flog.fail()) { cerr << "Error: write log file failed!\n"; return 10; }
    flog << "lowMap\t" << st.lowMap << "\nmanyHits\t" << st.manyHits << "\nunpaired\t" << st.unpaired << "\nselfCircle\t" << st.selfCircle
         << "\ntrans\t" << st.trans << "\ncis10K\t" << st.cis10K << "\ncis1K\t" << st.cis1K << "\ncis0\t" << st.cis0 << '\n';
    flog.close();
    mk_destroy(ctx);
    return 0;
}
