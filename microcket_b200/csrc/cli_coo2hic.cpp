// cli_coo2hic.cpp — packs the per-resolution COO files of `pairs2bins` into one `.hic` (version 8) container (SURVEY.md
// §8(f)-3; the file the driver gets from `java -jar juicer_tools.jar pre -r <res,...> <sid>.final.pairs <sid>.hic <genome>.info`,
// microcket:525-529).  Host-only tool (no CUDA call: the counting was done by pairs2bins on the GPU); `pairs2bins -H <out.hic>`
// runs the same writer in-process on the arrays it already holds.
//   coo2hic [-g <genomeId>] -r <res[,res...]> <coo.prefix> <out.hic> <genome.info>
// reads <coo.prefix>.<res>.coo (`bin1<TAB>bin2<TAB>count`, upper triangle, sorted; bin = offset[chr] + pos / res in .info order).
// PARITY UNPINNED (hic_writer.hpp): written from the published format description, read back by tests/hic_reader.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include "hic_writer.hpp"
using namespace std;

static int read_coo(const string &path, vector<uint32_t> &b1, vector<uint32_t> &b2, vector<uint32_t> &ct) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { cerr << "Error: cannot read " << path << "\n"; return 10; }
    static char buf[1 << 22];
    uint64_t v[3] = {0, 0, 0}; int k = 0; bool digits = false; size_t got, line = 1;
    while ((got = fread(buf, 1, sizeof buf, f)) > 0)
        for (size_t i = 0; i < got; ++i) {
            const char c = buf[i];
            if (c >= '0' && c <= '9') { v[k] = v[k] > 0xFFFFFFFFull ? v[k] : v[k] * 10 + (uint64_t)(c - '0'); digits = true; }
            else if (c == '\t' && k < 2 && digits) { ++k; digits = false; }
            else if (c == '\n') {
                if (k == 2 && digits && (v[0] | v[1] | v[2]) > 0xFFFFFFFFull) { cerr << "Error: " << path << ": line " << line << " holds a number past 32 bits\n"; fclose(f); return 10; }
                if (k == 2 && digits) { b1.push_back((uint32_t)v[0]); b2.push_back((uint32_t)v[1]); ct.push_back((uint32_t)v[2]); }
                else if (k || digits) { cerr << "Error: " << path << ": line " << line << " is not bin1<TAB>bin2<TAB>count\n"; fclose(f); return 10; }
                v[0] = v[1] = v[2] = 0; k = 0; digits = false; ++line;
            } else { cerr << "Error: " << path << ": line " << line << " is not bin1<TAB>bin2<TAB>count\n"; fclose(f); return 10; }
        }
    fclose(f);
    return 0;
}

int main(int argc, char *argv[]) {
    string genome, reslist; int a = 1;
    while (a < argc && argv[a][0] == '-' && argv[a][1]) {
        if (!strcmp(argv[a], "-g") && a + 1 < argc) { genome = argv[a + 1]; a += 2; }
        else if (!strcmp(argv[a], "-r") && a + 1 < argc) { reslist = argv[a + 1]; a += 2; }
        else break;
    }
    if (argc - a != 3 || reslist.empty()) {
        cerr << "\nUsage: " << argv[0] << " [-g <genomeId>] -r <res[,res...]> <coo.prefix> <out.hic> <genome.info>\n\n";
        return 2;
    }
    vector<uint32_t> res;
    { stringstream ss(reslist); string t; while (getline(ss, t, ',')) if (!t.empty()) res.push_back((uint32_t)strtoul(t.c_str(), NULL, 10)); }
    for (uint32_t r : res) if (r == 0) { cerr << "Error: resolution 0\n"; return 2; }
    vector<string> names; vector<uint32_t> len;
    { ifstream fi(argv[a + 2]); if (fi.fail()) { cerr << "Error: cannot read " << argv[a + 2] << "\n"; return 10; }
      string n; uint32_t l; while (fi >> n >> l) { names.push_back(n); len.push_back(l); } }
    if (names.empty()) { cerr << "Error: no chromosomes in " << argv[a + 2] << "\n"; return 10; }
    if (genome.empty()) {                                                // hg38.info -> hg38, as the driver names its genomes
        genome = argv[a + 2];
        const size_t s = genome.find_last_of('/'); if (s != string::npos) genome = genome.substr(s + 1);
        const size_t d = genome.find_last_of('.'); if (d != string::npos && d > 0) genome = genome.substr(0, d);
    }
    hic::Writer w(genome, names, len, res);
    size_t cells = 0; string err;
    for (uint32_t r : res) {
        vector<uint32_t> b1, b2, ct;
        if (int rc = read_coo(string(argv[a]) + "." + to_string(r) + ".coo", b1, b2, ct)) return rc;
        if (w.add(r, b1.data(), b2.data(), ct.data(), b1.size(), &err)) { cerr << "Error: " << err << "\n"; return 10; }
        cells += b1.size();
    }
    if (w.write(argv[a + 1], &err)) { cerr << "Error: " << err << "\n"; return 10; }
    cerr << "INFO: " << cells << " cells at " << res.size() << " resolutions written to " << argv[a + 1] << ".\n";
    return 0;
}
