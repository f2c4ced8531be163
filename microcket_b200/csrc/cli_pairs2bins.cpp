// cli_pairs2bins.cpp — contact binning of a .pairs file on the GPU (new tool; stands where the driver calls
// `java -jar juicer_tools.jar pre -r <res,...> <sid>.final.pairs <sid>.hic <genome>.info`, microcket:525-529).
//   pairs2bins [-d] -r <res[,res...]> <in.pairs> <out.prefix> <genome.info>
// Writes <out.prefix>.<res>.coo with `bin1<TAB>bin2<TAB>count` (upper triangle, sorted), bins numbered in .info
// order with bin = offset[chr] + pos / res.  -d removes coordinate duplicates first (first occurrence wins).
// Writing .hic itself is out of scope.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include "../../include/microcket_b200.h"
using namespace std;

int main(int argc, char *argv[]) {
    bool dedup = false; string reslist;
    int a = 1;
    while (a < argc && argv[a][0] == '-' && argv[a][1]) {
        if (!strcmp(argv[a], "-d")) { dedup = true; ++a; }
        else if (!strcmp(argv[a], "-r") && a + 1 < argc) { reslist = argv[a + 1]; a += 2; }
        else break;
    }
    if (argc - a < 3 || reslist.empty()) {
        cerr << "\nUsage: " << argv[0] << " [-d] -r <res[,res...]> <in.pairs> <out.prefix> <genome.info>\n\n";
        return 2;
    }
    vector<uint32_t> res;
    { stringstream ss(reslist); string t; while (getline(ss, t, ',')) if (!t.empty()) res.push_back((uint32_t)strtoul(t.c_str(), NULL, 10)); }
    map<string, int> chr_id; vector<uint32_t> chr_len;
    { ifstream fi(argv[a + 2]); if (fi.fail()) { cerr << "Error: cannot read " << argv[a + 2] << "\n"; return 10; }
      string n; uint32_t l; while (fi >> n >> l) { chr_id[n] = (int)chr_len.size(); chr_len.push_back(l); } }
    ifstream fp(argv[a]);
    if (fp.fail()) { cerr << "Error: cannot read " << argv[a] << "\n"; return 10; }
    vector<mk_pair> pairs; string line;
    while (getline(fp, line)) {
        if (line.empty() || line[0] == '#') continue;
        stringstream ss(line); string id, c1, c2, s1, s2; uint32_t p1, p2;
        if (!(ss >> id >> c1 >> p1 >> c2 >> p2 >> s1 >> s2)) continue;
        auto i1 = chr_id.find(c1), i2 = chr_id.find(c2);
        if (i1 == chr_id.end() || i2 == chr_id.end()) continue;
        mk_pair r; memset(&r, 0, sizeof r);
        r.chr1 = (uint16_t)i1->second; r.chr2 = (uint16_t)i2->second; r.pos1 = p1; r.pos2 = p2;
        r.strands = (uint8_t)((s1 == "-" ? 1 : 0) | (s2 == "-" ? 2 : 0));
        pairs.push_back(r);
    }
    int dev = getenv("MICROCKET_DEVICE") ? atoi(getenv("MICROCKET_DEVICE")) : 0;
    mk_pairs_ws *ws = NULL;
    if (mk_pairs_ws_create(dev, pairs.size() + 1, &ws) != MK_OK) { cerr << "Error: " << mk_last_error() << "\n"; return 20; }
    size_t n = pairs.size();
    vector<uint32_t> b1(n + 1), b2(n + 1), ct(n + 1);
    for (size_t k = 0; k < res.size(); ++k) {
        size_t kept = 0, nnz = 0;
        if (mk_pairs_dedup_bin_host(ws, pairs.data(), n, dedup && k == 0, chr_len.data(), (int)chr_len.size(), NULL, 0, res[k],
                                    b1.data(), b2.data(), ct.data(), n + 1, &kept, &nnz) != MK_OK) { cerr << "Error: " << mk_last_error() << "\n"; return 20; }
        n = kept;
        string out = string(argv[a + 1]) + "." + to_string(res[k]) + ".coo";
        FILE *fo = fopen(out.c_str(), "w");
        if (!fo) { cerr << "Error: cannot write " << out << "\n"; return 10; }
        for (size_t i = 0; i < nnz; ++i) fprintf(fo, "%u\t%u\t%u\n", b1[i], b2[i], ct[i]);
        fclose(fo);
    }
    mk_pairs_ws_destroy(ws);
    return 0;
}
