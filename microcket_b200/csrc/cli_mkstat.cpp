// make.stat — the driver's statistics glue (bin/make.stat.pl:24-130, called at microcket:517) without perl: sums the
// key<TAB>value logs of the run (ktrim, krmdup, stitching, flash2pairs / unc2pairs) into <sid>.final.stat's table.
// usage: make.stat <sid> <concat=yes|no>      (stdout, like the perl script)
// Host-only text glue: nothing here runs on the GPU; it exists so that a Microcket install without perl still gets the
// same `.final.stat` from the logs the drop-in sam2pairs / krmdup write.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <regex>
#include <sstream>
#include <string>
#include <sys/stat.h>
#include <vector>

typedef std::map<std::string, double> Tab;

static bool add_log(const std::string &path, Tab &t, bool with_all) {          // make.stat.pl:24-31,33-39,103-121
    std::ifstream in(path);
    if (!in) { fprintf(stderr, "%s: %s\n", path.c_str(), strerror(errno)); return false; }
    std::string line;
    while (std::getline(in, line)) {
        size_t tab = line.find('\t');
        if (tab == std::string::npos) continue;
        std::string key = line.substr(0, tab), rest = line.substr(tab + 1);
        size_t tab2 = rest.find('\t');
        double v = atof(rest.substr(0, tab2).c_str());
        t[key] += v;
        if (with_all) t["all"] += v;
    }
    return true;
}

static std::string d(double v) {                                               // make.stat.pl:143-147: thousands separators
    char buf[64]; snprintf(buf, sizeof buf, "%.0f", v);
    std::string s = buf, out;
    const bool neg = !s.empty() && s[0] == '-';
    std::string digits = neg ? s.substr(1) : s;
    for (size_t i = 0; i < digits.size(); ++i) { if (i && (digits.size() - i) % 3 == 0) out.push_back(','); out.push_back(digits[i]); }
    return neg ? "-" + out : out;
}

static bool nonempty(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0 && st.st_size > 0; }

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "\nUsage: %s <sid> <concat=yes|no>\n\n", argv[0]); return 2; }
    const std::string sid = argv[1], concat = argv[2];
    printf("#Category\tCount\tFraction(%%)\n");
    Tab trim, rmdup, align;
    if (!add_log(sid + ".trim.log", trim, false)) return 1;
    if (!add_log(sid + ".rmdup.log", rmdup, false)) return 1;
    printf("## Preprocessing and alignment\n");
    printf("Total\t%s\t100.0\nKtrim\t%s\t%.1f\nUnique\t%s\t%.1f\n", d(trim["Total"]).c_str(), d(rmdup["Total"]).c_str(),
           rmdup["Total"] / trim["Total"] * 100, d(rmdup["Uniq"]).c_str(), rmdup["Uniq"] / rmdup["Total"] * 100);
    double prealign;
    if (concat == "yes") {
        double cat = 0, unc = 0, cut = 0;
        if (nonempty(sid + ".flash.log")) {                                    // old version: FLASH's own log
            std::ifstream in(sid + ".flash.log"); std::string line; std::smatch m;
            const std::regex re("\\sCombined pairs:\\s+(\\d+)");
            while (std::getline(in, line)) if (std::regex_search(line, m, re)) { cat = atof(m[1].str().c_str()); break; }
            if (nonempty(sid + ".cut.log")) {
                std::ifstream in2(sid + ".cut.log");
                const std::regex rt("Total\\s+(\\d+)"), rp("Pass\\s+(\\d+)");
                while (std::getline(in2, line)) {
                    if (std::regex_search(line, m, rt)) unc = atof(m[1].str().c_str());
                    if (std::regex_search(line, m, rp)) cut = atof(m[1].str().c_str());
                }
            } else { unc = rmdup["Uniq"] - cat; cut = unc; }
        } else {                                                               // new version: <sid>.stitch.stat, one line of key/value columns
            std::ifstream in(sid + ".stitch.stat");
            if (!in) { fprintf(stderr, "%s.stitch.stat: %s\n", sid.c_str(), strerror(errno)); return 1; }
            std::string line; std::getline(in, line);
            std::vector<std::string> f; std::stringstream ss(line); std::string tok;
            while (std::getline(ss, tok, '\t')) f.push_back(tok);
            f.resize(6);
            cat = atof(f[1].c_str()); unc = atof(f[3].c_str()); cut = atof(f[5].c_str());
        }
        printf("Stitched\t%s\t%.1f\nUnstitched\t%s\t%.1f\n  Discarded(too-short)\t%s\t%.1f\n", d(cat).c_str(), cat / rmdup["Uniq"] * 100,
               d(cut).c_str(), cut / rmdup["Uniq"] * 100, d(unc - cut).c_str(), (unc - cut) / rmdup["Uniq"] * 100);
        prealign = cat + cut;
        if (!add_log(sid + ".flash2pairs.log", align, true)) return 1;
    } else prealign = rmdup["Uniq"];
    if (!add_log(sid + ".unc2pairs.log", align, true)) return 1;
    const double all = align["all"];
    printf("Mappable\t%s\t%.1f\n", d(all).c_str(), all / prealign * 100);
    printf("## Interactions\n");
    const double uncalled = align["lowMap"] + align["manyHits"] + align["unpaired"] + align["selfCircle"];
    printf("Uncalled\t%s\t%.1f\n", d(uncalled).c_str(), uncalled / all * 100);
    printf("  Incomplete-mapping\t%s\t%.1f\n", d(align["lowMap"]).c_str(), align["lowMap"] / all * 100);
    printf("  Too-many-segments\t%s\t%.1f\n", d(align["manyHits"]).c_str(), align["manyHits"] / all * 100);
    printf("  Unpairable\t%s\t%.1f\n", d(align["unpaired"]).c_str(), align["unpaired"] / all * 100);
    printf("  Self-circle\t%s\t%.1f\n", d(align["selfCircle"]).c_str(), align["selfCircle"] / all * 100);
    const double valid = align["trans"] + align["cis10K"] + align["cis1K"] + align["cis0"];
    printf("Reported\t%s\t%.1f\n", d(valid).c_str(), valid / all * 100);
    printf("  Cis(<1K)\t%s\t%.1f\n", d(align["cis0"]).c_str(), align["cis0"] / all * 100);
    printf("  Cis(1-10K)\t%s\t%.1f\n", d(align["cis1K"]).c_str(), align["cis1K"] / all * 100);
    printf("  Cis(>=10K)\t%s\t%.1f\n", d(align["cis10K"]).c_str(), align["cis10K"] / all * 100);
    printf("  Trans\t%s\t%.1f\n", d(align["trans"]).c_str(), align["trans"] / all * 100);
    return 0;
}
