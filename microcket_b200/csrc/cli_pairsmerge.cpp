// cli_pairsmerge.cpp — merges already sorted `.pairs` files (host-only; second half of SURVEY.md §8(f)-1).  Stands where the
// driver pools the stitched and un-stitched pairs with
//   LANG=C sort -k2,2d -k4,4d -k3,3n -k5,5n --parallel=$thread -S 50% -m $sid.flash.pairs $sid.unc.pairs >>$sid.final.pairs
// (microcket:514): same order, byte-identical output, without GNU sort's per-comparison field splitting — every line's four
// keys are located once, when the line reaches the head of its stream.
//   pairsmerge <a.pairs> <b.pairs> [...]  > merged.pairs        (inputs sorted by sam2pairs ... sorted, or by the sort above)
// Order (GNU sort, C locale, no -b / -t / -s): field 2 then field 4 under -d (only blanks and alphanumerics take part; a field
// starts right behind the previous one, so it carries its leading blank), field 3 then field 5 under -n, then the whole line
// bytewise.  Equal lines come out in argument order.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

struct Stream {
    FILE *f = NULL; std::vector<char> buf; size_t lo = 0, hi = 0; bool eof = false;
    const char *line = NULL; size_t len = 0;                            // current line without its newline
    const char *k[4] = {0, 0, 0, 0}; size_t kl[4] = {0, 0, 0, 0};       // fields 2, 4 (text, with the leading blanks) and 3, 5 (numeric)
    long double num[2] = {0, 0};

    bool next() {                                                       // false at the end of the stream
        lo += line ? len + 1 : 0;
        while (true) {
            const char *nl = lo < hi ? (const char *)memchr(buf.data() + lo, '\n', hi - lo) : NULL;
            if (nl) { line = buf.data() + lo; len = (size_t)(nl - line); break; }
            if (eof) {
                if (lo < hi) { buf[hi] = '\n'; line = buf.data() + lo; len = hi - lo; ++hi; break; }   // last line without a newline
                line = NULL; return false;
            }
            if (lo > 0) { memmove(buf.data(), buf.data() + lo, hi - lo); hi -= lo; lo = 0; }
            if (hi + 1 >= buf.size()) buf.resize(buf.size() * 2);
            const size_t got = fread(buf.data() + hi, 1, buf.size() - hi - 1, f);
            hi += got; if (got == 0) eof = true;
        }
        // fields as GNU sort cuts them without -t: a field is a run of blanks followed by a run of non-blanks
        const char *p = line, *e = line + len;
        const char *fs[5], *fe[5];
        for (int i = 0; i < 5; ++i) {
            fs[i] = p;
            while (p < e && (*p == ' ' || *p == '\t')) ++p;
            while (p < e && *p != ' ' && *p != '\t') ++p;
            fe[i] = p;
        }
        k[0] = fs[1]; kl[0] = (size_t)(fe[1] - fs[1]); k[1] = fs[3]; kl[1] = (size_t)(fe[3] - fs[3]);
        k[2] = fs[2]; kl[2] = (size_t)(fe[2] - fs[2]); k[3] = fs[4]; kl[3] = (size_t)(fe[4] - fs[4]);
        for (int i = 0; i < 2; ++i) {                                   // -n: blanks, an optional '-', digits, an optional fraction
            const char *q = k[2 + i], *qe = q + kl[2 + i];
            while (q < qe && (*q == ' ' || *q == '\t')) ++q;
            bool neg = false; if (q < qe && *q == '-') { neg = true; ++q; }
            long double v = 0; while (q < qe && *q >= '0' && *q <= '9') v = v * 10 + (*q++ - '0');
            if (q < qe && *q == '.') { long double s = 0.1L; ++q; while (q < qe && *q >= '0' && *q <= '9') { v += s * (*q++ - '0'); s /= 10; } }
            num[i] = neg ? -v : v;
        }
        return true;
    }
};

static inline bool dict(unsigned char c) { return c == ' ' || c == '\t' || (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z'); }

static int cmp_dict(const char *a, size_t na, const char *b, size_t nb) {
    size_t i = 0, j = 0;
    while (true) {
        while (i < na && !dict((unsigned char)a[i])) ++i;
        while (j < nb && !dict((unsigned char)b[j])) ++j;
        if (i == na || j == nb) return (i < na) - (j < nb);
        if (a[i] != b[j]) return (unsigned char)a[i] < (unsigned char)b[j] ? -1 : 1;
        ++i; ++j;
    }
}

static int cmp(const Stream &a, const Stream &b) {
    if (int c = cmp_dict(a.k[0], a.kl[0], b.k[0], b.kl[0])) return c;
    if (int c = cmp_dict(a.k[1], a.kl[1], b.k[1], b.kl[1])) return c;
    if (a.num[0] != b.num[0]) return a.num[0] < b.num[0] ? -1 : 1;
    if (a.num[1] != b.num[1]) return a.num[1] < b.num[1] ? -1 : 1;
    const size_t n = a.len < b.len ? a.len : b.len;
    if (int c = memcmp(a.line, b.line, n)) return c;
    return (a.len > b.len) - (a.len < b.len);
}

int main(int argc, char *argv[]) {
    if (argc < 2) { std::cerr << "\nUsage: " << argv[0] << " <sorted.a.pairs> [<sorted.b.pairs> ...] > merged.pairs\n\n"; return 2; }
    std::vector<Stream> s(argc - 1);
    for (int i = 1; i < argc; ++i) {
        s[i - 1].f = strcmp(argv[i], "-") ? fopen(argv[i], "rb") : stdin;
        if (!s[i - 1].f) { std::cerr << "Error: cannot read " << argv[i] << "\n"; return 10; }
        s[i - 1].buf.resize((size_t)8 << 20);
    }
    std::vector<int> live;
    for (size_t i = 0; i < s.size(); ++i) if (s[i].next()) live.push_back((int)i);
    std::vector<char> out; out.reserve((size_t)9 << 20);
    while (!live.empty()) {
        size_t best = 0;
        for (size_t j = 1; j < live.size(); ++j) if (cmp(s[live[j]], s[live[best]]) < 0) best = j;      // ties: the earlier argument
        Stream &t = s[live[best]];
        out.insert(out.end(), t.line, t.line + t.len + 1);
        if (out.size() >= ((size_t)8 << 20)) { if (fwrite(out.data(), 1, out.size(), stdout) != out.size()) { std::cerr << "Error: write failed\n"; return 10; } out.clear(); }
        if (!t.next()) live.erase(live.begin() + (long)best);
    }
    if (!out.empty() && fwrite(out.data(), 1, out.size(), stdout) != out.size()) { std::cerr << "Error: write failed\n"; return 10; }
    return fflush(stdout) ? 10 : 0;
}
