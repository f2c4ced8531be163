// stub_cabi.cpp — TEST DOUBLE of libmicrocket_b200.so.  Never shipped, never built into the package: tests/test_cli_glue_stub.py
// compiles it into a temporary directory next to COPIES of bin/sam2pairs and bin/pairs2bins, so that the host-side glue of the
// executables (argument handling, the SAM / BAM input source, streaming loops, file writing, the .hic container) runs on a box
// without a GPU.  "Device" memory is host memory here; sam2pairs' compute is replaced by an ECHO (the pulled pair text is the
// pushed SAM text), the pair parser / binning by straightforward host loops.  It says nothing about the CUDA path.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>
#include "../../include/microcket_b200.h"

struct mk_ctx { std::string buf; size_t off = 0; };
struct mk_pairs_ws { int dummy; };
struct mk_hist { std::vector<uint32_t> len, res; std::vector<mk_pair> pairs; };

extern "C" {
const char *mk_last_error(void) { return "stub: call not emulated"; }
int mk_device_count(void) { return 1; }
int mk_dev_alloc(int, size_t n, void **p) { *p = malloc(n ? n : 1); return *p ? 0 : -1; }
void mk_dev_free(void *p) { free(p); }
int mk_host_alloc(size_t n, void **p) { *p = malloc(n ? n : 1); return *p ? 0 : -1; }
void mk_host_free(void *p) { free(p); }
int mk_copy_to_device(void *d, const void *s, size_t n) { memcpy(d, s, n); return 0; }
int mk_copy_to_host(void *d, const void *s, size_t n) { memcpy(d, s, n); return 0; }
int mk_copy_device(void *d, const void *s, size_t n) { memmove(d, s, n); return 0; }

void mk_s2p_default_cfg(mk_s2p_cfg *c) { memset(c, 0, sizeof *c); c->min_mapped_ratio = 0.5f; c->min_mapq = 10; c->emit_text = 1; c->hskip1 = c->hskip2 = 5; c->klen1 = c->klen2 = 16; }
int mk_s2p_create(const mk_s2p_cfg *, const char *const *, int, mk_ctx **out) { *out = new mk_ctx; return 0; }
void mk_destroy(mk_ctx *c) { delete c; }
int mk_s2p_push(mk_ctx *c, const char *p, size_t n, int) { if (n) c->buf.append(p, n); return 0; }
int mk_s2p_pull(mk_ctx *c, char *out, size_t cap, size_t *n, char *, size_t, size_t *n2) {
    const size_t m = std::min(cap, c->buf.size() - c->off);
    memcpy(out, c->buf.data() + c->off, m); c->off += m; *n = m; if (n2) *n2 = 0;
    if (c->off == c->buf.size()) { c->buf.clear(); c->off = 0; }
    return 0;
}
int mk_s2p_finish(mk_ctx *, mk_s2p_stats *st) { memset(st, 0, sizeof *st); return 0; }
int mk_s2p_rmdup_stats(mk_ctx *, struct mk_dedup_stats_s *) { return -1; }
int mk_s2p_chrom_count(mk_ctx *) { return 0; }
int mk_s2p_chrom_name(mk_ctx *, int, char *, size_t) { return -1; }
int mk_s2p_run_device(mk_ctx *, const char *, size_t, int, mk_s2p_dev_io *, void *) { return -1; }
int mk_pairs_chrom_ranks(const char *const *, int, uint16_t *) { return -1; }
int mk_pairs_sort_text_device(mk_pairs_ws *, const mk_pair *, size_t, const uint8_t *, const char *, const uint64_t *, const uint16_t *, int, uint32_t,
                              char *, size_t, size_t *, size_t *, void *) { return -1; }
int mk_pairs_dedup_bin_indexed_device(mk_pairs_ws *, mk_pair *, size_t, const uint32_t *, int, const uint16_t *, int, uint32_t, uint16_t,
                                      uint32_t *, uint32_t *, uint32_t *, size_t, uint8_t *, uint32_t *, size_t *, size_t *, void *) { return -1; }

int mk_pairs_ws_create(int, size_t, mk_pairs_ws **w) { *w = new mk_pairs_ws; return 0; }
void mk_pairs_ws_destroy(mk_pairs_ws *w) { delete w; }

int mk_pairs_parse_text_device(mk_pairs_ws *, const char *t, size_t nb, const char *const *names, int n_chrom, mk_pair *out, size_t cap,
                               size_t *n_lines, size_t *n_skipped, void *) {
    size_t n = 0, skipped = 0, p = 0;
    while (p < nb) {
        const char *nl = (const char *)memchr(t + p, '\n', nb - p);
        const size_t e = nl ? (size_t)(nl - t) : nb;
        std::vector<std::string> f; size_t a = p;
        for (size_t i = p; i <= e; ++i) if (i == e || t[i] == '\t') { f.emplace_back(t + a, i - a); a = i + 1; }
        p = e + 1;
        int c1 = -1, c2 = -1;
        if (f.size() >= 7 && f[0][0] != '#') for (int k = 0; k < n_chrom; ++k) { if (f[1] == names[k]) c1 = k; if (f[3] == names[k]) c2 = k; }
        if (c1 < 0 || c2 < 0) { ++skipped; continue; }
        if (n >= cap) return -1;
        mk_pair q; memset(&q, 0, sizeof q);
        q.chr1 = (uint16_t)c1; q.chr2 = (uint16_t)c2; q.pos1 = (uint32_t)strtoul(f[2].c_str(), NULL, 10); q.pos2 = (uint32_t)strtoul(f[4].c_str(), NULL, 10);
        q.strands = (uint8_t)((f[5] == "-") | ((f[6] == "-") << 1));
        out[n++] = q;
    }
    *n_lines = n; *n_skipped = skipped;
    return 0;
}

static size_t coo(const mk_pair *p, size_t n, const uint32_t *len, int nc, uint32_t res, uint32_t *b1, uint32_t *b2, uint32_t *ct) {
    std::vector<uint64_t> off(nc + 1, 0);
    for (int c = 0; c < nc; ++c) off[c + 1] = off[c] + len[c] / res + 1;
    std::map<std::pair<uint32_t, uint32_t>, uint32_t> m;
    for (size_t i = 0; i < n; ++i) {
        uint32_t a = (uint32_t)(off[p[i].chr1] + p[i].pos1 / res), b = (uint32_t)(off[p[i].chr2] + p[i].pos2 / res);
        if (a > b) std::swap(a, b);
        ++m[{a, b}];
    }
    size_t k = 0;
    for (auto &e : m) { b1[k] = e.first.first; b2[k] = e.first.second; ct[k] = e.second; ++k; }
    return k;
}
int mk_pairs_bin_device(mk_pairs_ws *, const mk_pair *p, size_t n, const uint32_t *len, int nc, const uint16_t *, int, uint32_t res,
                        uint32_t *b1, uint32_t *b2, uint32_t *ct, size_t, size_t *nnz, void *) { *nnz = coo(p, n, len, nc, res, b1, b2, ct); return 0; }
int mk_pairs_dedup_bin_device(mk_pairs_ws *, mk_pair *p, size_t n, const uint32_t *len, int nc, const uint16_t *, int, uint32_t res, uint16_t,
                              uint32_t *b1, uint32_t *b2, uint32_t *ct, size_t, size_t *n_kept, size_t *nnz, void *) {
    auto key = [](const mk_pair &q) { return std::make_tuple(q.lane, q.chr1, q.pos1, q.chr2, q.pos2, q.strands); };
    std::vector<mk_pair> v(p, p + n);
    std::stable_sort(v.begin(), v.end(), [&](const mk_pair &a, const mk_pair &b) { return key(a) < key(b); });
    size_t k = 0;
    for (size_t i = 0; i < n; ++i) if (i == 0 || key(v[i]) != key(v[i - 1])) p[k++] = v[i];
    *n_kept = k; *nnz = coo(p, k, len, nc, res, b1, b2, ct);
    return 0;
}
int mk_hist_cells(const uint32_t *len, int nc, uint32_t res, uint64_t *nb, uint64_t *ncells) {
    uint64_t b = 0; for (int c = 0; c < nc; ++c) b += len[c] / res + 1;
    *nb = b; *ncells = b * (b + 1) / 2; return 0;
}
int mk_hist_create(int, const uint32_t *len, int nc, const uint32_t *res, int nr, uint32_t *const *, mk_hist **h) {
    *h = new mk_hist; (*h)->len.assign(len, len + nc); (*h)->res.assign(res, res + nr); return 0;
}
void mk_hist_destroy(mk_hist *h) { delete h; }
int mk_hist_add_device(mk_hist *h, const mk_pair *p, size_t n, const uint16_t *, int, void *) { h->pairs.insert(h->pairs.end(), p, p + n); return 0; }
int mk_hist_coo_device(mk_hist *h, int ri, uint32_t *b1, uint32_t *b2, uint32_t *ct, size_t, size_t *nnz, uint64_t *total, void *) {
    *nnz = coo(h->pairs.data(), h->pairs.size(), h->len.data(), (int)h->len.size(), h->res[ri], b1, b2, ct);
    *total = h->pairs.size(); return 0;
}
}
