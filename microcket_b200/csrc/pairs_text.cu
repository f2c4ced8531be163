// pairs_text.cu — the .pairs TEXT after sam2pairs, on the device: the driver's sort and the deduplicated file.
//
//   mk_pairs_sort_text_device    the lines of the kept pairs in the order of
//                                `LANG=C sort -k2,2d -k4,4d -k3,3n -k5,5n` (microcket:480,484,502,506,514): chromosome names
//                                in dictionary order, then the two positions numerically, then — GNU sort's last resort — the
//                                whole line bytewise.  One radix sort of packed (rank1, rank2, pos1, pos2 | index) records,
//                                a fix-up of the runs of equal keys by line comparison, one gather of the lines.
//   mk_pairs_filter_text_device  the lines of the kept pairs in input order (what a deduplicated, still unsorted stream is).
//   mk_pairs_chrom_ranks         host helper: rank of every chromosome name under sort's `-d` rule in the C locale.
//
// Lines are addressed through the offsets k_emit records (mk_s2p_dev_io.d_line_off): line e = text[off[e], off[e+1]).
// Line format: unc2pairs.h:327-347 (rid \t chr1 \t pos1 \t chr2 \t pos2 \t s1 \t s2 \n).
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>
#include "mk_common.cuh"
#include "radix_sort.cuh"
#include "pairs_ws.h"

#define PT_T 256                       // lines per tile = threads per CTA
#define PT_STAGE 24576                 // bytes of one tile staged in shared memory (96 B per line on average)

struct SortCfg { u32 nbr, nbp, total_bits; };

// record = [ key : total_bits, left-aligned below bit 128 .. above bit 32 ][ input index : 32 ]
__global__ void __launch_bounds__(256) k_sort_keys(const mk_pair *p, u64 n, const u8 *keep, const u16 *rank_by_id, u32 n_ids, SortCfg c,
                                                   uint4 *key, unsigned long long *counters /* [0] kept */) {
    u32 nk = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const uint4 r = ((const uint4 *)p)[i];
        const u32 c1 = r.z & 0xFFFFu, c2 = r.z >> 16;
        uint4 k = make_uint4((u32)i, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        const bool in_range = c.nbp >= 32 || ((r.x >> c.nbp) == 0 && (r.y >> c.nbp) == 0);
        if ((!keep || keep[i]) && !in_range) atomicAdd(&counters[2], 1ull);     // a position above the caller's max_pos: reported
        if ((!keep || keep[i]) && c1 < n_ids && c2 < n_ids && in_range) {
            // 96 key bits: hi = rank1 : rank2 : top of pos1 ..., packed most significant first
            u64 hi = 0, lo = 0;                                          // (hi:lo) as a 128-bit integer, built by shifting
            auto put = [&](u64 v, u32 w) { hi = (hi << w) | (lo >> (64 - w)); lo = (lo << w) | v; };
            put(rank_by_id[c1], c.nbr); put(rank_by_id[c2], c.nbr); put(r.x, c.nbp); put(r.y, c.nbp);
            // left-align inside 96 bits, then place above the index word
            const u32 sh = 96u - c.total_bits;
            if (sh >= 64) { hi = lo << (sh - 64); lo = 0; }
            else if (sh) { hi = (hi << sh) | (lo >> (64 - sh)); lo <<= sh; }
            k.y = (u32)lo; k.z = (u32)(lo >> 32); k.w = (u32)hi;
            ++nk;
        }
        key[i] = k;
    }
    nk = __reduce_add_sync(0xffffffffu, nk);
    if ((threadIdx.x & 31) == 0 && nk) atomicAdd(&counters[0], (unsigned long long)nk);
}

__device__ __forceinline__ bool key_eq(const uint4 &a, const uint4 &b) { return a.y == b.y && a.z == b.z && a.w == b.w; }

// bytewise order of two lines (memcmp over the common length, the shorter line first)
__device__ __forceinline__ int line_cmp(const char *text, const u64 *off, u32 a, u32 b) {
    const u64 oa = off[a], ob = off[b];
    const u32 la = (u32)(off[a + 1] - oa), lb = (u32)(off[b + 1] - ob), m = la < lb ? la : lb;
    for (u32 k = 0; k < m; ++k) {
        const int d = (int)(unsigned char)text[oa + k] - (int)(unsigned char)text[ob + k];
        if (d) return d;
    }
    return la < lb ? -1 : (la > lb ? 1 : 0);
}

// runs of equal keys are in input order (stable sort): put them in whole-line order.  One thread per run head.
__global__ void __launch_bounds__(256) k_tie_fix(uint4 *b0, uint4 *b1, const RadixPlan *plan, const unsigned long long *counters,
                                                 const char *text, const u64 *off) {
    uint4 *s = plan->final_buf ? b1 : b0;
    const u64 m = counters[0];
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i + 1 < m; i += (u64)gridDim.x * blockDim.x) {
        const uint4 k = s[i];
        if (i > 0 && key_eq(s[i - 1], k)) continue;                     // not a run head
        if (!key_eq(s[i + 1], k)) continue;                             // run of one
        u64 e = i + 2;
        while (e < m && key_eq(s[e], k)) ++e;
        for (u64 a = i + 1; a < e; ++a) {                               // insertion sort of the run's indices
            const u32 x = s[a].x;
            u64 b = a;
            while (b > i && line_cmp(text, off, s[b - 1].x, x) > 0) { s[b].x = s[b - 1].x; --b; }
            s[b].x = x;
        }
    }
}

// bytes of every tile of PT_T lines, in output order (order == NULL: identity; keep == NULL: every line)
__global__ void __launch_bounds__(PT_T) k_line_tile_sums(const uint4 *b0, const uint4 *b1, const RadixPlan *plan, const u8 *keep,
                                                         const u64 *off, u64 m, const unsigned long long *m_dev, u32 *tile_sum) {
    __shared__ u32 s_w[PT_T / 32];
    const uint4 *order = b0 ? (plan->final_buf ? b1 : b0) : nullptr;
    if (m_dev) m = *m_dev;
    const u64 i = (u64)blockIdx.x * PT_T + threadIdx.x;
    u32 len = 0;
    if (i < m) {
        const u64 src = order ? order[i].x : i;
        if (order || !keep || keep[src]) len = (u32)(off[src + 1] - off[src]);
    }
    len = __reduce_add_sync(0xffffffffu, len);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = len;
    __syncthreads();
    if (threadIdx.x == 0) { u32 t = 0; for (int w = 0; w < PT_T / 32; ++w) t += s_w[w]; tile_sum[blockIdx.x] = t; }
}

// exclusive prefix of the tile sums (one CTA), total -> counters[1]
__global__ void __launch_bounds__(1024) k_line_tile_prefix(const u32 *tile_sum, u64 *tile_off, u64 n_tiles, unsigned long long *counters) {
    __shared__ u64 s_w[32];
    const u32 tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const u64 per = (n_tiles + 1023) / 1024;
    const u64 lo = (u64)tid * per < n_tiles ? (u64)tid * per : n_tiles, hi = lo + per < n_tiles ? lo + per : n_tiles;
    u64 sum = 0;
    for (u64 i = lo; i < hi; ++i) sum += tile_sum[i];
    u64 inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const u64 t = __shfl_up_sync(0xffffffffu, inc, d); if ((int)lane >= d) inc += t; }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const u64 v = s_w[lane]; u64 vi = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const u64 t = __shfl_up_sync(0xffffffffu, vi, d); if ((int)lane >= d) vi += t; }
        s_w[lane] = vi - v;
        if (lane == 31) counters[1] = vi;
    }
    __syncthreads();
    u64 run = s_w[wid] + inc - sum;
    for (u64 i = lo; i < hi; ++i) { tile_off[i] = run; run += tile_sum[i]; }
}

// the lines of one tile, concatenated in output order: staged in shared memory, written with 16-byte stores
__global__ void __launch_bounds__(PT_T) k_gather_lines(const uint4 *b0, const uint4 *b1, const RadixPlan *plan, const u8 *keep,
                                                       const char *text, const u64 *off, u64 m, const unsigned long long *m_dev,
                                                       const u64 *tile_off, const u32 *tile_sum, char *out, u64 out_cap, u32 *err) {
    __shared__ __align__(16) char s_stage[PT_STAGE + 16];
    __shared__ u32 s_scan[PT_T / 32 + 1];
    const uint4 *order = b0 ? (plan->final_buf ? b1 : b0) : nullptr;
    if (m_dev) m = *m_dev;
    const int tid = threadIdx.x;
    const u64 i = (u64)blockIdx.x * PT_T + tid;
    if ((u64)blockIdx.x * PT_T >= m) return;
    u64 so = 0; u32 len = 0;
    if (i < m) {
        const u64 src = order ? order[i].x : i;
        if (order || !keep || keep[src]) { so = off[src]; len = (u32)(off[src + 1] - so); }
    }
    u32 tot;
    const u32 ex = block_excl_scan<PT_T>(len, s_scan, &tot);
    const u64 t_off = tile_off[blockIdx.x];
    if (t_off + tot > out_cap) { if (tid == 0) atomicOr(err, 1u); return; }
    const bool staged = tot <= PT_STAGE;
    const u32 phase = (u32)((uintptr_t)(out + t_off) & 15u);            // stage with the destination's 16-byte phase
    char *dst = staged ? s_stage + phase + ex : out + t_off + ex;
    if (len) {
        // unaligned source: aligned 8-byte loads, realigned in registers
        const u64 a8 = so & ~(u64)7; const u32 sh = (u32)(so & 7u) * 8u;
        const u64 *src = (const u64 *)(text + a8);
        const u32 last_word = (u32)((so + len - 1 - a8) >> 3);          // index of the last 8-byte word the line touches
        u64 cur = __ldg(src);
        for (u32 k = 0; k < len; k += 8) {
            const u64 nxt = (k >> 3) + 1 <= last_word ? __ldg(src + (k >> 3) + 1) : 0ull;
            const u64 x = sh ? (cur >> sh) | (nxt << (64u - sh)) : cur;
            cur = nxt;
            const u32 nb = len - k < 8u ? len - k : 8u;
#pragma unroll
            for (int b = 0; b < 8; ++b) if ((u32)b < nb) dst[k + b] = (char)(x >> (8 * b));
        }
    }
    if (staged && tot) {
        __syncthreads();
        char *o = out + t_off;
        const u32 head = phase ? (16u - phase < tot ? 16u - phase : tot) : 0u;
        if ((u32)tid < head) o[tid] = s_stage[phase + tid];
        const u32 body = (tot - head) >> 4;
        for (u32 w = tid; w < body; w += PT_T) st_stream_v4((uint4 *)(o + head + ((u64)w << 4)), *(const uint4 *)(s_stage + phase + head + (w << 4)));
        const u32 tail0 = head + (body << 4);
        if (tail0 + tid < tot) o[tail0 + tid] = s_stage[phase + tail0 + tid];
    }
}

static u32 bits_for_v(u64 max_value) { u32 b = 1; while (b < 64 && (max_value >> b)) ++b; return b; }

// tile sums -> prefix -> gather, for m lines (m_dev: count on the device, with m an upper bound for the grid)
static int gather_lines(mk_pairs_ws *w, const uint4 *b0, const uint4 *b1, const u8 *keep, const char *text, const u64 *off, u64 m_bound,
                        const unsigned long long *m_dev, char *out, u64 out_cap, cudaStream_t s) {
    const u64 n_tiles = (m_bound + PT_T - 1) / PT_T;
    if (n_tiles == 0) return MK_OK;
    MK_TRY(w->text_scratch(n_tiles));
    const RadixPlan *plan = w->rws.plan.as<RadixPlan>();
    unsigned long long *cnt = w->counter.as<unsigned long long>();
    k_line_tile_sums<<<(unsigned)n_tiles, PT_T, 0, s>>>(b0, b1, plan, keep, off, m_bound, m_dev, w->tile_sum.as<u32>());
    k_line_tile_prefix<<<1, 1024, 0, s>>>(w->tile_sum.as<u32>(), w->tile_off.as<u64>(), n_tiles, cnt);
    k_gather_lines<<<(unsigned)n_tiles, PT_T, 0, s>>>(b0, b1, plan, keep, text, off, m_bound, m_dev, w->tile_off.as<u64>(), w->tile_sum.as<u32>(),
                                                      out, out_cap, (u32 *)(cnt + 5));
    w->launches += 3;
    return MK_OK;
}

extern "C" int mk_pairs_sort_text_device(mk_pairs_ws *w, const mk_pair *d_pairs, size_t n, const uint8_t *d_keep,
                                         const char *d_text, const uint64_t *d_line_off, const uint16_t *chrom_rank, int n_ids,
                                         uint32_t max_pos, char *d_out, size_t out_cap, size_t *out_len, size_t *n_lines, void *stream) {
    if (!w || !out_len || !n_lines || !chrom_rank || n_ids <= 0 || n_ids > 16384 || (n && (!d_pairs || !d_text || !d_line_off || !d_out))) {
        mk_set_error("mk_pairs_sort_text_device: bad argument"); return MK_ERR_ARG;
    }
    if (n > w->max_pairs) { mk_set_error("mk_pairs_sort_text_device: workspace holds %zu pairs, got %zu", w->max_pairs, n); return MK_ERR_CAPACITY; }
    MK_CUDA(cudaSetDevice(w->device));
    cudaStream_t s = (cudaStream_t)stream;
    *out_len = 0; *n_lines = 0;
    if (n == 0) return MK_OK;
    u32 max_rank = 0;
    for (int i = 0; i < n_ids; ++i) max_rank = std::max<u32>(max_rank, chrom_rank[i]);
    SortCfg sc; sc.nbr = bits_for_v(max_rank); sc.nbp = max_pos ? bits_for_v(max_pos) : 32; sc.total_bits = 2 * sc.nbr + 2 * sc.nbp;
    if (sc.total_bits > 95) { mk_set_error("mk_pairs_sort_text_device: key needs %u bits, 95 available", sc.total_bits); return MK_ERR_CAPACITY; }
    if (!w->sort2.p) MK_TRY(w->sort2.alloc(w->max_pairs * 16));
    u16 *d_rank = (u16 *)w->chr_off.p;
    MK_CUDA(cudaMemcpyAsync(d_rank, chrom_rank, (size_t)n_ids * 2, cudaMemcpyHostToDevice, s));
    MK_CUDA(cudaMemsetAsync(w->counter.p, 0, 64, s));
    unsigned long long *cnt = w->counter.as<unsigned long long>();
    uint4 *k0 = w->alt.as<uint4>(), *k1 = w->sort2.as<uint4>();
    k_sort_keys<<<w->sms * 8, 256, 0, s>>>(d_pairs, n, d_keep, d_rank, (u32)n_ids, sc, k0, cnt);
    w->launches += 1;
    // LSD over the key bytes only: the key sits left-aligned in bytes 4..15
    RadixSchedule sch; sch.n_pass = 0;
    const int first_byte = 16 - (int)((sc.total_bits + 7) / 8);
    for (int b = first_byte; b < 16; ++b) sch.byte_of[sch.n_pass++] = b;
    Rec16::Bufs b; b.k[0] = k0; b.k[1] = k1; b.v[0] = b.v[1] = nullptr;
    MK_TRY(radix_sort<Rec16>(b, n, sch, w->rws, 0, w->sms, s, &w->launches));
    k_tie_fix<<<w->sms * 8, 256, 0, s>>>(k0, k1, w->rws.plan.as<RadixPlan>(), cnt, d_text, d_line_off);
    w->launches += 1;
    MK_TRY(gather_lines(w, k0, k1, nullptr, d_text, d_line_off, n, cnt, d_out, out_cap, s));
    unsigned long long h[6];
    MK_CUDA(cudaMemcpyAsync(h, cnt, 48, cudaMemcpyDeviceToHost, s));
    MK_CUDA(cudaStreamSynchronize(s));
    if ((u32)h[5]) { mk_set_error("mk_pairs_sort_text_device: output capacity %zu too small (%llu bytes)", out_cap, h[1]); return MK_ERR_CAPACITY; }
    if (h[2]) { mk_set_error("mk_pairs_sort_text_device: %llu positions above max_pos %u", h[2], max_pos); return MK_ERR_INPUT; }
    *n_lines = (size_t)h[0]; *out_len = (size_t)h[1];
    return MK_OK;
}

extern "C" int mk_pairs_filter_text_device(mk_pairs_ws *w, size_t n, const uint8_t *d_keep, const char *d_text, const uint64_t *d_line_off,
                                           char *d_out, size_t out_cap, size_t *out_len, void *stream) {
    if (!w || !out_len || (n && (!d_text || !d_line_off || !d_out))) { mk_set_error("mk_pairs_filter_text_device: bad argument"); return MK_ERR_ARG; }
    MK_CUDA(cudaSetDevice(w->device));
    cudaStream_t s = (cudaStream_t)stream;
    *out_len = 0;
    if (n == 0) return MK_OK;
    MK_CUDA(cudaMemsetAsync(w->counter.p, 0, 64, s));
    MK_TRY(gather_lines(w, nullptr, nullptr, d_keep, d_text, d_line_off, n, nullptr, d_out, out_cap, s));
    unsigned long long h[6];
    MK_CUDA(cudaMemcpyAsync(h, w->counter.p, 48, cudaMemcpyDeviceToHost, s));
    MK_CUDA(cudaStreamSynchronize(s));
    if ((u32)h[5]) { mk_set_error("mk_pairs_filter_text_device: output capacity %zu too small (%llu bytes)", out_cap, h[1]); return MK_ERR_CAPACITY; }
    *out_len = (size_t)h[1];
    return MK_OK;
}

// sort's `-d`: only blanks and alphanumerics take part in the comparison (C locale: bytewise on what is left).  Names that
// compare equal get the same rank; the whole-line comparison then separates them, exactly as sort's last resort does.
extern "C" int mk_pairs_chrom_ranks(const char *const *names, int n, uint16_t *rank) {
    if (!names || !rank || n < 0 || n > 16384) { mk_set_error("mk_pairs_chrom_ranks: bad argument"); return MK_ERR_ARG; }
    std::vector<std::string> key(n);
    for (int i = 0; i < n; ++i)
        for (const char *c = names[i]; c && *c; ++c)
            if (*c == ' ' || *c == '\t' || (*c >= '0' && *c <= '9') || (*c >= 'A' && *c <= 'Z') || (*c >= 'a' && *c <= 'z')) key[i].push_back(*c);
    std::vector<int> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    std::sort(idx.begin(), idx.end(), [&](int a, int b) { return key[a] < key[b]; });   // std::string compares as unsigned bytes
    u16 r = 0;
    for (int k = 0; k < n; ++k) {
        if (k > 0 && key[idx[k]] != key[idx[k - 1]]) ++r;
        rank[idx[k]] = r;
    }
    return MK_OK;
}

// ------------------------------------------------------------------------------------------------ .pairs text -> packed pairs
// What pairs2bins needs from a `.final.pairs` file (anno/4DN.DCIC.header:2: readID chr1 pos1 chr2 pos2 strand1 strand2):
// newline index (count per 4 KiB block, prefix, positions), then one thread per line.  Header lines ('#'), short lines and
// chromosome names outside the table become records with chromosome id 0xFFFF: the dedup / binning calls leave those out
// and count them, so no compaction pass is needed here.
#define NL_T 256
__device__ __forceinline__ u32 nl_mask16(const uint4 &w) {              // bit q <-> byte q of the 16-byte word is '\n'
    auto y = [](u32 x) { const u32 v = x ^ 0x0A0A0A0Au; return ~(((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v | 0x7F7F7F7Fu); };   // 0x80 where the byte is '\n' (exact)
    return gather_flags4(y(w.x)) | (gather_flags4(y(w.y)) << 4) | (gather_flags4(y(w.z)) << 8) | (gather_flags4(y(w.w)) << 12);
}
__device__ __forceinline__ uint4 ld16_bounded(const char *text, u64 off, u64 n) {   // bytes at and beyond n read as 0
    if (off + 16 <= n) return *(const uint4 *)(text + off);
    u32 w[4] = {0, 0, 0, 0};
    for (u32 b = 0; b < 16 && off + b < n; ++b) w[b >> 2] |= (u32)(unsigned char)text[off + b] << (8 * (b & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}
__global__ void __launch_bounds__(NL_T) k_nl_count(const char *text, u64 n, u32 *blk_cnt) {
    __shared__ u32 s_w[NL_T / 32];
    const u64 off = ((u64)blockIdx.x * NL_T + threadIdx.x) * 16;
    u32 c = off < n ? __popc(nl_mask16(ld16_bounded(text, off, n))) : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { u32 t = 0; for (int w = 0; w < NL_T / 32; ++w) t += s_w[w]; blk_cnt[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(NL_T) k_nl_write(const char *text, u64 n, const u64 *blk_off, u32 *nl_pos, u64 cap) {
    __shared__ u32 s_scan[NL_T / 32 + 1];
    const u64 off = ((u64)blockIdx.x * NL_T + threadIdx.x) * 16;
    u32 m = off < n ? nl_mask16(ld16_bounded(text, off, n)) : 0;
    u32 tot;
    const u32 ex = block_excl_scan<NL_T>(__popc(m), s_scan, &tot);
    u64 o = blk_off[blockIdx.x] + ex;
    while (m) { const u32 q = __ffs(m) - 1; m &= m - 1; if (o < cap) nl_pos[o] = (u32)(off + q); ++o; }
}

struct NameSlot { unsigned long long key; u32 id; u32 len; char name[48]; };   // open addressing by FNV-1a of the name
__global__ void __launch_bounds__(256) k_pairs_parse(const char *text, const u32 *nl_pos, const unsigned long long *n_lines_p,
                                                     const NameSlot *tab, u32 mask, mk_pair *out, u64 cap, unsigned long long *bad) {
    const u64 n_lines = *n_lines_p < cap ? *n_lines_p : cap;
    u32 nbad = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_lines; i += (u64)gridDim.x * blockDim.x) {
        u32 p = i ? nl_pos[i - 1] + 1 : 0;
        const u32 e = nl_pos[i];
        mk_pair r; r.pos1 = r.pos2 = 0; r.chr1 = r.chr2 = 0xFFFF; r.strands = 0; r.cls = 0; r.lane = 0;
        bool ok = p < e && text[p] != '#';
        u32 pos[2] = {0, 0}, chr[2] = {0xFFFF, 0xFFFF};
        if (ok) {
            while (p < e && text[p] != '\t') ++p;                      // read id
            for (int k = 0; k < 2 && ok; ++k) {
                ok = p < e; ++p;                                         // the tab
                const u32 s0 = p;
                u64 h = 0xCBF29CE484222325ull;
                while (p < e && text[p] != '\t') { h = (h ^ (u64)(unsigned char)text[p]) * 0x100000001B3ull; ++p; }
                const u32 len = p - s0;
                if (h == 0) h = 0x9E3779B97F4A7C15ull;
                ok = ok && len > 0 && len <= 48;
                if (ok) {
                    u32 s = (u32)(h ^ (h >> 29)) & mask; bool found = false;
                    for (u32 probe = 0; probe <= mask; ++probe, s = (s + 1) & mask) {
                        const NameSlot *sl = &tab[s];
                        if (sl->key == 0) break;
                        if (sl->key == h && sl->len == len) {
                            bool same = true;
                            for (u32 b = 0; b < len && same; ++b) same = sl->name[b] == text[s0 + b];
                            if (same) { chr[k] = sl->id; found = true; break; }
                        }
                    }
                    ok = found;
                }
                ok = ok && p < e; ++p;                                   // the tab behind the name
                u32 v = 0, nd = 0;
                while (p < e && (u32)(text[p] - '0') <= 9u) { v = v * 10u + (u32)(text[p] - '0'); ++p; ++nd; }
                ok = ok && nd > 0 && nd <= 10;
                pos[k] = v;
            }
            // \t s1 \t s2
            ok = ok && p + 3 < e + 1 && text[p] == '\t' && text[p + 2] == '\t';
            if (ok) r.strands = (u8)((text[p + 1] == '-' ? 1 : 0) | (text[p + 3] == '-' ? 2 : 0));
        }
        if (ok) {
            r.pos1 = pos[0]; r.pos2 = pos[1]; r.chr1 = (u16)chr[0]; r.chr2 = (u16)chr[1];
            if (r.chr1 != r.chr2) r.cls = 0; else { const u32 d = r.pos2 - r.pos1; r.cls = d >= 10000u ? 1 : (d >= 1000u ? 2 : 3); }
        } else ++nbad;
        out[i] = r;
    }
    nbad = __reduce_add_sync(0xffffffffu, nbad);
    if ((threadIdx.x & 31) == 0 && nbad) atomicAdd(bad, (unsigned long long)nbad);
}

// d_text: n_bytes of .pairs text on the device ending with '\n' (16-byte aligned).  d_out[i] = pair of line i; *n_lines lines,
// *n_skipped of them not pairs of known chromosomes (headers included; chromosome id 0xFFFF).
extern "C" int mk_pairs_parse_text_device(mk_pairs_ws *w, const char *d_text, size_t n_bytes, const char *const *chrom_names, int n_chrom,
                                          mk_pair *d_out, size_t cap, size_t *n_lines, size_t *n_skipped, void *stream) {
    if (!w || !n_lines || !chrom_names || n_chrom <= 0 || n_chrom > 16384 || (n_bytes && (!d_text || !d_out))) { mk_set_error("mk_pairs_parse_text_device: bad argument"); return MK_ERR_ARG; }
    if (((uintptr_t)d_text & 15) != 0) { mk_set_error("mk_pairs_parse_text_device: d_text must be 16-byte aligned"); return MK_ERR_ARG; }
    if (n_bytes >= (1ull << 32)) { mk_set_error("mk_pairs_parse_text_device: at most 4 GiB of text per call"); return MK_ERR_CAPACITY; }
    MK_CUDA(cudaSetDevice(w->device));
    cudaStream_t s = (cudaStream_t)stream;
    *n_lines = 0; if (n_skipped) *n_skipped = 0;
    if (n_bytes == 0) return MK_OK;
    // chromosome table (rebuilt per call: a few KB)
    u32 slots = 64; while (slots < (u32)n_chrom * 4) slots <<= 1;
    std::vector<NameSlot> tab(slots);
    memset(tab.data(), 0, tab.size() * sizeof(NameSlot));
    for (int i = 0; i < n_chrom; ++i) {
        const char *nm = chrom_names[i]; const size_t len = nm ? strlen(nm) : 0;
        if (len == 0 || len > 48) { mk_set_error("mk_pairs_parse_text_device: chromosome name %d is empty or longer than 48 bytes", i); return MK_ERR_ARG; }
        u64 h = 0xCBF29CE484222325ull;
        for (size_t b = 0; b < len; ++b) h = (h ^ (u64)(unsigned char)nm[b]) * 0x100000001B3ull;
        if (h == 0) h = 0x9E3779B97F4A7C15ull;
        u32 sl = (u32)(h ^ (h >> 29)) & (slots - 1);
        while (tab[sl].key) sl = (sl + 1) & (slots - 1);
        tab[sl].key = h; tab[sl].id = (u32)i; tab[sl].len = (u32)len; memcpy(tab[sl].name, nm, len);
    }
    const u64 n_blk = (n_bytes + NL_T * 16 - 1) / (NL_T * 16);
    MK_TRY(w->text_scratch(n_blk));
    if (w->names.n < tab.size() * sizeof(NameSlot)) MK_TRY(w->names.alloc(tab.size() * sizeof(NameSlot)));
    const size_t nl_cap = cap + 1;
    if (w->nl.n < nl_cap * 4) MK_TRY(w->nl.alloc(nl_cap * 4));
    MK_CUDA(cudaMemcpyAsync(w->names.p, tab.data(), tab.size() * sizeof(NameSlot), cudaMemcpyHostToDevice, s));
    MK_CUDA(cudaStreamSynchronize(s));                                  // `tab` is a local
    MK_CUDA(cudaMemsetAsync(w->counter.p, 0, 64, s));
    unsigned long long *cnt = w->counter.as<unsigned long long>();
    k_nl_count<<<(unsigned)n_blk, NL_T, 0, s>>>(d_text, n_bytes, w->tile_sum.as<u32>());
    k_line_tile_prefix<<<1, 1024, 0, s>>>(w->tile_sum.as<u32>(), w->tile_off.as<u64>(), n_blk, cnt);      // total -> cnt[1]
    k_nl_write<<<(unsigned)n_blk, NL_T, 0, s>>>(d_text, n_bytes, w->tile_off.as<u64>(), w->nl.as<u32>(), nl_cap);
    k_pairs_parse<<<w->sms * 8, 256, 0, s>>>(d_text, w->nl.as<u32>(), cnt + 1, w->names.as<NameSlot>(), slots - 1, d_out, cap, cnt + 2);
    w->launches += 4;
    unsigned long long h[3];
    MK_CUDA(cudaMemcpyAsync(h, cnt, 24, cudaMemcpyDeviceToHost, s));
    MK_CUDA(cudaStreamSynchronize(s));
    MK_CUDA(cudaGetLastError());
    if (h[1] > cap) { mk_set_error("mk_pairs_parse_text_device: %llu lines, output capacity %zu", h[1], cap); return MK_ERR_CAPACITY; }
    *n_lines = (size_t)h[1];
    if (n_skipped) *n_skipped = (size_t)h[2];
    return MK_OK;
}
