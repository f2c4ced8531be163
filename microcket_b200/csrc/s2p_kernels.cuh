// s2p_kernels.cuh — sm_100a kernels of the sam2pairs path.
//
// One *window* of SAM text (<= 1 GiB, device resident) goes through
//   k_win_begin   set up the window from the device-side cursor, clear tile descriptors
//   k_scan_lines  K1: byte-parallel newline index, single pass, decoupled look-back    (replaces getline, pairutil.h:152)
//   k_parse       K2: one thread per line: first six fields, filter, CIGAR walk         (pairutil.h:63-126,155-161; unc2pairs.h:34-36)
//   k_group       K3: group heads among kept records + per-group resolution             (pairutil.h:163-173; flash2pairs.h; unc2pairs.h)
//   k_emit        K4: look-back scan of output sizes, .pairs text + packed records      (unc2pairs.h:310-348)
//   k_copy_sam    K5: SAM passthrough of the kept lines of emitted groups               (unc2pairs.h:351-356)
//   k_win_end     advance the cursor to the window's last (unprocessed) read group
// All window geometry lives in device memory (WinState), so consecutive windows are
// enqueued back to back without a host round trip.
#pragma once
#include "mk_common.cuh"

#define S2P_TILE_BYTES 32768
#define S2P_SCAN_THREADS 256
#define S2P_NAME_MAX 42

enum { ST_NONE = 0, ST_LOWMAP, ST_MANYHITS, ST_UNPAIRED, ST_SELFCIRCLE, ST_TRANS, ST_CIS10K, ST_CIS1K, ST_CIS0, ST_CIGARERR, ST_NCOUNTER };

// line meta bits
#define LM_KEEP 1u
#define LM_EQ 2u
#define LM_HEAD 4u
#define LM_PROC 8u
#define LM_EMIT 16u

struct __align__(16) LineRec {       // 48 B, written by K2 for kept lines
    u32 pos, right0, left1, right1;
    u32 leftClip, rightClip, mappable, line_len;
    u16 flag, qname_len, chr_slot; u8 segCnt /* 0 = cigar error, 3 = more than 2 */, pad0;
    u32 qname_off;                   // offset of the QNAME's first byte from the line start
    u32 pad1;
};

struct __align__(16) GroupRes {      // 32 B, written by K3 for group heads
    u32 posA, posB;
    u16 chrA, chrB;                  // chromosome ids
    u8 status, strands; u16 n_kept;
    u32 rid_line, last_line, text_len, sam_len;
};

struct __align__(64) ChrSlot {       // open-addressing table keyed by a 64-bit hash of the name
    unsigned long long key;          // 0 = empty
    unsigned long long name8;        // first 8 bytes of the name, zero padded (fast exact check for short names)
    int id;                          // -1 until published
    u16 len; char name[S2P_NAME_MAX];
};

struct WinState {
    // stream / buffer geometry
    u64 cursor, total;               // next window start; bytes available
    u64 ws, we;                      // this window
    u32 n_lines, carry_line, first_tile, is_last;
    u32 err, pad;
    // per-window emit totals
    u32 w_groups, w_emit, w_text, w_sam;
    // running output offsets (device-resident multi-window runs)
    u64 out_text, out_pairs, out_sam;
    // stream totals
    u64 groups_done, lines_done;
    unsigned long long counters[ST_NCOUNTER];
    u32 sc_count, n_chrom;
};

#define S2P_ERR_LINES 1u
#define S2P_ERR_TEXT 2u
#define S2P_ERR_PAIRS 4u
#define S2P_ERR_SAM 8u
#define S2P_ERR_SCLIST 16u
#define S2P_ERR_NOPROGRESS 32u
#define S2P_ERR_CHRTABLE 64u

struct S2PParams {
    const char *buf;          // SAM text base (16-byte aligned)
    WinState *st;
    u32 *nl_pos;              // newline offsets relative to ws
    u8 *lmeta;
    LineRec *rec;
    GroupRes *res;
    u32 *sam_dst;
    u64 *desc_scan, *desc_emitA, *desc_emitB;
    ChrSlot *chr; u32 chr_mask; int *id_to_slot; u32 chr_cap;
    u64 *sc_list; u32 sc_cap;
    char *out_text; u64 out_text_cap;
    mk_pair *out_pairs; u64 out_pairs_cap;
    char *out_sam; u64 out_sam_cap;
    u64 window_bytes; u32 cap_lines;
    int mode, min_mapq, write_sam, emit_text, emit_packed; float ratio; u16 lane;
    int running_offsets;      // 1: append at st->out_* (device-resident runs); 0: every window writes at 0
};

// ------------------------------------------------------------------------------------------------ begin / end
static __global__ void k_win_begin(S2PParams p, u32 n_desc) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_desc) { p.desc_scan[i] = 0; p.desc_emitA[i] = 0; p.desc_emitB[i] = 0; }
    if (i == 0) {
        WinState *s = p.st;
        s->ws = s->cursor;
        u64 we = s->cursor + p.window_bytes;
        s->we = we < s->total ? we : s->total;
        s->n_lines = 0; s->carry_line = 0xFFFFFFFFu;
        s->first_tile = (u32)(s->ws / S2P_TILE_BYTES);
        s->w_groups = s->w_emit = s->w_text = s->w_sam = 0;
        if (!p.running_offsets) { s->out_text = s->out_pairs = s->out_sam = 0; s->sc_count = 0; }
    }
}

static __global__ void k_win_end(S2PParams p) {
    WinState *s = p.st;
    u64 ws = s->ws, we = s->we;
    u32 n = s->n_lines;
    u64 next;
    if (s->carry_line != 0xFFFFFFFFu) {
        u32 c = s->carry_line;
        next = ws + (c ? (u64)p.nl_pos[c - 1] + 1 : 0);
    } else {
        next = ws + (n ? (u64)p.nl_pos[n - 1] + 1 : 0);   // no kept record: everything up to the last complete line is consumed
    }
    bool final_win = (we == s->total) && s->is_last;
    if (final_win) next = s->total;                        // the stream's last group is never processed (pairutil.h:176)
    else if (next == ws && we > ws && we - ws >= p.window_bytes) s->err |= S2P_ERR_NOPROGRESS;  // one group (or line) fills the window
    s->cursor = next;
    s->lines_done += (s->carry_line != 0xFFFFFFFFu && !final_win) ? s->carry_line : n;
    s->groups_done += s->w_groups;
    s->out_text += s->w_text; s->out_pairs += s->w_emit; s->out_sam += s->w_sam;
}

// ------------------------------------------------------------------------------------------------ K1: newline index
// Tile = 32 KiB at absolute 32 KiB boundaries of the buffer.  Loads are coalesced 128-bit streaming
// loads; the 16 newline flags of every 16-byte word go through shared memory so that each thread then
// owns 128 CONTIGUOUS bytes (8 words), which makes ranks a single block scan.
// Shared by the SAM and FASTQ paths: positions (relative to ws) of every '\n' in [ws, we).
__device__ __forceinline__ void scan_lines_body(const char *buf, const u64 ws, const u64 we, const int first_tile,
                                                u32 *nl_pos, const u32 cap_lines, u64 *desc, u32 *n_lines_out, u32 *err_out, u32 err_bit) {
    __shared__ __align__(16) u16 s_mask[S2P_TILE_BYTES / 16];
    __shared__ u32 s_scan[S2P_SCAN_THREADS / 32 + 1];
    __shared__ u32 s_base;
    if (we <= ws) return;
    const int last_tile = (int)((we - 1) / S2P_TILE_BYTES);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int tile = first_tile + blockIdx.x; tile <= last_tile; tile += gridDim.x) {
        const u64 tbase = (u64)tile * S2P_TILE_BYTES;
        const uint4 *src = (const uint4 *)(buf + tbase);
        uint4 w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            u64 off = tbase + ((u64)(j * S2P_SCAN_THREADS + tid) << 4);
            w[j] = (off < we && off + 16 > ws) ? ld_stream_v4(src + j * S2P_SCAN_THREADS + tid) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            u32 m = gather_flags4(byte_eq_mask(w[j].x, 0x0A0A0A0Au)) | (gather_flags4(byte_eq_mask(w[j].y, 0x0A0A0A0Au)) << 4) |
                    (gather_flags4(byte_eq_mask(w[j].z, 0x0A0A0A0Au)) << 8) | (gather_flags4(byte_eq_mask(w[j].w, 0x0A0A0A0Au)) << 12);
            u64 off = tbase + ((u64)(j * S2P_SCAN_THREADS + tid) << 4);
            // partial words at the window edges
            if (off < ws) { u64 d = ws - off; m = d >= 16 ? 0 : (m >> d) << d; }
            if (off + 16 > we) { u64 keep = we > off ? we - off : 0; m = keep >= 16 ? m : (m & ((1u << keep) - 1)); }
            s_mask[j * S2P_SCAN_THREADS + tid] = (u16)m;
        }
        __syncthreads();
        uint4 mm = ((const uint4 *)s_mask)[tid];          // flags of bytes [tid*128, tid*128+128)
        u32 cnt = __popc(mm.x) + __popc(mm.y) + __popc(mm.z) + __popc(mm.w);
        u32 total;
        u32 excl = block_excl_scan<S2P_SCAN_THREADS>(cnt, s_scan, &total);
        if (wid == 0) {
            u64 b = lookback_exclusive(desc - first_tile, tile, first_tile, total, lane);
            if (lane == 0) {
                s_base = (u32)b;
                if (tile == last_tile) {
                    u64 nl = b + total;
                    if (nl > cap_lines) { atomicOr(err_out, err_bit); nl = cap_lines; }
                    *n_lines_out = (u32)nl;
                }
            }
        }
        __syncthreads();
        u32 idx = s_base + excl;
        u32 rel = (u32)(tbase + (u64)tid * 128 - ws);     // may wrap for bytes before ws: those have no flags
        u32 parts[4] = {mm.x, mm.y, mm.z, mm.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            u32 m = parts[q];
            while (m) {
                int b = __ffs(m) - 1; m &= m - 1;
                if (idx < cap_lines) nl_pos[idx] = rel + q * 32 + b;
                ++idx;
            }
        }
        __syncthreads();
    }
}

static __global__ void __launch_bounds__(S2P_SCAN_THREADS) k_scan_lines(S2PParams p) {
    WinState *st = p.st;
    scan_lines_body(p.buf, st->ws, st->we, (int)st->first_tile, p.nl_pos, p.cap_lines, p.desc_scan, &st->n_lines, &st->err, S2P_ERR_LINES);
}

// ------------------------------------------------------------------------------------------------ chromosome table
__device__ __forceinline__ u64 hash_step(u64 h, int c) { return (h ^ (u64)c) * 0x100000001B3ull; }

// Returns the slot of the name; inserts it when unseen (lock-free; ids are published by the inserter).
static __device__ int chr_lookup_insert(const S2PParams &p, u64 h, u64 name8, const char *buf, u64 name_off, u32 len) {
    if (h == 0) h = 0x9E3779B97F4A7C15ull;
    const u32 l = len < S2P_NAME_MAX ? len : S2P_NAME_MAX;
    u32 s = (u32)(h ^ (h >> 29)) & p.chr_mask;
    for (u32 probe = 0; probe <= p.chr_mask; ++probe, s = (s + 1) & p.chr_mask) {
        ChrSlot *sl = &p.chr[s];
        unsigned long long k = *(volatile unsigned long long *)&sl->key;
        if (k == 0) {
            unsigned long long old = atomicCAS(&sl->key, 0ull, (unsigned long long)h);
            if (old == 0) {
                sl->len = (u16)l; sl->name8 = name8;
                for (u32 i = 0; i < l; ++i) sl->name[i] = buf[name_off + i];
                int id = (int)atomicAdd(&p.st->n_chrom, 1u);
                if ((u32)id < p.chr_cap) p.id_to_slot[id] = (int)s; else atomicOr(&p.st->err, S2P_ERR_CHRTABLE);
                __threadfence();
                *(volatile int *)&sl->id = id;
                return (int)s;
            }
            k = old;
        }
        if (k == h) {
            int id = *(volatile int *)&sl->id;
            if (id < 0) return (int)s;                       // being inserted by a concurrent thread: identity by hash
            __threadfence();
            bool same = sl->len == l && sl->name8 == name8;
            if (same && l > 8) for (u32 i = 8; same && i < l; ++i) same = sl->name[i] == buf[name_off + i];
            if (same) return (int)s;
        }
    }
    atomicOr(&p.st->err, S2P_ERR_CHRTABLE);
    return 0;
}

// ------------------------------------------------------------------------------------------------ K2: parse
__device__ __forceinline__ u32 parse_uint_tok(ByteReader &r, int &c) {
    while (is_blank(c)) c = r.next();
    u32 v = 0; bool bad = false; int n = 0;
    while (!is_ws(c)) { u32 d = (u32)(c - '0'); bad |= d > 9u; v = v * 10u + d; ++n; c = r.next(); }
    return (bad || n == 0) ? 0u : v;
}

static __global__ void __launch_bounds__(256) k_parse(S2PParams p) {
    const WinState *st = p.st;
    const u32 n_lines = st->n_lines;
    const u64 ws = st->ws;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n_lines; i += gridDim.x * blockDim.x) {
        const u32 start = i ? p.nl_pos[i - 1] + 1 : 0;
        const u32 end = p.nl_pos[i];
        if (p.write_sam) p.sam_dst[i] = 0xFFFFFFFFu;
        ByteReader r; r.init(p.buf, ws + start);
        int c = r.next();
        if (c == '@') { p.lmeta[i] = 0; continue; }          // header line (QNAME cannot contain '@')
        while (is_blank(c)) c = r.next();
        const u32 qoff = (u32)(r.pos - 1 - (ws + start));
        u32 meta = 0;
        // QNAME, compared on the fly with the previous line's first token
        if (i > 0) {
            const u32 pstart = i > 1 ? p.nl_pos[i - 2] + 1 : 0;
            ByteReader q; q.init(p.buf, ws + pstart);
            int d = q.next();
            while (is_blank(d)) d = q.next();
            bool eq = true;
            while (!is_ws(c)) { eq &= (c == d); if (!is_ws(d)) d = q.next(); c = r.next(); }
            eq &= is_ws(d);
            if (eq) meta |= LM_EQ;
        } else {
            while (!is_ws(c)) c = r.next();
        }
        const u32 qlen = (u32)(r.pos - 1 - (ws + start)) - qoff;
        const u32 flag = parse_uint_tok(r, c);
        // RNAME
        while (is_blank(c)) c = r.next();
        const u64 name_off = r.pos - 1;
        u64 h = 0xCBF29CE484222325ull, name8 = 0;
        u32 name_len = 0;
        while (!is_ws(c)) { h = hash_step(h, c); if (name_len < 8) name8 |= (u64)c << (8 * name_len); ++name_len; c = r.next(); }
        const u32 pos = parse_uint_tok(r, c);
        const u32 mapq = parse_uint_tok(r, c);
        if (mapq < (u32)p.min_mapq || (flag & 0x700u)) { p.lmeta[i] = (u8)meta; continue; }   // pairutil.h:157-161
        meta |= LM_KEEP;
        // CIGAR walk (pairutil.h:63-126)
        while (is_blank(c)) c = r.next();
        u32 val = 0, idx = 0, leftClip = 0, rightClip = 0, mappable = 0;
        u32 cur = pos, right0 = 0, left1 = 0, right1 = 0, last_right = 0;
        bool err = false;
        while (!is_ws(c)) {
            u32 d = (u32)(c - '0');
            if (d <= 9u) { val = val * 10u + d; c = r.next(); continue; }
            const int nxt = r.next();                       // one byte of look-ahead: is this op the last character?
            if (c == 'H' || c == 'S') {
                if (is_ws(nxt)) rightClip = val;
                else if (idx == 0) leftClip = val;          // overwrites: 5H30S100M leaves leftClip = 30
                else err = true;
            } else if (c == 'M' || c == 'D') {
                if (c == 'M') mappable += val;
                cur += val; last_right = cur - 1;
                if (idx == 0) right0 = last_right; else if (idx == 1) right1 = last_right;
            } else if (c == 'N') {
                cur += val; ++idx; last_right = 0;
                if (idx == 1) left1 = cur;
            } else if (c != 'I') err = true;
            val = 0; c = nxt;
        }
        LineRec rec;
        rec.pos = pos; rec.right0 = right0; rec.left1 = left1; rec.right1 = right1;
        rec.leftClip = leftClip; rec.rightClip = rightClip; rec.mappable = mappable; rec.line_len = end - start;
        rec.flag = (u16)flag; rec.qname_len = (u16)qlen;
        rec.chr_slot = (u16)chr_lookup_insert(p, h, name8, p.buf, name_off, name_len);
        u32 segCnt = idx + 1;
        if (last_right == 0) err = true;                             // pairutil.h:119
        rec.segCnt = err ? 0 : (u8)(segCnt > 2 ? 3 : segCnt);
        rec.pad0 = 0; rec.qname_off = qoff; rec.pad1 = 0;
        p.rec[i] = rec;
        p.lmeta[i] = (u8)meta;
    }
}

// ------------------------------------------------------------------------------------------------ K3: groups
struct Seg { u32 pos, right0, left1, right1, leftClip, rightClip, mappable; int segCnt; bool minus; u16 chr; };

__device__ __forceinline__ Seg seg_of(const LineRec &r) {
    Seg s; s.pos = r.pos; s.right0 = r.right0; s.left1 = r.left1; s.right1 = r.right1;
    s.leftClip = r.leftClip; s.rightClip = r.rightClip; s.mappable = r.mappable; s.segCnt = r.segCnt;
    s.minus = (r.flag & 16u) != 0; s.chr = r.chr_slot;
    return s;
}
// pairutil.h:180-188.  fp32 on purpose: __fmul_rn keeps the product un-fused and un-contracted.
__device__ __forceinline__ bool integrity1(const Seg &s, float ratio) {
    int total = (int)s.mappable;
    if ((int)s.leftClip > 20) total += (int)s.leftClip;
    if ((int)s.rightClip > 20) total += (int)s.rightClip;
    return (float)(int)s.mappable >= __fmul_rn((float)total, ratio);
}
// pairutil.h:190-208, including the s1.rightClip test of line 200
__device__ __forceinline__ bool integrity2(const Seg &a, const Seg &b, float ratio) {
    int ta = (int)a.mappable, tb = (int)b.mappable;
    if ((int)a.leftClip > 20) ta += (int)a.leftClip;
    if ((int)a.rightClip > 20) ta += (int)a.rightClip;
    if ((int)b.leftClip > 20) tb += (int)b.leftClip;
    if ((int)a.rightClip > 20) tb += (int)b.rightClip;
    int big = ta > tb ? ta : tb;
    return (float)(int)(a.mappable + b.mappable) >= __fmul_rn((float)big, ratio);
}
// the mate test of unc2pairs.h:191-308 (positions compared as signed int like the reference)
__device__ __forceinline__ bool mates(const Seg &lone, const Seg &c) {
    if (lone.chr != c.chr) return false;
    if (!lone.minus) return c.minus && (int)lone.pos < (int)c.pos && (int)c.right0 - (int)lone.pos <= 1000;
    return !c.minus && (int)c.pos < (int)lone.pos && (int)lone.right0 - (int)c.pos <= 1000;
}
__device__ __forceinline__ u32 distal_end(const Seg &s) { return (int)s.leftClip > (int)s.rightClip ? s.right0 : s.pos; }

static __device__ bool qname_equal_slow(const S2PParams &p, u64 ws, u32 a, u32 b) {
    ByteReader x, y;
    x.init(p.buf, ws + (a ? p.nl_pos[a - 1] + 1 : 0));
    y.init(p.buf, ws + (b ? p.nl_pos[b - 1] + 1 : 0));
    int c = x.next(), d = y.next();
    while (is_blank(c)) c = x.next();
    while (is_blank(d)) d = y.next();
    while (!is_ws(c) && !is_ws(d)) { if (c != d) return false; c = x.next(); d = y.next(); }
    return is_ws(c) && is_ws(d);
}

// bytewise order of two chromosome names (std::string::compare)
static __device__ int chr_name_cmp(const S2PParams &p, u16 sa, u16 sb) {
    if (sa == sb) return 0;
    const ChrSlot *a = &p.chr[sa], *b = &p.chr[sb];
    u32 la = a->len, lb = b->len, m = la < lb ? la : lb;
    for (u32 i = 0; i < m; ++i) {
        int d = (int)(u8)a->name[i] - (int)(u8)b->name[i];
        if (d) return d;
    }
    return la < lb ? -1 : (la > lb ? 1 : 0);
}

__device__ __forceinline__ u32 dec_digits(u32 v) {
    return v >= 1000000000u ? 10 : v >= 100000000u ? 9 : v >= 10000000u ? 8 : v >= 1000000u ? 7 : v >= 100000u ? 6 :
           v >= 10000u ? 5 : v >= 1000u ? 4 : v >= 100u ? 3 : v >= 10u ? 2 : 1;
}

static __global__ void __launch_bounds__(256) k_group(S2PParams p) {
    __shared__ u32 s_cnt[ST_NCOUNTER];
    if (threadIdx.x < ST_NCOUNTER) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    WinState *st = p.st;
    const u32 n_lines = st->n_lines;
    const u64 ws = st->ws;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n_lines; i += gridDim.x * blockDim.x) {
        const u32 mi = p.lmeta[i];
        if (!(mi & LM_KEEP)) continue;
        // ---- is this kept line the head of a group?  (pairutil.h:163-173: currId != lastId among kept records)
        bool head;
        {
            bool chain = (mi & LM_EQ) != 0;
            long j = (long)i - 1;
            while (j >= 0) { u32 mj = p.lmeta[j]; if (mj & LM_KEEP) break; chain = chain && (mj & LM_EQ); --j; }
            if (j < 0) head = true;
            else if (chain) head = false;
            else if (j == (long)i - 1) head = true;
            else head = !qname_equal_slow(p, ws, i, (u32)j);
        }
        if (!head) continue;
        // ---- collect the group's kept records
        u32 first[2] = {i, 0}, r1[2] = {0, 0}, r2[2] = {0, 0};
        u32 n = 0, n1 = 0, n2 = 0, sam_len = 0;
        u32 k = i, prev = i;
        bool chain = true, off_end = false;
        while (true) {
            // k is a member
            const LineRec *rk = &p.rec[k];
            u32 fl = rk->flag;
            if (n < 2) first[n] = k;
            ++n;
            if (fl & 64u) { if (n1 < 2) r1[n1] = k; ++n1; } else if (fl & 128u) { if (n2 < 2) r2[n2] = k; ++n2; }
            sam_len += rk->line_len + 1;
            prev = k;
            // next kept line
            u32 q = k + 1; chain = true;
            while (q < n_lines) { u32 mq = p.lmeta[q]; chain = chain && (mq & LM_EQ); if (mq & LM_KEEP) break; ++q; }
            if (q >= n_lines) { off_end = true; break; }
            bool same = chain ? true : (q == prev + 1 ? false : qname_equal_slow(p, ws, q, prev));
            if (!same) break;
            k = q;
        }
        if (off_end) {                       // the window's last group: carried to the next window (or dropped at EOF)
            st->carry_line = i;
            p.lmeta[i] = (u8)(mi | LM_HEAD);
            continue;
        }
        // ---- resolve
        GroupRes g; g.status = ST_NONE; g.posA = g.posB = 0; g.chrA = g.chrB = 0; g.strands = 0;
        g.n_kept = (u16)(n > 65535u ? 65535u : n); g.rid_line = prev; g.last_line = prev; g.text_len = 0; g.sam_len = sam_len;
        u32 p1 = 0, p2 = 0; u16 c1 = 0, c2 = 0; bool m1 = false, m2 = false; bool have = false, ordered = false;
        const float ratio = p.ratio;
        if (p.mode == 0) {                   // flash2pairs.h:25-154
            if (n == 1) {
                Seg a = seg_of(p.rec[first[0]]);
                if (a.segCnt == 0) g.status = ST_CIGARERR;
                else if (a.segCnt > 2) g.status = ST_MANYHITS;
                else if (!integrity1(a, ratio)) g.status = ST_LOWMAP;
                else {
                    p1 = a.pos; p2 = a.segCnt == 2 ? a.right1 : a.right0;
                    u32 d = p2 - p1;
                    g.status = d >= 10000u ? ST_CIS10K : (d >= 1000u ? ST_CIS1K : ST_CIS0);
                    c1 = c2 = a.chr; m1 = false; m2 = true; have = true; ordered = true;   // always "+ -", never swapped
                }
            } else if (n == 2) {
                Seg a = seg_of(p.rec[first[0]]), b = seg_of(p.rec[first[1]]);
                if (a.segCnt == 0 || b.segCnt == 0) g.status = ST_CIGARERR;
                else if (a.segCnt != 1 || b.segCnt != 1) g.status = ST_MANYHITS;
                else if (!integrity2(a, b, ratio)) g.status = ST_LOWMAP;
                else {
                    p1 = (int)a.leftClip > (int)a.rightClip ? a.right0 : a.pos;
                    p2 = (int)b.leftClip > (int)b.rightClip ? b.right0 : b.pos;
                    c1 = a.chr; c2 = b.chr; m1 = a.minus; m2 = b.minus; have = true;
                }
            } else g.status = ST_MANYHITS;
        } else {                             // unc2pairs.h:29-358
            if (n1 == 0 || n2 == 0 || n1 + n2 > 3) g.status = ST_NONE;        // silent drops, unc2pairs.h:52-59
            else if (n1 == 1 && n2 == 1) {
                Seg a = seg_of(p.rec[r1[0]]), b = seg_of(p.rec[r2[0]]);
                if (a.segCnt == 0) g.status = ST_CIGARERR;
                else if (!integrity1(a, ratio)) g.status = ST_LOWMAP;
                else if (b.segCnt == 0) g.status = ST_CIGARERR;
                else if (!integrity1(b, ratio)) g.status = ST_LOWMAP;
                else if (a.segCnt + b.segCnt > 3) g.status = ST_MANYHITS;
                else {
                    c1 = a.chr; c2 = b.chr; m1 = a.minus; m2 = b.minus; have = true;
                    if (a.segCnt == 1 && b.segCnt == 1) {
                        p1 = a.minus ? a.right0 : a.pos; p2 = b.minus ? b.right0 : b.pos;
                    } else if (a.segCnt == 2) {                                   // unc2pairs.h:146-167
                        if (!a.minus) {
                            if (b.minus && a.chr == b.chr && (int)a.left1 < (int)b.pos && (int)b.right0 - (int)a.left1 <= 1000) { p1 = a.pos; p2 = b.right0; }
                            else { g.status = ST_UNPAIRED; have = false; }
                        } else {
                            if (!b.minus && a.chr == b.chr && (int)b.pos < (int)a.pos && (int)a.right0 - (int)b.pos <= 1000) { p1 = a.right1; p2 = b.pos; }
                            else { g.status = ST_UNPAIRED; have = false; }
                        }
                    } else {                                                      // unc2pairs.h:168-189
                        if (!a.minus) {
                            if (b.minus && a.chr == b.chr && (int)a.pos < (int)b.pos && (int)b.right0 - (int)a.pos <= 1000) { p1 = a.pos; p2 = b.right1; }
                            else { g.status = ST_UNPAIRED; have = false; }
                        } else {
                            if (!b.minus && a.chr == b.chr && (int)b.left1 < (int)a.pos && (int)a.right0 - (int)b.left1 <= 1000) { p1 = a.right0; p2 = b.pos; }
                            else { g.status = ST_UNPAIRED; have = false; }
                        }
                    }
                }
            } else if (n1 == 1) {            // 1 + 2
                Seg a = seg_of(p.rec[r1[0]]), b = seg_of(p.rec[r2[0]]), c = seg_of(p.rec[r2[1]]);
                if (a.segCnt == 0) g.status = ST_CIGARERR;
                else if (!integrity1(a, ratio)) g.status = ST_LOWMAP;
                else if (b.segCnt == 0 || c.segCnt == 0) g.status = ST_CIGARERR;
                else if (!integrity2(b, c, ratio)) g.status = ST_LOWMAP;
                else if (a.segCnt != 1 || b.segCnt != 1 || c.segCnt != 1) g.status = ST_MANYHITS;
                else {
                    c1 = a.chr; m1 = a.minus; p1 = a.minus ? a.right0 : a.pos;
                    if (mates(a, b)) { c2 = c.chr; m2 = c.minus; p2 = distal_end(c); have = true; }
                    else if (mates(a, c)) { c2 = b.chr; m2 = b.minus; p2 = distal_end(b); have = true; }
                    else g.status = ST_UNPAIRED;
                }
            } else {                         // 2 + 1
                Seg a = seg_of(p.rec[r1[0]]), b = seg_of(p.rec[r1[1]]), c = seg_of(p.rec[r2[0]]);
                if (a.segCnt == 0 || b.segCnt == 0) g.status = ST_CIGARERR;
                else if (!integrity2(a, b, ratio)) g.status = ST_LOWMAP;
                else if (c.segCnt == 0) g.status = ST_CIGARERR;
                else if (!integrity1(c, ratio)) g.status = ST_LOWMAP;
                else if (a.segCnt != 1 || b.segCnt != 1 || c.segCnt != 1) g.status = ST_MANYHITS;
                else {
                    c2 = c.chr; m2 = c.minus; p2 = c.minus ? c.right0 : c.pos;
                    if (mates(c, a)) { c1 = b.chr; m1 = b.minus; p1 = distal_end(b); have = true; }
                    else if (mates(c, b)) { c1 = a.chr; m1 = a.minus; p1 = distal_end(a); have = true; }
                    else g.status = ST_UNPAIRED;
                }
            }
        }
        u32 meta = mi | LM_HEAD | LM_PROC;
        if (have) {
            u16 sA = c1, sB = c2;
            if (!ordered) {                  // unc2pairs.h:310-348
                int cc = chr_name_cmp(p, c1, c2);
                if (!(cc < 0 || (cc == 0 && p1 < p2))) {
                    u32 t = p1; p1 = p2; p2 = t; sA = c2; sB = c1; bool tm = m1; m1 = m2; m2 = tm;
                }
                if (cc == 0) {
                    u32 d = p2 - p1;
                    g.status = d <= 10u ? ST_SELFCIRCLE : (d >= 10000u ? ST_CIS10K : (d >= 1000u ? ST_CIS1K : ST_CIS0));
                } else g.status = ST_TRANS;
            }
            g.posA = p1; g.posB = p2; g.strands = (u8)((m1 ? 1 : 0) | (m2 ? 2 : 0));
            const ChrSlot *ca = &p.chr[sA], *cb = &p.chr[sB];
            g.chrA = (u16)ca->id; g.chrB = (u16)cb->id;
            if (g.status != ST_SELFCIRCLE) {
                meta |= LM_EMIT;
                g.text_len = (u32)p.rec[prev].qname_len + ca->len + cb->len + dec_digits(p1) + dec_digits(p2) + 9u;
            }
        }
        if (!(meta & LM_EMIT)) g.sam_len = 0;
        if (g.status != ST_NONE) atomicAdd(&s_cnt[g.status], 1u);
        p.res[i] = g;
        p.lmeta[i] = (u8)meta;
    }
    __syncthreads();
    if (threadIdx.x < ST_NCOUNTER && s_cnt[threadIdx.x]) atomicAdd(&st->counters[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------ K4: emit
#define EMIT_THREADS 256
#define EMIT_STAGE 24576

__device__ __forceinline__ u32 put_uint(char *dst, u32 v) {
    u32 n = dec_digits(v);
    for (int i = (int)n - 1; i >= 0; --i) { dst[i] = (char)('0' + v % 10u); v /= 10u; }
    return n;
}

template <class Sink>
__device__ __forceinline__ void write_pair_line(const S2PParams &p, u64 ws, const GroupRes &g, Sink &out) {
    const u32 rl = g.rid_line;
    const LineRec *rr = &p.rec[rl];
    const u64 qsrc = ws + (rl ? p.nl_pos[rl - 1] + 1 : 0) + rr->qname_off;
    const u32 ql = rr->qname_len;
    for (u32 i = 0; i < ql; ++i) out.put(p.buf[qsrc + i]);
    out.put('\t');
    const ChrSlot *ca = &p.chr[p.id_to_slot[g.chrA]], *cb = &p.chr[p.id_to_slot[g.chrB]];
    for (u32 i = 0; i < ca->len; ++i) out.put(ca->name[i]);
    out.put('\t');
    char tmp[10]; u32 n = put_uint(tmp, g.posA);
    for (u32 i = 0; i < n; ++i) out.put(tmp[i]);
    out.put('\t');
    for (u32 i = 0; i < cb->len; ++i) out.put(cb->name[i]);
    out.put('\t');
    n = put_uint(tmp, g.posB);
    for (u32 i = 0; i < n; ++i) out.put(tmp[i]);
    out.put('\t'); out.put((g.strands & 1) ? '-' : '+'); out.put('\t'); out.put((g.strands & 2) ? '-' : '+'); out.put('\n');
}
struct PtrSink { char *p; __device__ __forceinline__ void put(char c) { *p++ = c; } };

static __global__ void __launch_bounds__(EMIT_THREADS) k_emit(S2PParams p) {
    __shared__ __align__(16) char s_stage[EMIT_STAGE + 16];
    __shared__ u32 s_scanA[EMIT_THREADS / 32 + 1], s_scanT[EMIT_THREADS / 32 + 1], s_scanS[EMIT_THREADS / 32 + 1];
    __shared__ u64 s_baseA, s_baseB;
    WinState *st = p.st;
    const u32 n_lines = st->n_lines;
    const u64 ws = st->ws;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n_tiles = (int)((n_lines + EMIT_THREADS - 1) / EMIT_THREADS);
    const u64 base_text = st->out_text, base_pairs = st->out_pairs, base_sam = st->out_sam, base_groups = st->groups_done;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const u32 i = (u32)tile * EMIT_THREADS + tid;
        u32 m = i < n_lines ? p.lmeta[i] : 0;
        const bool proc = (m & LM_HEAD) && (m & LM_PROC);
        const bool emit = proc && (m & LM_EMIT);
        GroupRes g;
        if (proc) g = p.res[i];
        u32 vA = (proc ? 1u : 0u) | (emit ? 1u << 16 : 0u);
        u32 vT = emit ? g.text_len : 0u, vS = (emit && p.write_sam) ? g.sam_len : 0u;
        u32 totA, totT, totS;
        u32 exA = block_excl_scan<EMIT_THREADS>(vA, s_scanA, &totA);
        u32 exT = block_excl_scan<EMIT_THREADS>(vT, s_scanT, &totT);
        u32 exS = block_excl_scan<EMIT_THREADS>(vS, s_scanS, &totS);
        if (wid == 0) {
            u64 agg = (u64)(totA & 0xFFFFu) | ((u64)(totA >> 16) << 31);
            u64 b = lookback_exclusive(p.desc_emitA, tile, 0, agg, lane);
            if (lane == 0) { s_baseA = b; if (tile == n_tiles - 1) { u64 t = b + agg; st->w_groups = (u32)(t & 0x7FFFFFFFu); st->w_emit = (u32)(t >> 31); } }
        } else if (wid == 1) {
            u64 agg = (u64)totT | ((u64)totS << 31);
            u64 b = lookback_exclusive(p.desc_emitB, tile, 0, agg, lane);
            if (lane == 0) { s_baseB = b; if (tile == n_tiles - 1) { u64 t = b + agg; st->w_text = (u32)(t & 0x7FFFFFFFu); st->w_sam = (u32)(t >> 31); } }
        }
        __syncthreads();
        const u64 bA = s_baseA, bB = s_baseB;
        const u32 g_idx = (u32)(bA & 0x7FFFFFFFu) + (exA & 0xFFFFu);          // processed-group index inside the window
        const u32 e_idx = (u32)(bA >> 31) + (exA >> 16);
        const u64 t_off = base_text + (bB & 0x7FFFFFFFu), s_off = base_sam + (bB >> 31);
        if (proc && g.status == ST_SELFCIRCLE) {                               // for the thread-0-share emulation on the host
            u32 slot = atomicAdd(&st->sc_count, 1u);
            if (slot < p.sc_cap) p.sc_list[slot] = base_groups + g_idx; else atomicOr(&st->err, S2P_ERR_SCLIST);
        }
        if (emit && p.emit_packed) {
            u64 o = base_pairs + e_idx;
            if (o < p.out_pairs_cap) {
                mk_pair r; r.pos1 = g.posA; r.pos2 = g.posB; r.chr1 = g.chrA; r.chr2 = g.chrB; r.strands = g.strands;
                r.cls = (u8)(g.status - ST_TRANS); r.lane = p.lane;
                p.out_pairs[o] = r;
            } else atomicOr(&st->err, S2P_ERR_PAIRS);
        }
        if (p.emit_text && totT) {
            if (t_off + totT > p.out_text_cap) { if (tid == 0) atomicOr(&st->err, S2P_ERR_TEXT); }
            else if (totT <= EMIT_STAGE) {
                // stage the tile's lines in shared memory with the destination's 16-byte phase, then flush with wide stores
                const u32 phase = (u32)((u64)(p.out_text + t_off) & 15u);
                if (emit) { PtrSink s{s_stage + phase + exT}; write_pair_line(p, ws, g, s); }
                __syncthreads();
                char *dst = p.out_text + t_off;
                const u32 head = phase ? (16u - phase < totT ? 16u - phase : totT) : 0u;
                if ((u32)tid < head) dst[tid] = s_stage[phase + tid];
                const u32 body = (totT - head) >> 4;
                for (u32 w = tid; w < body; w += EMIT_THREADS)
                    *(uint4 *)(dst + head + ((u64)w << 4)) = *(const uint4 *)(s_stage + phase + head + (w << 4));
                const u32 tail0 = head + (body << 4);
                if (tail0 + tid < totT) dst[tail0 + tid] = s_stage[phase + tail0 + tid];
            } else if (emit) {
                PtrSink s{p.out_text + t_off + exT}; write_pair_line(p, ws, g, s);
            }
        }
        if (p.write_sam && emit) {
            // destination of every kept line of the group; K5 does the copying
            u64 o = s_off + exS;
            if (o + g.sam_len > p.out_sam_cap) atomicOr(&st->err, S2P_ERR_SAM);
            else {
                u32 k = i;
                while (true) {
                    p.sam_dst[k] = (u32)(o - base_sam);
                    o += p.rec[k].line_len + 1;
                    if (k == g.last_line) break;
                    ++k; while (!(p.lmeta[k] & LM_KEEP)) ++k;
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ K5: SAM passthrough
// One warp per line: the lines of emitted groups are copied verbatim (with their '\n').
static __global__ void __launch_bounds__(256) k_copy_sam(S2PParams p) {
    const WinState *st = p.st;
    const u32 n_lines = st->n_lines;
    const u64 ws = st->ws, base = st->out_sam;
    const int lane = threadIdx.x & 31;
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (u32 i = warp; i < n_lines; i += nwarps) {
        if (!(p.lmeta[i] & LM_KEEP)) continue;
        const u32 d = p.sam_dst[i];
        if (d == 0xFFFFFFFFu) continue;
        const u64 src = ws + (i ? p.nl_pos[i - 1] + 1 : 0);
        const u32 len = p.nl_pos[i] - (i ? p.nl_pos[i - 1] + 1 : 0) + 1;
        char *dst = p.out_sam + base + d;
        for (u32 b = lane; b < len; b += 32) dst[b] = p.buf[src + b];
    }
}
