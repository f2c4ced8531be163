// s2p_fused.cuh — the single-pass sam2pairs path: every SAM byte is read from HBM once.
//
//   k_ft_strip   warp-private pipelines, no block barrier, no dependency between warps.  A warp claims a 128 KiB strip of the
//                window (absolute 128 KiB boundaries, atomic ticket) and streams it with the same coalesced 128-bit loads and
//                SWAR newline test as k_scan_chunks, eight 512-byte rows in flight.  Line starts go to a ring in shared memory,
//                never to HBM.  Whenever 32 complete lines are pending, the warp parses them one per lane (their first 112
//                bytes come back out of L2, staged in shared-memory columns so that unaligned 8-byte fetches are conflict-free
//                32-bit loads), keeps the 52-byte records of the last 64 lines in shared memory, resolves the read groups whose
//                eight-line look-ahead is complete (resolve_group) and writes .pairs text, packed pairs and passthrough copy
//                entries into the strip's own scratch segment.  No per-line record, line index or group result goes to HBM.
//                A strip depends on no other strip: the group that starts in it is followed into the next strip (2 KiB halo,
//                then a byte-level walk from global memory), and whether its first kept line continues a group of the previous
//                strip is decided by looking back the same way.
//   k_ft_prefix  one CTA: exclusive prefixes of the strips' sizes, window totals, counters committed, self-circle group indices
//   k_ft_gather  one CTA per strip: the scratch segments copied to their final, dense positions (128-bit stores at any
//                source/destination alignment), passthrough lines copied straight from the SAM text
//
// Replaces pairutil.h:136-177 (load_batch: getline, filter, grouping) + the flash2pairs.h / unc2pairs.h bodies + the string
// appends of unc2pairs.h:311-356 in one pass.  Anything the geometry cannot hold (more than ~160 newlines in 4 KiB, more than
// FT_LMAX emitted pairs / passthrough lines or FT_TEXT_CAP bytes of pair text per strip) gives the WINDOW back to the
// multi-kernel path (k_scan_chunks ... k_emit), which handles any input; results are identical by construction (same
// tokenisation rules, same generic parser for anything unusual, same resolver).
#pragma once
#include "s2p_kernels.cuh"

#define FT_TILE 131072u               // strip size; ft_tot / ft_pre / scratch segments are indexed by window-local strip
#define FT_LMAX 1024u                 // emitted pairs / passthrough entries per strip
#define FT_TEXT_CAP 49152u            // bytes of pair text per strip in the scratch
#define FS_THREADS 256
#define FS_WARPS 8
#define FS_HALO 4096u                 // bytes scanned on either side of the strip (8 rows of 512 B, one scan step)
#define FS_RING 256u                  // line starts (relative to the strip's scan origin) known to the warp
#define FS_RECS 64u                   // records / meta bytes of the last 64 parsed lines
#define FS_LOOKAHEAD 8u               // lines parsed beyond the heads a round resolves
#define FS_STAGE 4064u                // pair text of one round staged in the (then dead) line-head columns

struct __align__(4) FtRec {           // LineRec's fields at a 13-word stride: conflict-free across the lanes of a warp
    u32 pos, right0, left1, right1, leftClip, rightClip, mappable, line_len;
    u16 flag, qname_len, chr_slot; u8 segCnt, pad0;
    u32 qname_off, pad1, pad2;
};
static_assert(sizeof(FtRec) == 52, "record stride");
#define FS_OFF_REC 4096
#define FS_OFF_RING (FS_OFF_REC + 52 * FS_RECS)
#define FS_OFF_META (FS_OFF_RING + 4 * FS_RING)
#define FS_WARP_SMEM (FS_OFF_META + FS_RECS)
#define FS_SMEM (FS_WARPS * FS_WARP_SMEM + 64)

// ---------------------------------------------------------------------------------------------- byte-level helpers (rare paths)
#define FT_NONE (~(u64)0)
// first '\n' in [a, we), or FT_NONE
static __device__ __noinline__ u64 ft_find_nl(const char *buf, u64 a, u64 we) {
    while (a < we && (a & 7)) { if (buf[a] == '\n') return a; ++a; }
    while (a + 8 <= we) {
        const u64 x = __ldg((const u64 *)(buf + a));
        const u32 lo = nl_y((u32)x), hi = nl_y((u32)(x >> 32));
        if (lo) return a + ((u32)(__ffs(lo) - 1) >> 3);
        if (hi) return a + 4 + ((u32)(__ffs(hi) - 1) >> 3);
        a += 8;
    }
    while (a < we) { if (buf[a] == '\n') return a; ++a; }
    return FT_NONE;
}
// start of the line that precedes the line starting at `a` (a > ws, buf[a - 1] == '\n')
static __device__ __noinline__ u64 ft_prev_line_start(const char *buf, u64 a, u64 ws) {
    u64 i = a - 1;
    while (i > ws) { --i; if (buf[i] == '\n') return i + 1; }
    return ws;
}
// is the kept line at `a` the first kept record of its QNAME run?  (walks back over the lines before it)
static __device__ __noinline__ bool ft_head_slow(const S2PParams &p, u64 ws, u64 a) {
    u64 cur = a;
    while (cur > ws) {
        const u64 pa = ft_prev_line_start(p.buf, cur, ws);
        LineRec rec;
        const u32 meta = parse_line_slow_abs(p, pa, false, 0, rec);
        if (meta & LM_KEEP) return !qname_equal_abs(p, a, pa);
        cur = pa;
    }
    return true;
}

// the generic parser, out of line, into a shared-memory record
static __device__ __noinline__ u32 fs_parse_slow(const S2PParams &p, u64 a, bool has_prev, u64 pa, FtRec *out) {
    LineRec r;
    const u32 meta = parse_line_slow_abs(p, a, has_prev, pa, r);
    if (meta & LM_KEEP) {
        out->pos = r.pos; out->right0 = r.right0; out->left1 = r.left1; out->right1 = r.right1; out->leftClip = r.leftClip; out->rightClip = r.rightClip;
        out->mappable = r.mappable; out->flag = r.flag; out->qname_len = r.qname_len; out->chr_slot = r.chr_slot; out->segCnt = r.segCnt; out->qname_off = r.qname_off;
    }
    return meta;
}

// one out-of-line copy of the resolver for this path (records in shared or local memory)
static __device__ __noinline__ Resolved fs_resolve(const S2PParams &p, u32 n, u32 n1, u32 n2, const FtRec *f0, const FtRec *f1,
                                                   const FtRec *a1, const FtRec *b1, const FtRec *a2, const FtRec *b2) {
    return resolve_group(p, n, n1, n2, f0, f1, a1, b1, a2, b2);
}
struct FtGroup { Resolved r; u32 sam_len, n_members, qname_len, qname_off; bool off_end, kept; };
// The whole group of the head line at `a_head`, line by line from global memory (any number of lines, any line length).
// entries != nullptr: also writes one passthrough copy entry per member (src relative to ws, length with '\n', destination).
static __device__ __noinline__ void ft_group_slow(const S2PParams &p, u64 ws, u64 we, u64 a_head, FtGroup &out, uint4 *entries, u32 ent_cap, u32 dst0) {
    FtRec f[2], r1[2], r2[2], rec;
    u32 n = 0, n1 = 0, n2 = 0, sam_len = 0;
    out.off_end = false;
    u64 cur = a_head;
    u64 e = ft_find_nl(p.buf, cur, we);
    rec.qname_len = 0; rec.qname_off = 0;
    const u32 meta = fs_parse_slow(p, cur, false, 0, &rec);
    out.kept = (meta & LM_KEEP) != 0; out.qname_len = rec.qname_len; out.qname_off = rec.qname_off;
    out.sam_len = 0; out.n_members = 0;
    if (!out.kept || e == FT_NONE) { out.off_end = true; return; }
    while (true) {
        // `rec` (line [cur, e]) is a member
        if (n < 2) f[n] = rec;
        if (rec.flag & 64u) { if (n1 < 2) r1[n1] = rec; ++n1; } else if (rec.flag & 128u) { if (n2 < 2) r2[n2] = rec; ++n2; }
        const u32 len = (u32)(e - cur) + 1u;
        if (entries && n < ent_cap) entries[n] = make_uint4((u32)(cur - ws), len, dst0 + sam_len, 0u);
        sam_len += len; ++n;
        // next kept line
        bool more = false;
        while (true) {
            const u64 nxt = e + 1;
            if (nxt >= we) { out.off_end = true; break; }
            const u64 e2 = ft_find_nl(p.buf, nxt, we);
            if (e2 == FT_NONE) { out.off_end = true; break; }
            const u32 m2 = fs_parse_slow(p, nxt, true, a_head, &rec);
            cur = nxt; e = e2;
            if (!(m2 & LM_KEEP)) continue;
            more = (m2 & LM_EQ) != 0;
            break;
        }
        if (out.off_end || !more) break;
    }
    out.sam_len = sam_len; out.n_members = n;
    if (!out.off_end) out.r = fs_resolve(p, n, n1, n2, &f[0], &f[1], &r1[0], &r1[1], &r2[0], &r2[1]);
}

// ---------------------------------------------------------------------------------------------- scan of 512-byte rows
// Line starts (newline offset + 1, relative to the scan origin; rel0 = offset of rbase from it) of `nrows` rows appended, in byte
// order, to the warp's ring; returns the new number of known starts.  Same tests as k_scan_chunks.
// a 512-byte row in which some 16-byte word holds several newlines (empty or very short lines): ranks by a warp scan
static __device__ __noinline__ u32 fs_push_multi(u32 *ring, u32 n, u32 z, u32 rel, u32 lane) {
    const u32 c = __popc(z);
    const u32 inc = warp_incl_scan(c, (int)lane);
    const u32 tot = __shfl_sync(0xFFFFFFFFu, inc, 31);
    if (c) {
        u32 m = 0, idx = n + inc - c;
        while (z) { const u32 b = __ffs(z) - 1; z &= z - 1; m |= 1u << byte_of_perm_bit(b); }
        while (m) { const u32 q = __ffs(m) - 1; m &= m - 1; ring[idx & (FS_RING - 1u)] = rel + q; ++idx; }
    }
    return n + tot;
}
template <int U>
__device__ __forceinline__ u32 fs_scan_push(const char *buf, long long rbase, u32 nrows, u32 rel0, u64 ws, u64 we, u32 *ring, u32 n, u32 lane) {
    const bool edge = rbase < (long long)ws || rbase + (long long)nrows * 512 > (long long)we;
    const uint4 *src = (const uint4 *)(buf + rbase) + lane;
    const u32 lt = (1u << lane) - 1u;
#pragma unroll 1
    for (u32 it = 0; it < nrows; it += U) {
        uint4 w[U];
        if (!edge) {
#pragma unroll
            for (int u = 0; u < U; ++u) w[u] = ld_stream_v4(src + (it + u) * 32u);
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long off = rbase + (long long)((it + u) * 32u + lane) * 16;
                w[u] = (off < (long long)we && off + 16 > (long long)ws) ? ld_stream_v4(src + (it + u) * 32u) : make_uint4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const u32 e0 = nl_raw(w[u].x), e1 = nl_raw(w[u].y), e2 = nl_raw(w[u].z), e3 = nl_raw(w[u].w);
            u32 z = ((e0 >> 7) & 0x01010101u) | ((e1 >> 6) & 0x02020202u) | ((e2 >> 5) & 0x04040404u) | ((e3 >> 4) & 0x08080808u);
            if (edge && z) {
                const long long off = rbase + (long long)((it + u) * 32u + lane) * 16;
                u32 keep = 0;
#pragma unroll 1
                for (u32 q = 0; q < 16; ++q) if (off + q >= (long long)ws && off + q < (long long)we) keep |= 1u << perm_bit_of_byte(q);
                z &= keep;
            }
            const u32 bal = __ballot_sync(0xFFFFFFFFu, z != 0);
            if (bal == 0) continue;
            const u32 multi = __ballot_sync(0xFFFFFFFFu, (z & (z - 1u)) != 0);
            const u32 rel = rel0 + (it + u) * 512u + lane * 16u + 1u;
            if (multi == 0) {
                if (z) ring[(n + __popc(bal & lt)) & (FS_RING - 1u)] = rel + byte_of_perm_bit(__ffs(z) - 1);
                n += __popc(bal);
                continue;
            }
            n = fs_push_multi(ring, n, z, rel, lane);
        }
    }
    return n;
}

// ---------------------------------------------------------------------------------------------- lean per-line parser
// The lane's line head sits in a shared-memory column of 32-bit words, word w at col[32 * w]; four zero words follow the
// 28 staged ones, so eight bytes at any offset below 112 are three conflict-free loads and two funnel shifts.
struct ColFetch {
    const u32 *col;
    __device__ __forceinline__ u32 byter(u32 r) const { return (col[(r >> 2) * 32] >> (8 * (r & 3u))) & 0xFFu; }
    __device__ __forceinline__ void f8(u32 r, u32 &lo, u32 &hi) const {
        const u32 *q = col + (r >> 2) * 32; const u32 sh = (r & 3u) * 8u;
        const u32 a = q[0], b = q[32], c = q[64];
        lo = __funnelshift_r(a, b, sh); hi = __funnelshift_r(b, c, sh);
    }
};
// `len` (1..8) decimal characters in the low bytes of hi:lo, first character lowest; ok is cleared on a non-digit
__device__ __forceinline__ u32 dec8(u32 lo, u32 hi, u32 len, bool &ok) {
    // right-align: the last character goes to byte 7
    const u32 sh = (8u - len) * 8u;
    u32 l2, h2, ml, mh;                                                // ml / mh: 0xFF in the real characters' bytes
    if (sh >= 32u) { h2 = lo << (sh - 32u); l2 = 0; mh = 0xFFFFFFFFu << (sh - 32u); ml = 0; }
    else { h2 = __funnelshift_l(lo, hi, sh); l2 = lo << sh; mh = 0xFFFFFFFFu; ml = 0xFFFFFFFFu << sh; }
    l2 = (l2 ^ 0x30303030u) & ml; h2 = (h2 ^ 0x30303030u) & mh;         // digit values; anything else has a byte > 9
    const u32 bad = ((((l2 & 0x7F7F7F7Fu) + 0x76767676u) | l2) | (((h2 & 0x7F7F7F7Fu) + 0x76767676u) | h2)) & 0x80808080u;
    ok = ok && bad == 0;
    u32 t = (l2 * 10u + (l2 >> 8)) & 0x00FF00FFu;
    const u32 vl = (t & 0xFFu) * 100u + (t >> 16);
    t = (h2 * 10u + (h2 >> 8)) & 0x00FF00FFu;
    const u32 vh = (t & 0xFFu) * 100u + (t >> 16);
    return vl * 10000u + vh;
}
// decimal field of 1..10 characters at offset r of the column
__device__ __forceinline__ u32 fs_dec_field(const ColFetch &f, u32 r, u32 len, bool &ok) {
    u32 lo, hi;
    f.f8(r, lo, hi);
    if (len <= 8u) return dec8(lo, hi, len, ok);
    const u32 head = dec8(lo, hi, len - 8u, ok);                         // one or two leading characters
    f.f8(r + len - 8u, lo, hi);
    return head * 100000000u + dec8(lo, hi, 8u, ok);
}

// Well-formed lines only (single tabs between the first six fields, nothing else below 0x21 in front of them, digits where
// numbers belong, RNAME <= 8 bytes); anything else returns false and goes through the generic byte parser, so the result is
// identical by construction.  m0..m3: bit j set iff byte j of the line is below 0x21 (the first 112 - s bytes).
__device__ __forceinline__ bool fs_parse_fast(const S2PParams &p, const ColFetch &f, const u32 s, u32 m0, u32 m1, u32 m2, u32 m3,
                                              const u64 a, FtRec &rec, u32 &meta, u32 &t0_out) {
    const u32 t0 = pop_lowest128(m0, m1, m2, m3), t1 = pop_lowest128(m0, m1, m2, m3), t2 = pop_lowest128(m0, m1, m2, m3);
    const u32 t3 = pop_lowest128(m0, m1, m2, m3), t4 = pop_lowest128(m0, m1, m2, m3), t5 = pop_lowest128(m0, m1, m2, m3);
    if (t5 >= 112u - s) return false;                                 // six separators inside the bytes we looked at
    if (t0 == 0 || t1 == t0 + 1 || t2 == t1 + 1 || t3 == t2 + 1 || t4 == t3 + 1 || t5 == t4 + 1) return false;
    if (f.byter(s + t0) != '\t' || f.byter(s + t1) != '\t' || f.byter(s + t2) != '\t' || f.byter(s + t3) != '\t' || f.byter(s + t4) != '\t' ||
        f.byter(s + t5) != '\t') return false;
    if (f.byter(s) == '@') return false;
    const u32 l_flag = t1 - t0 - 1, l_name = t2 - t1 - 1, l_pos = t3 - t2 - 1, l_mapq = t4 - t3 - 1;
    if (l_flag > 5 || l_name > 8 || l_pos > 10 || l_mapq > 3) return false;
    bool ok = true;
    const u32 flag = fs_dec_field(f, s + t0 + 1, l_flag, ok);
    const u32 pos = fs_dec_field(f, s + t2 + 1, l_pos, ok);
    const u32 mapq = fs_dec_field(f, s + t3 + 1, l_mapq, ok);
    if (!ok) return false;
    t0_out = t0;
    meta = 0;
    if (mapq < (u32)p.min_mapq || (flag & 0x700u)) return true;       // pairutil.h:157-161
    meta = LM_KEEP;
    // RNAME: FNV-1a over its bytes, same as the byte loop
    u32 nlo, nhi;
    f.f8(s + t1 + 1, nlo, nhi);
    u64 name8 = (u64)nlo | ((u64)nhi << 32);
    if (l_name < 8) name8 &= (1ull << (8 * l_name)) - 1;
    u64 h = 0xCBF29CE484222325ull;
#pragma unroll
    for (int k = 0; k < 8; ++k) if ((u32)k < l_name) h = hash_step(h, (int)((name8 >> (8 * k)) & 0xFF));
    // CIGAR walk (pairutil.h:63-126), 8 characters per fetch
    u32 val = 0, idx = 0, leftClip = 0, rightClip = 0, mappable = 0;
    u32 cur = pos, right0 = 0, left1 = 0, right1 = 0, last_right = 0;
    bool err = false;
    const u32 l_cig = t5 - t4 - 1;
    u32 xlo = 0, xhi = 0;
    for (u32 k = 0; k < l_cig; ++k) {
        if ((k & 7u) == 0) f.f8(s + t4 + 1 + k, xlo, xhi);
        const u32 c = xlo & 0xFFu; xlo = __funnelshift_r(xlo, xhi, 8); xhi >>= 8;
        const u32 d = c - '0';
        if (d <= 9u) { val = val * 10u + d; continue; }
        if (c == 'H' || c == 'S') {
            if (k + 1 == l_cig) rightClip = val;
            else if (idx == 0) leftClip = val;
            else err = true;
        } else if (c == 'M' || c == 'D') {
            if (c == 'M') mappable += val;
            cur += val; last_right = cur - 1;
            if (idx == 0) right0 = last_right; else if (idx == 1) right1 = last_right;
        } else if (c == 'N') {
            cur += val; ++idx; last_right = 0;
            if (idx == 1) left1 = cur;
        } else if (c != 'I') err = true;
        val = 0;
    }
    rec.pos = pos; rec.right0 = right0; rec.left1 = left1; rec.right1 = right1;
    rec.leftClip = leftClip; rec.rightClip = rightClip; rec.mappable = mappable;
    rec.flag = (u16)flag; rec.qname_len = (u16)t0;
    rec.chr_slot = (u16)chr_lookup_insert(p, h, name8, p.buf, a + t1 + 1, l_name);
    const u32 segCnt = idx + 1;
    if (last_right == 0) err = true;
    rec.segCnt = err ? 0 : (u8)(segCnt > 2 ? 3 : segCnt);
    rec.qname_off = 0;
    return true;
}
// ---------------------------------------------------------------------------------------------- the strip kernel
static __global__ void __launch_bounds__(FS_THREADS, 3) k_ft_strip(S2PParams p) {
    extern __shared__ __align__(16) unsigned char fs_smem[];
    const u32 tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    unsigned char *wb = fs_smem + wid * FS_WARP_SMEM;
    u32 *col = (u32 *)wb + lane;                                      // this lane's column of line-head words (word w at col[32 w])
    FtRec *recs = (FtRec *)(wb + FS_OFF_REC);
    u32 *ring = (u32 *)(wb + FS_OFF_RING);
    u8 *metas = (u8 *)(wb + FS_OFF_META);
    u32 *s_cnt = (u32 *)(fs_smem + FS_WARPS * FS_WARP_SMEM);
    WinState *st = p.st;
    if (st->halt || st->path_old) return;
    const u64 ws = st->ws, we = st->we;
    if (we <= ws) return;
    const u64 limit = st->total;
    const u64 strip0 = ws / FT_TILE;
    const u32 n_strips = (u32)((we - 1) / FT_TILE - strip0 + 1);
    if (tid < 16) s_cnt[tid] = 0;
#pragma unroll
    for (int j = 28; j < 32; ++j) col[32 * j] = 0;                    // pad words behind the 112 staged bytes
    __syncthreads();
    const S2PParams &ps = *p.self;                                     // what out-of-line callees get
    const u32 ltmask = (1u << lane) - 1u;
    u64 last_end = 0;

    while (true) {
        u32 sl = 0;
        if (lane == 0) sl = atomicAdd(&st->tickets[0], 1u);
        sl = __shfl_sync(0xFFFFFFFFu, sl, 0);
        if (sl >= n_strips || sl >= p.n_tiles_cap) break;
        if (*(volatile u32 *)&st->path_old) break;
        const u64 sbase = (strip0 + sl) * FT_TILE, send = sbase + FT_TILE;
        const long long ebase = (long long)sbase - (long long)FS_HALO; // line starts below are relative to ebase
        const bool has_initial = ebase <= (long long)ws;                // the window's first line starts inside the scanned range
        char *sc_text = p.ft_text + (size_t)sl * FT_TEXT_CAP;
        mk_pair *sc_pairs = p.ft_pairs + (size_t)sl * FT_LMAX;
        uint4 *sc_sam = p.ft_sam + (size_t)sl * FT_LMAX;
        u32 n_known = 0;
        if (has_initial) { if (lane == 0) ring[0] = (u32)((long long)ws - ebase); n_known = 1; }
        // scan steps of 8 rows (4 KiB): step 0 in front of the strip (only its last lines matter: look-back of the strip's first
        // group), steps 1..32 the strip, step 33 behind it (look-ahead of its last group)
        u32 n_main0 = n_known, n_drop = 0, n_main1 = n_known;
        u32 n_parsed = 0, n_res = 0;
        u32 run_g = 0, run_e = 0, run_t = 0, run_s = 0, run_n = 0;     // strip totals so far (warp-uniform)
        u32 stage = 0;
        bool ovf = false;
        while (true) {
            const u32 pending = n_known - n_parsed;                    // starts not parsed yet (the last one has no known end)
            if (stage <= 33u && (pending <= 32u || stage == 0u)) {     // not enough for a full round of 32 complete lines: scan on
                n_known = fs_scan_push<8>(p.buf, ebase + (long long)stage * 4096, 8u, stage * 4096u, ws, we, ring, n_known, lane);
                ++stage;
                if (n_known - n_parsed > FS_RING - 64u) { ovf = true; break; }
                if (stage == 1u) {                                     // lines from here on start behind a newline of this strip
                    n_main0 = n_known; n_drop = n_known > 3u ? n_known - 3u : 0u;   // lines below n_drop are never parsed
                    n_parsed = n_res = n_drop;
                }
                if (stage <= 33u) n_main1 = n_known;
                if (stage == 33u && n_main1 > n_main0) {               // end of the strip's last complete line (a window without kept records ends there)
                    __syncwarp();
                    const u64 e = (u64)(ebase + (long long)ring[(n_main1 - 1u) & (FS_RING - 1u)]);
                    if (e > last_end) last_end = e;
                }
                continue;
            }
            const bool final = stage > 33u;
            const u32 cnt = pending > 32u ? 32u : (pending ? pending - 1u : 0u);
            __syncwarp();
            // ---- parse `cnt` lines, one per lane.  A start is a complete line iff it has a successor in the ring.
            if (cnt) {
                const u32 line = n_parsed + lane;
                const bool act = lane < cnt;
                u64 a = 0, A = ~(u64)0;
                u32 m0 = 0, m1 = 0, m2 = 0, m3 = 0, s = 0;
                if (act) {
                    a = (u64)(ebase + (long long)ring[line & (FS_RING - 1u)]);
                    if (a + 144 <= limit) {
                        A = a & ~(u64)15; s = (u32)(a - A);
                        const uint4 *src = (const uint4 *)(p.buf + A);
                        uint4 w[7];
#pragma unroll
                        for (int j = 0; j < 7; ++j) w[j] = __ldg(src + j);
#pragma unroll
                        for (int j = 0; j < 7; ++j) { col[32 * (4 * j)] = w[j].x; col[32 * (4 * j + 1)] = w[j].y; col[32 * (4 * j + 2)] = w[j].z; col[32 * (4 * j + 3)] = w[j].w; }
                        m0 = lt21_mask16(w[0]) | (lt21_mask16(w[1]) << 16); m1 = lt21_mask16(w[2]) | (lt21_mask16(w[3]) << 16);
                        m2 = lt21_mask16(w[4]) | (lt21_mask16(w[5]) << 16); m3 = lt21_mask16(w[6]);
                        if (s) { m0 = __funnelshift_r(m0, m1, s); m1 = __funnelshift_r(m1, m2, s); m2 = __funnelshift_r(m2, m3, s); m3 >>= s; }
                    }
                }
                const u64 An = __shfl_up_sync(0xFFFFFFFFu, A, 1);       // the previous line's staged base (lane - 1)
                __syncwarp();
                if (act) {
                    const bool has_prev = line > 0;
                    const u64 pa = has_prev ? (u64)(ebase + (long long)ring[(line - 1u) & (FS_RING - 1u)]) : 0;
                    FtRec &rec = recs[line & (FS_RECS - 1u)];
                    ColFetch f; f.col = col;
                    u32 meta = 0, t0 = 0;
                    if (A != ~(u64)0 && fs_parse_fast(p, f, s, m0, m1, m2, m3, a, rec, meta, t0)) {
                        if (has_prev) {
                            bool eq;
                            if (is_blank((int)(unsigned char)p.buf[pa])) eq = qname_equal_abs(ps, a, pa);   // operator>> skips leading blanks
                            else if (lane > 0 && An != ~(u64)0 && (u32)(pa - An) + t0 + 1 <= 112) {        // inside the neighbour's staged bytes
                                ColFetch fp; fp.col = col - 1;
                                const u32 sp = (u32)(pa - An);
                                eq = true;
                                for (u32 k = 0; k < t0 && eq; k += 8) {
                                    u32 xl, xh, yl, yh;
                                    f.f8(s + k, xl, xh); fp.f8(sp + k, yl, yh);
                                    xl ^= yl; xh ^= yh;
                                    const u32 r = t0 - k;
                                    if (r < 8u) { if (r <= 4u) { xh = 0; if (r < 4u) xl &= (1u << (8u * r)) - 1u; } else xh &= (1u << (8u * (r - 4u))) - 1u; }
                                    eq = (xl | xh) == 0;
                                }
                                eq = eq && is_ws((int)fp.byter(sp + t0));
                            } else {
                                GlobalFetch gf; gf.buf = p.buf; gf.A = 0;
                                eq = qname_eq_fetch(gf, a, pa, t0);
                            }
                            if (eq) meta |= LM_EQ;
                        }
                    } else meta = fs_parse_slow(ps, a, has_prev, pa, &rec);
                    if (!has_prev && !has_initial) meta |= LM_EQ_UNK;   // the line before the first one we know of
                    metas[line & (FS_RECS - 1u)] = (u8)meta;
                }
                n_parsed += cnt;
                __syncwarp();
            }
            // ---- resolve the heads whose look-ahead is parsed, 32 lines per step
            const bool last = final && (n_known - n_parsed <= 1u);
            const u32 R = last ? n_known : (n_parsed > FS_LOOKAHEAD ? n_parsed - FS_LOOKAHEAD : 0u);
            while (n_res < R && !ovf) {
                const u32 line = n_res + lane;
                const u32 step = R - n_res < 32u ? R - n_res : 32u;
                bool proc = false, emit = false, use_slow = false;
                Resolved rs; rs.status = ST_NONE; rs.have = false; rs.p1 = rs.p2 = 0; rs.sA = rs.sB = 0; rs.strands = 0;
                u32 text_len = 0, sam_len = 0, n_mem = 0, rid_len = 0, rid_off = 0;
                u64 a = 0;
                if (lane < step) {
                    const u32 rel = ring[line & (FS_RING - 1u)];
                    a = (u64)(ebase + (long long)rel);
                    const bool own = rel >= FS_HALO && rel < FS_HALO + FT_TILE && a < we;
                    const bool parsed = line < n_parsed;
                    u32 meta = parsed ? metas[line & (FS_RECS - 1u)] : 0u;
                    bool kept = own && (meta & LM_KEEP);
                    if (own && !parsed) {                              // the last start we know: a line only if it ends before `we`
                        use_slow = true;
                        kept = false;
                        if (ft_find_nl(p.buf, a, we) != FT_NONE) {
                            FtRec tmp;
                            const bool has_prev = line > 0;
                            meta = fs_parse_slow(ps, a, has_prev, has_prev ? (u64)(ebase + (long long)ring[(line - 1u) & (FS_RING - 1u)]) : 0, &tmp);
                            if (!has_prev && !has_initial) meta |= LM_EQ_UNK;
                            kept = (meta & LM_KEEP) != 0;
                        }
                    }
                    if (kept) {
                        // head?  (pairutil.h:163-173: currId != lastId among kept records)
                        int head = (meta & LM_EQ_UNK) ? 2 : -1;        // 0 no, 1 yes, 2 walk the bytes
                        {
                            bool chain = (meta & LM_EQ) != 0;
                            const u32 low = n_parsed > FS_RECS ? max(n_drop, n_parsed - FS_RECS) : n_drop;   // oldest line whose meta is still held
                            u32 j = line;
                            while (head < 0 && j > low) {
                                --j;
                                const u32 mj = metas[j & (FS_RECS - 1u)];
                                if (mj & LM_KEEP) {
                                    if (chain) head = 0;
                                    else if (j + 1 == line) head = 1;
                                    else head = qname_equal_abs(ps, a, (u64)(ebase + (long long)ring[j & (FS_RING - 1u)])) ? 0 : 1;
                                    break;
                                }
                                if (mj & LM_EQ_UNK) { head = 2; break; }
                                chain = chain && (mj & LM_EQ);
                            }
                            if (head < 0) head = (low == 0 && has_initial) ? 1 : 2;   // nothing kept before it in the window / out of sight
                            if (head == 2) head = ft_head_slow(ps, ws, a) ? 1 : 0;
                        }
                        if (head) {
                            u32 f0 = line, f1 = line, r1a = line, r1b = line, r2a = line, r2b = line;
                            u32 n = 0, n1 = 0, n2 = 0;
                            if (!use_slow) {
                                u32 k = line;
                                while (true) {
                                    const u32 fl = recs[k & (FS_RECS - 1u)].flag;
                                    if (n == 0) f0 = k; else if (n == 1) f1 = k;
                                    ++n;
                                    if (fl & 64u) { if (n1 == 0) r1a = k; else if (n1 == 1) r1b = k; ++n1; }
                                    else if (fl & 128u) { if (n2 == 0) r2a = k; else if (n2 == 1) r2b = k; ++n2; }
                                    sam_len += ring[(k + 1u) & (FS_RING - 1u)] - ring[k & (FS_RING - 1u)];
                                    u32 q = k + 1; bool chain = true; u32 mq = 0;
                                    while (q < n_parsed) { mq = metas[q & (FS_RECS - 1u)]; chain = chain && (mq & LM_EQ); if (mq & LM_KEEP) break; ++q; }
                                    if (q >= n_parsed) { use_slow = true; break; }
                                    const bool same = chain ? true : (q == k + 1 ? false :
                                        qname_equal_abs(ps, (u64)(ebase + (long long)ring[q & (FS_RING - 1u)]), (u64)(ebase + (long long)ring[k & (FS_RING - 1u)])));
                                    if (!same) break;
                                    k = q;
                                }
                                n_mem = n;
                            }
                            if (use_slow) {
                                FtGroup g;
                                ft_group_slow(ps, ws, we, a, g, nullptr, 0, 0);
                                if (g.off_end) {                       // the window's last group: carried to the next window (or dropped at EOF)
                                    st->ft_carry_pos = a; st->ft_carry_tile = sl;
                                    st->ft_carry_nl = line >= n_main0 ? line - n_main0 + 1u : 0u;   // newlines of this strip before the line
                                } else { proc = true; rs = g.r; sam_len = g.sam_len; n_mem = g.n_members; rid_len = g.qname_len; rid_off = g.qname_off; }
                            } else {
                                proc = true;
                                rs = fs_resolve(ps, n, n1, n2, &recs[f0 & (FS_RECS - 1u)], &recs[f1 & (FS_RECS - 1u)], &recs[r1a & (FS_RECS - 1u)],
                                                   &recs[r1b & (FS_RECS - 1u)], &recs[r2a & (FS_RECS - 1u)], &recs[r2b & (FS_RECS - 1u)]);
                                rid_len = recs[line & (FS_RECS - 1u)].qname_len; rid_off = recs[line & (FS_RECS - 1u)].qname_off;
                            }
                            if (proc) {
                                if (rs.status != ST_NONE) atomicAdd(&s_cnt[rs.status], 1u);
                                if (rs.have && rs.status != ST_SELFCIRCLE) {
                                    emit = true;
                                    text_len = rid_len + p.chr[rs.sA].len + p.chr[rs.sB].len + dec_digits(rs.p1) + dec_digits(rs.p2) + 9u;
                                }
                            }
                        }
                    }
                }
                if (!emit || !p.write_sam) { sam_len = 0; n_mem = 0; }   // (also the partial sums of a walk that went to the byte level)
                // -- offsets inside the step
                const u32 vT = p.emit_text ? text_len : 0u;
                const u32 bal_p = __ballot_sync(0xFFFFFFFFu, proc), bal_e = __ballot_sync(0xFFFFFFFFu, emit);
                const u32 iT = warp_incl_scan(vT, (int)lane);
                const u32 tT = __shfl_sync(0xFFFFFFFFu, iT, 31), bT = iT - vT;
                u32 bS = 0, bE = 0, tS = 0, tE = 0;
                if (p.write_sam) {
                    const u32 iS = warp_incl_scan(sam_len, (int)lane), iE = warp_incl_scan(n_mem, (int)lane);
                    tS = __shfl_sync(0xFFFFFFFFu, iS, 31); tE = __shfl_sync(0xFFFFFFFFu, iE, 31); bS = iS - sam_len; bE = iE - n_mem;
                }
                const u32 bG = __popc(bal_p & ltmask), bP = __popc(bal_e & ltmask), tG = __popc(bal_p), tP = __popc(bal_e);
                if (run_t + tT > FT_TEXT_CAP || run_n + tE > FT_LMAX || run_e + tP > FT_LMAX) { ovf = true; break; }
                const bool staged = tT <= FS_STAGE;
                const u32 phase = (u32)((size_t)(sc_text + run_t) & 15u);
                char *s_stage = (char *)wb;                            // the line-head columns are dead until the next parse round
                if (p.emit_text && tT && staged) {                     // the text is OR-ed into a zeroed stage
#pragma unroll
                    for (int j = 0; j < 8; ++j) ((uint4 *)wb)[j * 32 + lane] = make_uint4(0, 0, 0, 0);
                    __syncwarp();
                }
                if (proc) {
                    if (rs.status == ST_SELFCIRCLE) {                  // (strip, group index inside the strip): k_ft_prefix makes it global
                        const u32 slot = atomicAdd(&st->sc_count, 1u);
                        if (slot < p.sc_cap) p.sc_list[slot] = ((u64)sl << 32) | (u64)(run_g + bG); else atomicOr(&st->err, S2P_ERR_SCLIST);
                    }
                    if (emit) {
                        const ChrSlot *ca = &p.chr[rs.sA], *cb = &p.chr[rs.sB];
                        if (p.emit_packed) {
                            uint4 r;
                            r.x = rs.p1; r.y = rs.p2; r.z = (u32)(u16)ca->id | ((u32)(u16)cb->id << 16);
                            r.w = (u32)rs.strands | ((u32)(rs.status - ST_TRANS) << 8) | ((u32)p.lane << 16);
                            ((uint4 *)sc_pairs)[run_e + bP] = r;
                        }
                        if (p.emit_text) {
                            if (staged && ca->len <= 8 && cb->len <= 8) fs_write_pair_line(p.buf, a + rid_off, rid_len, ca, cb, rs.p1, rs.p2, rs.strands, s_stage + phase + bT);
                            else fs_write_pair_line_bytes(p.buf, a + rid_off, rid_len, ca, cb, rs.p1, rs.p2, rs.strands, staged ? s_stage + phase + bT : sc_text + run_t + bT, staged);
                        }
                        if (p.write_sam) {                             // one copy entry per kept line of the group
                            uint4 *ent = sc_sam + run_n + bE;
                            if (use_slow) { FtGroup g; ft_group_slow(ps, ws, we, a, g, ent, n_mem, run_s + bS); }
                            else {
                                u32 k = line, d = run_s + bS, c = 0;
                                while (true) {
                                    const u32 r0 = ring[k & (FS_RING - 1u)], len = ring[(k + 1u) & (FS_RING - 1u)] - r0;
                                    ent[c++] = make_uint4((u32)((u64)(ebase + (long long)r0) - ws), len, d, 0u);
                                    d += len;
                                    if (c == n_mem) break;
                                    ++k; while (!(metas[k & (FS_RECS - 1u)] & LM_KEEP)) ++k;
                                }
                            }
                        }
                    }
                }
                if (p.emit_text && tT && staged) {
                    __syncwarp();
                    char *dst = sc_text + run_t;
                    const u32 head = phase ? (16u - phase < tT ? 16u - phase : tT) : 0u;
                    if (lane < head) dst[lane] = s_stage[phase + lane];
                    const u32 body = (tT - head) >> 4;
                    for (u32 w = lane; w < body; w += 32u)
                        *(uint4 *)(dst + head + ((size_t)w << 4)) = *(const uint4 *)(s_stage + phase + head + (w << 4));
                    const u32 tail0 = head + (body << 4);
                    if (tail0 + lane < tT) dst[tail0 + lane] = s_stage[phase + tail0 + lane];
                    __syncwarp();
#pragma unroll
                    for (int j = 28; j < 32; ++j) col[32 * j] = 0;     // the stage covered the pad words of the columns
                }
                run_g += tG; run_e += tP; run_t += tT; run_s += tS; run_n += tE;
                n_res += step;
                __syncwarp();
            }
            if (last || ovf) break;
        }
        if (ovf) { if (lane == 0) atomicOr(&st->path_old, 1u); break; }
        if (lane == 0) {                                               // (groups | emitted << 16, text bytes, passthrough bytes, newlines of the strip)
            p.ft_tot[sl] = make_uint4(run_g | (run_e << 16), run_t, run_s, n_main1 - n_main0);
            p.ft_nent[sl] = run_n;
        }
    }
    __syncthreads();
    if (tid < ST_NCOUNTER && s_cnt[tid]) atomicAdd(&st->w_counters[tid], (unsigned long long)s_cnt[tid]);
    if (lane == 0 && last_end) atomicMax((unsigned long long *)&st->ft_last_end, (unsigned long long)last_end);
}

// ---------------------------------------------------------------------------------------------- prefixes over the tiles
static __global__ void __launch_bounds__(1024) k_ft_prefix(S2PParams p) {
    __shared__ u32 s_w[5][32];
    WinState *st = p.st;
    if (st->halt) return;
    const u32 tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    if (st->path_old) { if (tid == 0) st->sc_count = st->sc_count0; return; }   // the multi-kernel path redoes the window
    const u64 ws = st->ws, we = st->we;
    u32 nt = 0;
    if (we > ws) { const u64 n64 = (we - 1) / FT_TILE - ws / FT_TILE + 1; nt = (u32)(n64 < p.n_tiles_cap ? n64 : p.n_tiles_cap); }
    const u32 per = (nt + 1023u) / 1024u;
    const u32 lo = tid * per < nt ? tid * per : nt, hi = lo + per < nt ? lo + per : nt;
    const uint4 *__restrict__ tot = p.ft_tot;
    uint4 *__restrict__ pre = p.ft_pre;
    u32 G = 0, E = 0, T = 0, S = 0, L = 0;
    for (u32 i = lo; i < hi; ++i) { const uint4 v = tot[i]; G += v.x & 0xFFFFu; E += v.x >> 16; T += v.y; S += v.z; L += v.w; }
    const u32 iG = warp_incl_scan(G, (int)lane), iE = warp_incl_scan(E, (int)lane), iT = warp_incl_scan(T, (int)lane),
              iS = warp_incl_scan(S, (int)lane), iL = warp_incl_scan(L, (int)lane);
    if (lane == 31) { s_w[0][wid] = iG; s_w[1][wid] = iE; s_w[2][wid] = iT; s_w[3][wid] = iS; s_w[4][wid] = iL; }
    __syncthreads();
    if (wid < 5) { const u32 v = s_w[wid][lane]; const u32 vi = warp_incl_scan(v, (int)lane); s_w[wid][lane] = vi - v; }
    __syncthreads();
    u32 bG = s_w[0][wid] + iG - G, bE = s_w[1][wid] + iE - E, bT = s_w[2][wid] + iT - T, bS = s_w[3][wid] + iS - S, bL = s_w[4][wid] + iL - L;
    const bool carried = st->ft_carry_pos != ~(u64)0;
    const u32 ct = st->ft_carry_tile;
    for (u32 i = lo; i < hi; ++i) {
        const uint4 v = tot[i];
        pre[i] = make_uint4(bG, bE, bT, bS);
        if (carried && i == ct) st->ft_lines_before_carry = bL + st->ft_carry_nl;
        bG += v.x & 0xFFFFu; bE += v.x >> 16; bT += v.y; bS += v.z; bL += v.w;
    }
    if (tid == 1023) { st->w_groups = bG; st->w_emit = bE; st->w_text = bT; st->w_sam = bS; st->n_lines = bL; }
    if (tid < ST_NCOUNTER) st->counters[tid] += st->w_counters[tid];
    __syncthreads();
    // self-circle entries of this window: (tile, index inside the tile) -> index in the stream
    const u32 s0 = st->sc_count0, s1 = st->sc_count < p.sc_cap ? st->sc_count : p.sc_cap;
    const u64 gbase = st->groups_done;
    for (u32 i = s0 + tid; i < s1; i += 1024u) {
        const u64 e = p.sc_list[i];
        p.sc_list[i] = gbase + pre[(u32)(e >> 32)].x + (u32)e;
    }
}

// ---------------------------------------------------------------------------------------------- scratch -> dense outputs
// n bytes from src to dst, any alignment of either, by the `nthr` threads of a group (idx = thread index in the group):
// 16-byte stores to the aligned body of dst, each built from two aligned 16-byte loads of src.  src must be readable up to
// the next 16-byte boundary past src + n.
__device__ __forceinline__ void ft_copy_bytes(char *dst, const char *src, u32 n, u32 idx, u32 nthr) {
    const u32 head = min(n, (u32)((16u - (u32)((size_t)dst & 15u)) & 15u));
    if (idx < head) dst[idx] = src[idx];
    const u32 body = (n - head) >> 4;
    const char *s0 = src + head;
    const u32 m = (u32)((size_t)s0 & 15u);
    const uint4 *sa = (const uint4 *)(s0 - m);
    uint4 *da = (uint4 *)(dst + head);
    if (m == 0) {
        for (u32 w = idx; w < body; w += nthr) da[w] = __ldg(sa + w);
    } else {
        const u32 ws_ = m >> 2, sh = (m & 3u) * 8u;
        for (u32 w = idx; w < body; w += nthr) {
            const uint4 a = __ldg(sa + w), b = __ldg(sa + w + 1);
            u32 x0, x1, x2, x3, x4;
            if (ws_ == 0) { x0 = a.x; x1 = a.y; x2 = a.z; x3 = a.w; x4 = b.x; }
            else if (ws_ == 1) { x0 = a.y; x1 = a.z; x2 = a.w; x3 = b.x; x4 = b.y; }
            else if (ws_ == 2) { x0 = a.z; x1 = a.w; x2 = b.x; x3 = b.y; x4 = b.z; }
            else { x0 = a.w; x1 = b.x; x2 = b.y; x3 = b.z; x4 = b.w; }
            da[w] = make_uint4(__funnelshift_r(x0, x1, sh), __funnelshift_r(x1, x2, sh), __funnelshift_r(x2, x3, sh), __funnelshift_r(x3, x4, sh));
        }
    }
    const u32 tail0 = head + (body << 4);
    if (tail0 + idx < n) dst[tail0 + idx] = src[tail0 + idx];
}

static __global__ void __launch_bounds__(256) k_ft_gather(S2PParams p) {
    WinState *st = p.st;
    if (st->halt || st->path_old) return;
    const u64 ws = st->ws, we = st->we;
    if (we <= ws) return;
    const u32 tl = blockIdx.x;
    if ((u64)tl > (we - 1) / FT_TILE - ws / FT_TILE || tl >= p.n_tiles_cap) return;
    const u32 tid = threadIdx.x;
    const uint4 tot = p.ft_tot[tl], pre = p.ft_pre[tl];
    const u32 n_emit = tot.x >> 16;
    if (p.emit_text && tot.y) {
        const u64 o = st->out_text + pre.z;
        if (o + tot.y > p.out_text_cap) { if (tid == 0) atomicOr(&st->err, S2P_ERR_TEXT); }
        else ft_copy_bytes(p.out_text + o, p.ft_text + (size_t)tl * FT_TEXT_CAP, tot.y, tid, 256u);
    }
    if (p.emit_packed && n_emit) {
        const u64 o = st->out_pairs + pre.y;
        if (o + n_emit > p.out_pairs_cap) { if (tid == 0) atomicOr(&st->err, S2P_ERR_PAIRS); }
        else for (u32 i = tid; i < n_emit; i += 256u) ((uint4 *)p.out_pairs)[o + i] = ((const uint4 *)p.ft_pairs)[(size_t)tl * FT_LMAX + i];
    }
    if (p.write_sam && tot.z) {
        const u64 o = st->out_sam + pre.w;
        if (o + tot.z > p.out_sam_cap) { if (tid == 0) atomicOr(&st->err, S2P_ERR_SAM); }
        else {
            const u32 ne = p.ft_nent[tl];
            const uint4 *ent = p.ft_sam + (size_t)tl * FT_LMAX;
            for (u32 i = tid >> 5; i < ne; i += 8u) {
                const uint4 e = ent[i];
                ft_copy_bytes(p.out_sam + o + e.z, p.buf + ws + e.x, e.y, tid & 31u, 32u);
            }
        }
    }
}
