// s2p_fused.cuh — the single-pass sam2pairs path: every SAM byte is read from HBM once.
//
//   k_ft_tile    one CTA per 128 KiB tile of the window (absolute 128 KiB boundaries).  Eight warps stream the tile (16 KiB
//                each, the same 128-bit SWAR newline test as k_scan_chunks) plus 2 KiB on either side; the newline positions
//                go to shared memory, never to HBM.  The lines that start in the tile are then parsed one per thread out of L2
//                (their first 112 bytes, staged in shared memory like k_parse), grouped by QNAME, resolved (resolve_group) and
//                turned into .pairs text, packed pairs and passthrough copy entries in the tile's own scratch segment — no
//                per-line record, line index or group result ever goes to HBM.  A tile depends on no other tile: the group that
//                starts in it is followed into the next tile (halo lines, then a byte-level walk), and whether its first kept
//                line continues a group of the previous tile is decided by looking back the same way.
//   k_ft_prefix  one CTA: exclusive prefixes of the tiles' sizes, window totals, counters committed, self-circle group indices
//   k_ft_gather  one CTA per tile: the scratch segments copied to their final, dense positions (128-bit stores at any
//                source/destination alignment), passthrough lines copied straight from the SAM text
//
// Replaces pairutil.h:136-177 (load_batch: getline, filter, grouping) + flash2pairs.h / unc2pairs.h bodies + the string
// appends of unc2pairs.h:311-356 in one pass.  Anything the tile geometry cannot hold (more than FT_WCAP newlines in a 16 KiB
// chunk, more than FT_LMAX lines or FT_TEXT_CAP bytes of pair text per tile) gives the WINDOW back to the multi-kernel path
// (k_scan_chunks ... k_emit), which handles any input; results are identical by construction (same parsers, same resolver).
#pragma once
#include "s2p_kernels.cuh"

#define FT_THREADS 256
#define FT_WARPS 8
#define FT_CHUNK 16384u
#define FT_TILE (FT_WARPS * FT_CHUNK)
#define FT_HALO 2048u                 // bytes scanned on either side of the tile (4 rows of 512 B)
#define FT_WCAP 384u                  // newline slots per 16 KiB chunk (lines of >= 43 B on average)
#define FT_HCAP 32u                   // newline slots per 512-byte halo row
#define FT_LMAX 1024u                 // line starts per tile (halo included)
#define FT_LOOKBACK 2u                // lines parsed before / after the heads a round resolves
#define FT_LOOKAHEAD 6u
#define FT_TEXT_CAP 49152u            // bytes of pair text per tile in the scratch
#define FT_STAGE_CAP 32752u           // text of one round staged in shared memory (the line-head columns are dead by then)
#define FT_M_END 0x80u                // s_meta: no (complete, parsed) line here

#define FT_OFF_LINE 0
#define FT_OFF_REC 32768
#define FT_OFF_START (FT_OFF_REC + 12288)
#define FT_OFF_A (FT_OFF_START + 4 * (FT_LMAX + 4))
#define FT_OFF_META (FT_OFF_A + 2048)
#define FT_OFF_MISC (FT_OFF_META + 256)
#define FT_SMEM (FT_OFF_MISC + 512)

static_assert(FT_WARPS * FT_WCAP * 2 + FT_WARPS * FT_HCAP * 2 <= 12288, "newline lists alias the record array");
static_assert(sizeof(LineRec) * 256 == 12288, "record array");

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

struct FtGroup { Resolved r; u32 sam_len, n_members; bool off_end; };
// The whole group of the head line at `a_head`, line by line from global memory (any number of lines, any line length).
// entries != nullptr: also writes one passthrough copy entry per member (src relative to ws, length with '\n', destination).
static __device__ __noinline__ void ft_group_slow(const S2PParams &p, u64 ws, u64 we, u64 a_head, FtGroup &out, uint4 *entries, u32 ent_cap, u32 dst0) {
    LineRec f[2], r1[2], r2[2], rec;
    u32 n = 0, n1 = 0, n2 = 0, sam_len = 0;
    out.off_end = false;
    u64 cur = a_head;
    u64 e = ft_find_nl(p.buf, cur, we);
    u32 meta = parse_line_slow_abs(p, cur, false, 0, rec);
    (void)meta;
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
            const u32 m2 = parse_line_slow_abs(p, nxt, true, a_head, rec);
            cur = nxt; e = e2;
            if (!(m2 & LM_KEEP)) continue;
            more = (m2 & LM_EQ) != 0;
            break;
        }
        if (out.off_end || !more) break;
    }
    out.sam_len = sam_len; out.n_members = n;
    if (!out.off_end) out.r = resolve_group(p, n, n1, n2, &f[0], &f[1], &r1[0], &r1[1], &r2[0], &r2[1]);
}

// ---------------------------------------------------------------------------------------------- scan of 512-byte rows
// Newline offsets (relative to rbase) of `nrows` rows appended, in byte order, to a shared-memory list; returns their number
// (which may exceed cap: the caller gives the window up).  Same tests as k_scan_chunks.
template <int U>
__device__ __forceinline__ u32 ft_scan_rows(const char *buf, long long rbase, u32 nrows, u64 ws, u64 we, u16 *list, u32 cap, u32 lane) {
    const bool edge = rbase < (long long)ws || rbase + (long long)nrows * 512 > (long long)we;
    const uint4 *src = (const uint4 *)(buf + rbase) + lane;
    const u32 lt = (1u << lane) - 1u;
    u32 n = 0;
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
            const u32 rel = (it + u) * 512u + lane * 16u;
            if (multi == 0) {
                if (z) {
                    const u32 idx = n + __popc(bal & lt);
                    if (idx < cap) list[idx] = (u16)(rel + byte_of_perm_bit(__ffs(z) - 1));
                }
                n += __popc(bal);
                continue;
            }
            const u32 c = __popc(z);
            const u32 inc = warp_incl_scan(c, (int)lane);
            const u32 tot = __shfl_sync(0xFFFFFFFFu, inc, 31);
            if (c) {
                u32 m = 0, idx = n + inc - c;
#pragma unroll 1
                while (z) { const u32 b = __ffs(z) - 1; z &= z - 1; m |= 1u << byte_of_perm_bit(b); }
#pragma unroll 1
                while (m) { const u32 q = __ffs(m) - 1; m &= m - 1; if (idx < cap) list[idx] = (u16)(rel + q); ++idx; }
            }
            n += tot;
        }
    }
    return n;
}

// ---------------------------------------------------------------------------------------------- the tile kernel
// s_misc layout
#define FM_OVF 0
#define FM_GROUPS 1
#define FM_EMIT 2
#define FM_TEXT 3
#define FM_SAM 4
#define FM_NENT 5
#define FM_CNT 8          // ST_NCOUNTER counters
#define FM_SEG 24         // 16 segment counts
#define FM_SCAN 40        // 4 x 8 warp totals

static __global__ void __launch_bounds__(FT_THREADS, 3) k_ft_tile(S2PParams p) {
    extern __shared__ __align__(16) unsigned char ft_smem[];
    u32 (*s_line)[256] = (u32 (*)[256])(ft_smem + FT_OFF_LINE);       // line heads, one column of 32-bit words per thread; text stage later
    LineRec *s_rec = (LineRec *)(ft_smem + FT_OFF_REC);               // this round's records; the newline lists before the first round
    u16 *s_own = (u16 *)(ft_smem + FT_OFF_REC);
    u16 *s_halo = s_own + FT_WARPS * FT_WCAP;
    u32 *s_start = (u32 *)(ft_smem + FT_OFF_START);                   // line starts relative to ebase, ascending
    u64 *s_A = (u64 *)(ft_smem + FT_OFF_A);
    u8 *s_meta = (u8 *)(ft_smem + FT_OFF_META);
    u32 *s_misc = (u32 *)(ft_smem + FT_OFF_MISC);
    WinState *st = p.st;
    if (st->halt || st->path_old) return;
    const u64 ws = st->ws, we = st->we;
    if (we <= ws) return;
    const u64 tile = ws / FT_TILE + blockIdx.x;
    if (tile > (we - 1) / FT_TILE || blockIdx.x >= p.n_tiles_cap) return;
    const u64 limit = st->total;
    const u32 tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const u64 tbase = tile * FT_TILE, tend = tbase + FT_TILE;
    const long long ebase = (long long)tbase - (long long)FT_HALO;    // positions below are relative to ebase
    if (tid < 128) s_misc[tid] = 0;
#pragma unroll
    for (int j = 28; j < LF_WORDS; ++j) s_line[j][tid] = 0;           // pad words behind the 112 staged bytes
    __syncthreads();

    // ---- 1. newline scan: own 16 KiB chunk, then one halo row per warp (warps 0-3 before the tile, 4-7 behind it)
    {
        const u32 n_own = ft_scan_rows<8>(p.buf, (long long)tbase + (long long)wid * FT_CHUNK, FT_CHUNK / 512u, ws, we, s_own + wid * FT_WCAP, FT_WCAP, lane);
        const long long hb = wid < 4 ? ebase + (long long)wid * 512 : (long long)tend + (long long)(wid - 4) * 512;
        const u32 n_halo = ft_scan_rows<1>(p.buf, hb, 1u, ws, we, s_halo + wid * FT_HCAP, FT_HCAP, lane);
        if (lane == 0) {
            s_misc[FM_SEG + 4 + wid] = n_own;
            s_misc[FM_SEG + (wid < 4 ? wid : 8 + wid)] = n_halo;
            if (n_own > FT_WCAP || n_halo > FT_HCAP) s_misc[FM_OVF] = 1;
        }
    }
    __syncthreads();
    // ---- 2. dense, ascending list of line starts (a start at `we` is kept as the end marker of the last complete line)
    const u32 has_initial = ebase <= (long long)ws ? 1u : 0u;           // the window's first line starts at ws
    u32 Ltot, n_own_nl = 0;
    {
        u32 pre[17]; pre[0] = has_initial;
#pragma unroll
        for (int s = 0; s < 16; ++s) { const u32 c = s_misc[FM_SEG + s]; pre[s + 1] = pre[s] + c; if (s >= 4 && s < 12) n_own_nl += c; }
        Ltot = pre[16];
        if (Ltot > FT_LMAX) { if (tid == 0) s_misc[FM_OVF] = 1; }
        else if (!s_misc[FM_OVF]) {
            if (tid == 0 && has_initial) s_start[0] = (u32)((long long)ws - ebase);
            // every warp copies the two lists it produced
            const u32 so = 4 + wid, sh = wid < 4 ? wid : 8 + wid;
            const u32 bo = FT_HALO + wid * FT_CHUNK, bh = wid < 4 ? wid * 512u : FT_HALO + FT_TILE + (wid - 4) * 512u;
            u32 po = 0, ph = 0, co = 0, ch = 0;
#pragma unroll
            for (int s = 0; s < 16; ++s) { if ((u32)s == so) { po = pre[s]; co = pre[s + 1] - pre[s]; } if ((u32)s == sh) { ph = pre[s]; ch = pre[s + 1] - pre[s]; } }
            for (u32 i = lane; i < co; i += 32) s_start[po + i] = bo + s_own[wid * FT_WCAP + i] + 1u;
            for (u32 i = lane; i < ch; i += 32) s_start[ph + i] = bh + s_halo[wid * FT_HCAP + i] + 1u;
        }
        if (tid == 0 && n_own_nl) {                                    // end of the window's last complete line (a window without kept records)
            u32 s = 11; while (s_misc[FM_SEG + s] == 0) --s;
            const u64 x = tbase + (u64)(s - 4) * FT_CHUNK + s_own[(s - 4) * FT_WCAP + (s_misc[FM_SEG + s] <= FT_WCAP ? s_misc[FM_SEG + s] - 1 : 0)];
            if (!s_misc[FM_OVF]) atomicMax((unsigned long long *)&st->ft_last_end, (unsigned long long)(x + 1));
        }
    }
    __syncthreads();
    if (s_misc[FM_OVF]) { if (tid == 0) atomicOr(&st->path_old, 1u); return; }
    // own lines = starts in [tbase, tend): dense indices [j0, j1)
    u32 j0, j1;
    {
        u32 lo = 0, hi = Ltot;
        while (lo < hi) { const u32 m = (lo + hi) >> 1; if (s_start[m] < FT_HALO) lo = m + 1; else hi = m; }
        j0 = lo; hi = Ltot;
        while (lo < hi) { const u32 m = (lo + hi) >> 1; if (s_start[m] < FT_HALO + FT_TILE) lo = m + 1; else hi = m; }
        j1 = lo;
    }
    const u32 tl = blockIdx.x;                                         // window-local tile index
    char *sc_text = p.ft_text + (size_t)tl * FT_TEXT_CAP;
    mk_pair *sc_pairs = p.ft_pairs + (size_t)tl * FT_LMAX;
    uint4 *sc_sam = p.ft_sam + (size_t)tl * FT_LMAX;
    __syncthreads();                                                   // the newline lists (aliasing s_rec) are dead from here

    // ---- 3. rounds: 256 lines parsed (one per thread), heads in [lo, hi) resolved and emitted
    u32 base = j0 >= FT_LOOKBACK ? j0 - FT_LOOKBACK : 0u;
    u32 lo = j0;
    while (lo < j1) {
        const bool last_round = base + 256u >= Ltot;
        const u32 hi = last_round ? j1 : min(j1, base + 256u - FT_LOOKAHEAD);
        const u32 li = base + tid;
        // -- parse.  A dense entry is a complete line iff it has a successor; the last entry is only known to be a line start.
        const bool parsed = li + 1 < Ltot;
        u64 a = 0;
        LineFetch lf; lf.buf = p.buf; lf.col = &s_line[0][tid]; lf.A = 0;
        bool staged = false;
        if (parsed) {
            a = (u64)(ebase + (long long)s_start[li]);
            lf.A = a & ~(u64)15;
            staged = a + 144 <= limit;
            if (staged) {
                const uint4 *src = (const uint4 *)(p.buf + lf.A);
                uint4 w[7];
#pragma unroll
                for (int j = 0; j < 7; ++j) w[j] = __ldg(src + j);
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    s_line[4 * j][tid] = w[j].x; s_line[4 * j + 1][tid] = w[j].y; s_line[4 * j + 2][tid] = w[j].z; s_line[4 * j + 3][tid] = w[j].w;
                    lf.w[j] = w[j];
                }
            }
        }
        s_A[tid] = staged ? lf.A : ~(u64)0;
        __syncthreads();
        u32 meta = FT_M_END;
        LineRec rec;
        if (parsed) {
            const bool has_prev = li > 0;
            const u64 pa = has_prev ? (u64)(ebase + (long long)s_start[li - 1]) : 0;
            FastTok tok;
            meta = 0;
            if (staged && parse_line_fast<LineFetch, false>(p, lf, a, limit, tok, rec, meta)) {
                if (has_prev) {
                    bool eq;
                    if (is_blank((int)(unsigned char)p.buf[pa])) eq = qname_equal_abs(p, a, pa);
                    else if (tid > 0 && s_A[tid - 1] != ~(u64)0 && (u32)(pa - s_A[tid - 1]) + tok.t0 + 1 <= 112) {
                        LineFetch lp; lp.buf = p.buf; lp.col = &s_line[0][tid - 1]; lp.A = s_A[tid - 1];
                        const u32 so = (u32)(a - lf.A), sp = (u32)(pa - lp.A);
                        eq = true;
                        for (u32 k = 0; k < tok.t0 && eq; k += 8) {
                            u64 x = fetch8r(lf, so + k), y = fetch8r(lp, sp + k);
                            if (tok.t0 - k < 8) { const u64 m = (1ull << (8 * (tok.t0 - k))) - 1; x &= m; y &= m; }
                            eq = x == y;
                        }
                        eq = eq && is_ws(lp.byter(sp + tok.t0));
                    } else {
                        GlobalFetch gf; gf.buf = p.buf; gf.A = 0;
                        eq = qname_eq_fetch(gf, a, pa, tok.t0);
                    }
                    if (eq) meta |= LM_EQ;
                }
            } else meta = parse_line_slow_abs(p, a, has_prev, pa, rec);
            if (!has_prev && !has_initial) meta |= LM_EQ_UNK;          // the line before the first one we know of
            if (meta & LM_KEEP) s_rec[tid] = rec;
        }
        s_meta[tid] = (u8)meta;
        __syncthreads();

        // -- group: is this line the head of a read group, and what does the group resolve to?
        bool proc = false, emit = false, use_slow = false;
        Resolved rs; rs.status = ST_NONE; rs.have = false; rs.p1 = rs.p2 = 0; rs.sA = rs.sB = 0; rs.strands = 0;
        u32 text_len = 0, sam_len = 0, n_mem = 0, rid_len = 0, rid_off = 0;
        if (li >= lo && li < hi) {
            bool kept = (meta & LM_KEEP) != 0 && parsed;
            if (!parsed && li < Ltot) {                                // the list's last entry: a line only if it ends before `we`
                a = (u64)(ebase + (long long)s_start[li]);
                if (a < we && ft_find_nl(p.buf, a, we) != FT_NONE) {
                    const bool has_prev = li > 0;
                    meta = parse_line_slow_abs(p, a, has_prev, has_prev ? (u64)(ebase + (long long)s_start[li - 1]) : 0, rec);
                    if (!has_prev && !has_initial) meta |= LM_EQ_UNK;
                    kept = (meta & LM_KEEP) != 0;
                    use_slow = true;
                }
            }
            if (kept) {
                // head?  (pairutil.h:163-173: currId != lastId among kept records)
                int head;                                              // 0 no, 1 yes, 2 walk the bytes
                {
                    bool chain = (meta & LM_EQ) != 0;
                    head = (meta & LM_EQ_UNK) ? 2 : -1;
                    int j = (int)tid - 1;
                    while (head < 0 && j >= 0) {
                        const u32 mj = s_meta[j];
                        if (mj & LM_KEEP) break;
                        if (mj & LM_EQ_UNK) { head = 2; break; }
                        chain = chain && (mj & LM_EQ);
                        --j;
                    }
                    if (head < 0) {
                        if (j < 0) head = (base == 0 && has_initial) ? 1 : 2;
                        else if (chain) head = 0;
                        else if (j == (int)tid - 1) head = 1;
                        else head = qname_equal_abs(p, a, (u64)(ebase + (long long)s_start[base + j])) ? 0 : 1;
                    }
                    if (head == 2) head = ft_head_slow(p, ws, a) ? 1 : 0;
                }
                if (head) {
                    u32 first[2] = {tid, tid}, r1[2] = {tid, tid}, r2[2] = {tid, tid};
                    u32 n = 0, n1 = 0, n2 = 0;
                    if (!use_slow) {
                        u32 k = tid;
                        while (true) {
                            const u32 fl = s_rec[k].flag;
                            if (n < 2) first[n] = k;
                            ++n;
                            if (fl & 64u) { if (n1 < 2) r1[n1] = k; ++n1; } else if (fl & 128u) { if (n2 < 2) r2[n2] = k; ++n2; }
                            sam_len += s_start[base + k + 1] - s_start[base + k];
                            u32 q = k + 1; bool chain = true; u32 mq = FT_M_END;
                            while (q < 256u) { mq = s_meta[q]; if (mq & FT_M_END) break; chain = chain && (mq & LM_EQ); if (mq & LM_KEEP) break; ++q; }
                            if (q >= 256u || (mq & FT_M_END)) { use_slow = true; break; }
                            const bool same = chain ? true : (q == k + 1 ? false :
                                qname_equal_abs(p, (u64)(ebase + (long long)s_start[base + q]), (u64)(ebase + (long long)s_start[base + k])));
                            if (!same) break;
                            k = q;
                        }
                        n_mem = n;
                    }
                    if (use_slow) {
                        FtGroup g;
                        ft_group_slow(p, ws, we, a, g, nullptr, 0, 0);
                        if (g.off_end) {                               // the window's last group: carried to the next window (or dropped at EOF)
                            st->ft_carry_pos = a; st->ft_carry_tile = tl;
                            u32 c = 0;                                 // newlines of this tile before the line
                            for (u32 k = has_initial; k <= li; ++k) if (s_start[k] > FT_HALO) ++c;
                            st->ft_carry_nl = c;
                        } else { proc = true; rs = g.r; sam_len = g.sam_len; n_mem = g.n_members; }
                    } else {
                        proc = true;
                        rs = resolve_group(p, n, n1, n2, &s_rec[first[0]], &s_rec[first[1]], &s_rec[r1[0]], &s_rec[r1[1]], &s_rec[r2[0]], &s_rec[r2[1]]);
                    }
                    if (proc) {
                        rid_len = parsed ? s_rec[tid].qname_len : rec.qname_len;
                        rid_off = parsed ? s_rec[tid].qname_off : rec.qname_off;
                        if (rs.status != ST_NONE) atomicAdd(&s_misc[FM_CNT + rs.status], 1u);
                        if (rs.have && rs.status != ST_SELFCIRCLE) {
                            emit = true;
                            text_len = rid_len + p.chr[rs.sA].len + p.chr[rs.sB].len + dec_digits(rs.p1) + dec_digits(rs.p2) + 9u;
                        }
                    }
                }
            }
        }
        if (!emit || !p.write_sam) { sam_len = 0; n_mem = 0; }           // (also the partial sums of a walk that went to the byte level)
        // -- offsets inside the round: (groups | emitted << 16), text bytes, passthrough bytes, passthrough entries
        u32 vA = (proc ? 1u : 0u) | (emit ? 1u << 16 : 0u), vT = p.emit_text ? text_len : 0u, vS = sam_len, vE = n_mem;
        const u32 iA = warp_incl_scan(vA, (int)lane), iT = warp_incl_scan(vT, (int)lane), iS = warp_incl_scan(vS, (int)lane), iE = warp_incl_scan(vE, (int)lane);
        __syncthreads();                                               // every thread is done with s_line / s_rec of this round
        if (lane == 31) { s_misc[FM_SCAN + wid] = iA; s_misc[FM_SCAN + 8 + wid] = iT; s_misc[FM_SCAN + 16 + wid] = iS; s_misc[FM_SCAN + 24 + wid] = iE; }
        __syncthreads();
        u32 bA = iA - vA, bT = iT - vT, bS = iS - vS, bE = iE - vE, tA = 0, tT = 0, tS = 0, tE = 0;
#pragma unroll
        for (u32 w = 0; w < 8; ++w) {
            const u32 xa = s_misc[FM_SCAN + w], xt = s_misc[FM_SCAN + 8 + w], xs = s_misc[FM_SCAN + 16 + w], xe = s_misc[FM_SCAN + 24 + w];
            tA += xa; tT += xt; tS += xs; tE += xe;
            if (w < wid) { bA += xa; bT += xt; bS += xs; bE += xe; }
        }
        const u32 run_g = s_misc[FM_GROUPS], run_e = s_misc[FM_EMIT], run_t = s_misc[FM_TEXT], run_s = s_misc[FM_SAM], run_n = s_misc[FM_NENT];
        const bool fits = run_t + tT <= FT_TEXT_CAP && run_n + tE <= FT_LMAX && run_e + (tA >> 16) <= FT_LMAX;
        if (!fits) { if (tid == 0) atomicOr(&st->path_old, 1u); return; }   // uniform: the window goes to the multi-kernel path
        const bool stage = tT <= FT_STAGE_CAP;
        const u32 phase = (u32)((size_t)(sc_text + run_t) & 15u);
        char *s_stage = (char *)ft_smem;
        if (proc) {
            if (rs.status == ST_SELFCIRCLE) {                          // (tile, group index inside the tile): k_ft_prefix makes it global
                const u32 slot = atomicAdd(&st->sc_count, 1u);
                if (slot < p.sc_cap) p.sc_list[slot] = ((u64)tl << 32) | (u64)(run_g + (bA & 0xFFFFu)); else atomicOr(&st->err, S2P_ERR_SCLIST);
            }
            if (emit) {
                const ChrSlot *ca = &p.chr[rs.sA], *cb = &p.chr[rs.sB];
                if (p.emit_packed) {
                    mk_pair r; r.pos1 = rs.p1; r.pos2 = rs.p2; r.chr1 = (u16)ca->id; r.chr2 = (u16)cb->id; r.strands = rs.strands;
                    r.cls = (u8)(rs.status - ST_TRANS); r.lane = p.lane;
                    sc_pairs[run_e + (bA >> 16)] = r;
                }
                if (p.emit_text) {
                    RidInfo rid; rid.abs = a + rid_off; rid.len = rid_len;
                    write_pair_line_slots(p, rid, ca, cb, rs.p1, rs.p2, rs.strands, stage ? s_stage + phase + bT : sc_text + run_t + bT);
                }
                if (p.write_sam) {                                     // one copy entry per kept line of the group
                    uint4 *ent = sc_sam + run_n + bE;
                    if (use_slow) { FtGroup g; ft_group_slow(p, ws, we, a, g, ent, n_mem, run_s + bS); }
                    else {
                        u32 k = tid, d = run_s + bS, c = 0;
                        while (true) {
                            const u32 len = s_start[base + k + 1] - s_start[base + k];
                            ent[c++] = make_uint4((u32)((u64)(ebase + (long long)s_start[base + k]) - ws), len, d, 0u);
                            d += len;
                            if (c == n_mem) break;
                            ++k; while (!(s_meta[k] & LM_KEEP)) ++k;
                        }
                    }
                }
            }
        }
        if (p.emit_text && tT && stage) {
            __syncthreads();
            char *dst = sc_text + run_t;
            const u32 head = phase ? (16u - phase < tT ? 16u - phase : tT) : 0u;
            if (tid < head) dst[tid] = s_stage[phase + tid];
            const u32 body = (tT - head) >> 4;
            for (u32 w = tid; w < body; w += FT_THREADS)
                *(uint4 *)(dst + head + ((size_t)w << 4)) = *(const uint4 *)(s_stage + phase + head + (w << 4));
            const u32 tail0 = head + (body << 4);
            if (tail0 + tid < tT) dst[tail0 + tid] = s_stage[phase + tail0 + tid];
        }
        __syncthreads();
        if (tid == 0) {
            s_misc[FM_GROUPS] = run_g + (tA & 0xFFFFu); s_misc[FM_EMIT] = run_e + (tA >> 16); s_misc[FM_TEXT] = run_t + tT;
            s_misc[FM_SAM] = run_s + tS; s_misc[FM_NENT] = run_n + tE;
        }
        if (stage && tT) {                                             // the stage overwrote the pad words of the line columns
#pragma unroll
            for (int j = 28; j < LF_WORDS; ++j) s_line[j][tid] = 0;
        }
        __syncthreads();
        lo = hi;
        base = hi - FT_LOOKBACK;
    }
    __syncthreads();
    if (tid == 0) {
        p.ft_tot[tl] = make_uint4(s_misc[FM_GROUPS] | (s_misc[FM_EMIT] << 16), s_misc[FM_TEXT], s_misc[FM_SAM], n_own_nl);
        p.ft_nent[tl] = s_misc[FM_NENT];
    }
    if (tid < ST_NCOUNTER && s_misc[FM_CNT + tid]) atomicAdd(&st->w_counters[tid], (unsigned long long)s_misc[FM_CNT + tid]);
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
