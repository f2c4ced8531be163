// dedup.cu — krmdup on the GPU: FASTQ-level duplicate removal keyed on 2-bit packed read bases.
// Replaces src/preprocess/krmdup.cpp (and krmdup.pipe.cpp, same algorithm) of the reference:
//   load_batch :88-149   -> k_scan_lines (newline index) + k_fq_keys (lengths, first key base, bucket)
//   key packing :168-193 -> k_fq_keys
//   set probe :201-212   -> radix sort of (key, index) inside the window, first of every run probes a persistent
//                           device hash set (the reference's unordered_sets live for the whole process, :322)
//   ordered write :206-226 -> k_fq_layout (per 65 536-pair batch: buckets A, C, G, T) + k_fq_copy
//   log :383-389         -> mk_dedup_finish
#include <algorithm>
#include <cstring>
#include <deque>
#include <vector>
#include <chrono>
#include "ctx.h"
#include "s2p_kernels.cuh"
#include "radix_sort.cuh"

#define FQ_BATCH 65536u            // krmdup.cpp:19
#define FQ_ERR_LINES 1u
#define FQ_ERR_OUT 2u
#define FQ_ERR_TABLE 4u
#define HS_EMPTY 0xFFFFFFFFFFFFFFFFull

struct FqState {
    u64 total; u32 n_lines, err, is_last, pad;
    u64 n_pairs, consumed, pair_base;
    u64 out1, out2;
    unsigned long long uniq, dup, discard;
    u32 allones_seen[2], inserted[2];
    u32 ticket, pad2;
};

struct FqParams {
    const char *buf; FqState *st;
    u32 *nl_pos; u32 cap_lines; u64 *desc;
    uint4 *rec0, *rec1;            // sort records {key lo, key hi, pair idx, tag}
    u8 *cls;                       // per pair: bits0-1 bucket, bit2 valid, bit3 keep
    u32 *eoff;                     // per pair: [2p] / [2p+1] exclusive kept bytes of the pair's own bucket (mate 1 / mate 2)
    u32 *btab;                     // per batch (+1): 8 exclusive byte counts (4 buckets x 2 mates) at the batch start
    u64 *hset[2]; u64 hmask[2];
    char *out1, *out2; u64 out_cap;
    int hskip1, klen1, hskip2, klen2;
};

__global__ void k_fq_begin(FqParams p, u32 n_desc, u64 total, u32 is_last) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_desc) p.desc[i] = 0;
    if (i == 0) { p.st->total = total; p.st->n_lines = 0; p.st->is_last = is_last; p.st->n_pairs = 0; p.st->consumed = 0; p.st->out1 = p.st->out2 = 0; p.st->ticket = 0; }
}

__global__ void __launch_bounds__(S2P_SCAN_THREADS) k_fq_scan(FqParams p) {
    scan_lines_body<4>(p.buf, 0, p.st->total, p.nl_pos, p.cap_lines, p.desc, &p.st->n_lines, &p.st->err, FQ_ERR_LINES);
}

__device__ __forceinline__ u32 fq_line_start(const u32 *nl, u32 line) { return line ? nl[line - 1] + 1 : 0; }

// thread per pair: validity, bucket, key (krmdup.cpp:103-141,156-198)
__global__ void __launch_bounds__(256) k_fq_keys(FqParams p) {
    FqState *st = p.st;
    const u32 n_lines = st->n_lines;
    u64 np = n_lines / 8;
    if (!st->is_last) np = (np / FQ_BATCH) * FQ_BATCH;            // whole batches only: the output order is per batch
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->n_pairs = np;
        st->consumed = np ? (u64)p.nl_pos[np * 8 - 1] + 1 : 0;
    }
    const u32 e1 = (u32)(p.hskip1 + p.klen1), e2 = (u32)(p.hskip2 + p.klen2);
    u32 n_disc = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (u64)gridDim.x * blockDim.x) {
        const u32 l0 = (u32)i * 8;
        const u32 s1 = fq_line_start(p.nl_pos, l0 + 1), len1 = p.nl_pos[l0 + 1] - s1;
        const u32 s2 = fq_line_start(p.nl_pos, l0 + 5), len2 = p.nl_pos[l0 + 5] - s2;
        const char first = len1 >= e1 ? p.buf[s1 + p.hskip1] : 'N';
        u32 bucket = first == 'A' ? 0u : first == 'C' ? 1u : first == 'G' ? 2u : 3u;
        bool valid = first != 'N' && len1 >= e1 && len2 >= e2;
        u64 key = 0;
        if (valid) {
            for (u32 k = (u32)p.hskip1; k < e1 && valid; ++k) {
                const char c = p.buf[s1 + k] & 0xDF;                  // upper-case: the reference accepts both cases
                const u64 code = c == 'A' ? 1 : c == 'T' ? 2 : c == 'C' ? 0 : c == 'G' ? 3 : 4;
                valid = code < 4; key = (key << 2) | (code & 3);
            }
            for (u32 k = (u32)p.hskip2; k < e2 && valid; ++k) {
                const char c = p.buf[s2 + k] & 0xDF;
                const u64 code = c == 'A' ? 1 : c == 'T' ? 2 : c == 'C' ? 0 : c == 'G' ? 3 : 4;
                valid = code < 4; key = (key << 2) | (code & 3);
            }
        }
        // the T bucket also takes lower-case first bases (krmdup.cpp:134-141): its set is distinct from A/C/G's,
        // so such keys get their own identity space (tag 1)
        const u32 top = (u32)(key >> (2 * (p.klen1 + p.klen2) - 2)) & 3u;     // code of the first key base
        const u32 tag = !valid ? 2u : (bucket == 3u && top != 2u) ? 1u : 0u;
        p.rec0[i] = make_uint4((u32)key, (u32)(key >> 32), (u32)i, tag);
        p.cls[i] = (u8)(bucket | (valid ? 4u : 0u));
        n_disc += !valid;
    }
    n_disc = __reduce_add_sync(0xffffffffu, n_disc);
    if ((threadIdx.x & 31) == 0 && n_disc) atomicAdd(&st->discard, (unsigned long long)n_disc);
}

__device__ __forceinline__ u64 hs_hash(u64 k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k;
}
// returns true when the key was not in the set (and inserts it).  The probe is bounded by the table size: a full table
// raises *full instead of spinning (the host sizes the tables so that this cannot happen; it is the safety net).
__device__ __forceinline__ bool hs_insert(u64 *tab, u64 mask, u64 key, u32 *full) {
    u64 s = hs_hash(key) & mask;
    for (u64 probe = 0; probe <= mask; ++probe) {
        u64 cur = tab[s];
        if (cur == key) return false;
        if (cur == HS_EMPTY) {
            u64 old = atomicCAS((unsigned long long *)&tab[s], (unsigned long long)HS_EMPTY, (unsigned long long)key);
            if (old == HS_EMPTY) return true;
            if (old == key) return false;
        }
        s = (s + 1) & mask;
    }
    atomicOr(full, FQ_ERR_TABLE);
    return false;
}

// sorted records: the first of every (key, tag) run is this window's earliest occurrence; it survives iff the key was
// never seen in earlier windows (krmdup.cpp:201-212)
__global__ void __launch_bounds__(256) k_fq_mark(FqParams p, const RadixPlan *plan) {
    FqState *st = p.st;
    const u64 np = st->n_pairs;
    const uint4 *r = plan->final_buf ? p.rec1 : p.rec0;
    u32 n_uniq = 0, n_dup = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (u64)gridDim.x * blockDim.x) {
        const uint4 a = r[i];
        if (a.w == 2u) continue;
        bool first = true;
        if (i > 0) { const uint4 b = r[i - 1]; first = a.x != b.x || a.y != b.y || a.w != b.w; }
        bool keep = false;
        if (first) {
            const u64 key = (u64)a.x | ((u64)a.y << 32);
            if (key == HS_EMPTY) keep = atomicExch(&st->allones_seen[a.w], 1u) == 0u;      // poly-G key: the table's empty marker
            else { keep = hs_insert(p.hset[a.w], p.hmask[a.w], key, &st->err); if (keep) atomicAdd(&st->inserted[a.w], 1u); }
        }
        if (keep) { p.cls[a.z] |= 8u; ++n_uniq; } else ++n_dup;
    }
    n_uniq = __reduce_add_sync(0xffffffffu, n_uniq); n_dup = __reduce_add_sync(0xffffffffu, n_dup);
    if ((threadIdx.x & 31) == 0) { if (n_uniq) atomicAdd(&st->uniq, (unsigned long long)n_uniq); if (n_dup) atomicAdd(&st->dup, (unsigned long long)n_dup); }
}

__global__ void k_hs_rehash(const u64 *old_tab, u64 old_slots, u64 *new_tab, u64 new_mask, u32 *err) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < old_slots; i += (u64)gridDim.x * blockDim.x) {
        u64 k = old_tab[i];
        if (k != HS_EMPTY) hs_insert(new_tab, new_mask, k, err);
    }
}

// output layout: per batch, kept pairs of bucket A, then C, G, T, each in input order (krmdup.cpp:216-226,328-364).
// One look-back scan of 8 byte counters (4 buckets x 2 mates); the batch table samples it at batch starts.
#define FL_T 256
__global__ void __launch_bounds__(FL_T) k_fq_layout(FqParams p, u64 *descs /* 4 arrays of n_tiles */, u32 n_tiles_cap) {
    __shared__ u32 s_scan[8][FL_T / 32 + 1];
    __shared__ u64 s_base[4];
    FqState *st = p.st;
    const u64 np = st->n_pairs;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n_tiles = (int)((np + FL_T - 1) / FL_T);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const u64 i = (u64)tile * FL_T + tid;
        u32 v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        u32 bucket = 0; bool keep = false;
        if (i < np) {
            const u32 c = p.cls[i];
            bucket = c & 3u; keep = (c & 8u) != 0;
            if (keep) {
                const u32 l0 = (u32)i * 8;
                const u32 a0 = fq_line_start(p.nl_pos, l0), a1 = p.nl_pos[l0 + 1], q0 = fq_line_start(p.nl_pos, l0 + 3), q1 = p.nl_pos[l0 + 3];
                const u32 b0 = fq_line_start(p.nl_pos, l0 + 4), b1 = p.nl_pos[l0 + 5], r0 = fq_line_start(p.nl_pos, l0 + 7), r1 = p.nl_pos[l0 + 7];
                v[bucket] = (a1 - a0 + 1) + 2 + (q1 - q0 + 1);            // id\nseq\n  +\n  qual\n
                v[4 + bucket] = (b1 - b0 + 1) + 2 + (r1 - r0 + 1);
            }
        }
        u32 ex[8], tot[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) ex[k] = block_excl_scan<FL_T>(v[k], s_scan[k], &tot[k]);
        if (wid < 4) {
            // chains: (A1,C1) (G1,T1) (A2,C2) (G2,T2), two 31-bit fields per descriptor
            const int k0 = wid * 2;
            u64 agg = (u64)tot[k0] | ((u64)tot[k0 + 1] << 31);
            u64 b = lookback_exclusive(descs + (u64)wid * n_tiles_cap, tile, 0, agg, lane);
            if (lane == 0) s_base[wid] = b;
        }
        __syncthreads();
        u32 E[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) E[k] = (u32)((s_base[k >> 1] >> ((k & 1) ? 31 : 0)) & 0x7FFFFFFFu) + ex[k];
        if (i < np) {
            p.eoff[2 * i] = E[bucket]; p.eoff[2 * i + 1] = E[4 + bucket];
            if ((i % FQ_BATCH) == 0) { u32 *t = p.btab + (i / FQ_BATCH) * 8; for (int k = 0; k < 8; ++k) t[k] = E[k]; }
            if (i == np - 1) {                                           // totals = the entry after the last batch
                u32 *t = p.btab + ((np + FQ_BATCH - 1) / FQ_BATCH) * 8;
                u32 o1 = 0, o2 = 0;
                for (int k = 0; k < 8; ++k) { u32 e = E[k] + v[k]; t[k] = e; if (k < 4) o1 += e; else o2 += e; }
                st->out1 = o1; st->out2 = o2;
                if (o1 > p.out_cap || o2 > p.out_cap) atomicOr(&st->err, FQ_ERR_OUT);
            }
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void warp_copy(char *dst, const char *src, u32 n, int lane) {
    for (u32 b = lane; b < n; b += 32) dst[b] = src[b];
}

// warp per kept pair: "%s\n%s\n+\n%s\n" for each mate (krmdup.cpp:206-209)
__global__ void __launch_bounds__(256) k_fq_copy(FqParams p) {
    const FqState *st = p.st;
    if (st->err & FQ_ERR_OUT) return;
    const u64 np = st->n_pairs;
    const int lane = threadIdx.x & 31;
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 i = warp; i < np; i += nwarps) {
        const u32 c = p.cls[i];
        if (!(c & 8u)) continue;
        const u32 bucket = c & 3u;
        const u32 *t0 = p.btab + (i / FQ_BATCH) * 8, *t1 = t0 + 8;
        for (int m = 0; m < 2; ++m) {
            u32 off = 0;
            for (int k = 0; k < 4; ++k) off += t0[4 * m + k];                       // everything before this batch
            for (u32 k = 0; k < bucket; ++k) off += t1[4 * m + k] - t0[4 * m + k];    // earlier buckets of this batch
            off += p.eoff[2 * i + m] - t0[4 * m + bucket];                          // earlier kept pairs of this bucket
            const u32 l0 = (u32)i * 8 + 4 * m;
            const u32 a0 = fq_line_start(p.nl_pos, l0), a1 = p.nl_pos[l0 + 1];
            const u32 q0 = fq_line_start(p.nl_pos, l0 + 3), q1 = p.nl_pos[l0 + 3];
            char *dst = (m ? p.out2 : p.out1) + off;
            const u32 n0 = a1 - a0 + 1, n1 = q1 - q0 + 1;
            warp_copy(dst, p.buf + a0, n0, lane);
            if (lane == 0) { dst[n0] = '+'; dst[n0 + 1] = '\n'; }
            warp_copy(dst + n0 + 2, p.buf + q0, n1, lane);
        }
    }
}

// ------------------------------------------------------------------------------------------------ host side
struct DedupCtx : mk_ctx {
    mk_dedup_cfg cfg;
    size_t W = 0; u32 cap_lines = 0, cap_pairs = 0, n_desc = 0, n_tiles_cap = 0;
    int sms = 148;
    cudaStream_t s = nullptr;
    DevBuf d_in, d_state, d_nl, d_desc, d_rec0, d_rec1, d_cls, d_eoff, d_btab, d_ldesc, d_out1, d_out2;
    DevBuf d_hset[2]; u64 hslots[2] = {0, 0}; u64 hcount[2] = {0, 0};
    RadixWs rws;
    PinBuf h_in;                                   // pinned: the window being filled by push() (input that is already pinned is DMA'd from the caller's memory)
    size_t fill = 0;                               // bytes of h_in filled so far
    const char *ext_ptr = nullptr; size_t ext_len = 0;   // unconsumed tail of the last push, still in the caller's PINNED memory (fill == 0 then)
    size_t ho_len[2] = {0, 0}, ho_off[2] = {0, 0}; // undrained part of d_out1 / d_out2: the last window's kept records stay on the device until pulled
    // cfg.async_pull: mk_dedup_pull only enqueues its device-to-host copies (stream s_out), so that they run while the next
    // push's host-to-device copy does; bytes [0, fly_end) of the output buffers may still be being read
    cudaStream_t s_out = nullptr; cudaEvent_t ev_out = nullptr; size_t fly_end[2] = {0, 0}; bool fly = false;
    size_t out_cap = 0;                            // bytes of d_out1 / d_out2: 2 W (4 W with async_pull: room behind the copies in flight)
    bool finished = false;
    std::deque<std::vector<char>> q1, q2; size_t q1_off = 0, q2_off = 0;
    u64 pairs_total = 0;
    DedupCtx() { kind = MK_CTX_DEDUP; }
    ~DedupCtx() override { cudaSetDevice(cfg.device); if (s_out) { cudaStreamSynchronize(s_out); cudaStreamDestroy(s_out); } if (ev_out) cudaEventDestroy(ev_out); if (s) cudaStreamDestroy(s); }
};

static double dd_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

extern "C" void mk_dedup_default_cfg(mk_dedup_cfg *c) {
    memset(c, 0, sizeof *c);
    c->hskip1 = 5; c->klen1 = 16; c->hskip2 = 5; c->klen2 = 16; c->device = 0; c->window_bytes = 0;   // krmdup.cpp:231-234
}

static int hs_alloc(DedupCtx *c, int which, u64 slots) {
    DevBuf nb;
    MK_TRY(nb.alloc(slots * 8));
    MK_CUDA(cudaMemsetAsync(nb.p, 0xFF, slots * 8, c->s));
    if (c->d_hset[which].p && c->hcount[which]) {
        k_hs_rehash<<<c->sms * 8, 256, 0, c->s>>>(c->d_hset[which].as<u64>(), c->hslots[which], nb.as<u64>(), slots - 1, &c->d_state.as<FqState>()->err);
        c->launches_generic += 1;
    }
    MK_CUDA(cudaStreamSynchronize(c->s));
    c->d_hset[which].free();
    c->d_hset[which].p = nb.p; c->d_hset[which].n = nb.n; nb.p = nullptr; nb.n = 0;
    c->hslots[which] = slots;
    return MK_OK;
}

extern "C" int mk_dedup_create(const mk_dedup_cfg *cfg, mk_ctx **out) {
    if (!cfg || !out) { mk_set_error("mk_dedup_create: null argument"); return MK_ERR_ARG; }
    if (cfg->klen1 + cfg->klen2 > 32 || cfg->klen1 + cfg->klen2 < 16 || cfg->klen1 < 1 || cfg->klen2 < 0 || cfg->hskip1 < 0 || cfg->hskip2 < 0) {
        mk_set_error("Error: invalid key sizes!"); return MK_ERR_ARG;                              // krmdup.cpp:259-262
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { mk_set_error("no CUDA device: microcket_b200 has no CPU fallback"); return MK_ERR_CUDA; }
    const bool trace = getenv("MICROCKET_TRACE") != nullptr;
    const double t_start = dd_now();
    MK_CUDA(cudaSetDevice(cfg->device));
    MK_CUDA(cudaFree(0));
    if (trace) fprintf(stderr, "[krmdup create] CUDA context %.3f s\n", dd_now() - t_start);
    DedupCtx *c = new DedupCtx();
    c->cfg = *cfg;
    c->W = cfg->window_bytes ? cfg->window_bytes : (size_t)256 << 20;
    if (c->W < (1u << 20)) c->W = 1u << 20;
    if (c->W > ((size_t)2040 << 20)) c->W = (size_t)2040 << 20;
    c->W = (c->W + 15) & ~(size_t)15;
    c->sms = mk_sm_count(cfg->device);
    c->cap_lines = (u32)(c->W / 16 + 1024);
    c->cap_pairs = c->cap_lines / 8 + 8;
    c->n_desc = (u32)(c->W / S2P_TILE_BYTES + 4);
    c->n_tiles_cap = c->cap_pairs / FL_T + 4;
    int rc = MK_OK;
#define A(x) do { if (rc == MK_OK) rc = (x); } while (0)
    A(c->d_in.alloc(c->W + 64)); A(c->d_state.alloc(sizeof(FqState))); A(c->d_nl.alloc((size_t)c->cap_lines * 4));
    A(c->d_desc.alloc((size_t)c->n_desc * 8)); A(c->d_rec0.alloc((size_t)c->cap_pairs * 16)); A(c->d_rec1.alloc((size_t)c->cap_pairs * 16));
    A(c->d_cls.alloc(c->cap_pairs)); A(c->d_eoff.alloc((size_t)c->cap_pairs * 8)); A(c->d_btab.alloc((size_t)(c->cap_pairs / FQ_BATCH + 4) * 32));
    A(c->d_ldesc.alloc((size_t)c->n_tiles_cap * 4 * 8)); c->out_cap = (cfg->async_pull ? 4 : 2) * c->W;
    A(c->d_out1.alloc(c->out_cap + 64)); A(c->d_out2.alloc(c->out_cap + 64));
    A(c->rws.alloc(c->cap_pairs));
    if (trace) fprintf(stderr, "[krmdup create] + device buffers %.3f s\n", dd_now() - t_start);
    A(c->h_in.alloc(c->W + 64));
    if (trace) fprintf(stderr, "[krmdup create] + pinned buffers %.3f s\n", dd_now() - t_start);
#undef A
    if (rc != MK_OK) { delete c; return rc; }
    cudaStreamCreateWithFlags(&c->s, cudaStreamNonBlocking);
    if (cfg->async_pull) { cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking); cudaEventCreateWithFlags(&c->ev_out, cudaEventDisableTiming); }
    MK_CUDA(cudaMemsetAsync(c->d_state.p, 0, sizeof(FqState), c->s));
    MK_CUDA(cudaMemsetAsync(c->d_in.p, '\n', c->W + 64, c->s));
    rc = hs_alloc(c, 0, 1ull << 22);
    if (rc == MK_OK) rc = hs_alloc(c, 1, 1ull << 16);
    if (rc != MK_OK) { delete c; return rc; }
    *out = c;
    return MK_OK;
}

static int dd_check(mk_ctx *x, DedupCtx **c) {
    if (!x || x->kind != MK_CTX_DEDUP) { mk_set_error("not a krmdup context"); return MK_ERR_ARG; }
    *c = (DedupCtx *)x;
    cudaSetDevice((*c)->cfg.device);
    return MK_OK;
}

static bool dd_ptr_is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// one window of complete lines: the first n_stage bytes of h_in followed by n_direct bytes DMA'd straight from `direct`
// (caller memory that is already pinned); returns bytes consumed (whole batches unless last)
static int dd_window(DedupCtx *c, size_t n_stage, const char *direct, size_t n_direct, bool is_last, size_t *consumed) {
    cudaStream_t s = c->s;
    const size_t n = n_stage + n_direct;
    // kept records of earlier windows that nobody pulled yet (several windows inside one push): this window's records are
    // appended behind them while the output buffers (2 W each; a window's output is smaller than its input) have room;
    // only then are they moved off the device before the kernels overwrite them
    const bool pending = c->ho_off[0] < c->ho_len[0] || c->ho_off[1] < c->ho_len[1];
    const bool append = pending && c->ho_len[0] + n <= c->out_cap && c->ho_len[1] + n <= c->out_cap;
    for (int w = 0; w < 2 && !append; ++w) {
        if (c->ho_off[w] < c->ho_len[w]) {
            std::vector<char> v(c->ho_len[w] - c->ho_off[w]);
            MK_CUDA(cudaMemcpyAsync(v.data(), (w ? c->d_out2 : c->d_out1).as<char>() + c->ho_off[w], v.size(), cudaMemcpyDeviceToHost, s));
            MK_CUDA(cudaStreamSynchronize(s));
            (w ? c->q2 : c->q1).emplace_back(std::move(v));
        }
        c->ho_len[w] = c->ho_off[w] = 0;
    }
    const bool trace = getenv("MICROCKET_TRACE") != nullptr;
    const double t0 = dd_now();
    // the window can insert at most one key per pair into EITHER set (tag 1 = lower-case first key base, e.g. soft-masked
    // reads: every pair of such a file lands there): keep both tables at most ~60 % full for the same bound
    const u64 bound = n / 40 + 16;                       // a pair of records needs at least 8 newlines + ids + bases
    for (int w = 0; w < 2; ++w) {
        u64 need = c->hcount[w] + bound;
        u64 slots = c->hslots[w];
        while (need * 10 > slots * 6) slots *= 2;
        if (slots != c->hslots[w]) MK_TRY(hs_alloc(c, w, slots));
    }
    FqParams p; memset(&p, 0, sizeof p);
    p.buf = c->d_in.as<char>(); p.st = c->d_state.as<FqState>(); p.nl_pos = c->d_nl.as<u32>(); p.cap_lines = c->cap_lines;
    p.desc = c->d_desc.as<u64>(); p.rec0 = c->d_rec0.as<uint4>(); p.rec1 = c->d_rec1.as<uint4>(); p.cls = c->d_cls.as<u8>();
    p.eoff = c->d_eoff.as<u32>(); p.btab = c->d_btab.as<u32>();
    for (int w = 0; w < 2; ++w) { p.hset[w] = c->d_hset[w].as<u64>(); p.hmask[w] = c->hslots[w] - 1; }
    size_t base1 = append ? c->ho_len[0] : 0, base2 = append ? c->ho_len[1] : 0;
    bool wait_out = false;
    if (c->fly && !append) {
        // copies of earlier pulls may still be reading the front of the buffers: write behind them while there is room, else
        // the kernels (not this window's host-to-device copy, which is already enqueued) wait for those copies
        // (room for two windows: a second window of the same push appends behind this one)
        if (c->fly_end[0] + 2 * n <= c->out_cap && c->fly_end[1] + 2 * n <= c->out_cap) { base1 = c->fly_end[0]; base2 = c->fly_end[1]; }
        else wait_out = true;
    }
    p.out1 = c->d_out1.as<char>() + base1; p.out2 = c->d_out2.as<char>() + base2; p.out_cap = c->out_cap - std::max(base1, base2);
    p.hskip1 = c->cfg.hskip1; p.klen1 = c->cfg.klen1; p.hskip2 = c->cfg.hskip2; p.klen2 = c->cfg.klen2;
    if (n_stage) MK_CUDA(cudaMemcpyAsync(c->d_in.p, c->h_in.p, n_stage, cudaMemcpyHostToDevice, s));
    if (n_direct) MK_CUDA(cudaMemcpyAsync(c->d_in.as<char>() + n_stage, direct, n_direct, cudaMemcpyHostToDevice, s));
    if (wait_out) { MK_CUDA(cudaStreamWaitEvent(s, c->ev_out, 0)); c->fly = false; c->fly_end[0] = c->fly_end[1] = 0; }
    k_fq_begin<<<(c->n_desc + 255) / 256, 256, 0, s>>>(p, c->n_desc, n, is_last ? 1u : 0u);
    int occ = 1; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fq_scan, S2P_SCAN_THREADS, 4 * 8192);
    k_fq_scan<<<c->sms * std::max(1, std::min(occ, 4)), S2P_SCAN_THREADS, 4 * 8192, s>>>(p);
    k_fq_keys<<<c->sms * 8, 256, 0, s>>>(p);
    c->launches_generic += 3;
    // the pair count lives on the device; the sort is sized by a host read of it (one small sync per window)
    FqState hst;
    MK_CUDA(cudaMemcpyAsync(&hst, c->d_state.p, sizeof hst, cudaMemcpyDeviceToHost, s));
    MK_CUDA(cudaStreamSynchronize(s));
    if (hst.err & FQ_ERR_LINES) { mk_set_error("krmdup: too many lines in one window"); return MK_ERR_CAPACITY; }
    const double t1 = dd_now();
    const u64 np = hst.n_pairs;
    *consumed = (size_t)hst.consumed;
    if (np == 0) return MK_OK;
    RadixSchedule sch; sch.n_pass = 9;
    const int order[9] = {12, 0, 1, 2, 3, 4, 5, 6, 7};
    for (int i = 0; i < 9; ++i) sch.byte_of[i] = order[i];
    Rec16::Bufs b; b.k[0] = p.rec0; b.k[1] = p.rec1; b.v[0] = b.v[1] = nullptr;
    MK_TRY(radix_sort<Rec16>(b, np, sch, c->rws, 0, c->sms, s, &c->launches_generic));
    k_fq_mark<<<c->sms * 8, 256, 0, s>>>(p, c->rws.plan.as<RadixPlan>());
    MK_CUDA(cudaMemsetAsync(c->d_ldesc.p, 0, (size_t)c->n_tiles_cap * 4 * 8, s));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fq_layout, FL_T, 0);
    k_fq_layout<<<c->sms * std::max(1, std::min(occ, 4)), FL_T, 0, s>>>(p, c->d_ldesc.as<u64>(), c->n_tiles_cap);
    k_fq_copy<<<c->sms * 8, 256, 0, s>>>(p);
    c->launches_generic += 3;
    MK_CUDA(cudaMemcpyAsync(&hst, c->d_state.p, sizeof hst, cudaMemcpyDeviceToHost, s));
    MK_CUDA(cudaStreamSynchronize(s));
    if (hst.err & FQ_ERR_OUT) { mk_set_error("krmdup: output buffer too small"); return MK_ERR_CAPACITY; }
    if (hst.err & FQ_ERR_TABLE) { mk_set_error("krmdup: key set full (device hash table)"); return MK_ERR_CAPACITY; }
    const double t2 = dd_now();
    c->hcount[0] += hst.inserted[0]; c->hcount[1] += hst.inserted[1];
    MK_CUDA(cudaMemsetAsync((char *)c->d_state.p + offsetof(FqState, inserted), 0, 8, s));
    c->pairs_total += np;
    // the kept records stay in d_out1 / d_out2 until mk_dedup_pull copies them straight into the caller's buffers
    if (!append) { c->ho_off[0] = base1; c->ho_off[1] = base2; }
    c->ho_len[0] = base1 + hst.out1; c->ho_len[1] = base2 + hst.out2;
    if (trace) fprintf(stderr, "[krmdup window] %zu bytes (%zu staged), %llu pairs: h2d+scan+keys %.1f ms, sort+mark+layout+copy %.1f ms\n",
                       n, n_stage, (unsigned long long)np, (t1 - t0) * 1e3, (t2 - t1) * 1e3);
    return MK_OK;
}

extern "C" int mk_dedup_push(mk_ctx *x, const char *bytes, size_t n, int is_last) {
    DedupCtx *c; MK_TRY(dd_check(x, &c));
    if (c->finished) { mk_set_error("mk_dedup_push after the last chunk"); return MK_ERR_STATE; }
    char *h = c->h_in.as<char>();
    const double t_push = dd_now();
    bool pinned = n >= (1u << 20) && dd_ptr_is_pinned(bytes);
    if (c->ext_len) {
        // the tail the last push left in the caller's pinned memory: one region with this push when it continues it (a caller
        // walking through one big pinned buffer), else it moves to the staging buffer now
        if (pinned && bytes == c->ext_ptr + c->ext_len) { bytes = c->ext_ptr; n += c->ext_len; }
        else { memcpy(h, c->ext_ptr, c->ext_len); c->fill = c->ext_len; }
        c->ext_len = 0; c->ext_ptr = nullptr;
    }
    size_t off = 0;
    while (true) {
        const size_t take = std::min(c->W - c->fill, n - off);
        const bool final_chunk = is_last && off + take == n;
        if (pinned && !is_last && c->fill == 0 && n - off < c->W) {     // not a window's worth: it stays where it is until the next push
            c->ext_ptr = bytes + off; c->ext_len = n - off;
            break;
        }
        // pinned caller memory: the window is [what h_in holds][bytes + off, + take), the second part DMA'd in place
        if (pinned && take && c->fill + take == c->W && !final_chunk) {
            const void *nl = memrchr(bytes + off, '\n', take);
            if (nl) {
                const size_t nd = (size_t)((const char *)nl - (bytes + off)) + 1, staged = c->fill;
                size_t consumed = 0;
                MK_TRY(dd_window(c, staged, bytes + off, nd, false, &consumed));
                if (consumed == 0) { mk_set_error("krmdup: one 65 536-pair batch does not fit the window (%zu bytes)", c->W); return MK_ERR_CAPACITY; }
                if (consumed >= staged) { off += consumed - staged; c->fill = 0; }
                else { memmove(h, h + consumed, staged - consumed); c->fill = staged - consumed; }     // (a batch never ends inside the staged part in practice)
                continue;
            }
        }
        if (take) { memcpy(h + c->fill, bytes + off, take); c->fill += take; off += take; }
        if (c->fill < c->W && !final_chunk) break;                       // wait for more input
        if (c->fill == 0) break;
        if (final_chunk && h[c->fill - 1] != '\n') h[c->fill++] = '\n';   // getline accepts a last line without '\n' (room: W + 64)
        size_t len = c->fill;
        if (!final_chunk) {
            const void *nl = memrchr(h, '\n', len);
            if (!nl) { mk_set_error("krmdup: a line longer than the window"); return MK_ERR_CAPACITY; }
            len = (size_t)((const char *)nl - h) + 1;
        }
        size_t consumed = 0;
        MK_TRY(dd_window(c, len, nullptr, 0, final_chunk, &consumed));
        if (final_chunk) { c->fill = 0; break; }
        if (consumed == 0) { mk_set_error("krmdup: one 65 536-pair batch does not fit the window (%zu bytes)", c->W); return MK_ERR_CAPACITY; }
        memmove(h, h + consumed, c->fill - consumed);                    // whole batches only: the rest opens the next window
        c->fill -= consumed;
    }
    if (is_last) c->finished = true;
    if (getenv("MICROCKET_TRACE")) fprintf(stderr, "[krmdup push] %zu bytes in %.1f ms, %zu staged, %zu left in caller memory, %zu + %zu spilled\n",
                                           n, (dd_now() - t_push) * 1e3, c->fill, c->ext_len, c->q1.size(), c->q2.size());
    return MK_OK;
}

static size_t dd_drain(std::deque<std::vector<char>> &q, size_t &qoff, char *out, size_t cap) {
    size_t n = 0;
    while (out && n < cap && !q.empty()) {
        std::vector<char> &f = q.front();
        size_t m = std::min(cap - n, f.size() - qoff);
        memcpy(out + n, f.data() + qoff, m);
        n += m; qoff += m;
        if (qoff == f.size()) { q.pop_front(); qoff = 0; }
    }
    return n;
}

extern "C" int mk_dedup_pull(mk_ctx *x, char *r1, size_t cap1, size_t *n1, char *r2, size_t cap2, size_t *n2) {
    DedupCtx *c; MK_TRY(dd_check(x, &c));
    size_t a = dd_drain(c->q1, c->q1_off, r1, cap1), b = dd_drain(c->q2, c->q2_off, r2, cap2);
    // then the last window's records, from the device straight into the caller's buffers (both copies in flight together)
    const bool async = c->s_out != nullptr;
    if (async && c->fly) MK_CUDA(cudaStreamSynchronize(c->s_out));      // what earlier pulls reported has landed now
    cudaStream_t cs = async ? c->s_out : c->s;
    bool copied = false;
    if (r1 && c->q1.empty() && a < cap1) {
        const size_t m = std::min(cap1 - a, c->ho_len[0] - c->ho_off[0]);
        if (m) { MK_CUDA(cudaMemcpyAsync(r1 + a, c->d_out1.as<char>() + c->ho_off[0], m, cudaMemcpyDeviceToHost, cs)); copied = true; }
        a += m; c->ho_off[0] += m;
    }
    if (r2 && c->q2.empty() && b < cap2) {
        const size_t m = std::min(cap2 - b, c->ho_len[1] - c->ho_off[1]);
        if (m) { MK_CUDA(cudaMemcpyAsync(r2 + b, c->d_out2.as<char>() + c->ho_off[1], m, cudaMemcpyDeviceToHost, cs)); copied = true; }
        b += m; c->ho_off[1] += m;
    }
    const double t_p = copied && getenv("MICROCKET_TRACE") ? dd_now() : 0;
    if (copied && async) {                                              // the copies run on while the caller pushes the next window
        MK_CUDA(cudaEventRecord(c->ev_out, c->s_out));
        c->fly = true; c->fly_end[0] = c->ho_off[0]; c->fly_end[1] = c->ho_off[1];
    } else if (copied) MK_CUDA(cudaStreamSynchronize(c->s));
    if (t_p != 0) fprintf(stderr, "[krmdup pull] %zu + %zu bytes to the host in %.1f ms\n", a, b, (dd_now() - t_p) * 1e3);
    if (n1) *n1 = a;
    if (n2) *n2 = b;
    return MK_OK;
}

// Forget the stream (key sets, counters, pending output) but keep every allocation: the context then stands for a new krmdup process.
extern "C" int mk_dedup_reset(mk_ctx *x) {
    DedupCtx *c; MK_TRY(dd_check(x, &c));
    MK_CUDA(cudaStreamSynchronize(c->s));
    if (c->s_out) MK_CUDA(cudaStreamSynchronize(c->s_out));
    c->fly = false; c->fly_end[0] = c->fly_end[1] = 0;
    MK_CUDA(cudaMemsetAsync(c->d_state.p, 0, sizeof(FqState), c->s));
    for (int w = 0; w < 2; ++w) { MK_CUDA(cudaMemsetAsync(c->d_hset[w].p, 0xFF, c->hslots[w] * 8, c->s)); c->hcount[w] = 0; c->ho_len[w] = c->ho_off[w] = 0; }
    MK_CUDA(cudaStreamSynchronize(c->s));
    c->fill = 0; c->ext_ptr = nullptr; c->ext_len = 0; c->finished = false; c->pairs_total = 0;
    c->q1.clear(); c->q2.clear(); c->q1_off = c->q2_off = 0;
    return MK_OK;
}

extern "C" int mk_dedup_finish(mk_ctx *x, mk_dedup_stats *out) {
    DedupCtx *c; MK_TRY(dd_check(x, &c));
    if (!out) { mk_set_error("mk_dedup_finish: null stats"); return MK_ERR_ARG; }
    if (!c->finished) MK_TRY(mk_dedup_push(x, nullptr, 0, 1));
    if (c->s_out) MK_CUDA(cudaStreamSynchronize(c->s_out));             // async_pull: everything reported so far has landed
    FqState st;
    MK_CUDA(cudaMemcpyAsync(&st, c->d_state.p, sizeof st, cudaMemcpyDeviceToHost, c->s));
    MK_CUDA(cudaStreamSynchronize(c->s));
    out->uniq = (u32)st.uniq; out->dup = (u32)st.dup; out->discard = (u32)st.discard; out->pairs = c->pairs_total;
    return MK_OK;
}
