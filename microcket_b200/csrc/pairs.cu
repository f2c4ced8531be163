// pairs.cu — duplicate removal by sort + adjacent-unique, and contact binning to COO, over device-resident data.
//
//   mk_dedup_keys_device   64-bit keys (krmdup's 2-bit packed read bases, src/preprocess/krmdup.cpp:168-193):
//                          stable radix sort of (key, index), first of every run is the occurrence the reference's
//                          unordered_set keeps (krmdup.cpp:201-212)
//   mk_pairs_dedup_device  mk_pair records sorted directly on (lane, chr1, pos1, chr2, pos2, strands), runs collapsed
//   mk_pairs_bin_device    bin = offset[chr] + pos / res (util/analyze.EBV/calc.loop2EBV.pl:28), sort of (bin1,bin2)
//                          keys, run-length encoding to COO triplets — stands in for `juicer_tools pre` (microcket:525-529)
#include <algorithm>
#include <vector>
#include "mk_common.cuh"
#include "radix_sort.cuh"
#include "pairs_ws.h"

struct K64 {                       // keys only
    typedef u64 Key;
    static constexpr int ITEMS = 16;
    static constexpr int MIN_BLOCKS = 1;
    static constexpr bool HAS_VAL = false;
    struct Bufs { u64 *k[2]; u32 *v[2]; };
    __device__ static __forceinline__ u32 digit(const u64 &k, int byte) { return (u32)(k >> (8 * byte)) & 255u; }
    __device__ static __forceinline__ u64 load_key(const Bufs &b, int which, u64 i) { return b.k[which][i]; }
    __device__ static __forceinline__ void store_key(const Bufs &b, int which, u64 i, const u64 &k) { b.k[which][i] = k; }
};

// ------------------------------------------------------------------------------------------------ unique / compaction kernels
#define UQ_T 256
#define UQ_ITEMS 8

// keep[idx] = first of run, for sorted (key, idx); counts uniques
__global__ void __launch_bounds__(256) k_mark_first(const u64 *k0, const u64 *k1, const u32 *v0, const u32 *v1, const RadixPlan *plan,
                                                    u64 n, u8 *keep, unsigned long long *n_unique) {
    const u64 *k = plan->final_buf ? k1 : k0;
    const u32 *v = plan->final_buf ? v1 : v0;
    u32 local = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        bool first = i == 0 || k[i] != k[i - 1];
        keep[plan->n_exec ? v[i] : (u32)i] = first ? 1 : 0;    // no pass ran (all keys equal): values are still the iota
        local += first;
    }
    local = __reduce_add_sync(0xffffffffu, local);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(n_unique, (unsigned long long)local);
}

__device__ __forceinline__ bool pair_key_differs(const uint4 &a, const uint4 &b) {
    // identity = every field except cls (byte 13), which is a function of the others
    return a.x != b.x || a.y != b.y || a.z != b.z || ((a.w ^ b.w) & 0xFFFF00FFu) != 0;
}

// stream compaction of the first record of every run (sorted records); single pass with look-back
// (tiles are claimed with a ticket, here and in the kernels below: a tile's predecessors are then always owned by CTAs that
// are already running, so the look-back never waits on a CTA that another stream's kernel keeps from being scheduled)
__global__ void __launch_bounds__(UQ_T) k_unique_rec16(const uint4 *b0, const uint4 *b1, const RadixPlan *plan, u64 n,
                                                       uint4 *o0, uint4 *o1, u64 *desc, unsigned long long *n_out, u32 *ticket) {
    __shared__ u32 s_scan[UQ_T / 32 + 1];
    __shared__ u64 s_base;
    __shared__ u32 s_tile;
    const uint4 *in = plan->final_buf ? b1 : b0;
    uint4 *out = plan->final_buf ? o0 : o1;                   // the other buffer
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n_tiles = (int)((n + UQ_T * UQ_ITEMS - 1) / (UQ_T * UQ_ITEMS));
    while (true) {
        if (tid == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const int tile = (int)s_tile;
        if (tile >= n_tiles) break;
        const u64 base = (u64)tile * UQ_T * UQ_ITEMS + (u64)tid * UQ_ITEMS;
        uint4 r[UQ_ITEMS]; u32 f = 0, cnt = 0;
        uint4 prev = base > 0 && base <= n ? in[base - 1] : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < UQ_ITEMS; ++k) {
            if (base + k < n) {
                r[k] = in[base + k];
                bool first = (base + k == 0) || pair_key_differs(r[k], prev);
                prev = r[k];
                if (first) { f |= 1u << k; ++cnt; }
            }
        }
        u32 tot;
        u32 ex = block_excl_scan<UQ_T>(cnt, s_scan, &tot);
        if (wid == 0) {
            u64 b = lookback_exclusive(desc, tile, 0, tot, lane);
            if (lane == 0) { s_base = b; if (tile == n_tiles - 1) *n_out = b + tot; }
        }
        __syncthreads();
        u64 o = s_base + ex;
#pragma unroll
        for (int k = 0; k < UQ_ITEMS; ++k) if (f & (1u << k)) out[o++] = r[k];
        __syncthreads();
    }
}

// run heads of sorted 64-bit keys: head positions + keys, compacted (single pass with look-back)
__global__ void __launch_bounds__(UQ_T) k_rle_heads(const u64 *k0, const u64 *k1, const RadixPlan *plan, u64 n,
                                                    u64 *out_key, u32 *out_pos, u64 cap, u64 *desc, unsigned long long *n_out, u32 *ticket) {
    __shared__ u32 s_scan[UQ_T / 32 + 1];
    __shared__ u64 s_base;
    __shared__ u32 s_tile;
    const u64 *in = plan->final_buf ? k1 : k0;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n_tiles = (int)((n + UQ_T * UQ_ITEMS - 1) / (UQ_T * UQ_ITEMS));
    while (true) {
        if (tid == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const int tile = (int)s_tile;
        if (tile >= n_tiles) break;
        const u64 base = (u64)tile * UQ_T * UQ_ITEMS + (u64)tid * UQ_ITEMS;
        u64 r[UQ_ITEMS]; u32 f = 0, cnt = 0;
        u64 prev = base > 0 && base <= n ? in[base - 1] : 0;
#pragma unroll
        for (int k = 0; k < UQ_ITEMS; ++k) {
            if (base + k < n) {
                r[k] = in[base + k];
                bool first = (base + k == 0) || r[k] != prev;
                prev = r[k];
                if (first) { f |= 1u << k; ++cnt; }
            }
        }
        u32 tot;
        u32 ex = block_excl_scan<UQ_T>(cnt, s_scan, &tot);
        if (wid == 0) {
            u64 b = lookback_exclusive(desc, tile, 0, tot, lane);
            if (lane == 0) { s_base = b; if (tile == n_tiles - 1) *n_out = b + tot; }
        }
        __syncthreads();
        u64 o = s_base + ex;
#pragma unroll
        for (int k = 0; k < UQ_ITEMS; ++k)
            if (f & (1u << k)) { if (o < cap) { out_key[o] = r[k]; out_pos[o] = (u32)(base + k); } ++o; }
        __syncthreads();
    }
}

// the run of sentinel keys (pairs that failed validation), if any, is the last one: it is not a cell
__global__ void k_rle_trim(const u64 *key, const u32 *pos, unsigned long long *counters /* [0] runs -> cells, [2] <- valid elements */, u64 n, u64 cap) {
    const u64 runs = counters[0];
    counters[2] = n;
    if (runs && runs <= cap && key[runs - 1] == 0xFFFFFFFFFFFFFFFFull) { counters[0] = runs - 1; counters[2] = pos[runs - 1]; }
}
__global__ void k_rle_finish(const u64 *key, const u32 *pos, const unsigned long long *nnz_p, u64 n_unused, u64 cap,
                             u32 *bin1, u32 *bin2, u32 *cnt) {
    const u64 n = nnz_p[2];
    const u64 nnz = *nnz_p < cap ? *nnz_p : cap;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (u64)gridDim.x * blockDim.x) {
        u64 k = key[i];
        bin1[i] = (u32)(k >> 32); bin2[i] = (u32)k;
        u32 nxt = (i + 1 < *nnz_p && i + 1 < cap) ? pos[i + 1] : (u32)n;
        if (i + 1 < *nnz_p && i + 1 >= cap) nxt = pos[i] + 1;   // truncated output: count unknowable, caller gets an error
        cnt[i] = nxt - pos[i];
    }
}

// (bin1,bin2) key of every pair at one resolution.  Every field is checked (chromosome id inside the table, position not
// past the chromosome's end — one rule for every resolution, so a pair is in all matrices or in none): an offender is counted in *bad, the host turns that into MK_ERR_INPUT, and its key is the all-ones
// sentinel (sorted last, dropped by the run-length step) so that nothing is read or counted out of range.
#define BIN_BAD_KEY 0xFFFFFFFFFFFFFFFFull
__global__ void k_bin_keys(const mk_pair *p, u64 n, const u32 *chr_off /* per pair-chr id */, const u32 *chr_len /* length per id */, u32 n_ids,
                           u32 res, u64 *key, unsigned long long *bad) {
    u32 nbad = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const uint4 r = ((const uint4 *)p)[i];
        const u32 pos1 = r.x, pos2 = r.y, c1 = r.z & 0xFFFFu, c2 = r.z >> 16;
        u64 k = BIN_BAD_KEY;
        if (c1 < n_ids && c2 < n_ids && pos1 <= chr_len[c1] && pos2 <= chr_len[c2]) {
            u32 a = chr_off[c1] + pos1 / res, b = chr_off[c2] + pos2 / res;
            if (a > b) { u32 t = a; a = b; b = t; }              // upper triangle
            k = ((u64)a << 32) | b;
        }
        nbad += k == BIN_BAD_KEY;
        key[i] = k;
    }
    nbad = __reduce_add_sync(0xffffffffu, nbad);
    if ((threadIdx.x & 31) == 0 && nbad) atomicAdd(bad, (unsigned long long)nbad);
}

// ------------------------------------------------------------------------------------------------ workspace (pairs_ws.h)
extern "C" int mk_pairs_ws_create(int device, size_t max_pairs, mk_pairs_ws **out) {
    if (!out || max_pairs == 0) { mk_set_error("mk_pairs_ws_create: bad argument"); return MK_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { mk_set_error("no CUDA device: microcket_b200 has no CPU fallback"); return MK_ERR_CUDA; }
    MK_CUDA(cudaSetDevice(device));
    mk_pairs_ws *w = new mk_pairs_ws();
    w->device = device; w->sms = mk_sm_count(device); w->max_pairs = max_pairs;
    int rc = w->rws.alloc(max_pairs);
    if (rc == MK_OK) rc = w->alt.alloc(max_pairs * 16);
    if (rc == MK_OK) rc = w->keys2.alloc(max_pairs * 8);
    if (rc == MK_OK) rc = w->heads_key.alloc(max_pairs * 8);
    if (rc == MK_OK) rc = w->heads_pos.alloc(max_pairs * 4);
    if (rc == MK_OK) rc = w->desc.alloc((max_pairs / (UQ_T * UQ_ITEMS) + 4) * 8 + 2048);
    if (rc == MK_OK) rc = w->counter.alloc(64);
    if (rc == MK_OK) rc = w->chr_off.alloc(65536 * 8 + 16384 * 8);
    if (rc == MK_OK && cudaMemset(w->chr_off.p, 0, w->chr_off.n) != cudaSuccess) rc = MK_ERR_CUDA;
    if (rc != MK_OK) { delete w; return rc; }
    *out = w;
    return MK_OK;
}
extern "C" void mk_pairs_ws_destroy(mk_pairs_ws *w) { if (w) { cudaSetDevice(w->device); delete w; } }
extern "C" uint64_t mk_pairs_launch_count(mk_pairs_ws *w) { return w ? w->launches : 0; }
extern "C" uint64_t mk_pairs_dropped(mk_pairs_ws *w) { return w ? w->dropped : 0; }

static int lookback_grid(const void *kernel, int threads, int sms) {
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, 0);
    return sms * std::max(1, std::min(occ, 4));
}

extern "C" int mk_pairs_dedup_device(mk_pairs_ws *w, mk_pair *d_pairs, size_t n, size_t *n_kept, void *stream) {
    if (!w || !n_kept) { mk_set_error("mk_pairs_dedup_device: bad argument"); return MK_ERR_ARG; }
    if (n > w->max_pairs) { mk_set_error("mk_pairs_dedup_device: workspace holds %zu pairs, got %zu", w->max_pairs, n); return MK_ERR_CAPACITY; }
    MK_CUDA(cudaSetDevice(w->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) { *n_kept = 0; return MK_OK; }
    // LSD byte schedule over the mk_pair layout: strands(12) pos2(4..7) chr2(10,11) pos1(0..3) chr1(8,9) lane(14,15)
    RadixSchedule sch; sch.n_pass = 15;
    const int order[15] = {12, 4, 5, 6, 7, 10, 11, 0, 1, 2, 3, 8, 9, 14, 15};
    for (int i = 0; i < 15; ++i) sch.byte_of[i] = order[i];
    Rec16::Bufs b; b.k[0] = (uint4 *)d_pairs; b.k[1] = w->alt.as<uint4>(); b.v[0] = b.v[1] = nullptr;
    MK_TRY(radix_sort<Rec16>(b, n, sch, w->rws, 0, w->sms, s, &w->launches));
    const int n_tiles = (int)((n + UQ_T * UQ_ITEMS - 1) / (UQ_T * UQ_ITEMS));
    MK_CUDA(cudaMemsetAsync(w->desc.p, 0, (size_t)(n_tiles + 1) * 8, s));
    MK_CUDA(cudaMemsetAsync(w->counter.p, 0, 64, s));
    k_unique_rec16<<<lookback_grid((const void *)k_unique_rec16, UQ_T, w->sms), UQ_T, 0, s>>>(
        b.k[0], b.k[1], w->rws.plan.as<RadixPlan>(), n, b.k[0], b.k[1], w->desc.as<u64>(), w->counter.as<unsigned long long>(),
        (u32 *)(w->counter.as<unsigned long long>() + 4));
    w->launches += 1;
    struct { unsigned long long kept; } hc;
    u32 final_buf = 0;
    MK_CUDA(cudaMemcpyAsync(&hc, w->counter.p, 8, cudaMemcpyDeviceToHost, s));
    MK_CUDA(cudaMemcpyAsync(&final_buf, (char *)w->rws.plan.p + offsetof(RadixPlan, final_buf), 4, cudaMemcpyDeviceToHost, s));
    MK_CUDA(cudaStreamSynchronize(s));
    // sorted data sat in buffer final_buf, the compacted run heads went to the other one
    if (final_buf == 0) MK_CUDA(cudaMemcpyAsync(d_pairs, w->alt.p, (size_t)hc.kept * 16, cudaMemcpyDeviceToDevice, s));
    *n_kept = (size_t)hc.kept;
    MK_CUDA(cudaStreamSynchronize(s));
    return MK_OK;
}

extern "C" int mk_pairs_bin_device(mk_pairs_ws *w, const mk_pair *d_pairs, size_t n, const uint32_t *chrom_len, int n_chrom,
                                   const uint16_t *chrom_id_map, int n_map, uint32_t res, uint32_t *d_bin1, uint32_t *d_bin2,
                                   uint32_t *d_cnt, size_t cap, size_t *nnz, void *stream) {
    if (!w || !nnz || !chrom_len || n_chrom <= 0 || res == 0) { mk_set_error("mk_pairs_bin_device: bad argument"); return MK_ERR_ARG; }
    if (n > w->max_pairs) { mk_set_error("mk_pairs_bin_device: workspace holds %zu pairs, got %zu", w->max_pairs, n); return MK_ERR_CAPACITY; }
    MK_CUDA(cudaSetDevice(w->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) { *nnz = 0; return MK_OK; }
    // bin offset of every chromosome in the given order; then per pair-chromosome-id through the map
    std::vector<u64> off(n_chrom + 1, 0);
    for (int c = 0; c < n_chrom; ++c) off[c + 1] = off[c] + chrom_len[c] / res + 1;
    if (off[n_chrom] >= (1ull << 32)) { mk_set_error("mk_pairs_bin_device: more than 2^32 bins"); return MK_ERR_CAPACITY; }
    const int n_ids = chrom_id_map ? n_map : n_chrom;
    if (n_ids > 65536) { mk_set_error("mk_pairs_bin_device: too many chromosome ids"); return MK_ERR_ARG; }
    std::vector<u32> by_id(2 * (size_t)n_ids);                          // [offsets | lengths per id], one upload
    for (int i = 0; i < n_ids; ++i) {
        int c = chrom_id_map ? chrom_id_map[i] : i;
        if (c < 0 || c >= n_chrom) { mk_set_error("mk_pairs_bin_device: chromosome map entry %d out of range", i); return MK_ERR_ARG; }
        by_id[i] = (u32)off[c]; by_id[n_ids + i] = chrom_len[c];
    }
    MK_CUDA(cudaMemcpyAsync(w->chr_off.p, by_id.data(), by_id.size() * 4, cudaMemcpyHostToDevice, s));
    MK_CUDA(cudaMemsetAsync(w->counter.p, 0, 64, s));
    unsigned long long *cnt = w->counter.as<unsigned long long>();
    u64 *k0 = w->alt.as<u64>(), *k1 = w->keys2.as<u64>();
    k_bin_keys<<<w->sms * 8, 256, 0, s>>>(d_pairs, n, w->chr_off.as<u32>(), w->chr_off.as<u32>() + n_ids, (u32)n_ids, res, k0, cnt + 3);
    w->launches += 1;
    int nbytes = 1; while (nbytes < 4 && (off[n_chrom] >> (8 * nbytes))) ++nbytes;
    RadixSchedule sch; sch.n_pass = 0;
    for (int i = 0; i < nbytes; ++i) sch.byte_of[sch.n_pass++] = i;
    for (int i = 0; i < nbytes; ++i) sch.byte_of[sch.n_pass++] = 4 + i;
    K64::Bufs b; b.k[0] = k0; b.k[1] = k1; b.v[0] = b.v[1] = nullptr;
    MK_TRY(radix_sort<K64>(b, n, sch, w->rws, 0, w->sms, s, &w->launches));
    const int n_tiles = (int)((n + UQ_T * UQ_ITEMS - 1) / (UQ_T * UQ_ITEMS));
    MK_CUDA(cudaMemsetAsync(w->desc.p, 0, (size_t)(n_tiles + 1) * 8, s));
    k_rle_heads<<<lookback_grid((const void *)k_rle_heads, UQ_T, w->sms), UQ_T, 0, s>>>(
        k0, k1, w->rws.plan.as<RadixPlan>(), n, w->heads_key.as<u64>(), w->heads_pos.as<u32>(), w->max_pairs, w->desc.as<u64>(), cnt, (u32 *)(cnt + 4));
    k_rle_trim<<<1, 1, 0, s>>>(w->heads_key.as<u64>(), w->heads_pos.as<u32>(), cnt, n, w->max_pairs);
    k_rle_finish<<<w->sms * 4, 256, 0, s>>>(w->heads_key.as<u64>(), w->heads_pos.as<u32>(), cnt, n, cap, d_bin1, d_bin2, d_cnt);
    w->launches += 3;
    unsigned long long hc[4] = {0, 0, 0, 0};
    MK_CUDA(cudaMemcpyAsync(hc, cnt, 32, cudaMemcpyDeviceToHost, s));
    MK_CUDA(cudaStreamSynchronize(s));
    const unsigned long long h = hc[0];
    *nnz = (size_t)h; w->dropped = hc[3];
    if (h > cap) { mk_set_error("mk_pairs_bin_device: %llu non-zero cells, output capacity %zu", h, cap); return MK_ERR_CAPACITY; }
    return MK_OK;
}

extern "C" int mk_dedup_keys_device(int device, const uint64_t *d_keys, size_t n, uint8_t *d_keep, uint64_t *n_unique, void *stream) {
    if (!d_keys || !d_keep || !n_unique) { mk_set_error("mk_dedup_keys_device: bad argument"); return MK_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { mk_set_error("no CUDA device: microcket_b200 has no CPU fallback"); return MK_ERR_CUDA; }
    MK_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    *n_unique = 0;
    if (n == 0) return MK_OK;
    RadixWs rws; MK_TRY(rws.alloc(n));
    DevBuf k0, k1, v0, v1, cnt;
    MK_TRY(k0.alloc(n * 8)); MK_TRY(k1.alloc(n * 8)); MK_TRY(v0.alloc(n * 4)); MK_TRY(v1.alloc(n * 4)); MK_TRY(cnt.alloc(8));
    MK_CUDA(cudaMemcpyAsync(k0.p, d_keys, n * 8, cudaMemcpyDeviceToDevice, s));
    MK_CUDA(cudaMemsetAsync(cnt.p, 0, 8, s));
    RadixSchedule sch; sch.n_pass = 8;
    for (int i = 0; i < 8; ++i) sch.byte_of[i] = i;
    KV64::Bufs b; b.k[0] = k0.as<u64>(); b.k[1] = k1.as<u64>(); b.v[0] = v0.as<u32>(); b.v[1] = v1.as<u32>();
    u64 launches = 0;
    const int sms = mk_sm_count(device);
    MK_TRY(radix_sort<KV64>(b, n, sch, rws, 1, sms, s, &launches));
    k_mark_first<<<sms * 8, 256, 0, s>>>(b.k[0], b.k[1], b.v[0], b.v[1], rws.plan.as<RadixPlan>(), n, d_keep, cnt.as<unsigned long long>());
    unsigned long long h = 0;
    MK_CUDA(cudaMemcpyAsync(&h, cnt.p, 8, cudaMemcpyDeviceToHost, s));
    MK_CUDA(cudaStreamSynchronize(s));
    *n_unique = h;
    return MK_OK;
}

// Host-buffer convenience over the two device entry points: packed pairs in host memory -> kept pairs (sorted, in place
// in `pairs`) and COO counts at one resolution in host memory.  What the pairs2bins CLI and bench.py's e2e leg call.
extern "C" int mk_pairs_dedup_bin_host(mk_pairs_ws *w, mk_pair *pairs, size_t n, int do_dedup, const uint32_t *chrom_len, int n_chrom,
                                       const uint16_t *chrom_id_map, int n_map, uint32_t res,
                                       uint32_t *bin1, uint32_t *bin2, uint32_t *cnt, size_t cap, size_t *n_kept, size_t *nnz) {
    if (!w || !pairs || !n_kept || !nnz) { mk_set_error("mk_pairs_dedup_bin_host: bad argument"); return MK_ERR_ARG; }
    if (n > w->max_pairs) { mk_set_error("mk_pairs_dedup_bin_host: workspace holds %zu pairs, got %zu", w->max_pairs, n); return MK_ERR_CAPACITY; }
    MK_CUDA(cudaSetDevice(w->device));
    *n_kept = n; *nnz = 0;
    if (n == 0) return MK_OK;
    DevBuf &d_pairs = w->h_pairs, &d_b1 = w->h_b1, &d_b2 = w->h_b2, &d_c = w->h_c;
    if (!d_pairs.p) MK_TRY(d_pairs.alloc(w->max_pairs * sizeof(mk_pair)));
    MK_CUDA(cudaMemcpy(d_pairs.p, pairs, n * sizeof(mk_pair), cudaMemcpyHostToDevice));
    size_t kept = n;
    if (do_dedup) {
        MK_TRY(mk_pairs_dedup_device(w, d_pairs.as<mk_pair>(), n, &kept, nullptr));
        MK_CUDA(cudaMemcpy(pairs, d_pairs.p, kept * sizeof(mk_pair), cudaMemcpyDeviceToHost));
    }
    *n_kept = kept;
    if (res && bin1 && bin2 && cnt) {
        const size_t ocap = std::min(cap, kept);
        if (!d_b1.p) { MK_TRY(d_b1.alloc(w->max_pairs * 4)); MK_TRY(d_b2.alloc(w->max_pairs * 4)); MK_TRY(d_c.alloc(w->max_pairs * 4)); }
        size_t z = 0;
        MK_TRY(mk_pairs_bin_device(w, d_pairs.as<mk_pair>(), kept, chrom_len, n_chrom, chrom_id_map, n_map, res,
                                   d_b1.as<u32>(), d_b2.as<u32>(), d_c.as<u32>(), ocap, &z, nullptr));
        MK_CUDA(cudaMemcpy(bin1, d_b1.p, z * 4, cudaMemcpyDeviceToHost));
        MK_CUDA(cudaMemcpy(bin2, d_b2.p, z * 4, cudaMemcpyDeviceToHost));
        MK_CUDA(cudaMemcpy(cnt, d_c.p, z * 4, cudaMemcpyDeviceToHost));
        *nnz = z;
    }
    return MK_OK;
}

// ------------------------------------------------------------------------------------------------ multi-GPU: owner partition
// owner(pair) = mix(chr1, chr2, pos1 / res) mod world: equal keys and equal (bin1,bin2) cells at resolution `res` land on
// one rank, so duplicate removal and COO counts need no further exchange after the all-to-all.
__host__ __device__ __forceinline__ u32 mk_owner_hash(u32 chr1, u32 chr2, u32 pbin) {
    u32 h = (chr1 * 0x9E3779B1u) ^ (chr2 * 0x85EBCA77u) ^ (pbin * 0xC2B2AE3Du);
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

__global__ void __launch_bounds__(256) k_owner_count(const mk_pair *p, u64 n, u32 world, u32 res, unsigned long long *counts) {
    __shared__ u32 s_c[64];
    if (threadIdx.x < 64) s_c[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const u64 n_round = (n + 31) & ~(u64)31;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += (u64)gridDim.x * blockDim.x) {
        u32 d = 0xFFFFu + lane;
        if (i < n) { const uint4 r = ((const uint4 *)p)[i]; d = mk_owner_hash(r.z & 0xFFFFu, r.z >> 16, r.x / res) % world; }
        u32 peers = __match_any_sync(0xffffffffu, d);
        if (i < n && lane == __ffs(peers) - 1) atomicAdd(&s_c[d], (u32)__popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < world && s_c[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)s_c[threadIdx.x]);
}

__global__ void __launch_bounds__(256) k_owner_scatter(const mk_pair *p, u64 n, u32 world, u32 res, unsigned long long *cursor /* start offsets, advanced */,
                                                       mk_pair *out) {
    const int lane = threadIdx.x & 31;
    const u64 n_round = (n + 31) & ~(u64)31;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += (u64)gridDim.x * blockDim.x) {
        u32 d = 0xFFFFu + lane; uint4 r = make_uint4(0, 0, 0, 0);
        if (i < n) { r = ((const uint4 *)p)[i]; d = mk_owner_hash(r.z & 0xFFFFu, r.z >> 16, r.x / res) % world; }
        u32 peers = __match_any_sync(0xffffffffu, d);
        int leader = __ffs(peers) - 1;
        unsigned long long base = 0;
        if (i < n && lane == leader) base = atomicAdd(&cursor[d], (unsigned long long)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (i < n) ((uint4 *)out)[base + __popc(peers & ((1u << lane) - 1u))] = r;
    }
}

// counts -> start offsets of the segments (exclusive prefix), in place behind the counts
__global__ void k_owner_prefix(unsigned long long *dc, u32 world) {
    unsigned long long run = 0;
    for (u32 r = 0; r < world; ++r) { dc[64 + r] = run; run += dc[r]; }
}

// Groups `n` pairs by owner rank into d_out (contiguous segments in rank order); counts[r] = pairs destined to rank r.
// Count, prefix and scatter are enqueued back to back; the only host round trip is the read-back of the counts.
extern "C" int mk_pairs_partition_device(mk_pairs_ws *w, const mk_pair *d_pairs, size_t n, int world, uint32_t res, mk_pair *d_out,
                                         uint64_t *counts, void *stream) {
    if (!w || !counts || world < 1 || world > 64 || res == 0 || (n && (!d_pairs || !d_out))) { mk_set_error("mk_pairs_partition_device: bad argument"); return MK_ERR_ARG; }
    if (n > w->max_pairs) { mk_set_error("mk_pairs_partition_device: workspace holds %zu pairs, got %zu", w->max_pairs, n); return MK_ERR_CAPACITY; }
    MK_CUDA(cudaSetDevice(w->device));
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long *dc = (unsigned long long *)w->desc.p;           // 2 x 64 counters: counts, then cursors
    MK_CUDA(cudaMemsetAsync(dc, 0, 128 * 8, s));
    if (n) {
        k_owner_count<<<w->sms * 8, 256, 0, s>>>(d_pairs, n, (u32)world, res, dc);
        k_owner_prefix<<<1, 1, 0, s>>>(dc, (u32)world);
    }
    unsigned long long h[64];
    MK_CUDA(cudaMemcpyAsync(h, dc, (size_t)world * 8, cudaMemcpyDeviceToHost, s));
    if (n) k_owner_scatter<<<w->sms * 8, 256, 0, s>>>(d_pairs, n, (u32)world, res, dc + 64, d_out);
    w->launches += n ? 3 : 0;
    MK_CUDA(cudaStreamSynchronize(s));
    for (int r = 0; r < world; ++r) counts[r] = h[r];
    return MK_OK;
}

extern "C" uint32_t mk_pairs_owner(uint32_t chr1, uint32_t chr2, uint32_t pos1, uint32_t res, uint32_t world) {
    return mk_owner_hash(chr1, chr2, pos1 / res) % world;
}

// ------------------------------------------------------------------------------------------------ fused dedup + binning
// One sort serves both steps.  Every pair becomes a tightly packed integer, most significant field first:
//   [ bin1 : nb ][ bin2 : nb ][ lane : nl ][ pos1 % res : nr ][ pos2 % res : nr ][ swapped : 1 ][ strands : 2 ]   (<= 96 bits)
// with the pair's INPUT INDEX in bits 96..127 of the 16-byte record, which the sort never looks at.  Equal pairs are equal
// integers (duplicate removal = adjacent unique) and all pairs of one (bin1,bin2) cell are contiguous (binning = run-length
// encoding of the top 2*nb bits), in ceil(bits/8) radix passes instead of the 11 + 6 of the two separate sorts.  The sort
// is stable, so the head of every run is the pair that came FIRST in the input — the occurrence the reference's
// unordered_set keeps (src/preprocess/krmdup.cpp:201-212) — and its index says which emitted .pairs line survives.
// The packing is invertible, so the kept pairs are decoded back into mk_pair.
// A pair that cannot be keyed (chromosome id outside the table, position past the chromosome's last bin, lane > max_lane)
// gets the all-ones key: it sorts last, is dropped from both outputs and is counted (mk_pairs_dropped).
struct PackCfg { u32 res, nb, nr, nl, total_bits, n_dec, n_ids, max_lane; };

__device__ __forceinline__ void put_bits(u64 &lo, u64 &hi, u64 v, u32 width) {      // key = (key << width) | v;  v < 2^width
    hi = width >= 64 ? lo << (width - 64) : (width ? (hi << width) | (lo >> (64 - width)) : hi);
    lo = width >= 64 ? 0 : lo << width;
    lo |= v;
}
__device__ __forceinline__ u64 take_bits(u64 &lo, u64 &hi, u32 width) {             // v = key & mask; key >>= width
    const u64 v = width >= 64 ? lo : (lo & ((1ull << width) - 1));
    lo = width >= 64 ? hi : (width ? (lo >> width) | (hi << (64 - width)) : lo);
    hi = width >= 64 ? 0 : hi >> width;
    return v;
}

// (the keys' digit histograms for the radix sort are taken here, while the key is still in registers: one read of the keys less)
__global__ void __launch_bounds__(256) k_pack_keys(const mk_pair *p, u64 n, const u32 *off_by_id, const u32 *len_by_id, PackCfg c, uint4 *key,
                                                   unsigned long long *bad, RadixSchedule sch, RadixPlan *plan) {
    extern __shared__ u32 s_hist[];                                      // sch.n_pass * 256
    for (int i = threadIdx.x; i < sch.n_pass * 256; i += 256) s_hist[i] = 0;
    __syncthreads();
    const int lane_id = threadIdx.x & 31;
    u32 nbad = 0;
    const u64 n_round = (n + 31) & ~(u64)31;                             // whole warps stay in the loop (ballots in the histogram step)
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += (u64)gridDim.x * blockDim.x) {
        const bool valid = i < n;
        uint4 k = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, (u32)i);
        if (valid) {
            const uint4 r = ((const uint4 *)p)[i];
            u32 pos1 = r.x, pos2 = r.y; const u32 c1 = r.z & 0xFFFFu, c2 = r.z >> 16;
            u32 st = r.w & 3u; const u32 lane = r.w >> 16;
            bool ok = c1 < c.n_ids && c2 < c.n_ids && lane <= c.max_lane;
            if (ok) {
                const u32 q1 = pos1 / c.res, q2 = pos2 / c.res;
                ok = pos1 <= len_by_id[c1] && pos2 <= len_by_id[c2];
                if (ok) {
                    u32 a = off_by_id[c1] + q1, b = off_by_id[c2] + q2;
                    u32 r1 = pos1 - q1 * c.res, r2 = pos2 - q2 * c.res;
                    u32 sw = 0;
                    if (a > b) { u32 t = a; a = b; b = t; t = r1; r1 = r2; r2 = t; st = ((st & 1u) << 1) | (st >> 1); sw = 1; }
                    u64 lo = 0, hi = 0;
                    put_bits(lo, hi, a, c.nb); put_bits(lo, hi, b, c.nb); put_bits(lo, hi, lane, c.nl);
                    put_bits(lo, hi, r1, c.nr); put_bits(lo, hi, r2, c.nr); put_bits(lo, hi, sw, 1); put_bits(lo, hi, st, 2);
                    k.x = (u32)lo; k.y = (u32)(lo >> 32); k.z = (u32)hi;
                }
            }
            nbad += !ok;
            key[i] = k;
        }
        radix_hist_add<Rec16>(s_hist, sch, k, valid, lane_id);
    }
    nbad = __reduce_add_sync(0xffffffffu, nbad);
    if ((threadIdx.x & 31) == 0 && nbad) atomicAdd(bad, (unsigned long long)nbad);
    __syncthreads();
    radix_hist_flush(s_hist, sch.n_pass, plan);
}

// chromosome (index into dec_off) of a bin: last entry of dec_off that is <= bin.  `hint` is tried first: keys are sorted by
// bin1, a thread's consecutive records mostly share their chromosome, and three quarters of all pairs are cis (bin2 on bin1's
// chromosome), so the binary search over the table (dependent loads) is the rare path.
__device__ __forceinline__ u32 chrom_of_bin(u32 bin, u32 hint, const PackCfg &c, const u32 *dec_off) {
    if (dec_off[hint] <= bin && (hint + 1 == c.n_dec || bin < dec_off[hint + 1])) return hint;
    u32 l = 0, h = c.n_dec - 1;
    while (l < h) { const u32 m = (l + h + 1) >> 1; if (dec_off[m] <= bin) l = m; else h = m - 1; }
    return l;
}
__device__ __forceinline__ mk_pair unpack_key(uint4 k, const PackCfg &c, const u32 *dec_off, const u16 *dec_id, u32 &hint) {
    u64 lo = (u64)k.x | ((u64)k.y << 32), hi = (u64)k.z;
    u32 st = (u32)take_bits(lo, hi, 2); const u32 sw = (u32)take_bits(lo, hi, 1);
    u32 r2 = (u32)take_bits(lo, hi, c.nr), r1 = (u32)take_bits(lo, hi, c.nr);
    const u32 lane = (u32)take_bits(lo, hi, c.nl);
    u32 b = (u32)take_bits(lo, hi, c.nb), a = (u32)take_bits(lo, hi, c.nb);
    u32 ka = chrom_of_bin(a, hint, c, dec_off);                  // (a <= b here: the key holds the bins in upper-triangle order)
    u32 kb = chrom_of_bin(b, ka, c, dec_off);
    hint = ka;
    if (sw) { u32 t = a; a = b; b = t; t = r1; r1 = r2; r2 = t; t = ka; ka = kb; kb = t; st = ((st & 1u) << 1) | (st >> 1); }
    mk_pair o;
    o.pos1 = (a - dec_off[ka]) * c.res + r1; o.pos2 = (b - dec_off[kb]) * c.res + r2;
    o.chr1 = dec_id[ka]; o.chr2 = dec_id[kb]; o.strands = (u8)st; o.lane = (u16)lane;
    if (o.chr1 != o.chr2) o.cls = 0;
    else { const u32 d = o.pos2 - o.pos1; o.cls = d >= 10000u ? 1 : (d >= 1000u ? 2 : 3); }
    return o;
}

// sorted packed keys -> kept pairs (first of every run, decoded) + cell heads (bin1, bin2, rank of the cell's first kept pair)
// + optionally keep[input index] = 1 and kept_idx[rank] = input index of every kept pair
__global__ void __launch_bounds__(UQ_T) k_uniq_cells(const uint4 *b0, const uint4 *b1, const RadixPlan *plan, u64 n, PackCfg c,
                                                     const u32 *dec_off, const u16 *dec_id, mk_pair *out0, mk_pair *out1,
                                                     u32 *cell_b1, u32 *cell_b2, u32 *cell_first, u64 cell_cap, u8 *keep, u32 *kept_idx,
                                                     u64 *desc, unsigned long long *counters /* [0] kept, [1] cells */, u32 *ticket) {
    __shared__ u32 s_w[2][UQ_T / 32];
    __shared__ u64 s_base;
    __shared__ u32 s_tile;
    const uint4 *in = plan->final_buf ? b1 : b0;
    mk_pair *out = plan->final_buf ? out0 : out1;                 // the buffer the sorted keys are NOT in
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n_tiles = (int)((n + UQ_T * UQ_ITEMS - 1) / (UQ_T * UQ_ITEMS));
    const u32 cell_shift = c.total_bits - 2 * c.nb;               // bits below the (bin1,bin2) prefix
    while (true) {
        // tiles are claimed with a ticket: a tile's predecessors are then always owned by CTAs that are already running, so the
        // look-back cannot wait on a CTA that was never scheduled (other streams may share the device)
        if (tid == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const int tile = (int)s_tile;
        if (tile >= n_tiles) break;
        const u64 base = (u64)tile * UQ_T * UQ_ITEMS + (u64)tid * UQ_ITEMS;
        uint4 r[UQ_ITEMS]; u32 fk = 0, fc = 0, nk = 0, nc = 0;
        uint4 prev = base > 0 && base <= n ? in[base - 1] : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < UQ_ITEMS; ++k) {
            if (base + k < n) {
                r[k] = in[base + k];
                const bool first = base + k == 0;
                const bool dk = first || r[k].x != prev.x || r[k].y != prev.y || r[k].z != prev.z;   // .w is the input index
                bool dc = first;
                if (!first && dk) {
                    u64 lo = (u64)r[k].x | ((u64)r[k].y << 32), hi = (u64)r[k].z;
                    u64 plo = (u64)prev.x | ((u64)prev.y << 32), phi = (u64)prev.z;
                    take_bits(lo, hi, cell_shift); take_bits(plo, phi, cell_shift);
                    dc = lo != plo || hi != phi;
                }
                prev = r[k];
                if (dk) { fk |= 1u << k; ++nk; }
                if (dc) { fc |= 1u << k; ++nc; }
            }
        }
        const u32 ik = warp_incl_scan(nk, lane), ic = warp_incl_scan(nc, lane);
        if (lane == 31) { s_w[0][wid] = ik; s_w[1][wid] = ic; }
        __syncthreads();
        u32 bk = 0, bc = 0, tk = 0, tc = 0;
#pragma unroll
        for (int w = 0; w < UQ_T / 32; ++w) { const u32 x = s_w[0][w], y = s_w[1][w]; tk += x; tc += y; if (w < wid) { bk += x; bc += y; } }
        if (wid == 0) {
            const u64 agg = (u64)tk | ((u64)tc << 31);
            const u64 b = lookback_exclusive(desc, tile, 0, agg, lane);
            if (lane == 0) { s_base = b; if (tile == n_tiles - 1) { counters[0] = (b & 0x7FFFFFFFu) + tk; counters[1] = (b >> 31) + tc; } }
        }
        __syncthreads();
        u64 ok = (s_base & 0x7FFFFFFFu) + bk + ik - nk, oc = (s_base >> 31) + bc + ic - nc;
        u32 hint = 0;
#pragma unroll
        for (int k = 0; k < UQ_ITEMS; ++k) {
            if (fc & (1u << k)) {
                if (oc < cell_cap) {
                    u64 lo = (u64)r[k].x | ((u64)r[k].y << 32), hi = (u64)r[k].z;
                    take_bits(lo, hi, cell_shift);
                    const u32 bb = (u32)take_bits(lo, hi, c.nb), aa = (u32)take_bits(lo, hi, c.nb);
                    cell_b1[oc] = aa; cell_b2[oc] = bb; cell_first[oc] = (u32)ok;
                }
                ++oc;
            }
            if (fk & (1u << k)) {
                out[ok] = unpack_key(r[k], c, dec_off, dec_id, hint);
                if (keep) keep[r[k].w] = 1;
                if (kept_idx) kept_idx[ok] = r[k].w;
                ++ok;
            }
        }
        __syncthreads();                                                // s_w / s_base / s_tile are reused by the next tile
    }
}

// pairs that could not be keyed form the last run (all-ones key): it is neither a kept pair nor a cell
__global__ void k_uniq_trim(const uint4 *b0, const uint4 *b1, const RadixPlan *plan, u64 n, u8 *keep, unsigned long long *counters) {
    const uint4 *in = plan->final_buf ? b1 : b0;
    if (!counters[3]) return;
    const u64 first_bad = n - counters[3];
    const uint4 k = in[first_bad];
    if (k.x == 0xFFFFFFFFu && k.y == 0xFFFFFFFFu && k.z == 0xFFFFFFFFu) {
        counters[0] -= 1; counters[1] -= 1;
        if (keep) keep[k.w] = 0;
    }
}

__global__ void k_cell_counts(const u32 *cell_first, const unsigned long long *counters, u64 cap, u32 *cnt) {
    const u64 kept = counters[0], cells = counters[1] < cap ? counters[1] : cap;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (u64)gridDim.x * blockDim.x)
        cnt[i] = (i + 1 < counters[1] && i + 1 < cap ? cell_first[i + 1] : (u32)kept) - cell_first[i];
}

// decoded pairs landed in the sort's other buffer: bring them home when that is not the caller's
__global__ void k_copy_if_alt(const RadixPlan *plan, u32 keys_in_home_when, const uint4 *alt, uint4 *home, const unsigned long long *counters) {
    if (plan->final_buf != keys_in_home_when) return;                  // the sorted keys ended in the workspace: the pairs are already home
    const u64 n = counters[0];
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) home[i] = alt[i];
}

static u32 bits_for(u64 max_value) { u32 b = 1; while (b < 64 && (max_value >> b)) ++b; return b; }

extern "C" int mk_pairs_dedup_bin_indexed_device(mk_pairs_ws *w, mk_pair *d_pairs, size_t n, const uint32_t *chrom_len, int n_chrom,
                                                 const uint16_t *chrom_id_map, int n_map, uint32_t res, uint16_t max_lane,
                                                 uint32_t *d_bin1, uint32_t *d_bin2, uint32_t *d_cnt, size_t cap,
                                                 uint8_t *d_keep, uint32_t *d_kept_idx, size_t *n_kept, size_t *nnz, void *stream) {
    if (!w || !n_kept || !nnz || !chrom_len || n_chrom <= 0 || res == 0 || !d_bin1 || !d_bin2 || !d_cnt) { mk_set_error("mk_pairs_dedup_bin_device: bad argument"); return MK_ERR_ARG; }
    if (n > w->max_pairs) { mk_set_error("mk_pairs_dedup_bin_device: workspace holds %zu pairs, got %zu", w->max_pairs, n); return MK_ERR_CAPACITY; }
    MK_CUDA(cudaSetDevice(w->device));
    cudaStream_t s = (cudaStream_t)stream;
    *n_kept = 0; *nnz = 0; w->dropped = 0;
    if (n == 0) return MK_OK;
    std::vector<u64> off(n_chrom + 1, 0);
    for (int c = 0; c < n_chrom; ++c) off[c + 1] = off[c] + chrom_len[c] / res + 1;
    if (off[n_chrom] >= (1ull << 32)) { mk_set_error("mk_pairs_dedup_bin_device: more than 2^32 bins"); return MK_ERR_CAPACITY; }
    const int n_ids = chrom_id_map ? n_map : n_chrom;
    if (n_ids > 16384 || n_chrom > 16384) { mk_set_error("mk_pairs_dedup_bin_device: too many chromosomes"); return MK_ERR_ARG; }
    // one upload: [by_id offsets (16384 u32)] [length per id (16384 u32)] [dec_off (16384 u32)] [dec_id (16384 u16)]
    std::vector<u32> tab(3 * 16384 + 8192, 0);
    u32 *by_id = tab.data(), *nb_id = by_id + 16384, *dec_off = nb_id + 16384; u16 *dec_id = (u16 *)(dec_off + 16384);
    for (int c = 0; c < n_chrom; ++c) { dec_off[c] = (u32)off[c]; dec_id[c] = 0xFFFF; }
    for (int i = 0; i < n_ids; ++i) {
        int c = chrom_id_map ? chrom_id_map[i] : i;
        if (c < 0 || c >= n_chrom) { mk_set_error("mk_pairs_dedup_bin_device: chromosome map entry %d out of range", i); return MK_ERR_ARG; }
        by_id[i] = (u32)off[c]; nb_id[i] = chrom_len[c];
        if (dec_id[c] == 0xFFFF) dec_id[c] = (u16)i;
    }
    PackCfg pc; pc.res = res; pc.nb = bits_for(off[n_chrom]) /* one value past the last bin stays free for the all-ones key */;
    pc.nr = bits_for(res - 1); pc.nl = max_lane ? bits_for(max_lane) : 0;
    pc.total_bits = 2 * pc.nb + 2 * pc.nr + pc.nl + 3; pc.n_dec = (u32)n_chrom; pc.n_ids = (u32)n_ids; pc.max_lane = max_lane;
    if (pc.total_bits > 96) { mk_set_error("mk_pairs_dedup_bin_device: key needs %u bits, 96 available", pc.total_bits); return MK_ERR_CAPACITY; }
    u32 *d_by_id = w->chr_off.as<u32>(), *d_nb_id = d_by_id + 16384, *d_dec_off = d_nb_id + 16384; u16 *d_dec_id = (u16 *)(d_dec_off + 16384);
    MK_CUDA(cudaMemcpyAsync(d_by_id, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, s));    // (pageable source: staged before the call returns)
    MK_CUDA(cudaMemsetAsync(w->counter.p, 0, 64, s));
    unsigned long long *cnt = w->counter.as<unsigned long long>();
    if (d_keep) MK_CUDA(cudaMemsetAsync(d_keep, 0, n, s));
    RadixSchedule sch; sch.n_pass = (int)((pc.total_bits + 7) / 8);
    for (int i = 0; i < sch.n_pass; ++i) sch.byte_of[i] = i;
    // The pairs buffer doubles as one of the two sort buffers.  The decoded pairs go to the buffer the sorted keys are NOT in:
    // with an odd number of passes the keys start IN the pairs buffer (packed in place: a thread reads pair i and writes key i),
    // end in the workspace, and the kept pairs land where the caller wants them without a copy (k_copy_if_alt covers the
    // case of a pass the device found trivial and skipped).
    const bool in_place = (sch.n_pass & 1) != 0;
    uint4 *k0 = in_place ? (uint4 *)d_pairs : w->alt.as<uint4>(), *k1 = in_place ? w->alt.as<uint4>() : (uint4 *)d_pairs;
    MK_CUDA(cudaMemsetAsync(w->rws.plan.p, 0, sizeof(RadixPlan), s));
    k_pack_keys<<<w->sms * 8, 256, sch.n_pass * 256 * 4, s>>>(d_pairs, n, d_by_id, d_nb_id, pc, k0, cnt + 3, sch, w->rws.plan.as<RadixPlan>());
    w->launches += 1;
    Rec16::Bufs b; b.k[0] = k0; b.k[1] = k1; b.v[0] = b.v[1] = nullptr;
    MK_TRY(radix_sort<Rec16>(b, n, sch, w->rws, 0, w->sms, s, &w->launches, /*hist_done=*/true));
    const int n_tiles = (int)((n + UQ_T * UQ_ITEMS - 1) / (UQ_T * UQ_ITEMS));
    MK_CUDA(cudaMemsetAsync(w->desc.p, 0, (size_t)(n_tiles + 1) * 8, s));
    const RadixPlan *plan = w->rws.plan.as<RadixPlan>();
    k_uniq_cells<<<lookback_grid((const void *)k_uniq_cells, UQ_T, w->sms), UQ_T, 0, s>>>(
        k0, k1, plan, n, pc, d_dec_off, d_dec_id, (mk_pair *)k0, (mk_pair *)k1,
        d_bin1, d_bin2, w->heads_pos.as<u32>(), cap, d_keep, d_kept_idx, w->desc.as<u64>(), cnt, (u32 *)(cnt + 4));
    k_uniq_trim<<<1, 1, 0, s>>>(k0, k1, plan, n, d_keep, cnt);
    k_cell_counts<<<w->sms * 4, 256, 0, s>>>(w->heads_pos.as<u32>(), cnt, cap, d_cnt);
    // sorted keys sat in buffer final_buf; the decoded pairs went to the other one: bring them home if that is the workspace
    k_copy_if_alt<<<w->sms * 4, 256, 0, s>>>(plan, in_place ? 0u : 1u, w->alt.as<uint4>(), (uint4 *)d_pairs, cnt);
    w->launches += 4;
    unsigned long long h[4];
    MK_CUDA(cudaMemcpyAsync(h, cnt, 32, cudaMemcpyDeviceToHost, s));
    MK_CUDA(cudaStreamSynchronize(s));                                  // the one host round trip: the two result counts
    *n_kept = (size_t)h[0]; *nnz = (size_t)h[1]; w->dropped = h[3];
    if (h[1] > cap) { mk_set_error("mk_pairs_dedup_bin_device: %llu non-zero cells, output capacity %zu", h[1], cap); return MK_ERR_CAPACITY; }
    return MK_OK;
}

extern "C" int mk_pairs_dedup_bin_device(mk_pairs_ws *w, mk_pair *d_pairs, size_t n, const uint32_t *chrom_len, int n_chrom,
                                         const uint16_t *chrom_id_map, int n_map, uint32_t res, uint16_t max_lane,
                                         uint32_t *d_bin1, uint32_t *d_bin2, uint32_t *d_cnt, size_t cap,
                                         size_t *n_kept, size_t *nnz, void *stream) {
    return mk_pairs_dedup_bin_indexed_device(w, d_pairs, n, chrom_len, n_chrom, chrom_id_map, n_map, res, max_lane, d_bin1, d_bin2, d_cnt, cap,
                                             nullptr, nullptr, n_kept, nnz, stream);
}
