// hist.cu — contact binning as a dense histogram, all resolutions from ONE read of the packed pairs.
//
// For every requested resolution r: bin = offset_r[chr] + pos / r (util/analyze.EBV/calc.loop2EBV.pl:28; chromosome order
// and lengths from anno/<genome>.info), cell (bin1 <= bin2) of an upper-triangle u32 matrix.  Stands in for the counting
// half of `juicer_tools pre -r <list>` (microcket:525-529; default list microcket:98) at the resolutions whose triangle fits
// in HBM (hg38: >= 100 kb = 1.9 GB, and coarser); finer ones go through the sort path (mk_pairs_bin_device).
//
//   k_hist_add   one pass over the pairs; per pair and resolution one increment.  Hi-C counts pile up on the diagonal
//                (half of all pairs are closer than 1 kb), so every CTA keeps the DIAGONALS of all resolutions privately in
//                shared memory (hg38, five resolutions: 53 802 counters = 215 KB) and only off-diagonal cells — spread over
//                the whole triangle, no contention — go to global memory as atomics ("spill").  The private diagonals are
//                flushed once per CTA.
//   k_hist_coo   non-zero cells of one triangle in (bin1, bin2) order (= storage order) as COO triplets: single pass,
//                tiles claimed by ticket, decoupled look-back for the output offsets.
// The matrices can live in caller memory, so that several GPUs add into their own copy and the copies are summed with one
// reduce (NCCL) before the extraction.
#include <algorithm>
#include <vector>
#include "mk_common.cuh"

#define HS_MAX_RES 12
#define HS_THREADS 1024
#define HC_T 256
#define HC_ITEMS 16

struct HistRes {
    u32 res, nb;                 // resolution, total bins
    u32 diag_off;                // first word of this resolution's private diagonal in shared memory, or 0xFFFFFFFF: all global
    u32 pad;
    u32 *cells;                  // triangle, row-major: cell (a, b >= a) at a * nb - a (a - 1) / 2 + (b - a)
    const u32 *off_by_id;        // bin offset of every pair-chromosome id at this resolution
};
struct HistCfg { int n_res; u32 n_ids; u32 diag_words; const u32 *len_by_id; HistRes r[HS_MAX_RES]; };

__host__ __device__ __forceinline__ u64 tri_row_start(u64 a, u64 nb) { return a * nb - (a * (a - 1)) / 2; }   // a = 0 -> 0 (a - 1 wraps, times 0)

__global__ void __launch_bounds__(HS_THREADS, 1) k_hist_add(const mk_pair *p, u64 n, HistCfg c, unsigned long long *bad) {
    extern __shared__ u32 s_diag[];
    for (u32 i = threadIdx.x; i < c.diag_words; i += HS_THREADS) s_diag[i] = 0;
    __syncthreads();
    u32 nbad = 0;
    for (u64 i = (u64)blockIdx.x * HS_THREADS + threadIdx.x; i < n; i += (u64)gridDim.x * HS_THREADS) {
        const uint4 r = ((const uint4 *)p)[i];
        const u32 pos1 = r.x, pos2 = r.y, c1 = r.z & 0xFFFFu, c2 = r.z >> 16;
        // one rule for every resolution: known chromosome, position inside it (so pos / res is always below its bin count)
        if (c1 >= c.n_ids || c2 >= c.n_ids || pos1 > c.len_by_id[c1] || pos2 > c.len_by_id[c2]) { ++nbad; continue; }
#pragma unroll 1
        for (int k = 0; k < c.n_res; ++k) {
            const HistRes &h = c.r[k];
            const u32 q1 = pos1 / h.res, q2 = pos2 / h.res;
            u32 a = h.off_by_id[c1] + q1, b = h.off_by_id[c2] + q2;
            if (a > b) { const u32 t = a; a = b; b = t; }
            if (a == b && h.diag_off != 0xFFFFFFFFu) atomicAdd(&s_diag[h.diag_off + a], 1u);
            else atomicAdd(&h.cells[tri_row_start(a, h.nb) + (b - a)], 1u);
        }
    }
    __syncthreads();
    for (int k = 0; k < c.n_res; ++k) {
        const HistRes &h = c.r[k];
        if (h.diag_off == 0xFFFFFFFFu) continue;
        for (u32 a = threadIdx.x; a < h.nb; a += HS_THREADS) {
            const u32 v = s_diag[h.diag_off + a];
            if (v) atomicAdd(&h.cells[tri_row_start(a, h.nb)], v);
        }
    }
    nbad = __reduce_add_sync(0xffffffffu, nbad);
    if ((threadIdx.x & 31) == 0 && nbad) atomicAdd(bad, (unsigned long long)nbad);
}

// (a, b) of linear triangle index L
__device__ __forceinline__ void tri_unrank(u64 L, u64 nb, u32 &a, u32 &b) {
    const double t = 2.0 * (double)nb + 1.0;
    double x = (t - sqrt(t * t - 8.0 * (double)L)) * 0.5;
    u64 r = x > 0 ? (u64)x : 0;
    if (r >= nb) r = nb - 1;
    while (r > 0 && tri_row_start(r, nb) > L) --r;
    while (r + 1 < nb && tri_row_start(r + 1, nb) <= L) ++r;
    a = (u32)r; b = (u32)(r + (L - tri_row_start(r, nb)));
}

__global__ void __launch_bounds__(HC_T) k_hist_coo(const u32 *cells, u64 n_cells, u32 nb, u32 *bin1, u32 *bin2, u32 *cnt, u64 cap,
                                                   u64 *desc, unsigned long long *counters /* [0] nnz, [1] sum */, u32 *ticket) {
    __shared__ u32 s_scan[HC_T / 32 + 1];
    __shared__ u64 s_base;
    __shared__ u32 s_tile;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const u64 n_tiles = (n_cells + HC_T * HC_ITEMS - 1) / (HC_T * HC_ITEMS);
    unsigned long long sum = 0;
    while (true) {
        if (tid == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const u64 tile = s_tile;
        if (tile >= n_tiles) break;
        const u64 base = tile * HC_T * HC_ITEMS + (u64)tid * HC_ITEMS;
        u32 v[HC_ITEMS]; u32 nz = 0;
        if (base + HC_ITEMS <= n_cells && ((uintptr_t)(cells + base) & 15) == 0) {
#pragma unroll
            for (int k = 0; k < HC_ITEMS; k += 4) { const uint4 w = ld_stream_v4((const uint4 *)(cells + base + k)); v[k] = w.x; v[k + 1] = w.y; v[k + 2] = w.z; v[k + 3] = w.w; }
        } else {
#pragma unroll
            for (int k = 0; k < HC_ITEMS; ++k) v[k] = base + k < n_cells ? cells[base + k] : 0u;
        }
#pragma unroll
        for (int k = 0; k < HC_ITEMS; ++k) { nz += v[k] != 0; sum += v[k]; }
        u32 tot;
        const u32 ex = block_excl_scan<HC_T>(nz, s_scan, &tot);
        if (wid == 0) {
            const u64 b = lookback_exclusive(desc, (int)tile, 0, tot, lane);
            if (lane == 0) { s_base = b; if (tile == n_tiles - 1) counters[0] = b + tot; }
        }
        __syncthreads();
        if (nz) {
            u64 o = s_base + ex;
            u32 a, b;
            tri_unrank(base, nb, a, b);
#pragma unroll
            for (int k = 0; k < HC_ITEMS; ++k) {
                if (v[k]) { if (o < cap) { bin1[o] = a; bin2[o] = b; cnt[o] = v[k]; } ++o; }
                if (++b == nb) { ++a; b = a; }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if (lane == 0 && sum) atomicAdd(&counters[1], sum);
}

struct mk_hist {
    int device = 0, sms = 148, n_res = 0, n_chrom = 0;
    std::vector<u32> chrom_len, res;
    std::vector<u64> n_cells, nb;
    DevBuf own[HS_MAX_RES];                  // matrices this object allocated (empty when the caller provided them)
    std::vector<u32 *> cells;
    DevBuf tables, desc, counter;            // per resolution: [off_by_id | nb_by_id] x 16384 ids; look-back descriptors; counters
    int n_ids_loaded = -1; std::vector<u16> map_loaded;
    u64 launches = 0, dropped = 0, added = 0;
    u32 diag_off[HS_MAX_RES]; u32 diag_words = 0;
};

static u64 hist_bins(const u32 *chrom_len, int n_chrom, u32 res) {
    u64 nb = 0;
    for (int c = 0; c < n_chrom; ++c) nb += chrom_len[c] / res + 1;
    return nb;
}

extern "C" int mk_hist_cells(const uint32_t *chrom_len, int n_chrom, uint32_t res, uint64_t *n_bins, uint64_t *n_cells) {
    if (!chrom_len || n_chrom <= 0 || res == 0) { mk_set_error("mk_hist_cells: bad argument"); return MK_ERR_ARG; }
    const u64 nb = hist_bins(chrom_len, n_chrom, res);
    if (n_bins) *n_bins = nb;
    if (n_cells) *n_cells = nb * (nb + 1) / 2;
    return MK_OK;
}

extern "C" int mk_hist_create(int device, const uint32_t *chrom_len, int n_chrom, const uint32_t *res, int n_res,
                              uint32_t *const *d_cells, mk_hist **out) {
    if (!out || !chrom_len || !res || n_chrom <= 0 || n_chrom > 16384 || n_res <= 0 || n_res > HS_MAX_RES) { mk_set_error("mk_hist_create: bad argument"); return MK_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { mk_set_error("no CUDA device: microcket_b200 has no CPU fallback"); return MK_ERR_CUDA; }
    MK_CUDA(cudaSetDevice(device));
    mk_hist *h = new mk_hist();
    h->device = device; h->sms = mk_sm_count(device); h->n_res = n_res; h->n_chrom = n_chrom;
    h->chrom_len.assign(chrom_len, chrom_len + n_chrom); h->res.assign(res, res + n_res);
    h->cells.resize(n_res); h->n_cells.resize(n_res); h->nb.resize(n_res);
    int rc = MK_OK;
    u64 max_cells = 0;
    for (int k = 0; k < n_res && rc == MK_OK; ++k) {
        if (res[k] == 0) { mk_set_error("mk_hist_create: resolution 0"); rc = MK_ERR_ARG; break; }
        h->nb[k] = hist_bins(chrom_len, n_chrom, res[k]);
        h->n_cells[k] = h->nb[k] * (h->nb[k] + 1) / 2;
        if (h->nb[k] >= (1ull << 32)) { mk_set_error("mk_hist_create: more than 2^32 bins"); rc = MK_ERR_CAPACITY; break; }
        max_cells = std::max(max_cells, h->n_cells[k]);
        if (d_cells && d_cells[k]) h->cells[k] = d_cells[k];
        else {
            rc = h->own[k].alloc(h->n_cells[k] * 4);
            if (rc == MK_OK) { h->cells[k] = h->own[k].as<u32>(); if (cudaMemset(h->cells[k], 0, h->n_cells[k] * 4) != cudaSuccess) rc = MK_ERR_CUDA; }
        }
    }
    if (rc == MK_OK) rc = h->tables.alloc((size_t)(n_res + 1) * 16384 * 4);
    if (rc == MK_OK) rc = h->desc.alloc((max_cells / (HC_T * HC_ITEMS) + 4) * 8);
    if (rc == MK_OK) rc = h->counter.alloc(64);
    if (rc != MK_OK) { delete h; return rc; }
    // private diagonals: coarsest resolutions first, as many as fit the shared memory of one CTA
    int smem_max = 0;
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    const u32 budget = (u32)std::max(0, smem_max - 1024) / 4;
    std::vector<int> order(n_res);
    for (int k = 0; k < n_res; ++k) order[k] = k;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return h->nb[a] < h->nb[b]; });
    u32 used = 0;
    for (int k = 0; k < n_res; ++k) h->diag_off[k] = 0xFFFFFFFFu;
    for (int k : order) if (used + h->nb[k] <= budget) { h->diag_off[k] = used; used += (u32)h->nb[k]; }
    h->diag_words = used;
    MK_CUDA(cudaFuncSetAttribute(k_hist_add, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(used * 4 + 16)));
    *out = h;
    return MK_OK;
}

extern "C" void mk_hist_destroy(mk_hist *h) { if (h) { cudaSetDevice(h->device); delete h; } }
extern "C" uint64_t mk_hist_dropped(mk_hist *h) { return h ? h->dropped : 0; }
extern "C" uint64_t mk_hist_launch_count(mk_hist *h) { return h ? h->launches : 0; }

extern "C" int mk_hist_matrix(mk_hist *h, int res_idx, uint32_t **d_cells, uint64_t *n_cells, uint64_t *n_bins) {
    if (!h || res_idx < 0 || res_idx >= h->n_res) { mk_set_error("mk_hist_matrix: bad argument"); return MK_ERR_ARG; }
    if (d_cells) *d_cells = h->cells[res_idx];
    if (n_cells) *n_cells = h->n_cells[res_idx];
    if (n_bins) *n_bins = h->nb[res_idx];
    return MK_OK;
}

extern "C" int mk_hist_reset(mk_hist *h, void *stream) {
    if (!h) { mk_set_error("mk_hist_reset: null"); return MK_ERR_ARG; }
    MK_CUDA(cudaSetDevice(h->device));
    for (int k = 0; k < h->n_res; ++k) MK_CUDA(cudaMemsetAsync(h->cells[k], 0, h->n_cells[k] * 4, (cudaStream_t)stream));
    h->dropped = 0; h->added = 0;
    return MK_OK;
}

// Accumulate n pairs into every resolution's matrix.  chrom_id_map (host, may be NULL) maps mk_pair chromosome ids to
// indices of chrom_len, as in mk_pairs_bin_device.  Unkeyable pairs are left out and counted (mk_hist_dropped).
extern "C" int mk_hist_add_device(mk_hist *h, const mk_pair *d_pairs, size_t n, const uint16_t *chrom_id_map, int n_map, void *stream) {
    if (!h || (n && !d_pairs)) { mk_set_error("mk_hist_add_device: bad argument"); return MK_ERR_ARG; }
    MK_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int n_ids = chrom_id_map ? n_map : h->n_chrom;
    if (n_ids <= 0 || n_ids > 16384) { mk_set_error("mk_hist_add_device: bad chromosome map"); return MK_ERR_ARG; }
    std::vector<u16> mp(n_ids);
    for (int i = 0; i < n_ids; ++i) {
        const int c = chrom_id_map ? chrom_id_map[i] : i;
        if (c < 0 || c >= h->n_chrom) { mk_set_error("mk_hist_add_device: chromosome map entry %d out of range", i); return MK_ERR_ARG; }
        mp[i] = (u16)c;
    }
    if (h->n_ids_loaded != n_ids || mp != h->map_loaded) {              // tables change only when the id map does
        std::vector<u32> tab((size_t)(h->n_res + 1) * 16384, 0);       // [lengths by id] then per resolution [offsets by id]
        for (int i = 0; i < n_ids; ++i) tab[i] = h->chrom_len[mp[i]];
        for (int k = 0; k < h->n_res; ++k) {
            std::vector<u64> off(h->n_chrom + 1, 0);
            for (int c = 0; c < h->n_chrom; ++c) off[c + 1] = off[c] + h->chrom_len[c] / h->res[k] + 1;
            u32 *o = tab.data() + (size_t)(k + 1) * 16384;
            for (int i = 0; i < n_ids; ++i) o[i] = (u32)off[mp[i]];
        }
        MK_CUDA(cudaMemcpyAsync(h->tables.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, s));
        MK_CUDA(cudaStreamSynchronize(s));                              // `tab` is a local
        h->n_ids_loaded = n_ids; h->map_loaded = mp;
    }
    if (n == 0) return MK_OK;
    HistCfg c; c.n_res = h->n_res; c.n_ids = (u32)n_ids; c.diag_words = h->diag_words; c.len_by_id = h->tables.as<u32>();
    for (int k = 0; k < h->n_res; ++k) {
        c.r[k].res = h->res[k]; c.r[k].nb = (u32)h->nb[k]; c.r[k].diag_off = h->diag_off[k]; c.r[k].pad = 0; c.r[k].cells = h->cells[k];
        c.r[k].off_by_id = h->tables.as<u32>() + (size_t)(k + 1) * 16384;
    }
    MK_CUDA(cudaMemsetAsync(h->counter.p, 0, 64, s));
    const int grid = (int)std::min<u64>((u64)h->sms, (n + HS_THREADS - 1) / HS_THREADS);
    k_hist_add<<<grid, HS_THREADS, h->diag_words * 4 + 16, s>>>(d_pairs, n, c, h->counter.as<unsigned long long>());
    h->launches += 1;
    unsigned long long bad = 0;
    MK_CUDA(cudaMemcpyAsync(&bad, h->counter.p, 8, cudaMemcpyDeviceToHost, s));
    MK_CUDA(cudaStreamSynchronize(s));
    MK_CUDA(cudaGetLastError());
    h->dropped += bad; h->added += n - bad;
    return MK_OK;
}

// Non-zero cells of one resolution as COO triplets sorted by (bin1, bin2), bin1 <= bin2; *total = sum of the counts.
extern "C" int mk_hist_coo_device(mk_hist *h, int res_idx, uint32_t *d_bin1, uint32_t *d_bin2, uint32_t *d_cnt, size_t cap,
                                  size_t *nnz, uint64_t *total, void *stream) {
    if (!h || res_idx < 0 || res_idx >= h->n_res || !nnz || (cap && (!d_bin1 || !d_bin2 || !d_cnt))) { mk_set_error("mk_hist_coo_device: bad argument"); return MK_ERR_ARG; }
    MK_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const u64 n_cells = h->n_cells[res_idx];
    const u64 n_tiles = (n_cells + HC_T * HC_ITEMS - 1) / (HC_T * HC_ITEMS);
    if (n_tiles >= (1ull << 31)) { mk_set_error("mk_hist_coo_device: matrix too large"); return MK_ERR_CAPACITY; }
    MK_CUDA(cudaMemsetAsync(h->desc.p, 0, (n_tiles + 1) * 8, s));
    MK_CUDA(cudaMemsetAsync(h->counter.p, 0, 64, s));
    unsigned long long *cnt = h->counter.as<unsigned long long>();
    const int grid = (int)std::min<u64>((u64)h->sms * 8, n_tiles);
    k_hist_coo<<<grid, HC_T, 0, s>>>(h->cells[res_idx], n_cells, (u32)h->nb[res_idx], d_bin1, d_bin2, d_cnt, cap, h->desc.as<u64>(), cnt, (u32 *)(cnt + 4));
    h->launches += 1;
    unsigned long long r[2];
    MK_CUDA(cudaMemcpyAsync(r, cnt, 16, cudaMemcpyDeviceToHost, s));
    MK_CUDA(cudaStreamSynchronize(s));
    MK_CUDA(cudaGetLastError());
    *nnz = (size_t)r[0];
    if (total) *total = r[1];
    if (r[0] > cap) { mk_set_error("mk_hist_coo_device: %llu non-zero cells, output capacity %zu", r[0], cap); return MK_ERR_CAPACITY; }
    return MK_OK;
}
