// synth.cu — synthetic SAM / FASTQ generation on host and device (identical bytes), see synth.h.
#include <vector>
#include "mk_common.cuh"
#include "synth.h"

// exclusive scan of u32 lengths into u64 offsets, single pass with decoupled look-back
#define SCAN_T 256
#define SCAN_ITEMS 8
__global__ void __launch_bounds__(SCAN_T) k_len_scan(const u32 *len, u64 *off, u64 n, u64 *desc, u64 *total) {
    __shared__ u32 s_scan[SCAN_T / 32 + 1];
    __shared__ u64 s_base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n_tiles = (int)((n + SCAN_T * SCAN_ITEMS - 1) / (SCAN_T * SCAN_ITEMS));
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        u64 base = (u64)tile * SCAN_T * SCAN_ITEMS + (u64)tid * SCAN_ITEMS;
        u32 v[SCAN_ITEMS]; u32 sum = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) { v[k] = base + k < n ? len[base + k] : 0; sum += v[k]; }
        u32 tot;
        u32 ex = block_excl_scan<SCAN_T>(sum, s_scan, &tot);
        if (wid == 0) {
            u64 b = lookback_exclusive(desc, tile, 0, tot, lane);
            if (lane == 0) { s_base = b; if (tile == n_tiles - 1) *total = b + tot; }
        }
        __syncthreads();
        u64 o = s_base + ex;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) { if (base + k < n) off[base + k] = o; o += v[k]; }
        __syncthreads();
    }
}

__global__ void k_synth_len(mk_synth_cfg cfg, int mode, u64 first, u64 count, u32 *len) {
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < count; i += (u64)gridDim.x * blockDim.x) {
        mk_sink w; w.p = nullptr; w.n = 0;
        if (mode == 2) mk_gen_fastq_pair(&cfg, first + i, &w); else mk_gen_group(&cfg, first + i, &w);
        len[i] = (u32)w.n;
    }
}
__global__ void k_synth_write(mk_synth_cfg cfg, int mode, u64 first, u64 count, const u64 *off, char *out, u64 cap) {
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < count; i += (u64)gridDim.x * blockDim.x) {
        mk_sink w; w.p = out + off[i]; w.n = 0;
        u64 end = (i + 1 < count) ? off[i + 1] : cap;
        if (end > cap) continue;
        if (mode == 2) mk_gen_fastq_pair(&cfg, first + i, &w); else mk_gen_group(&cfg, first + i, &w);
    }
}

static void synth_cfg(mk_synth_cfg *c, uint64_t seed, int mode, int genome, const mk_synth_opts *o = nullptr) {
    mk_synth_init(c, seed, mode == 0 ? 0 : 1, genome);
    if (!o) return;
    if (o->dup_per_1024 > 0) { c->sam_dup_per_1024 = o->dup_per_1024; c->sam_dup_universe = o->dup_universe; c->dup_per_1024 = o->dup_per_1024; }
    if (o->chimeric_per_1024 >= 0) c->w_chimeric = o->chimeric_per_1024;
    if (o->noise_per_1024 >= 0) c->w_noise = o->noise_per_1024;
    if (o->selfcircle_per_1024 >= 0) c->w_selfcircle = o->selfcircle_per_1024;
}

extern "C" void mk_synth_default_opts(mk_synth_opts *o) {
    if (!o) return;
    o->dup_per_1024 = 0; o->dup_universe = 0; o->chimeric_per_1024 = -1; o->noise_per_1024 = -1; o->selfcircle_per_1024 = -1;
}
extern "C" int mk_synth_host(uint64_t seed, int mode, int genome, uint64_t first, uint64_t count, char *buf, size_t cap, size_t *n_out) {
    return mk_synth_host_ex(seed, mode, genome, nullptr, first, count, buf, cap, n_out);
}
extern "C" int mk_synth_device(int device, uint64_t seed, int mode, int genome, uint64_t first, uint64_t count,
                               char *d_buf, size_t cap, size_t *n_out, void *stream) {
    return mk_synth_device_ex(device, seed, mode, genome, nullptr, first, count, d_buf, cap, n_out, stream);
}

extern "C" int mk_synth_host_ex(uint64_t seed, int mode, int genome, const mk_synth_opts *opts, uint64_t first, uint64_t count,
                                char *buf, size_t cap, size_t *n_out) {
    if (mode < 0 || mode > 2 || !n_out) { mk_set_error("mk_synth_host: bad argument"); return MK_ERR_ARG; }
    mk_synth_cfg c; synth_cfg(&c, seed, mode, genome, opts);
    mk_sink w; w.p = nullptr; w.n = 0;
    for (uint64_t i = 0; i < count; ++i) { if (mode == 2) mk_gen_fastq_pair(&c, first + i, &w); else mk_gen_group(&c, first + i, &w); }
    size_t need = w.n;
    *n_out = need;
    if (!buf) return MK_OK;
    if (need > cap) { mk_set_error("mk_synth_host: buffer too small (%zu needed)", need); return MK_ERR_CAPACITY; }
    w.p = buf; w.n = 0;
    for (uint64_t i = 0; i < count; ++i) { if (mode == 2) mk_gen_fastq_pair(&c, first + i, &w); else mk_gen_group(&c, first + i, &w); }
    return MK_OK;
}

extern "C" int mk_synth_device_ex(int device, uint64_t seed, int mode, int genome, const mk_synth_opts *opts, uint64_t first, uint64_t count,
                                  char *d_buf, size_t cap, size_t *n_out, void *stream) {
    if (mode < 0 || mode > 2 || !n_out) { mk_set_error("mk_synth_device: bad argument"); return MK_ERR_ARG; }
    MK_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    mk_synth_cfg c; synth_cfg(&c, seed, mode, genome, opts);
    if (count == 0) { *n_out = 0; return MK_OK; }
    DevBuf d_len, d_off, d_desc, d_total;
    const int n_tiles = (int)((count + SCAN_T * SCAN_ITEMS - 1) / (SCAN_T * SCAN_ITEMS));
    MK_TRY(d_len.alloc(count * 4)); MK_TRY(d_off.alloc(count * 8)); MK_TRY(d_desc.alloc((size_t)n_tiles * 8 + 8)); MK_TRY(d_total.alloc(8));
    MK_CUDA(cudaMemsetAsync(d_desc.p, 0, (size_t)n_tiles * 8 + 8, s));
    int sms = mk_sm_count(device);
    k_synth_len<<<sms * 8, 256, 0, s>>>(c, mode, first, count, d_len.as<u32>());
    int occ = 1; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_len_scan, SCAN_T, 0);
    k_len_scan<<<sms * (occ > 4 ? 4 : (occ < 1 ? 1 : occ)), SCAN_T, 0, s>>>(d_len.as<u32>(), d_off.as<u64>(), count, d_desc.as<u64>(), d_total.as<u64>());
    u64 total = 0;
    MK_CUDA(cudaMemcpyAsync(&total, d_total.p, 8, cudaMemcpyDeviceToHost, s));
    MK_CUDA(cudaStreamSynchronize(s));
    *n_out = total;
    if (!d_buf) return MK_OK;
    if (total > cap) { mk_set_error("mk_synth_device: buffer too small (%llu needed)", (unsigned long long)total); return MK_ERR_CAPACITY; }
    k_synth_write<<<sms * 8, 256, 0, s>>>(c, mode, first, count, d_off.as<u64>(), d_buf, total);
    MK_CUDA(cudaStreamSynchronize(s));
    MK_CUDA(cudaGetLastError());
    return MK_OK;
}
