// mk_common.cuh — shared device helpers and host-side error plumbing.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/microcket_b200.h"

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

// ---------------------------------------------------------------- host error plumbing
void mk_set_error(const char *fmt, ...);
#define MK_CUDA(call)                                                                        \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            mk_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return MK_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)
#define MK_TRY(call) do { int r__ = (call); if (r__ != MK_OK) return r__; } while (0)

struct DevBuf {   // owning device allocation
    void *p = nullptr; size_t n = 0;
    int alloc(size_t bytes) {
        free();
        if (bytes == 0) bytes = 16;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) { p = nullptr; mk_set_error("cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); return MK_ERR_NOMEM; }
        n = bytes; return MK_OK;
    }
    void free() { if (p) cudaFree(p); p = nullptr; n = 0; }
    template <class T> T *as() const { return (T *)p; }
    ~DevBuf() { free(); }
    DevBuf() {}
    DevBuf(const DevBuf &) = delete; DevBuf &operator=(const DevBuf &) = delete;
};
struct PinBuf {   // owning pinned host allocation
    void *p = nullptr; size_t n = 0;
    int alloc(size_t bytes) {
        free();
        if (bytes == 0) bytes = 16;
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e != cudaSuccess) { p = nullptr; mk_set_error("cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e)); return MK_ERR_NOMEM; }
        n = bytes; return MK_OK;
    }
    void free() { if (p) cudaFreeHost(p); p = nullptr; n = 0; }
    template <class T> T *as() const { return (T *)p; }
    ~PinBuf() { free(); }
    PinBuf() {}
    PinBuf(const PinBuf &) = delete; PinBuf &operator=(const PinBuf &) = delete;
};

int mk_sm_count(int device);

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__

__device__ __forceinline__ uint4 ld_stream_v4(const uint4 *p) {   // streaming 128-bit load, no L1 allocation
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_v4(uint4 *p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}

// 0x80 in every byte of x that equals the byte value replicated in `pat` (exact, no cross-byte carries)
__device__ __forceinline__ u32 byte_eq_mask(u32 x, u32 pat) {
    u32 v = x ^ pat;
    return ~(((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v | 0x7F7F7F7Fu);
}
// gather the four 0x80 flags of byte_eq_mask into bits 0..3
__device__ __forceinline__ u32 gather_flags4(u32 y) { return ((y >> 7) * 0x01020408u) >> 24; }

__device__ __forceinline__ u32 warp_incl_scan(u32 v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += t; }
    return v;
}

// Block-wide exclusive scan of one u32 per thread; returns exclusive prefix, *total = block sum.
// `sh` must hold THREADS/32 + 1 words.
template <int THREADS>
__device__ __forceinline__ u32 block_excl_scan(u32 v, u32 *sh, u32 *total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    u32 inc = warp_incl_scan(v, lane);
    if (lane == 31) sh[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        u32 w = lane < THREADS / 32 ? sh[lane] : 0;
        u32 winc = warp_incl_scan(w, lane);
        if (lane < THREADS / 32) sh[lane] = winc - w;
        if (lane == 31) sh[THREADS / 32] = winc;
    }
    __syncthreads();
    u32 r = inc - v + sh[wid];
    *total = sh[THREADS / 32];
    __syncthreads();
    return r;
}

// ---- decoupled look-back over tiles processed in index order by a fully resident grid.
// Descriptor word: [63:62] status (0 none, 1 aggregate, 2 inclusive prefix), [61:0] value.
#define LB_AGG (1ull << 62)
#define LB_INC (2ull << 62)
#define LB_VAL(x) ((x) & ((1ull << 62) - 1))
__device__ __forceinline__ u64 ld_volatile_u64(const u64 *p) {
    u64 v; asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p)); return v;
}
__device__ __forceinline__ void st_volatile_u64(u64 *p, u64 v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Called by ONE warp of the CTA that owns `tile`; returns the exclusive prefix of the tile
// (sum of aggregates of tiles first_tile..tile-1) and publishes this tile's inclusive prefix.
// Every round looks at LB_WIDE * 32 predecessors with all loads in flight at once: with a persistent grid the
// nearest published inclusive prefix is typically one wave (several hundred tiles) back, and a narrow window
// would turn that distance into a chain of dependent L2 round trips.
#define LB_WIDE 1
__device__ __forceinline__ u64 lookback_exclusive(u64 *desc, int tile, int first_tile, u64 aggregate, int lane) {
    if (tile == first_tile) {
        if (lane == 0) st_volatile_u64(&desc[tile], LB_INC | aggregate);
        return 0;
    }
    if (lane == 0) st_volatile_u64(&desc[tile], LB_AGG | aggregate);
    u64 excl = 0;
    int t = tile - 1;                                    // nearest predecessor not yet accounted for
    while (true) {
        u64 d[LB_WIDE];
#pragma unroll
        for (int j = 0; j < LB_WIDE; ++j) {
            const int mine = t - lane - 32 * j;
            d[j] = mine >= first_tile ? ld_volatile_u64(&desc[mine]) : LB_INC;   // virtual tiles before the first: prefix 0
        }
        bool done = false;
#pragma unroll
        for (int j = 0; j < LB_WIDE; ++j) {
            const u32 st = (u32)(d[j] >> 62);
            if (__any_sync(0xffffffffu, st == 0)) break;                         // not published yet: poll again from here
            const u32 inc_mask = __ballot_sync(0xffffffffu, st == 2);
            u64 v = LB_VAL(d[j]);
            if (inc_mask) {
                const int first_inc = __ffs(inc_mask) - 1;                       // nearest tile with an inclusive prefix
                if (lane > first_inc) v = 0;
            }
#pragma unroll
            for (int s = 16; s; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
            excl += v;
            t -= 32;
            if (inc_mask) { done = true; break; }
        }
        if (done) break;
    }
    if (lane == 0) st_volatile_u64(&desc[tile], LB_INC | (excl + aggregate));
    return excl;
}

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }

// ---- forward byte reader over global memory using aligned 8-byte loads
struct ByteReader {
    const u64 *base;   // 8-byte aligned buffer base
    u64 cur;           // unread bytes of the current word (next byte in the low 8 bits)
    u64 pos;           // absolute byte offset of the next byte
    int left;
    __device__ __forceinline__ void init(const char *buf, u64 off) {
        base = (const u64 *)buf; pos = off;
        int sh = (int)(off & 7);
        cur = __ldg(base + (off >> 3)) >> (sh * 8);
        left = 8 - sh;
    }
    __device__ __forceinline__ int next() {
        if (left == 0) { cur = __ldg(base + (pos >> 3)); left = 8; }
        int c = (int)(cur & 0xFF);
        cur >>= 8; --left; ++pos;
        return c;
    }
};

__device__ __forceinline__ bool is_ws(int c) { return c == ' ' || (unsigned)(c - 9) <= 4u; }       // includes '\n'
__device__ __forceinline__ bool is_blank(int c) { return c == ' ' || c == '\t' || (unsigned)(c - 11) <= 2u; }  // excludes '\n'

#endif  // __CUDACC__
