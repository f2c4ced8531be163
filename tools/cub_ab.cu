// cub_ab.cu — same-box A/B of the library's hand-written radix sort (csrc/radix_sort.cuh, Rec16 policy) against
// cub::DeviceRadixSort on the dedup + binning sort's input: n 16-byte records sorted on their low `bits` bits (69 for hg38 at
// 5 kb), the upper bits (the input index) carried along.  SURVEY.md K7 names the CUB sort "the baseline to beat".
//   build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I microcket_b200/csrc tools/cub_ab.cu \
//               -L microcket_b200 -lmicrocket_b200 -Xlinker -rpath=$PWD/microcket_b200 -o tools/scratch/cub_ab
//   run:   tools/scratch/cub_ab [n = 97800000] [bits = 69] [reps = 5]     -> one JSON line
#include <cstdio>
#include <cstdlib>
#include <cub/cub.cuh>
#include "mk_common.cuh"
#include "radix_sort.cuh"

__global__ void k_fill(uint4 *k, u64 n, u32 bits) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        u64 x = i * 0x9E3779B97F4A7C15ull; x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
        u64 y = (i + 1) * 0xD1342543DE82EF95ull; y ^= y >> 31; y *= 0x94D049BB133111EBull; y ^= y >> 29;
        u64 lo = x, hi = bits > 64 ? (y & ((1ull << (bits - 64)) - 1)) : 0;
        if (bits < 64) lo &= (1ull << bits) - 1;
        k[i] = make_uint4((u32)lo, (u32)(lo >> 32), (u32)hi, (u32)i);    // index in the top word, never sorted on
    }
}
__global__ void k_check(const uint4 *k, u64 n, u32 bits, unsigned long long *bad) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x + 1; i < n; i += (u64)gridDim.x * blockDim.x) {
        const uint4 a = k[i - 1], b = k[i];
        const u64 alo = (u64)a.x | ((u64)a.y << 32), blo = (u64)b.x | ((u64)b.y << 32);
        const u64 m = bits > 64 ? ((1ull << (bits - 64)) - 1) : 0;
        const u64 ahi = a.z & m, bhi = b.z & m;
        if (ahi > bhi || (ahi == bhi && alo > blo) || (ahi == bhi && alo == blo && a.w > b.w)) atomicAdd(bad, 1ull);   // sorted AND stable
    }
}

int main(int argc, char **argv) {
    const u64 n = argc > 1 ? strtoull(argv[1], 0, 10) : 97800000ull;
    const u32 bits = argc > 2 ? atoi(argv[2]) : 69;
    const int reps = argc > 3 ? atoi(argv[3]) : 5;
    uint4 *src, *a, *b; unsigned long long *bad;
    cudaMalloc(&src, n * 16); cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16); cudaMalloc(&bad, 8);
    k_fill<<<148 * 8, 256>>>(src, n, bits);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms_ours = 1e30f, ms_cub = 1e30f; unsigned long long bad_ours = 0, bad_cub = 0;
    // ---- ours
    RadixWs ws; ws.alloc(n);
    RadixSchedule sch; sch.n_pass = (bits + 7) / 8; for (int i = 0; i < sch.n_pass; ++i) sch.byte_of[i] = i;
    u64 launches = 0;
    for (int r = 0; r < reps + 1; ++r) {
        cudaMemcpy(a, src, n * 16, cudaMemcpyDeviceToDevice);
        Rec16::Bufs bufs; bufs.k[0] = a; bufs.k[1] = b; bufs.v[0] = bufs.v[1] = nullptr;
        cudaEventRecord(e0);
        radix_sort<Rec16>(bufs, n, sch, ws, 0, 148, 0, &launches);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r) ms_ours = ms < ms_ours ? ms : ms_ours;
    }
    {
        RadixPlan hp; cudaMemcpy(&hp, ws.plan.p, sizeof hp, cudaMemcpyDeviceToHost);
        cudaMemset(bad, 0, 8);
        k_check<<<148 * 8, 256>>>(hp.final_buf ? b : a, n, bits, bad);
        cudaMemcpy(&bad_ours, bad, 8, cudaMemcpyDeviceToHost);
    }
    // ---- cub: 128-bit keys, bits [0, bits)
    typedef unsigned __int128 K;
    size_t tmp_bytes = 0; void *tmp = nullptr;
    cub::DoubleBuffer<K> db((K *)a, (K *)b);
    cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, db, (long long)n, 0, (int)bits);
    cudaMalloc(&tmp, tmp_bytes);
    for (int r = 0; r < reps + 1; ++r) {
        cudaMemcpy(a, src, n * 16, cudaMemcpyDeviceToDevice);
        db = cub::DoubleBuffer<K>((K *)a, (K *)b);
        cudaEventRecord(e0);
        cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, db, (long long)n, 0, (int)bits);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r) ms_cub = ms < ms_cub ? ms : ms_cub;
    }
    cudaMemset(bad, 0, 8);
    k_check<<<148 * 8, 256>>>((const uint4 *)db.Current(), n, bits, bad);
    cudaMemcpy(&bad_cub, bad, 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    printf("{\"n\": %llu, \"key_bits\": %u, \"record_bytes\": 16, \"ours_ms\": %.3f, \"cub_ms\": %.3f, \"ours_over_cub\": %.3f, "
           "\"ours_unsorted_or_unstable\": %llu, \"cub_unsorted_or_unstable\": %llu, \"cub_temp_bytes\": %zu, \"cuda\": \"%s\"}\n",
           (unsigned long long)n, bits, ms_ours, ms_cub, ms_ours / ms_cub, bad_ours, bad_cub, tmp_bytes, cudaGetErrorString(e));
    return 0;
}
