// api.cu — error text, device queries, context destruction.
#include <cstdarg>
#include <cstdio>
#include "ctx.h"
#include "mk_common.cuh"

static thread_local char g_err[1024] = "";

void mk_set_error(const char *fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

extern "C" const char *mk_last_error(void) { return g_err; }
extern "C" int mk_version(void) { return 100; }

extern "C" int mk_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int mk_sm_count(int device) {
    int n = 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) n = 148;
    return n;
}

extern "C" void mk_destroy(mk_ctx *c) { delete c; }

// device-to-device copy between raw pointers (buffers owned by this library have no tensor on the host side)
extern "C" int mk_copy_device(void *d_dst, const void *d_src, size_t nbytes) {
    if (nbytes && (!d_dst || !d_src)) { mk_set_error("mk_copy_device: null pointer"); return MK_ERR_ARG; }
    MK_CUDA(cudaMemcpy(d_dst, d_src, nbytes, cudaMemcpyDeviceToDevice));
    return MK_OK;
}

// ---- raw device / pinned memory for host programs that are compiled without the CUDA toolkit (the drop-in CLIs)
extern "C" int mk_dev_alloc(int device, size_t nbytes, void **d_ptr) {
    if (!d_ptr) { mk_set_error("mk_dev_alloc: null"); return MK_ERR_ARG; }
    MK_CUDA(cudaSetDevice(device));
    cudaError_t e = cudaMalloc(d_ptr, nbytes ? nbytes : 16);
    if (e != cudaSuccess) { *d_ptr = nullptr; mk_set_error("cudaMalloc(%zu): %s", nbytes, cudaGetErrorString(e)); return MK_ERR_NOMEM; }
    return MK_OK;
}
extern "C" void mk_dev_free(void *d_ptr) { if (d_ptr) cudaFree(d_ptr); }
extern "C" int mk_host_alloc(size_t nbytes, void **h_ptr) {             // pinned
    if (!h_ptr) { mk_set_error("mk_host_alloc: null"); return MK_ERR_ARG; }
    cudaError_t e = cudaMallocHost(h_ptr, nbytes ? nbytes : 16);
    if (e != cudaSuccess) { *h_ptr = nullptr; mk_set_error("cudaMallocHost(%zu): %s", nbytes, cudaGetErrorString(e)); return MK_ERR_NOMEM; }
    return MK_OK;
}
extern "C" void mk_host_free(void *h_ptr) { if (h_ptr) cudaFreeHost(h_ptr); }
extern "C" int mk_copy_to_device(void *d_dst, const void *h_src, size_t nbytes) {
    MK_CUDA(cudaMemcpy(d_dst, h_src, nbytes, cudaMemcpyHostToDevice));
    return MK_OK;
}
extern "C" int mk_copy_to_host(void *h_dst, const void *d_src, size_t nbytes) {
    MK_CUDA(cudaMemcpy(h_dst, d_src, nbytes, cudaMemcpyDeviceToHost));
    return MK_OK;
}
