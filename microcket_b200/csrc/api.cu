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
