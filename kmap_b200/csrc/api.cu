// Error plumbing shared by every entry point of libkmap_b200.
#include <cstdarg>
#include <cstdio>
#include "common.cuh"

static thread_local char g_err[512] = "";

void kmap_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int kmap_check_launch(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return KMAP_OK;
    kmap_set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}

extern "C" const char* kmap_last_error(void) { return g_err; }
extern "C" int kmap_version(void) { return 100; }
