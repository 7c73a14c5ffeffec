// api.cu — status strings, thread-local error text and launch-sizing helpers of the C ABI.
#include <stdarg.h>
#include <stdio.h>

#include "common.h"

static thread_local char g_err[512] = "";

int vbq_fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int vbq_grid_for(long long total, int block, int *grid) {
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long need = (total + block - 1) / block;
    long long cap = (long long)sms * 16;
    *grid = (int)(need < 1 ? 1 : (need > cap ? cap : need));
    return VBQ_OK;
}

int vbq_check_depth(int N) {
    if (N < 0 || N > VBQ_MAX_DEPTH)
        return vbq_fail(VBQ_ERR_BAD_DEPTH, "max_bits_per_coord=%d outside [0,%d]", N, VBQ_MAX_DEPTH);
    return VBQ_OK;
}

extern "C" int vbq_version(void) { return VBQ_VERSION; }

extern "C" const char *vbq_status_string(int s) {
    switch (s) {
        case VBQ_OK: return "ok";
        case VBQ_ERR_NULL_POINTER: return "null pointer";
        case VBQ_ERR_BAD_SHAPE: return "bad shape";
        case VBQ_ERR_BAD_DEPTH: return "bad max_bits_per_coord";
        case VBQ_ERR_BAD_FLAGS: return "bad flags";
        case VBQ_ERR_WORKSPACE: return "workspace missing or too small";
        case VBQ_ERR_CUDA: return "CUDA error";
        case VBQ_ERR_MISALIGNED: return "misaligned pointer";
        default: return "unknown status";
    }
}

extern "C" const char *vbq_last_error(void) { return g_err; }
