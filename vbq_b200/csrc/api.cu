// api.cu — status strings, thread-local error text and launch-sizing helpers of the C ABI.
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

#include "common.h"

static thread_local char g_err[512] = "";

int vbq_fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// Device ordinal and SM count of the current device.  The SM count never changes, so it is cached per ordinal
// (a benign, idempotent cache: the only process-wide state of the library besides the thread-local error text).
int vbq_current_device(int *dev, int *sms) {
    static std::atomic<int> cache[64];
    CUDA_TRY(cudaGetDevice(dev));
    int n = (*dev >= 0 && *dev < 64) ? cache[*dev].load(std::memory_order_relaxed) : 0;
    if (n == 0) {
        CUDA_TRY(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, *dev));
        if (*dev >= 0 && *dev < 64) cache[*dev].store(n, std::memory_order_relaxed);
    }
    *sms = n;
    return VBQ_OK;
}

int vbq_grid_for(long long total, int block, int *grid) {
    int dev = 0, sms = 0;
    RETURN_IF(vbq_current_device(&dev, &sms));
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
