// common.h — shared host-side helpers of libvbq_b200 (error reporting, launch sizing).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <atomic>

#include "vbq_b200.h"

int vbq_fail(int code, const char *fmt, ...);
int vbq_grid_for(long long total, int block, int *grid);
int vbq_current_device(int *dev, int *sms);
int vbq_check_depth(int N);

#define CUDA_TRY(expr)                                                                            \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return vbq_fail(VBQ_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_));        \
    } while (0)

#define RETURN_IF(x)                 \
    do {                             \
        int s_ = (x);                \
        if (s_ != VBQ_OK) return s_; \
    } while (0)

// Raise a kernel's dynamic shared-memory limit to the sm_100 maximum once per (kernel, device): the attribute is
// sticky, so repeating the driver call on every launch only costs host time.
#define VBQ_MAX_SMEM_BYTES (227 * 1024)
#define VBQ_ENSURE_MAX_SMEM(kern, dev)                                                                   \
    do {                                                                                                 \
        static std::atomic<unsigned long long> done_{0};                                                 \
        const unsigned long long bit_ = 1ull << ((dev) & 63);                                            \
        if ((dev) >= 64 || !(done_.load(std::memory_order_relaxed) & bit_)) {                            \
            cudaFuncAttributes fa_;                                                                      \
            CUDA_TRY(cudaFuncGetAttributes(&fa_, kern));   /* static shared memory counts against the limit */ \
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,             \
                                          VBQ_MAX_SMEM_BYTES - (int)fa_.sharedSizeBytes));               \
            done_.fetch_or(bit_, std::memory_order_relaxed);                                             \
        }                                                                                                \
    } while (0)
