// common.h — shared host-side helpers of libvbq_b200 (error reporting, launch sizing).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "vbq_b200.h"

int vbq_fail(int code, const char *fmt, ...);
int vbq_grid_for(long long total, int block, int *grid);
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
