// priors.cu — prior models and code-point tables: the learned factorized prior's CDF / inverse CDF
// (reference img-compression/learned_prior.py:30-218), Gaussian priors (vae_models.py:14-43, notebook
// ipynb:383-390) and the heap-order (C, Q) tables of ChannelwisePriorCDFQuantizer.build_code_points
// (quantizer.py:25-37).
#include "common.h"

// ------------------------------------------------------------------------------------------------------------
// learned factorized prior (learned_prior.py:70-107): logits of the CDF and their derivative
// ------------------------------------------------------------------------------------------------------------
template <typename T>
struct Prior {
    T m0[3], b0[3], f0[3];
    T m1[9], b1[3], f1[3];
    T m2[9], b2[3], f2[3];
    T m3[3], b3;
};

template <typename T>
__device__ __forceinline__ void load_prior(const float *__restrict__ p, Prior<T> &P) {
    int k = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) P.m0[i] = (T)p[k++];
#pragma unroll
    for (int i = 0; i < 3; ++i) P.b0[i] = (T)p[k++];
#pragma unroll
    for (int i = 0; i < 3; ++i) P.f0[i] = (T)p[k++];
#pragma unroll
    for (int i = 0; i < 9; ++i) P.m1[i] = (T)p[k++];
#pragma unroll
    for (int i = 0; i < 3; ++i) P.b1[i] = (T)p[k++];
#pragma unroll
    for (int i = 0; i < 3; ++i) P.f1[i] = (T)p[k++];
#pragma unroll
    for (int i = 0; i < 9; ++i) P.m2[i] = (T)p[k++];
#pragma unroll
    for (int i = 0; i < 3; ++i) P.b2[i] = (T)p[k++];
#pragma unroll
    for (int i = 0; i < 3; ++i) P.f2[i] = (T)p[k++];
#pragma unroll
    for (int i = 0; i < 3; ++i) P.m3[i] = (T)p[k++];
    P.b3 = (T)p[k];
}

__device__ __forceinline__ float tanh_t(float x) { return tanhf(x); }
__device__ __forceinline__ double tanh_t(double x) { return tanh(x); }

// returns logits; *dl receives d logits / d x when WITH_D
template <typename T, bool WITH_D>
__device__ __forceinline__ T prior_logits(const Prior<T> &P, T x, T *dl) {
    T h[3], dh[3], g[3], dg[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        T v = P.m0[i] * x + P.b0[i];
        T th = tanh_t(v);
        h[i] = v + P.f0[i] * th;
        if (WITH_D) dh[i] = P.m0[i] * ((T)1 + P.f0[i] * ((T)1 - th * th));
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        T v = P.m1[3 * i] * h[0] + P.m1[3 * i + 1] * h[1] + P.m1[3 * i + 2] * h[2] + P.b1[i];
        T th = tanh_t(v);
        g[i] = v + P.f1[i] * th;
        if (WITH_D)
            dg[i] = (P.m1[3 * i] * dh[0] + P.m1[3 * i + 1] * dh[1] + P.m1[3 * i + 2] * dh[2]) *
                    ((T)1 + P.f1[i] * ((T)1 - th * th));
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        T v = P.m2[3 * i] * g[0] + P.m2[3 * i + 1] * g[1] + P.m2[3 * i + 2] * g[2] + P.b2[i];
        T th = tanh_t(v);
        h[i] = v + P.f2[i] * th;
        if (WITH_D)
            dh[i] = (P.m2[3 * i] * dg[0] + P.m2[3 * i + 1] * dg[1] + P.m2[3 * i + 2] * dg[2]) *
                    ((T)1 + P.f2[i] * ((T)1 - th * th));
    }
    if (WITH_D) *dl = P.m3[0] * dh[0] + P.m3[1] * dh[1] + P.m3[2] * dh[2];
    return P.m3[0] * h[0] + P.m3[1] * h[1] + P.m3[2] * h[2] + P.b3;
}

// Root of logits_c(z) = logit(xi) in float64 (bracket by doubling from [-1,1], then Newton kept inside the
// bracket), rounded to float32.  A pure function of (channel parameters, xi): no warm start, no dependence on
// the calling thread, so "the code point of (c,n,i)" is well defined.
__device__ float solve_inverse_cdf(const Prior<double> &P, double xi) {
    if (!(xi > 0.0)) return xi == 0.0 ? -CUDART_INF_F : CUDART_NAN_F;
    if (!(xi < 1.0)) return xi == 1.0 ? CUDART_INF_F : CUDART_NAN_F;
    const double target = log(xi) - log1p(-xi);
    double lo = -1.0, hi = 1.0, d;
    for (int k = 0; k < 1000 && !(prior_logits<double, false>(P, lo, &d) < target); ++k) lo *= 2.0;
    for (int k = 0; k < 1000 && !(prior_logits<double, false>(P, hi, &d) > target); ++k) hi *= 2.0;
    double x = 0.5 * (lo + hi);
    for (int it = 0; it < 200; ++it) {
        double df;
        const double f = prior_logits<double, true>(P, x, &df) - target;
        if (f == 0.0) break;
        if (f < 0.0) lo = x; else hi = x;
        double xn = x - f / df;
        if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
        const double step = fabs(xn - x);
        x = xn;
        if (step <= 1e-14 * fabs(x) + 1e-300 || !(hi - lo > 0.0)) break;
    }
    return (float)x;
}

__global__ void learned_cdf_kernel(const float *__restrict__ params, int C, const float *__restrict__ x,
                                   long long total, float *__restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int c = (int)(t % C);
        Prior<float> P;
        load_prior(params + (size_t)c * VBQ_PRIOR_PARAMS, P);
        float d;
        const float lg = prior_logits<float, false>(P, x[t], &d);
        out[t] = 1.0f / (1.0f + expf(-lg));
    }
}

__global__ void learned_inverse_cdf_kernel(const float *__restrict__ params, int C, const double *__restrict__ xi,
                                           long long total, float *__restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int c = (int)(t % C);
        Prior<double> P;
        load_prior(params + (size_t)c * VBQ_PRIOR_PARAMS, P);
        out[t] = solve_inverse_cdf(P, xi[t]);
    }
}

__global__ void gaussian_inverse_cdf_kernel(const double *__restrict__ mean, const double *__restrict__ stdv, int C,
                                            const double *__restrict__ xi, long long total,
                                            double *__restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int c = (int)(t % C);
        const double m = mean ? mean[c] : 0.0;
        const double s = stdv ? stdv[c] : 1.0;
        // scipy: ndtri(q) * scale + loc, two roundings
        out[t] = __dadd_rn(__dmul_rn(normcdfinv(xi[t]), s), m);
    }
}

// heap entry h -> xi = (i + 1/2) 2^-n, exact in float64 (utils.py:23-24)
__device__ __forceinline__ double heap_xi(int h) {
    const int n = 31 - __clz(h + 1);
    const int i = h + 1 - (1 << n);
    return ((double)i + 0.5) * exp2((double)-n);
}

__global__ void build_table_learned_kernel(const float *__restrict__ params, int C, int Q,
                                           float *__restrict__ table) {
    const long long total = (long long)C * Q;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int c = (int)(t / Q);
        const int h = (int)(t - (long long)c * Q);
        Prior<double> P;
        load_prior(params + (size_t)c * VBQ_PRIOR_PARAMS, P);
        table[t] = solve_inverse_cdf(P, heap_xi(h));
    }
}

__global__ void build_table_gaussian_kernel(const double *__restrict__ mean, const double *__restrict__ stdv, int C,
                                            int Q, float *__restrict__ table) {
    const long long total = (long long)C * Q;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int c = (int)(t / Q);
        const int h = (int)(t - (long long)c * Q);
        const double m = mean ? mean[c] : 0.0;
        const double s = stdv ? stdv[c] : 1.0;
        table[t] = (float)__dadd_rn(__dmul_rn(normcdfinv(heap_xi(h)), s), m);  // cast: quantizer.py:34
    }
}

// ------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------
extern "C" int vbq_learned_cdf(const float *d_params, int C, const float *d_x, long long rows, float *d_cdf,
                               void *stream) {
    if (!d_params || (rows > 0 && (!d_x || !d_cdf))) return vbq_fail(VBQ_ERR_NULL_POINTER, "vbq_learned_cdf: null pointer");
    if (C < 1 || rows < 0) return vbq_fail(VBQ_ERR_BAD_SHAPE, "vbq_learned_cdf: rows=%lld C=%d", rows, C);
    if (rows == 0) return VBQ_OK;
    int grid;
    RETURN_IF(vbq_grid_for(rows * C, 256, &grid));
    learned_cdf_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_params, C, d_x, rows * C, d_cdf);
    CUDA_TRY(cudaGetLastError());
    return VBQ_OK;
}

extern "C" int vbq_learned_inverse_cdf(const float *d_params, int C, const double *d_xi, long long rows, float *d_z,
                                       void *stream) {
    if (!d_params || (rows > 0 && (!d_xi || !d_z)))
        return vbq_fail(VBQ_ERR_NULL_POINTER, "vbq_learned_inverse_cdf: null pointer");
    if (C < 1 || rows < 0) return vbq_fail(VBQ_ERR_BAD_SHAPE, "vbq_learned_inverse_cdf: rows=%lld C=%d", rows, C);
    if (rows == 0) return VBQ_OK;
    int grid;
    RETURN_IF(vbq_grid_for(rows * C, 128, &grid));
    learned_inverse_cdf_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(d_params, C, d_xi, rows * C, d_z);
    CUDA_TRY(cudaGetLastError());
    return VBQ_OK;
}

extern "C" int vbq_gaussian_inverse_cdf(const double *d_mean, const double *d_std, int C, const double *d_xi,
                                        long long rows, double *d_z, void *stream) {
    if (rows > 0 && (!d_xi || !d_z)) return vbq_fail(VBQ_ERR_NULL_POINTER, "vbq_gaussian_inverse_cdf: null pointer");
    if (C < 1 || rows < 0) return vbq_fail(VBQ_ERR_BAD_SHAPE, "vbq_gaussian_inverse_cdf: rows=%lld C=%d", rows, C);
    if (rows == 0) return VBQ_OK;
    int grid;
    RETURN_IF(vbq_grid_for(rows * C, 256, &grid));
    gaussian_inverse_cdf_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_mean, d_std, C, d_xi, rows * C, d_z);
    CUDA_TRY(cudaGetLastError());
    return VBQ_OK;
}

extern "C" int vbq_build_code_points_learned(const float *d_params, int C, int N, float *d_table, void *stream) {
    if (!d_params || !d_table) return vbq_fail(VBQ_ERR_NULL_POINTER, "vbq_build_code_points_learned: null pointer");
    if (C < 1) return vbq_fail(VBQ_ERR_BAD_SHAPE, "vbq_build_code_points_learned: C=%d", C);
    RETURN_IF(vbq_check_depth(N));
    const int Q = (1 << (N + 1)) - 1;
    int grid;
    RETURN_IF(vbq_grid_for((long long)C * Q, 128, &grid));
    build_table_learned_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(d_params, C, Q, d_table);
    CUDA_TRY(cudaGetLastError());
    return VBQ_OK;
}

extern "C" int vbq_build_code_points_gaussian(const double *d_mean, const double *d_std, int C, int N, float *d_table,
                                              void *stream) {
    if (!d_table) return vbq_fail(VBQ_ERR_NULL_POINTER, "vbq_build_code_points_gaussian: null pointer");
    if (C < 1) return vbq_fail(VBQ_ERR_BAD_SHAPE, "vbq_build_code_points_gaussian: C=%d", C);
    RETURN_IF(vbq_check_depth(N));
    const int Q = (1 << (N + 1)) - 1;
    int grid;
    RETURN_IF(vbq_grid_for((long long)C * Q, 256, &grid));
    build_table_gaussian_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_mean, d_std, C, Q, d_table);
    CUDA_TRY(cudaGetLastError());
    return VBQ_OK;
}

