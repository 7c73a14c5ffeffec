// vbq_b200.cu — sm_100a kernels and C ABI of the VBQ rate-distortion quantization path.
//
// Reference behaviour reproduced here (paths relative to mandt-lab/vbq):
//   img-compression/quantizer.py:25-80,156-240   code-point tables, per-depth bracketing, candidate order
//   img-compression/utils.py:307-327,363-423     float32 score -0.5*((z-mu)/sigma)^2 - lambda*len, first argmax
//   img-compression/learned_prior.py:30-218      factorized-prior CDF and its inverse
//   img-compression/vae_models.py:14-43          Gaussian priors
// Design: the prior's quantile function tabulated on the dyadic grid is an implicit binary search
// tree in heap order (node h has children 2h+1, 2h+2).  A CTA keeps the tree of 16 channels interleaved in
// shared memory (bank = channel + 16*(h&1)) and every thread walks it once per coordinate: one compare per bit
// depth gives the bracket of mu at that depth, whose two ends are scored in registers.  Nothing on this path is
// a dense contraction, so tensor cores are not used.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "vbq_b200.h"

// ------------------------------------------------------------------------------------------------------------
// status / errors
// ------------------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(VBQ_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_));        \
    } while (0)

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

__global__ void pack_table_kernel(const float *__restrict__ table, int C, int Q, int Qs, int n_groups,
                                  float *__restrict__ packed) {
    const long long total = (long long)n_groups * Qs * VBQ_GROUP;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int j = (int)(t % VBQ_GROUP);
        const long long r = t / VBQ_GROUP;
        const int h = (int)(r % Qs);
        const int g = (int)(r / Qs);
        const int c = min(g * VBQ_GROUP + j, C - 1);
        packed[t] = table[(size_t)c * Q + h];
    }
}

// ------------------------------------------------------------------------------------------------------------
// the quantization kernel
// ------------------------------------------------------------------------------------------------------------
constexpr int kThreads = 512;
constexpr int kRowsPerPass = kThreads / VBQ_GROUP;  // 32
constexpr int kMaxGrid = 1024;

struct QArgs {
    const float *mu, *sigma;
    long long rows;
    int C;
    const float *table, *packed;
    int N, Q, S, Qs;
    const float *pen, *len;
    int n_lambda, pen_channels;
    const float *em;
    float *zhat;
    int *qidx, *level;
    float *bits, *em_bits;
    double *totals, *partials;
    unsigned *ticket;
    unsigned flags;
    int n_groups;
    long long n_tiles, total_units;
};

// a/b with a correctly rounded reciprocal r = RN(1/b): q0 = RN(a r), e = a - q0 b (exact in an FMA),
// q = RN(q0 + e r) is the IEEE quotient (Markstein); checked bit for bit by tests/test_gpu_parity.py.
__device__ __forceinline__ float div_rn(float a, float b, float r) {
    const float q0 = __fmul_rn(a, r);
    const float e = __fmaf_rn(-q0, b, a);
    return __fmaf_rn(e, r, q0);
}

// utils.py:318-320 then :393-396:  fl( fl(-0.5 * fl(t*t)) - pen ),  t = fl(fl(z-mu)/sigma).
// -0.5*t2 is exact, so one FMA reproduces the two roundings.  npen = -pen.
__device__ __forceinline__ float score_exact(float z, float mu, float sg, float rs, float npen) {
    const float t = div_rn(__fsub_rn(z, mu), sg, rs);
    return __fmaf_rn(__fmul_rn(t, t), -0.5f, npen);
}

template <int U, bool FAST>
__global__ void __launch_bounds__(kThreads, 1) vbq_quantize_kernel(const QArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int N = a.N;
    float *sT = smem;                                  // [Qs][16] code points, levels < S
    float *sPen = sT + (size_t)a.Qs * VBQ_GROUP;       // [N+1][16] negated penalties
    float *sSuf = sPen + (N + 1) * VBQ_GROUP;          // [N+1][16] max over deeper levels of the negated penalty
    float *sLen = sSuf + (N + 1) * VBQ_GROUP;          // [N+1][16] code lengths
    __shared__ double sRed[VBQ_TOTALS][kThreads / 32];
    __shared__ bool sLast;

    const int lam = blockIdx.y;
    const int col = threadIdx.x & (VBQ_GROUP - 1);
    const int rsub = threadIdx.x >> 4;
    const bool prune = !(a.flags & VBQ_FLAG_NO_PRUNE);
    const bool logvar = (a.flags & VBQ_FLAG_LOGVAR) != 0;
    const long long u0 = a.total_units * blockIdx.x / gridDim.x;
    const long long u1 = a.total_units * (blockIdx.x + 1) / gridDim.x;
    const size_t lam_off = (size_t)lam * (size_t)a.rows * (size_t)a.C;
    const int S = a.S;

    double acc_level = 0.0, acc_len = 0.0, acc_em = 0.0, acc_dist = 0.0;
    int cur_g = -1;

    for (long long unit = u0; unit < u1; ++unit) {
        const int g = (int)(unit / a.n_tiles);
        const long long tile = unit - (long long)g * a.n_tiles;
        if (g != cur_g) {
            __syncthreads();
            const float4 *src = reinterpret_cast<const float4 *>(a.packed + (size_t)g * a.Qs * VBQ_GROUP);
            float4 *dst = reinterpret_cast<float4 *>(sT);
            for (int k = threadIdx.x; k < a.Qs * (VBQ_GROUP / 4); k += kThreads) dst[k] = __ldg(src + k);
            if (threadIdx.x < VBQ_GROUP) {
                const int j = threadIdx.x;
                const int c = min(g * VBQ_GROUP + j, a.C - 1);
                const size_t po = ((size_t)lam * a.pen_channels + (a.pen_channels == 1 ? 0 : c)) * (N + 1);
                float suf = -CUDART_INF_F;
                for (int n = N; n >= 0; --n) {
                    const float np_ = -a.pen[po + n];
                    sPen[n * VBQ_GROUP + j] = np_;
                    sSuf[n * VBQ_GROUP + j] = suf;
                    suf = fmaxf(suf, np_);
                    sLen[n * VBQ_GROUP + j] = a.len ? a.len[po + n] : (float)n;
                }
            }
            __syncthreads();
            cur_g = g;
        }
        const int c = g * VBQ_GROUP + col;
        const int cc = min(c, a.C - 1);
        const float *gT = a.table + (size_t)cc * a.Q;

        float mu[U], sg[U], rs[U];
        bool valid[U];
        size_t off[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long row = tile * (kRowsPerPass * U) + u * kRowsPerPass + rsub;
            valid[u] = row < a.rows && c < a.C;
            off[u] = (size_t)row * a.C + c;
            mu[u] = valid[u] ? __ldg(a.mu + off[u]) : 0.0f;
            float s = valid[u] ? __ldg(a.sigma + off[u]) : 1.0f;
            if (logvar) s = sqrtf(expf(s));
            sg[u] = s;
            rs[u] = FAST ? 0.5f * __frcp_rn(s) * __frcp_rn(s) : __frcp_rn(s);
        }

        // depth 0: the single median point (left_0 == right_0; only left_0 is a candidate, quantizer.py:182-183)
        float bestL[U], bestR[U];
        int hL[U], hR[U], idx[U];
        {
            const float z0 = sT[col];
            const float np0 = sPen[col];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (FAST) {
                    const float d = z0 - mu[u];
                    bestL[u] = __fmaf_rn(-d * d, rs[u], np0);
                } else {
                    bestL[u] = score_exact(z0, mu[u], sg[u], rs[u], np0);
                }
                bestR[u] = -CUDART_INF_F;
                hL[u] = 0;
                hR[u] = 0;
                idx[u] = mu[u] > z0 ? 1 : 0;
            }
        }
        bool stop = false;
        if (prune) {
            bool done = true;
            const float th = sSuf[col];
#pragma unroll
            for (int u = 0; u < U; ++u) done = done && (bestL[u] > th);
            stop = __all_sync(0xffffffffu, done);
        }

        for (int n = 1; n <= N && !stop; ++n) {
            const int base = (1 << n) - 1;  // heap offset of the level == index of its last point
            const float npn = sPen[n * VBQ_GROUP + col];
            const bool in_smem = n < S;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int ip = idx[u];
                const float zp = in_smem ? sT[(base + ip) * VBQ_GROUP + col] : __ldg(gT + base + ip);
                const bool b = mu[u] > zp;
                int nb = b ? ip + 1 : ip - 1;
                nb = max(nb, 0);
                bool path_is_left = b;
                if (nb > base) {           // mu above the highest point of this depth
                    if (n < N) {
                        nb = base;         // edge padding: left = right = highest (quantizer.py:57)
                    } else {
                        nb = max(base - 1, 0);  // no padding at depth N: (second highest, highest)
                        path_is_left = false;
                    }
                }
                const float zn = in_smem ? sT[(base + nb) * VBQ_GROUP + col] : __ldg(gT + base + nb);
                const float zl = path_is_left ? zp : zn;
                const float zr = path_is_left ? zn : zp;
                const int il = path_is_left ? ip : nb;
                const int ir = path_is_left ? nb : ip;
                if (FAST) {
                    const float dl = mu[u] - zl, dr = zr - mu[u];
                    const bool use_r = dr < dl;
                    const float d = use_r ? dr : dl;
                    const float s = __fmaf_rn(-d * d, rs[u], npn);
                    if (s > bestL[u]) {
                        bestL[u] = s;
                        hL[u] = base + (use_r ? ir : il);
                    }
                } else {
                    const float sl = score_exact(zl, mu[u], sg[u], rs[u], npn);
                    const float sr = score_exact(zr, mu[u], sg[u], rs[u], npn);
                    if (sl > bestL[u]) {
                        bestL[u] = sl;
                        hL[u] = base + il;
                    }
                    if (sr > bestR[u]) {
                        bestR[u] = sr;
                        hR[u] = base + ir;
                    }
                }
                idx[u] = 2 * ip + (b ? 1 : 0);
            }
            if (prune) {
                bool done = true;
                const float th = sSuf[n * VBQ_GROUP + col];
#pragma unroll
                for (int u = 0; u < U; ++u) done = done && (fmaxf(bestL[u], bestR[u]) > th);
                stop = __all_sync(0xffffffffu, done);
            }
        }

#pragma unroll
        for (int u = 0; u < U; ++u) {
            // first maximum in candidate order left_0..left_N, right_1..right_N (utils.py:401)
            const int h = (bestR[u] > bestL[u]) ? hR[u] : hL[u];
            const int n = 31 - __clz(h + 1);
            const int i = h + 1 - (1 << n);
            const int q = (((2 * i + 1) << (N - n))) - 1;
            const float zh = h < a.Qs ? sT[h * VBQ_GROUP + col] : __ldg(gT + h);
            const float ln = sLen[n * VBQ_GROUP + col];
            float eb = 0.0f;
            if (valid[u]) {
                const size_t o = lam_off + off[u];
                if (a.em) eb = __ldg(a.em + ((size_t)lam * a.C + c) * a.Q + q);
                if (a.zhat) a.zhat[o] = zh;
                if (a.qidx) a.qidx[o] = q;
                if (a.level) a.level[o] = n;
                if (a.bits) a.bits[o] = ln;
                if (a.em_bits) a.em_bits[o] = eb;
                if (a.totals) {
                    const float r1 = FAST ? __frcp_rn(sg[u]) : rs[u];
                    const double t = (double)div_rn(__fsub_rn(zh, mu[u]), sg[u], r1);
                    acc_level += (double)n;
                    acc_len += (double)ln;
                    acc_em += (double)eb;
                    acc_dist += 0.5 * t * t;
                }
            }
        }
    }

    if (a.totals) {
        double v[VBQ_TOTALS] = {acc_level, acc_len, acc_em, acc_dist};
#pragma unroll
        for (int k = 0; k < VBQ_TOTALS; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
            if ((threadIdx.x & 31) == 0) sRed[k][threadIdx.x >> 5] = v[k];
        }
        __syncthreads();
        double *part = a.partials + ((size_t)lam * kMaxGrid + blockIdx.x) * VBQ_TOTALS;
        if (threadIdx.x < VBQ_TOTALS) {
            double s = 0.0;
            for (int w = 0; w < kThreads / 32; ++w) s += sRed[threadIdx.x][w];
            part[threadIdx.x] = s;
            __threadfence();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned t = atomicAdd(a.ticket + lam, 1u);
            sLast = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (sLast && threadIdx.x < VBQ_TOTALS) {
            __threadfence();
            const volatile double *p = a.partials + (size_t)lam * kMaxGrid * VBQ_TOTALS;
            double s = 0.0;
            for (unsigned b = 0; b < gridDim.x; ++b) s += p[b * VBQ_TOTALS + threadIdx.x];  // fixed order
            a.totals[lam * VBQ_TOTALS + threadIdx.x] = s;
            if (threadIdx.x == 0) a.ticket[lam] = 0u;
        }
    }
}

__global__ void selftest_divide_kernel(const float *__restrict__ x, const float *__restrict__ y, long long n,
                                       float *__restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride)
        out[t] = div_rn(x[t], y[t], __frcp_rn(y[t]));
}

// ------------------------------------------------------------------------------------------------------------
// host side of the C ABI
// ------------------------------------------------------------------------------------------------------------
static int grid_for(long long total, int block, int *grid) {
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long need = (total + block - 1) / block;
    long long cap = (long long)sms * 16;
    *grid = (int)(need < 1 ? 1 : (need > cap ? cap : need));
    return VBQ_OK;
}

static int check_depth(int N) {
    if (N < 0 || N > VBQ_MAX_DEPTH) return fail(VBQ_ERR_BAD_DEPTH, "max_bits_per_coord=%d outside [0,%d]", N, VBQ_MAX_DEPTH);
    return VBQ_OK;
}

#define RETURN_IF(x)            \
    do {                        \
        int s_ = (x);           \
        if (s_ != VBQ_OK) return s_; \
    } while (0)

extern "C" int vbq_learned_cdf(const float *d_params, int C, const float *d_x, long long rows, float *d_cdf,
                               void *stream) {
    if (!d_params || (rows > 0 && (!d_x || !d_cdf))) return fail(VBQ_ERR_NULL_POINTER, "vbq_learned_cdf: null pointer");
    if (C < 1 || rows < 0) return fail(VBQ_ERR_BAD_SHAPE, "vbq_learned_cdf: rows=%lld C=%d", rows, C);
    if (rows == 0) return VBQ_OK;
    int grid;
    RETURN_IF(grid_for(rows * C, 256, &grid));
    learned_cdf_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_params, C, d_x, rows * C, d_cdf);
    CUDA_TRY(cudaGetLastError());
    return VBQ_OK;
}

extern "C" int vbq_learned_inverse_cdf(const float *d_params, int C, const double *d_xi, long long rows, float *d_z,
                                       void *stream) {
    if (!d_params || (rows > 0 && (!d_xi || !d_z)))
        return fail(VBQ_ERR_NULL_POINTER, "vbq_learned_inverse_cdf: null pointer");
    if (C < 1 || rows < 0) return fail(VBQ_ERR_BAD_SHAPE, "vbq_learned_inverse_cdf: rows=%lld C=%d", rows, C);
    if (rows == 0) return VBQ_OK;
    int grid;
    RETURN_IF(grid_for(rows * C, 128, &grid));
    learned_inverse_cdf_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(d_params, C, d_xi, rows * C, d_z);
    CUDA_TRY(cudaGetLastError());
    return VBQ_OK;
}

extern "C" int vbq_gaussian_inverse_cdf(const double *d_mean, const double *d_std, int C, const double *d_xi,
                                        long long rows, double *d_z, void *stream) {
    if (rows > 0 && (!d_xi || !d_z)) return fail(VBQ_ERR_NULL_POINTER, "vbq_gaussian_inverse_cdf: null pointer");
    if (C < 1 || rows < 0) return fail(VBQ_ERR_BAD_SHAPE, "vbq_gaussian_inverse_cdf: rows=%lld C=%d", rows, C);
    if (rows == 0) return VBQ_OK;
    int grid;
    RETURN_IF(grid_for(rows * C, 256, &grid));
    gaussian_inverse_cdf_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_mean, d_std, C, d_xi, rows * C, d_z);
    CUDA_TRY(cudaGetLastError());
    return VBQ_OK;
}

extern "C" int vbq_build_code_points_learned(const float *d_params, int C, int N, float *d_table, void *stream) {
    if (!d_params || !d_table) return fail(VBQ_ERR_NULL_POINTER, "vbq_build_code_points_learned: null pointer");
    if (C < 1) return fail(VBQ_ERR_BAD_SHAPE, "vbq_build_code_points_learned: C=%d", C);
    RETURN_IF(check_depth(N));
    const int Q = (1 << (N + 1)) - 1;
    int grid;
    RETURN_IF(grid_for((long long)C * Q, 128, &grid));
    build_table_learned_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(d_params, C, Q, d_table);
    CUDA_TRY(cudaGetLastError());
    return VBQ_OK;
}

extern "C" int vbq_build_code_points_gaussian(const double *d_mean, const double *d_std, int C, int N, float *d_table,
                                              void *stream) {
    if (!d_table) return fail(VBQ_ERR_NULL_POINTER, "vbq_build_code_points_gaussian: null pointer");
    if (C < 1) return fail(VBQ_ERR_BAD_SHAPE, "vbq_build_code_points_gaussian: C=%d", C);
    RETURN_IF(check_depth(N));
    const int Q = (1 << (N + 1)) - 1;
    int grid;
    RETURN_IF(grid_for((long long)C * Q, 256, &grid));
    build_table_gaussian_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_mean, d_std, C, Q, d_table);
    CUDA_TRY(cudaGetLastError());
    return VBQ_OK;
}

static inline int smem_levels(int N) { return (N + 1 < VBQ_SMEM_LEVELS) ? N + 1 : VBQ_SMEM_LEVELS; }

extern "C" long long vbq_packed_table_floats(int C, int N) {
    if (C < 1 || N < 0 || N > VBQ_MAX_DEPTH) return -1;
    const long long groups = (C + VBQ_GROUP - 1) / VBQ_GROUP;
    const long long Qs = (1ll << smem_levels(N)) - 1;
    return groups * Qs * VBQ_GROUP;
}

extern "C" int vbq_pack_code_points(const float *d_table, int C, int N, float *d_packed, void *stream) {
    if (!d_table || !d_packed) return fail(VBQ_ERR_NULL_POINTER, "vbq_pack_code_points: null pointer");
    if (C < 1) return fail(VBQ_ERR_BAD_SHAPE, "vbq_pack_code_points: C=%d", C);
    RETURN_IF(check_depth(N));
    if (((uintptr_t)d_packed & 15) != 0) return fail(VBQ_ERR_MISALIGNED, "vbq_pack_code_points: d_packed not 16-byte aligned");
    const int Q = (1 << (N + 1)) - 1;
    const int Qs = (1 << smem_levels(N)) - 1;
    const int groups = (C + VBQ_GROUP - 1) / VBQ_GROUP;
    int grid;
    RETURN_IF(grid_for((long long)groups * Qs * VBQ_GROUP, 256, &grid));
    pack_table_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_table, C, Q, Qs, groups, d_packed);
    CUDA_TRY(cudaGetLastError());
    return VBQ_OK;
}

// workspace = [ticket counters, padded to 256 B][per-lambda, per-CTA partial totals]
static inline size_t ticket_bytes(int n_lambda) { return (((size_t)n_lambda * sizeof(unsigned)) + 255) & ~(size_t)255; }

extern "C" long long vbq_quantize_workspace_bytes(int n_lambda) {
    if (n_lambda < 1) return -1;
    return (long long)ticket_bytes(n_lambda) + (long long)n_lambda * kMaxGrid * VBQ_TOTALS * (long long)sizeof(double);
}

template <int U, bool FAST>
static int launch_quantize(const QArgs &a, int grid_x, size_t smem, cudaStream_t st) {
    CUDA_TRY(cudaFuncSetAttribute(vbq_quantize_kernel<U, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vbq_quantize_kernel<U, FAST><<<dim3(grid_x, a.n_lambda), kThreads, smem, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    return VBQ_OK;
}

extern "C" int vbq_quantize(const float *d_mu, const float *d_sigma, long long rows, int C, const float *d_table,
                            const float *d_packed, int N, const float *d_penalty, const float *d_length, int n_lambda,
                            int pen_channels, const float *d_entropy_model, float *d_zhat, int *d_qidx, int *d_level,
                            float *d_bits, float *d_em_bits, double *d_totals, void *d_workspace,
                            long long workspace_bytes, unsigned flags, void *stream) {
    if (rows < 0 || C < 1 || n_lambda < 1 || n_lambda > 65535 || (pen_channels != 1 && pen_channels != C))
        return fail(VBQ_ERR_BAD_SHAPE, "vbq_quantize: rows=%lld C=%d n_lambda=%d pen_channels=%d", rows, C, n_lambda,
                    pen_channels);
    RETURN_IF(check_depth(N));
    if (flags & ~(VBQ_FLAG_LOGVAR | VBQ_FLAG_NO_PRUNE | VBQ_FLAG_FAST))
        return fail(VBQ_ERR_BAD_FLAGS, "vbq_quantize: unknown flag bits 0x%x", flags);
    if (!d_table || !d_packed || !d_penalty || (rows > 0 && (!d_mu || !d_sigma)))
        return fail(VBQ_ERR_NULL_POINTER, "vbq_quantize: null input pointer");
    if (d_em_bits && !d_entropy_model)
        return fail(VBQ_ERR_NULL_POINTER, "vbq_quantize: d_em_bits requested without d_entropy_model");
    if (((uintptr_t)d_packed & 15) != 0) return fail(VBQ_ERR_MISALIGNED, "vbq_quantize: d_packed not 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;

    QArgs a;
    a.mu = d_mu; a.sigma = d_sigma; a.rows = rows; a.C = C;
    a.table = d_table; a.packed = d_packed;
    a.N = N; a.Q = (1 << (N + 1)) - 1; a.S = smem_levels(N); a.Qs = (1 << a.S) - 1;
    a.pen = d_penalty; a.len = d_length; a.n_lambda = n_lambda; a.pen_channels = pen_channels;
    a.em = d_entropy_model;
    a.zhat = d_zhat; a.qidx = d_qidx; a.level = d_level; a.bits = d_bits; a.em_bits = d_em_bits;
    a.totals = d_totals; a.partials = nullptr; a.ticket = nullptr;
    a.flags = flags;
    a.n_groups = (C + VBQ_GROUP - 1) / VBQ_GROUP;

    if (d_totals) {
        const long long need = vbq_quantize_workspace_bytes(n_lambda);
        if (!d_workspace || workspace_bytes < need)
            return fail(VBQ_ERR_WORKSPACE, "vbq_quantize: totals need a %lld-byte workspace (got %lld)", need,
                        workspace_bytes);
        if (((uintptr_t)d_workspace & 255) != 0)
            return fail(VBQ_ERR_MISALIGNED, "vbq_quantize: workspace not 256-byte aligned");
        a.ticket = (unsigned *)d_workspace;
        a.partials = (double *)((char *)d_workspace + ticket_bytes(n_lambda));
        CUDA_TRY(cudaMemsetAsync(a.ticket, 0, (size_t)n_lambda * sizeof(unsigned), st));
    }
    if (rows == 0) {
        if (d_totals) CUDA_TRY(cudaMemsetAsync(d_totals, 0, (size_t)n_lambda * VBQ_TOTALS * sizeof(double), st));
        return VBQ_OK;
    }

    constexpr int U = 2;
    a.n_tiles = (rows + kRowsPerPass * U - 1) / (kRowsPerPass * U);
    a.total_units = a.n_tiles * a.n_groups;
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long gx = a.total_units < sms ? a.total_units : sms;
    if (gx > kMaxGrid) gx = kMaxGrid;
    const size_t smem = ((size_t)a.Qs * VBQ_GROUP + 3 * (size_t)(N + 1) * VBQ_GROUP) * sizeof(float);
    if (flags & VBQ_FLAG_FAST) return launch_quantize<U, true>(a, (int)gx, smem, st);
    return launch_quantize<U, false>(a, (int)gx, smem, st);
}

extern "C" int vbq_selftest_divide(const float *d_a, const float *d_b, long long n, float *d_out, void *stream) {
    if (n > 0 && (!d_a || !d_b || !d_out)) return fail(VBQ_ERR_NULL_POINTER, "vbq_selftest_divide: null pointer");
    if (n < 0) return fail(VBQ_ERR_BAD_SHAPE, "vbq_selftest_divide: n=%lld", n);
    if (n == 0) return VBQ_OK;
    int grid;
    RETURN_IF(grid_for(n, 256, &grid));
    selftest_divide_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_a, d_b, n, d_out);
    CUDA_TRY(cudaGetLastError());
    return VBQ_OK;
}
