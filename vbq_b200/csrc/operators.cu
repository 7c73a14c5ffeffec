// operators.cu — the two stand-alone operator forms of the reference's hot path that callers can use directly:
//   * ChannelwisePriorCDFQuantizer.get_all_N_bit_intervals (quantizer.py:65-80): the bracket of mu at every bit depth
//   * utils.batch_quantize_indep_dims (utils.py:363-423): argmax of fun(P) - lambda*L over explicit candidates
// They are not on the critical path of vbq_quantize (which fuses both); they exist so that the reference's public
// surface is complete and so that tests can compare intermediate results with the oracle.
#include "common.h"

// ---- brackets -----------------------------------------------------------------------------------------------
__global__ void intervals_kernel(const float *__restrict__ mu, long long rows, int C, const float *__restrict__ table,
                                 int N, int Q, float *__restrict__ left, float *__restrict__ right) {
    const long long total = rows * C;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        // thread order (c, b): writes to the (C, N+1, B) outputs are coalesced along b
        const int c = (int)(t / rows);
        const long long b = t - (long long)c * rows;
        const float m = mu[b * C + c];
        const float *T = table + (size_t)c * Q;
        int ip = 0;
        for (int n = 0; n <= N; ++n) {
            const int base = (1 << n) - 1, last = base;
            const float zp = T[base + ip];
            const bool gt = m > zp;
            const int fg = ip + (gt ? 1 : 0);                  // first index with point >= mu  (searchsorted 'left')
            const int ir = min(fg, last);
            int il;
            if (fg == 0) il = 0;
            else if (fg > last) il = n < N ? last : max(last - 1, 0);   // edge padding except at depth N (quantizer.py:57)
            else il = fg - 1;
            if (n == 0) il = 0;
            const size_t o = ((size_t)c * (N + 1) + n) * rows + b;
            left[o] = T[base + il];
            right[o] = T[base + ir];
            ip = 2 * ip + (gt ? 1 : 0);
        }
    }
}

extern "C" int vbq_intervals(const float *d_mu, long long rows, int C, const float *d_table, int N, float *d_left,
                             float *d_right, void *stream) {
    if (rows < 0 || C < 1) return vbq_fail(VBQ_ERR_BAD_SHAPE, "vbq_intervals: rows=%lld C=%d", rows, C);
    RETURN_IF(vbq_check_depth(N));
    if (!d_table || (rows > 0 && (!d_mu || !d_left || !d_right)))
        return vbq_fail(VBQ_ERR_NULL_POINTER, "vbq_intervals: null pointer");
    if (rows == 0) return VBQ_OK;
    int grid;
    RETURN_IF(vbq_grid_for(rows * C, 256, &grid));
    intervals_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_mu, rows, C, d_table, N, (1 << (N + 1)) - 1, d_left,
                                                            d_right);
    CUDA_TRY(cudaGetLastError());
    return VBQ_OK;
}

// ---- generic candidate operator -------------------------------------------------------------------------------
// scores[m] = fun_P[m] - fl(lambda * L[m]); first maximum over m (utils.py:392-415).  fun_P is either given
// (arbitrary `fun`) or the float32 Gaussian log-density of curry_normal_logpdf(ignore_const=True) (utils.py:318-320),
// computed with IEEE division.
template <typename LT>
__global__ void argmax_candidates_kernel(const float *__restrict__ P, const LT *__restrict__ L, int l_per_lambda,
                                         const float *__restrict__ funP, const float *__restrict__ loc,
                                         const float *__restrict__ scale, const float *__restrict__ lambs,
                                         int n_lambda, int M, long long BK, float *__restrict__ zhat,
                                         LT *__restrict__ bits, int *__restrict__ index) {
    const long long total = BK * n_lambda;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int lam = (int)(t / BK);
        const long long e = t - (long long)lam * BK;
        const float lm = lambs[lam];
        const LT *Ll = L + (l_per_lambda ? (size_t)lam * M * BK : 0);
        float best = 0.0f;
        int k = 0;
        for (int m = 0; m < M; ++m) {
            const size_t o = (size_t)m * BK + e;
            float f;
            if (funP) {
                f = funP[o];
            } else {
                const float q = __fdiv_rn(__fsub_rn(P[o], loc[e]), scale[e]);
                f = __fmul_rn(-0.5f, __fmul_rn(q, q));
            }
            const float s = __fsub_rn(f, __fmul_rn(lm, (float)Ll[o]));
            // np.argmax / tf.argmax semantics: first maximum, NaN counts as the maximum
            if (m == 0 || (s > best && !(best != best)) || (s != s && !(best != best))) {
                best = s;
                k = m;
            }
        }
        const size_t o = (size_t)k * BK + e;
        zhat[t] = P[o];
        bits[t] = Ll[o];
        if (index) index[t] = k;
    }
}

template <typename LT>
static int launch_argmax(const float *P, const LT *L, int l_per_lambda, const float *funP, const float *loc,
                         const float *scale, const float *lambs, int n_lambda, int M, long long BK, float *zhat,
                         LT *bits, int *index, cudaStream_t st) {
    int grid;
    RETURN_IF(vbq_grid_for(BK * n_lambda, 256, &grid));
    argmax_candidates_kernel<LT><<<grid, 256, 0, st>>>(P, L, l_per_lambda, funP, loc, scale, lambs, n_lambda, M, BK,
                                                      zhat, bits, index);
    CUDA_TRY(cudaGetLastError());
    return VBQ_OK;
}

extern "C" int vbq_argmax_candidates(const float *d_P, const void *d_L, int l_is_float, int l_per_lambda,
                                     const float *d_funP, const float *d_loc, const float *d_scale,
                                     const float *d_lambs, int n_lambda, int M, long long BK, float *d_zhat,
                                     void *d_bits, int *d_index, void *stream) {
    if (n_lambda < 1 || M < 1 || BK < 0)
        return vbq_fail(VBQ_ERR_BAD_SHAPE, "vbq_argmax_candidates: n_lambda=%d M=%d BK=%lld", n_lambda, M, BK);
    if (!d_lambs || (BK > 0 && (!d_P || !d_L || !d_zhat || !d_bits || (!d_funP && (!d_loc || !d_scale)))))
        return vbq_fail(VBQ_ERR_NULL_POINTER, "vbq_argmax_candidates: null pointer");
    if (BK == 0) return VBQ_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (l_is_float)
        return launch_argmax<float>(d_P, (const float *)d_L, l_per_lambda, d_funP, d_loc, d_scale, d_lambs, n_lambda, M,
                                    BK, d_zhat, (float *)d_bits, d_index, st);
    return launch_argmax<int>(d_P, (const int *)d_L, l_per_lambda, d_funP, d_loc, d_scale, d_lambs, n_lambda, M, BK,
                              d_zhat, (int *)d_bits, d_index, st);
}
