// embeddings.cu — float64 search for a SHARED prior: the notebook's compress_coordinates
// (word-embeddings/compress-trained-word-embeddings.ipynb:429-443) with its exact arithmetic.
//
// The notebook minimises  (c - mu)^2 + (2 beta) sigma^2 len(c)  over all code points in float64 (code points are
// float64 `norm.ppf` values, mu/sigma float32) and takes the first minimum in heap order.  Because the squared error
// is unimodal, only the two code points bracketing mu at each bit depth can win (ipynb:482 claims, and the tests
// confirm, "exact same result"), so the kernel walks the heap-ordered table as a binary search tree and evaluates the
// 2N+1 bracket ends with the notebook's float64 roundings:
//     squared_errors      = fl((c - double(mu))^2)
//     weighted_penalties  = fl(double(fl32(fl32(2 beta) * fl32(sigma*sigma))) * len)      (NumPy: float32 product first)
//     loss                = fl(squared_errors + weighted_penalties)
// The float32 image-path kernel cannot reproduce these choices on near-ties; this one matches them bit for bit.
#include "common.h"

constexpr int kEmbThreads = 256;

__global__ void __launch_bounds__(kEmbThreads) embeddings_f64_kernel(
    const float *__restrict__ mu, const float *__restrict__ sigma, long long n, const double *__restrict__ codepoints,
    int N, const double *__restrict__ lengths, double two_beta, int pen_f32, float *__restrict__ optima,
    int *__restrict__ heap_index, int *__restrict__ level) {
    extern __shared__ double sT[];                 // heap-order table, Q doubles, then N+1 lengths
    const int Q = (1 << (N + 1)) - 1;
    double *sLen = sT + Q;
    for (int k = threadIdx.x; k < Q; k += kEmbThreads) sT[k] = codepoints[k];
    for (int k = threadIdx.x; k <= N; k += kEmbThreads) sLen[k] = lengths[k];
    __syncthreads();
    const float two_beta32 = (float)two_beta;
    const long long stride = (long long)gridDim.x * kEmbThreads;
    for (long long t = (long long)blockIdx.x * kEmbThreads + threadIdx.x; t < n; t += stride) {
        const float m32 = mu[t], s32 = sigma[t];
        const double m = (double)m32;
        // (2*beta) * stds**2: float32 under the reference's NumPy 1.17 and for Python-float beta under NumPy 2;
        // float64 when beta is a NumPy float64 scalar under NumPy >= 2 (pen_f32 == 0)
        const double pen_unit = pen_f32 ? (double)__fmul_rn(two_beta32, __fmul_rn(s32, s32))
                                        : __dmul_rn(two_beta, (double)__fmul_rn(s32, s32));
        double best = CUDART_INF;
        int best_h = 0, ip = 0;
        for (int lv = 0; lv <= N; ++lv) {
            const int base = (1 << lv) - 1, last = base;
            const double zp = sT[base + ip];
            const bool gt = m > zp;
            const int fg = ip + (gt ? 1 : 0);
            const int ir = min(fg, last), il = max(fg - 1, 0);
            const double pen = __dmul_rn(pen_unit, sLen[lv]);
            // heap order inside a level is ascending: left candidate first, strict '<' keeps the first minimum
            const double dl = __dsub_rn(sT[base + min(il, last)], m);
            const double ll = __dadd_rn(__dmul_rn(dl, dl), pen);
            if (ll < best) {
                best = ll;
                best_h = base + min(il, last);
            }
            const double dr = __dsub_rn(sT[base + ir], m);
            const double lr = __dadd_rn(__dmul_rn(dr, dr), pen);
            if (lr < best) {
                best = lr;
                best_h = base + ir;
            }
            ip = 2 * ip + (gt ? 1 : 0);
        }
        if (optima) optima[t] = (float)sT[best_h];
        if (heap_index) heap_index[t] = best_h;
        if (level) level[t] = 31 - __clz(best_h + 1);
    }
}

extern "C" int vbq_compress_coordinates_f64(const float *d_mu, const float *d_sigma, long long n,
                                            const double *d_codepoints, int N, const double *d_lengths, double beta,
                                            int pen_f32, float *d_optima, int *d_heap_index, int *d_level,
                                            void *stream) {
    if (n < 0) return vbq_fail(VBQ_ERR_BAD_SHAPE, "vbq_compress_coordinates_f64: n=%lld", n);
    RETURN_IF(vbq_check_depth(N));
    if (N > 12) return vbq_fail(VBQ_ERR_BAD_DEPTH, "vbq_compress_coordinates_f64: max_codepoint_length=%d > 12", N);
    if (!d_codepoints || !d_lengths || (n > 0 && (!d_mu || !d_sigma)))
        return vbq_fail(VBQ_ERR_NULL_POINTER, "vbq_compress_coordinates_f64: null pointer");
    if (n == 0) return VBQ_OK;
    int grid;
    RETURN_IF(vbq_grid_for(n, kEmbThreads, &grid));
    const size_t smem = ((size_t)((1 << (N + 1)) - 1) + N + 1) * sizeof(double);
    int dev = 0, sms = 0;
    RETURN_IF(vbq_current_device(&dev, &sms));
    VBQ_ENSURE_MAX_SMEM(embeddings_f64_kernel, dev);
    embeddings_f64_kernel<<<grid, kEmbThreads, smem, (cudaStream_t)stream>>>(
        d_mu, d_sigma, n, d_codepoints, N, d_lengths, 2.0 * beta, pen_f32, d_optima, d_heap_index, d_level);
    CUDA_TRY(cudaGetLastError());
    return VBQ_OK;
}
