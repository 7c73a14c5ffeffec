// host_pipeline.cu — the hot path for HOST-resident latents: a chunked, triple-buffered
// H2D -> vbq_quantize -> D2H pipeline on three streams, so that the PCIe upload of chunk k+1, the kernel of
// chunk k and the download of chunk k-1 overlap.  This is what a drop-in for the reference's host-side call
// (ChannelwisePriorCDFQuantizer.compress_batch_channel_latents on NumPy arrays, quantizer.py:156-188) runs.
#include <new>

#include "common.h"

namespace {
constexpr int kSlots = 3;

struct Slot {
    float *mu = nullptr, *sigma = nullptr;
    float *zhat = nullptr, *bits = nullptr, *em_bits = nullptr;
    int *qidx = nullptr, *level = nullptr;
    cudaEvent_t in = nullptr, done = nullptr, out = nullptr;
};
}  // namespace

struct vbq_host_ctx {
    int C = 0, N = 0, n_lambda = 0;
    long long chunk_rows = 0;
    unsigned outputs = 0;
    cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    Slot slot[kSlots];
    double *d_totals = nullptr;
    void *d_ws = nullptr;
    long long ws_bytes = 0;
};

static void free_ctx(vbq_host_ctx *c) {
    if (!c) return;
    for (auto &s : c->slot) {
        cudaFree(s.mu); cudaFree(s.sigma); cudaFree(s.zhat); cudaFree(s.bits); cudaFree(s.em_bits);
        cudaFree(s.qidx); cudaFree(s.level);
        if (s.in) cudaEventDestroy(s.in);
        if (s.done) cudaEventDestroy(s.done);
        if (s.out) cudaEventDestroy(s.out);
    }
    cudaFree(c->d_totals);
    cudaFree(c->d_ws);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_k) cudaStreamDestroy(c->s_k);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    delete c;
}

#define CTX_TRY(expr)                                                                         \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            free_ctx(c);                                                                      \
            return vbq_fail(VBQ_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_));    \
        }                                                                                     \
    } while (0)

extern "C" int vbq_host_ctx_create(int C, int N, int n_lambda, long long chunk_rows, unsigned outputs,
                                   vbq_host_ctx **out) {
    if (!out) return vbq_fail(VBQ_ERR_NULL_POINTER, "vbq_host_ctx_create: null out pointer");
    *out = nullptr;
    if (C < 1 || n_lambda < 1 || chunk_rows < 1 || chunk_rows * (long long)C >= (1ll << 31))
        return vbq_fail(VBQ_ERR_BAD_SHAPE, "vbq_host_ctx_create: C=%d n_lambda=%d chunk_rows=%lld", C, n_lambda,
                        chunk_rows);
    RETURN_IF(vbq_check_depth(N));
    if (outputs & ~(VBQ_OUT_ZHAT | VBQ_OUT_QIDX | VBQ_OUT_LEVEL | VBQ_OUT_BITS | VBQ_OUT_EM_BITS | VBQ_OUT_TOTALS))
        return vbq_fail(VBQ_ERR_BAD_FLAGS, "vbq_host_ctx_create: unknown output bits 0x%x", outputs);
    vbq_host_ctx *c = new (std::nothrow) vbq_host_ctx;
    if (!c) return vbq_fail(VBQ_ERR_CUDA, "vbq_host_ctx_create: out of host memory");
    c->C = C; c->N = N; c->n_lambda = n_lambda; c->chunk_rows = chunk_rows; c->outputs = outputs;
    const size_t in_bytes = (size_t)chunk_rows * C * sizeof(float);
    const size_t out_bytes = in_bytes * n_lambda;
    CTX_TRY(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
    CTX_TRY(cudaStreamCreateWithFlags(&c->s_k, cudaStreamNonBlocking));
    CTX_TRY(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
    for (auto &s : c->slot) {
        CTX_TRY(cudaMalloc(&s.mu, in_bytes));
        CTX_TRY(cudaMalloc(&s.sigma, in_bytes));
        if (outputs & VBQ_OUT_ZHAT) CTX_TRY(cudaMalloc(&s.zhat, out_bytes));
        if (outputs & VBQ_OUT_QIDX) CTX_TRY(cudaMalloc(&s.qidx, out_bytes));
        if (outputs & VBQ_OUT_LEVEL) CTX_TRY(cudaMalloc(&s.level, out_bytes));
        if (outputs & VBQ_OUT_BITS) CTX_TRY(cudaMalloc(&s.bits, out_bytes));
        if (outputs & VBQ_OUT_EM_BITS) CTX_TRY(cudaMalloc(&s.em_bits, out_bytes));
        CTX_TRY(cudaEventCreateWithFlags(&s.in, cudaEventDisableTiming));
        CTX_TRY(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        CTX_TRY(cudaEventCreateWithFlags(&s.out, cudaEventDisableTiming));
    }
    if (outputs & VBQ_OUT_TOTALS) {
        c->ws_bytes = vbq_quantize_workspace_bytes(n_lambda);
        CTX_TRY(cudaMalloc(&c->d_ws, (size_t)c->ws_bytes));
        CTX_TRY(cudaMemset(c->d_ws, 0, (size_t)c->ws_bytes));   // every call leaves the ticket counters zero again
        CTX_TRY(cudaMalloc(&c->d_totals, (size_t)n_lambda * VBQ_TOTALS * sizeof(double)));
    }
    *out = c;
    return VBQ_OK;
}

extern "C" int vbq_host_ctx_destroy(vbq_host_ctx *c) {
    free_ctx(c);
    return VBQ_OK;
}

extern "C" int vbq_quantize_host(vbq_host_ctx *c, const float *h_mu, const float *h_sigma, long long rows,
                                 const float *d_table, const float *d_packed, const float *d_penalty,
                                 const float *d_length, int pen_channels, const float *d_entropy_model, float *h_zhat,
                                 int *h_qidx, int *h_level, float *h_bits, float *h_em_bits, double *h_totals,
                                 unsigned flags) {
    if (!c) return vbq_fail(VBQ_ERR_NULL_POINTER, "vbq_quantize_host: null context");
    if (rows < 0) return vbq_fail(VBQ_ERR_BAD_SHAPE, "vbq_quantize_host: rows=%lld", rows);
    if (rows > 0 && (!h_mu || !h_sigma)) return vbq_fail(VBQ_ERR_NULL_POINTER, "vbq_quantize_host: null input");
    if ((h_zhat && !(c->outputs & VBQ_OUT_ZHAT)) || (h_qidx && !(c->outputs & VBQ_OUT_QIDX)) ||
        (h_level && !(c->outputs & VBQ_OUT_LEVEL)) || (h_bits && !(c->outputs & VBQ_OUT_BITS)) ||
        (h_em_bits && !(c->outputs & VBQ_OUT_EM_BITS)) || (h_totals && !(c->outputs & VBQ_OUT_TOTALS)))
        return vbq_fail(VBQ_ERR_BAD_FLAGS, "vbq_quantize_host: an output was requested that the context was not created for");
    if (flags & VBQ_FLAG_ACCUMULATE_TOTALS)
        return vbq_fail(VBQ_ERR_BAD_FLAGS, "vbq_quantize_host: VBQ_FLAG_ACCUMULATE_TOTALS is managed internally");
    const int C = c->C, L = c->n_lambda;
    const size_t row_bytes = (size_t)C * sizeof(float);
    // Channel-independent penalties also travel as launch constants (vbq_quantize_hp): fetch the few bytes once.  The
    // tables are the caller's: they must be complete when this (synchronous) call is made, see include/vbq_b200.h.
    float h_pen[64];
    const float *pen_host = nullptr;
    if (pen_channels == 1 && (size_t)L * (c->N + 1) <= 64 && d_penalty && rows > 0) {
        CUDA_TRY(cudaMemcpyAsync(h_pen, d_penalty, (size_t)L * (c->N + 1) * sizeof(float), cudaMemcpyDeviceToHost, c->s_k));
        CUDA_TRY(cudaStreamSynchronize(c->s_k));
        pen_host = h_pen;
    }
    long long k = 0;
    for (long long r0 = 0; r0 < rows; r0 += c->chunk_rows, ++k) {
        const long long nr = rows - r0 < c->chunk_rows ? rows - r0 : c->chunk_rows;
        Slot &s = c->slot[k % kSlots];
        // upload: the slot is free once the download of the chunk that last used it has finished
        if (k >= kSlots) CUDA_TRY(cudaStreamWaitEvent(c->s_in, s.out, 0));
        CUDA_TRY(cudaMemcpyAsync(s.mu, h_mu + (size_t)r0 * C, nr * row_bytes, cudaMemcpyHostToDevice, c->s_in));
        CUDA_TRY(cudaMemcpyAsync(s.sigma, h_sigma + (size_t)r0 * C, nr * row_bytes, cudaMemcpyHostToDevice, c->s_in));
        CUDA_TRY(cudaEventRecord(s.in, c->s_in));
        // kernel
        CUDA_TRY(cudaStreamWaitEvent(c->s_k, s.in, 0));
        if (k >= kSlots) CUDA_TRY(cudaStreamWaitEvent(c->s_k, s.out, 0));
        RETURN_IF(vbq_quantize_hp(s.mu, s.sigma, nr, C, d_table, d_packed, c->N, d_penalty, pen_host, d_length, L, pen_channels,
                               d_entropy_model, h_zhat ? s.zhat : nullptr, h_qidx ? s.qidx : nullptr,
                               h_level ? s.level : nullptr, h_bits ? s.bits : nullptr,
                               h_em_bits ? s.em_bits : nullptr, h_totals ? c->d_totals : nullptr, c->d_ws, c->ws_bytes,
                               flags | VBQ_FLAG_WORKSPACE_ZEROED | (h_totals && k > 0 ? VBQ_FLAG_ACCUMULATE_TOTALS : 0u),
                               c->s_k));
        CUDA_TRY(cudaEventRecord(s.done, c->s_k));
        // download: (n_lambda, nr, C) device block -> rows [r0, r0+nr) of each lambda plane of the host array
        CUDA_TRY(cudaStreamWaitEvent(c->s_out, s.done, 0));
        const size_t w = nr * row_bytes, hp = (size_t)rows * row_bytes;
        const size_t max_pitch = 0x7fffffffull;   // cudaDeviceProp::memPitch of every current device
#define D2H(hp_, dp_)                                                                                         \
    if (hp_) {                                                                                                \
        if (L == 1)                                                                                           \
            CUDA_TRY(cudaMemcpyAsync((char *)(hp_) + (size_t)r0 * row_bytes, dp_, w, cudaMemcpyDeviceToHost,  \
                                     c->s_out));                                                              \
        else if (hp <= max_pitch)                                                                             \
            CUDA_TRY(cudaMemcpy2DAsync((char *)(hp_) + (size_t)r0 * row_bytes, hp, dp_, w, w, (size_t)L,      \
                                       cudaMemcpyDeviceToHost, c->s_out));                                    \
        else /* lambda planes further apart than the largest 2-D pitch: one copy per plane */                \
            for (int l_ = 0; l_ < L; ++l_)                                                                    \
                CUDA_TRY(cudaMemcpyAsync((char *)(hp_) + (size_t)l_ * hp + (size_t)r0 * row_bytes,            \
                                         (const char *)(dp_) + (size_t)l_ * w, w, cudaMemcpyDeviceToHost, c->s_out)); \
    }
        D2H(h_zhat, s.zhat);
        D2H(h_qidx, s.qidx);
        D2H(h_level, s.level);
        D2H(h_bits, s.bits);
        D2H(h_em_bits, s.em_bits);
#undef D2H
        CUDA_TRY(cudaEventRecord(s.out, c->s_out));
    }
    if (h_totals) {
        if (rows == 0)
            for (int i = 0; i < L * VBQ_TOTALS; ++i) h_totals[i] = 0.0;
        else
            CUDA_TRY(cudaMemcpyAsync(h_totals, c->d_totals, (size_t)L * VBQ_TOTALS * sizeof(double),
                                     cudaMemcpyDeviceToHost, c->s_k));
    }
    CUDA_TRY(cudaStreamSynchronize(c->s_in));
    CUDA_TRY(cudaStreamSynchronize(c->s_k));
    CUDA_TRY(cudaStreamSynchronize(c->s_out));
    return VBQ_OK;
}
