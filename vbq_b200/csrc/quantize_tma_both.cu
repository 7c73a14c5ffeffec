// quantize_tma_both.cu — instantiations of vbq_bisect_tma_kernel (quantize_tma.cuh) for ARBITRARY non-negative penalties:
// both ends of the reference's bracket stay candidates at every bit depth.  This is the mode every reference call runs in
// after build_entropy_models (corrected code lengths n + R_lambda[c, n], quantizer.py:166-180; utils.py:392-396), with
// the entropy-model gather of compress_latents (quantizer.py:223-228) fused in.
#include "quantize_tma.cuh"

constexpr int kBothWarps = 17, kBothPairs = 2;   // measured (Kodak batch, one lambda, no entropy-model gather): 15x2 73, 17x2 71,
                                                 // 19x2 80, 23x1 74 us

template <bool EM, bool TOTALS, int OUT>
static int launch_both(const QArgs &a, const void *out0, const void *out1, int dev, int sms, cudaStream_t st) {
#ifdef VBQ_DEV_VARIANTS
    const char *v = getenv("VBQ_TMA_VARIANT");
    const int vi = v ? atoi(v) : 0;
    if (vi == 1) return launch_tma<true, EM, false, TOTALS, kSmemDepth, OUT, 19, 2>(a, out0, out1, dev, sms, st);
    if (vi == 2) return launch_tma<true, EM, false, TOTALS, kSmemDepth, OUT, 23, 1>(a, out0, out1, dev, sms, st);
    if (vi == 3) return launch_tma<true, EM, false, TOTALS, kSmemDepth, OUT, 17, 2>(a, out0, out1, dev, sms, st);
#endif
    return launch_tma<true, EM, false, TOTALS, kSmemDepth, OUT, kBothWarps, kBothPairs>(a, out0, out1, dev, sms, st);
}

// max_bits_per_coord == 10; penalties available on the host (vbq_quantize_hp), finite and non-negative; C % 4 == 0 and
// 16-byte aligned arrays (TMA); outputs {z_hat, code length} (+ entropy-model bits), {sorted index} or totals only.
// Returns -1 if not applicable (the caller falls back to the bracket-walk kernels).
int vbq_launch_quantize_tma_both(const QArgs &a, int dev, int sms, cudaStream_t st) {
    if (a.N != kSmemDepth || a.C % 4 != 0 || !a.h_pen) return -1;
    uintptr_t al = (uintptr_t)a.mu | (uintptr_t)a.sigma | (uintptr_t)a.zhat | (uintptr_t)a.qidx | (uintptr_t)a.level |
                   (uintptr_t)a.bits | (uintptr_t)a.packed;
    if (al & 15) return -1;
    if (a.rows * (long long)a.C >= (1ll << 31)) return -1;
    if (a.em_bits && !a.em) return -1;
    const size_t n_pen = (size_t)a.n_lambda * a.pen_channels * (a.N + 1);
    for (size_t i = 0; i < n_pen; ++i)   // keys are the bit patterns of non-negative floats
        if (!(a.h_pen[i] >= 0.0f && a.h_pen[i] < 3.0e38f)) return -1;
    const unsigned outs = a.outm & 15u;
    const bool em = a.em != nullptr;
    if (em && outs != (1u | 8u) && outs != 0u) return -1;
    if (a.em_bits && outs != (1u | 8u)) return -1;
#ifdef VBQ_DEV_ONE
    if (outs == (1u | 8u) && em && a.totals) return launch_both<true, true, 1 | 8>(a, a.zhat, a.bits, dev, sms, st);
    return -1;
#else
    if (outs == (1u | 8u)) {
        if (em) return a.totals ? launch_both<true, true, 1 | 8>(a, a.zhat, a.bits, dev, sms, st)
                                : launch_both<true, false, 1 | 8>(a, a.zhat, a.bits, dev, sms, st);
        return a.totals ? launch_both<false, true, 1 | 8>(a, a.zhat, a.bits, dev, sms, st)
                        : launch_both<false, false, 1 | 8>(a, a.zhat, a.bits, dev, sms, st);
    }
    if (outs == 2u)
        return a.totals ? launch_both<false, true, 2>(a, a.qidx, nullptr, dev, sms, st)
                        : launch_both<false, false, 2>(a, a.qidx, nullptr, dev, sms, st);
    if (outs == 0u && a.totals)
        return em ? launch_both<true, true, 0>(a, nullptr, nullptr, dev, sms, st)
                  : launch_both<false, true, 0>(a, nullptr, nullptr, dev, sms, st);
    return -1;
#endif
}
