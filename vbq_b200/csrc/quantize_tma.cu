// quantize_tma.cu — host side of vbq_bisect_tma_kernel (quantize_tma.cuh): tensor maps and the instantiations for raw code
// lengths (channel-independent, non-decreasing penalties: one candidate per bit depth).  The instantiations for arbitrary
// penalties (both bracket ends) live in quantize_tma_both.cu.
#include "quantize_tma.cuh"

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess ||
            qr != cudaDriverEntryPointSuccess)
            return nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

int vbq_make_tensor_map(CUtensorMap *out, const void *base, int C, long long rows, long long planes,
                        long long plane_stride, int box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return vbq_fail(VBQ_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)rows, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)C * 4u, (cuuint64_t)plane_stride * 4u};
    const cuuint32_t box[3] = {VBQ_GROUP, (cuuint32_t)box_rows, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return vbq_fail(VBQ_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return VBQ_OK;
}

static_assert(vbq_walk_tree_floats(1) == kWalkFloats, "layout of the walk tree (tree.cuh, quantize.cu)");

constexpr int kConsumerWarps = 19, kPairs = 2;   // measured: 31x1 46.7, 23x1 46.1, 15x2 45.0, 17x2 44.9, 19x2 43.4, 21x2 45.1 (spills), 11x4 48.5 us per Kodak step

template <bool PRUNE, bool TOTALS, int NT>
static int launch_tma3(const QArgs &a, int dev, int sms, cudaStream_t st) {
    constexpr int W = kConsumerWarps, P = kPairs;
#ifdef VBQ_DEV_VARIANTS   // development: geometry variants of the benchmark's kernel, chosen by VBQ_TMA_VARIANT
    if constexpr (!PRUNE && TOTALS && NT == 10) {
        const char *v = getenv("VBQ_TMA_VARIANT");
        const int vi = v ? atoi(v) : 0;
        if ((a.outm & 15u) == (2u | 8u) && vi > 0) {
            if (vi == 1) return launch_tma<false, false, false, true, 10, 2 | 8, 20, 2>(a, a.qidx, a.bits, dev, sms, st);
            if (vi == 2) return launch_tma<false, false, false, true, 10, 2 | 8, 22, 2>(a, a.qidx, a.bits, dev, sms, st);
            if (vi == 3) return launch_tma<false, false, false, true, 10, 2 | 8, 18, 2>(a, a.qidx, a.bits, dev, sms, st);
            if (vi == 4) return launch_tma<false, false, false, true, 10, 2 | 8, 21, 2>(a, a.qidx, a.bits, dev, sms, st);
        }
    }
#endif
#ifdef VBQ_DEV_ONE   // development builds: only the benchmark's variant (fast compile, small SASS listing)
    if constexpr (!PRUNE && TOTALS && NT == 10) {
        if ((a.outm & 15u) == (2u | 8u)) return launch_tma<false, false, false, true, 10, 2 | 8, W, P>(a, a.qidx, a.bits, dev, sms, st);
    }
    return -1;
#else
    switch (a.outm & 15u) {
        case 2u | 8u: return launch_tma<false, false, PRUNE, TOTALS, NT, 2 | 8, W, P>(a, a.qidx, a.bits, dev, sms, st);
        case 1u | 4u: return launch_tma<false, false, PRUNE, TOTALS, NT, 1 | 4, W, P>(a, a.zhat, a.level, dev, sms, st);
        case 1u: return launch_tma<false, false, PRUNE, TOTALS, NT, 1, W, P>(a, a.zhat, nullptr, dev, sms, st);
        case 2u: return launch_tma<false, false, PRUNE, TOTALS, NT, 2, W, P>(a, a.qidx, nullptr, dev, sms, st);
        case 0u:
            if constexpr (TOTALS) return launch_tma<false, false, PRUNE, TOTALS, NT, 0, W, P>(a, nullptr, nullptr, dev, sms, st);
            return -1;
        default: return -1;
    }
#endif
}

// raw code lengths that do not depend on the channel, available on the host (vbq_quantize_hp) and non-decreasing in the
// bit depth; max_bits_per_coord <= 10; C % 4 == 0 and 16-byte aligned arrays (TMA); at most two outputs of the compiled
// sets.  Returns -1 if not applicable (the caller falls back to vbq_bisect_kernel, the cp.async formulation).
int vbq_launch_quantize_tma(const QArgs &a, int dev, int sms, cudaStream_t st) {
    if (a.len || a.em || a.N > kSmemDepth || a.C % 4 != 0 || !a.h_pen || a.pen_channels != 1) return -1;
    uintptr_t al = (uintptr_t)a.mu | (uintptr_t)a.sigma | (uintptr_t)a.zhat | (uintptr_t)a.qidx | (uintptr_t)a.level |
                   (uintptr_t)a.bits | (uintptr_t)a.packed;
    if (al & 15) return -1;
    if (a.rows * (long long)a.C >= (1ll << 31)) return -1;
    for (int l = 0; l < a.n_lambda; ++l) {   // the certified ranking needs 0 <= pen_0 <= pen_1 <= ... (false for NaN)
        float prev = 0.0f;
        for (int n = 0; n <= a.N; ++n) {
            const float p = a.h_pen[(size_t)l * (a.N + 1) + n];
            if (!(p >= prev)) return -1;
            prev = p;
        }
    }
    // results do not depend on the early-exit tests, so only the common case (compile-time depth) is compiled without them
    const bool prune = !(a.flags & VBQ_FLAG_NO_PRUNE);
    if (a.N == kSmemDepth) {
        if (prune) return a.totals ? launch_tma3<true, true, kSmemDepth>(a, dev, sms, st) : launch_tma3<true, false, kSmemDepth>(a, dev, sms, st);
        return a.totals ? launch_tma3<false, true, kSmemDepth>(a, dev, sms, st) : launch_tma3<false, false, kSmemDepth>(a, dev, sms, st);
    }
    return a.totals ? launch_tma3<true, true, 0>(a, dev, sms, st) : launch_tma3<true, false, 0>(a, dev, sms, st);
}
