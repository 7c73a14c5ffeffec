// quantize_tma_both.cu — instantiations of vbq_bisect_tma_kernel (quantize_tma.cuh) for ARBITRARY non-negative penalties:
// both ends of the reference's bracket stay candidates at every bit depth.  This is the mode every reference call runs in
// after build_entropy_models (corrected code lengths n + R_lambda[c, n], quantizer.py:166-180; utils.py:392-396), with
// the entropy-model gather of compress_latents (quantizer.py:223-228) fused in.
#include "quantize_tma.cuh"

// consumer warps x coordinate pairs per thread, measured on the Kodak batch, one lambda, no entropy-model output: with the
// neighbour mask and the literal search out of line 17x2 53.8, 19x2 52.2, 23x1 55.1 us (both ends at every depth: 15x2 73,
// 17x2 71, 19x2 80, 23x1 74 us)
constexpr int kBothWarps = 19, kBothPairs = 2;

// Entropy-model bits (quantizer.py:226-228: entropy_models[lamb] gathered at the sorted index of z_hat).  Gathering inside
// the search kernel costs one 32-byte L2 sector per coordinate — 32 L1 wavefronts per warp load, 33 us per launch on the
// Kodak batch, and hiding the latency behind the next unit's search does not help (measured: it is the LSU, not the wait) —
// because the (C, Q) table does not fit beside the walk tree in shared memory.  Instead the search kernel (EM == 2) sends the
// winner's heap index through the TMA box of the code length, and this kernel — one CTA per (lambda, 16-channel group, row range)
// with the group's 16 x Q table in shared memory (heap order) — turns it into the code length and the entropy-model bits: 12 bytes of streaming traffic per coordinate in 16-byte accesses.  The sum of the
// entropy-model bits goes to column 2 of the totals (exact integer partials, last CTA by ticket, like the search kernel).
constexpr int kGatherThreads = 1024;
constexpr int kGatherStep = kGatherThreads / 4;   // rows per pass: a thread owns 4 channels of a row

// LEN: the heap indices sit in the code-length planes, which receive the code lengths (after vbq_bisect_tma_kernel, whose two
// TMA output boxes are taken); else they sit in the entropy-model planes themselves (after vbq_both_sweep_kernel, which writes
// the code lengths itself): 8 instead of 12 bytes and 4 instead of 8 table look-ups per coordinate.
template <bool TOTALS, bool LEN>
__global__ void __launch_bounds__(kGatherThreads, 1) em_gather_kernel(const float *__restrict__ em, const float *__restrict__ len,
                                                                     int pen_channels, float *bits, float *em_bits,
                                                                     long long rows, int C, long long rows_per_cta,
                                                                     long long lam_stride, double *totals, Acc128 *partials,
                                                                     unsigned *ticket) {
    extern __shared__ __align__(16) float sE[];   // [16][Q]; then the code lengths as [N + 1][4][32]: one bank per lane
    __shared__ Acc128 sW[kGatherThreads / 32];
    __shared__ bool sLast;
    constexpr int N = kSmemDepth, Q = (2 << N) - 1;   // the both-ends kernel serves max_bits_per_coord == 10
    const int g = blockIdx.y, lam = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *sLen = sE + VBQ_GROUP * Q;
    // programmatic stream serialization: the tables are inputs of the call (complete before the search kernel started);
    // the sorted indices are read after the wait
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    {   // the group's rows of the (L, C, Q) table are contiguous; sorted index q -> heap index K = 2^n + j, where
        // q + 1 = (2 j + 1) 2^(N - n): winners crowd the shallow depths, whose SORTED indices all fall into bank 31
        const int gc = min(VBQ_GROUP, C - g * VBQ_GROUP);
        const float *src = em + ((size_t)lam * C + (size_t)g * VBQ_GROUP) * Q;
#pragma unroll 8
        for (int i = threadIdx.x; i < gc * Q; i += kGatherThreads) {
            const int ch = i / Q, q1 = i - ch * Q + 1, t = __ffs(q1) - 1;
            sE[ch * Q + (1 << (N - t)) + (q1 >> (t + 1)) - 1] = __ldg(src + i);
        }
        for (int k = threadIdx.x; LEN && k < (N + 1) * 128; k += kGatherThreads) {
            const int c = min(g * VBQ_GROUP + 4 * (k & 3) + ((k >> 5) & 3), C - 1), n = k >> 7;
            sLen[k] = len ? __ldg(len + ((size_t)lam * pen_channels + (pen_channels == 1 ? 0 : c)) * (N + 1) + n) : (float)n;
        }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __syncthreads();
    const int col = 4 * (threadIdx.x & 3), c = g * VBQ_GROUP + col;
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    float *pb = bits + (size_t)lam * lam_stride, *pe = em_bits + (size_t)lam * lam_stride;
    Acc128 acc = {0, 0};
    auto fetch = [&](int4 &q, const long long r) {
        if (r < r1) q = __ldcs(reinterpret_cast<const int4 *>((LEN ? pb : pe) + (size_t)r * C + c));
    };
    auto emit = [&](const int4 q, const long long r) {
        if (r >= r1) return;
        const int qq[4] = {q.x, q.y, q.z, q.w};
        float b[4], e[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int K = min(max(qq[k], 1), Q);
            e[k] = sE[(col + k) * Q + K - 1];
            if (LEN) b[k] = sLen[((31 - __clz(K)) * 4 + k) * 32 + lane];
        }
        if (LEN) *reinterpret_cast<float4 *>(pb + (size_t)r * C + c) = make_float4(b[0], b[1], b[2], b[3]);
        *reinterpret_cast<float4 *>(pe + (size_t)r * C + c) = make_float4(e[0], e[1], e[2], e[3]);
        if (TOTALS) acc.add_q24((e[0] + e[1]) + (e[2] + e[3]));   // 4 terms per rounding, like the search kernel
    };
    if (c < C) {   // two passes in flight, the next two requested before the current ones are used (ping-pong registers)
        long long r = r0 + (threadIdx.x >> 2);
        int4 a0 = {0, 0, 0, 0}, a1 = a0, b0 = a0, b1 = a0;
        fetch(a0, r);
        fetch(a1, r + kGatherStep);
        while (r < r1) {
            fetch(b0, r + 2 * kGatherStep);
            fetch(b1, r + 3 * kGatherStep);
            emit(a0, r);
            emit(a1, r + kGatherStep);
            r += 2 * kGatherStep;
            if (r >= r1) break;
            fetch(a0, r + 2 * kGatherStep);
            fetch(a1, r + 3 * kGatherStep);
            emit(b0, r);
            emit(b1, r + kGatherStep);
            r += 2 * kGatherStep;
        }
    }
    if (!TOTALS) return;
    acc.warp_sum();
    if (lane == 0) sW[warp] = acc;
    __syncthreads();
    if (warp == 0) {
        acc = sW[lane];
        acc.warp_sum();
        const unsigned n_cta = gridDim.x * gridDim.y;
        Acc128 *mine = partials + (size_t)lam * n_cta;
        if (lane == 0) {
            mine[blockIdx.y * gridDim.x + blockIdx.x] = acc;
            unsigned t;
            asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(t) : "l"(ticket + lam) : "memory");
            sLast = (t == n_cta - 1);
        }
        __syncwarp();
        if (sLast) {
            __threadfence();
            const volatile unsigned long long *vp = reinterpret_cast<const volatile unsigned long long *>(mine);
            Acc128 x = {0, 0};
            for (unsigned k = lane; k < n_cta; k += 32) {
                Acc128 t_;
                t_.lo = vp[2 * k];
                t_.hi = vp[2 * k + 1];
                x.add(t_);
            }
            x.warp_sum();
            if (lane == 0) {
                totals[lam * VBQ_TOTALS + 2] += x.value();
                ticket[lam] = 0u;
            }
        }
    }
}
static_assert(kGatherThreads / 32 == 32, "one partial per lane of warp 0");

// one function per kernel instantiation: VBQ_ENSURE_MAX_SMEM keeps a static "attribute set" flag per call site
template <bool TOTALS, bool LEN>
static int launch_em_gather_t(const QArgs &a, int dev, long long slices, long long rows_per_cta, size_t smem, cudaStream_t st) {
    auto kern = em_gather_kernel<TOTALS, LEN>;
    VBQ_ENSURE_MAX_SMEM(kern, dev);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)slices, (unsigned)a.n_groups, (unsigned)a.n_lambda);
    cfg.blockDim = dim3(kGatherThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool pdl = !getenv("VBQ_NO_PDL");
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a.em, a.len, a.pen_channels, a.bits, a.em_bits, a.rows, a.C, rows_per_cta, a.lam_stride,
                                TOTALS ? a.totals : nullptr, TOTALS ? reinterpret_cast<Acc128 *>(a.partials) : nullptr,
                                TOTALS ? a.ticket : nullptr));
    return VBQ_OK;
}

static int launch_em_gather(const QArgs &a, int dev, int sms, cudaStream_t st, bool len_planes = true) {
    const size_t smem = ((size_t)VBQ_GROUP * a.Q + (size_t)(a.N + 1) * 128) * sizeof(float);
    // one CTA per SM (the table takes 131 KB): row slices such that the CTAs fill whole waves
    const long long per_slice = (long long)a.n_groups * a.n_lambda;
    long long max_slices = (a.rows + 2 * kGatherStep - 1) / (2 * kGatherStep);
    if (max_slices > 64) max_slices = 64;
    if (max_slices * a.n_groups > 2 * kMaxGrid) max_slices = 2 * kMaxGrid / a.n_groups;   // partials of a lambda fit the workspace
    if (max_slices < 1) max_slices = 1;
    long long slices = 1;
    double best = 0.0;
    for (long long s_ = 1; s_ <= max_slices; ++s_) {
        const long long ctas = s_ * per_slice, waves = (ctas + sms - 1) / sms;
        const double eff = (double)ctas / (double)(waves * sms);
        if (eff > best + 1e-9) { best = eff; slices = s_; }
        if (eff >= 0.93) break;
    }
    const long long rows_per_cta = (a.rows + slices - 1) / slices;
    if (len_planes)
        return a.totals ? launch_em_gather_t<true, true>(a, dev, slices, rows_per_cta, smem, st)
                        : launch_em_gather_t<false, true>(a, dev, slices, rows_per_cta, smem, st);
    return a.totals ? launch_em_gather_t<true, false>(a, dev, slices, rows_per_cta, smem, st)
                    : launch_em_gather_t<false, false>(a, dev, slices, rows_per_cta, smem, st);
}

// For vbq_both_sweep_kernel (sweep_both.cu), which leaves the winners' heap indices in the entropy-model planes of all lambdas
int vbq_launch_em_gather(const QArgs &a, int dev, int sms, cudaStream_t st) {
    if (a.N != kSmemDepth || !a.em || !a.em_bits || a.C % 4 != 0 || a.n_groups > 2 * kMaxGrid || a.n_lambda > 65535) return -1;
    if ((((uintptr_t)a.em_bits | (uintptr_t)a.em) & 15) != 0) return -1;
    return launch_em_gather(a, dev, sms, st, false);
}

template <int EM, bool TOTALS, int OUT>
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
    // the gather as a second, streaming kernel when the bits are an output
    static const bool in_kernel = getenv("VBQ_EM_IN_KERNEL") != nullptr;
    const bool defer = em && a.em_bits && outs == (1u | 8u) && !in_kernel && ((uintptr_t)a.em_bits & 15) == 0 &&
                       ((uintptr_t)a.em & 15) == 0 && a.n_groups <= 2 * kMaxGrid && a.n_lambda <= 65535;
#ifdef VBQ_DEV_ONE
    if (outs == (1u | 8u) && em && a.totals) {
        if (!defer) return launch_both<1, true, 1 | 8>(a, a.zhat, a.bits, dev, sms, st);
        RETURN_IF((launch_both<2, true, 1 | 8>(a, a.zhat, a.bits, dev, sms, st)));
        return launch_em_gather(a, dev, sms, st);
    }
    return -1;
#else
    if (outs == (1u | 8u)) {
        if (defer) {
            RETURN_IF(a.totals ? (launch_both<2, true, 1 | 8>(a, a.zhat, a.bits, dev, sms, st))
                               : (launch_both<2, false, 1 | 8>(a, a.zhat, a.bits, dev, sms, st)));
            return launch_em_gather(a, dev, sms, st);
        }
        if (em) return a.totals ? launch_both<1, true, 1 | 8>(a, a.zhat, a.bits, dev, sms, st)
                                : launch_both<1, false, 1 | 8>(a, a.zhat, a.bits, dev, sms, st);
        return a.totals ? launch_both<0, true, 1 | 8>(a, a.zhat, a.bits, dev, sms, st)
                        : launch_both<0, false, 1 | 8>(a, a.zhat, a.bits, dev, sms, st);
    }
    if (outs == 2u)
        return a.totals ? launch_both<0, true, 2>(a, a.qidx, nullptr, dev, sms, st)
                        : launch_both<0, false, 2>(a, a.qidx, nullptr, dev, sms, st);
    if (outs == 0u && a.totals)
        return em ? launch_both<1, true, 0>(a, nullptr, nullptr, dev, sms, st)
                  : launch_both<0, true, 0>(a, nullptr, nullptr, dev, sms, st);
    return -1;
#endif
}
