// quantize_bisect.cu — the single-lambda rate-distortion search as a certified BISECTION (sm_100a).
//
// Reference behaviour reproduced (paths relative to mandt-lab/vbq): the same as quantize.cu —
//   img-compression/quantizer.py:65-80, :156-188   bracket of mu at every bit depth, candidates left_0..left_N,
//                                                  right_1..right_N with code length n
//   img-compression/utils.py:318-320, :392-415     score fl(fl(-0.5 fl(t^2)) - fl(lambda n)), t = fl(fl(z-mu)/sigma),
//                                                  first argmax
//
// Why one candidate per depth is enough.  The code points of depths <= n form the nested dyadic grid
// G_n = {F^-1(k 2^-(n+1))}.  Walking the tree keeps the open interval (lo, hi) of G_(n-1) that contains mu; the PATH
// NODE z_n is the only depth-n point inside it.  The reference's depth-n bracket is {z_n, z'} where z' is the next
// depth-n point beyond lo or hi — or z_n itself / the clipped edge point when there is none.  z' lies beyond a bracket
// end `a` that (i) is a path node of a shallower depth m < n, (ii) is at least as close to mu, (iii) has
// pen_m <= pen_n when the penalties are non-decreasing in depth, and (iv) precedes z' in the reference's candidate
// order (same side, smaller depth).  Every float32 operation of the score is monotone, so score(z') <= score(a) and
// z' can never be the first maximiser.  Hence the winner is always one of the N+1 path nodes: one shared-memory load,
// one compare and one score per depth — the neighbour load, the nearer-end selection and the left/right decision of
// the bracket walk (quantize_kernel.cuh) disappear.
//
// Certified approximate scoring.  Only the IDENTITY of the winner is returned, so the path nodes are ranked with a
// cheap loss  A_n = fma(t, t, pen_n),  t = (z_n - mu) * (sqrt(1/2) / sigma)  (3 packed f32x2 instructions per two
// coordinates) instead of the 6-instruction IEEE-division chain.  A_n and the reference's -score E_n are both
// non-negative floats that differ by < 15 float32 roundings (error analysis in DESIGN.md §4), i.e. their bit patterns
// differ by < 30 as integers.  The depth is embedded in the 4 low bits of the pattern (key_n); the winner is the
// integer minimum (VIMNMX3), and a second pass (VIADDMNMX) measures the gap to the runner-up.  If the gap exceeds
// kKeyGuard = 192 > 2*(30+15) the reference's float32 scores are strictly ordered the same way and the result is
// certified identical; otherwise (about 1 coordinate in 10^4), or when the penalties are not non-decreasing and
// non-negative, the coordinate is redone by `reference_search`, the literal two-ended walk with IEEE arithmetic.
#include <stdlib.h>

#include "tree.cuh"

constexpr unsigned kKeyGuard = 192u;
constexpr unsigned kKeyMask = 0xfffffff0u;

// Literal restatement of the reference search for one coordinate (slow path): both bracket ends of every depth,
// IEEE float32 scores, first maximum in the order left_0..left_N, right_1..right_N.  Returns depth << 16 | index.
// sTc = this channel's column of the padded shared-memory tree, sPenc = its penalties (stride VBQ_GROUP).
static __device__ __noinline__ int reference_search(const float *sTc, const float *sPenc, float mu, float sg, int N) {
    const float rs = rcp_rn(sg);
    const float z0 = sTc[entry_of(0, 0) * VBQ_GROUP];
    float bestL = score_exact(z0, mu, sg, rs, -sPenc[0]), bestR = -CUDART_INF_F;
    int nL = 0, iL = 0, nR = 0, iR = 0;
    int ip = mu > z0 ? 1 : 0;   // index of the path node at the next depth
    for (int n = 1; n <= N; ++n) {
        const float zp = sTc[entry_of(n, ip) * VBQ_GROUP];
        const int b = mu > zp ? 1 : 0;
        const int fg = ip + b;   // number of depth-n points below mu = searchsorted(side='left'), quantizer.py:74
        const int il = clamp_index(fg, n, N, false), ir = clamp_index(fg, n, N, true);
        const float npn = -sPenc[n * VBQ_GROUP];
        const float sl = score_exact(sTc[entry_of(n, il) * VBQ_GROUP], mu, sg, rs, npn);
        const float sr = score_exact(sTc[entry_of(n, ir) * VBQ_GROUP], mu, sg, rs, npn);
        if (sl > bestL) { bestL = sl; nL = n; iL = il; }
        if (sr > bestR) { bestR = sr; nR = n; iR = ir; }
        ip = 2 * ip + b;
    }
    return bestR > bestL ? (nR << 16 | iR) : (nL << 16 | iL);
}

__device__ __forceinline__ float rcp_approx(float x) {   // MUFU.RCP, relative error <= 2^-23
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// (bits & ~15) | n as ONE LOP3 (the depth constant sits in a register)
__device__ __forceinline__ unsigned make_key(float loss, unsigned n) {
    unsigned k;
    asm("lop3.b32 %0, %1, 0xfffffff0, %2, 0xEA;" : "=r"(k) : "r"(__float_as_uint(loss)), "r"(n));
    return k;
}

// NT > 0: max_bits_per_coord == NT at compile time; NT == 0: runtime depth (<= kSmemDepth).
template <bool PRUNE, bool TOTALS, int NT, int U, int kThreads>
__global__ void __launch_bounds__(kThreads, 1) vbq_bisect_kernel(const QArgs a) {
    static_assert(U % 2 == 0, "coordinates are processed in f32x2 pairs");
    constexpr int RP = kThreads / VBQ_GROUP;              // rows covered by one pass of the CTA
    constexpr int P = U / 2;
    extern __shared__ __align__(16) float smem[];
    float *sT = smem;                                   // [kPadEntries][16] code points of depths 0..10
    float *sPen = sT + kPadEntries * VBQ_GROUP;         // [kSmemDepth+1][16] penalties (+inf beyond N)
    float *sStage = sPen + (kSmemDepth + 1) * VBQ_GROUP;   // [kStages][2][U][kThreads] thread-private staging ring
    float *myStage = sStage + threadIdx.x;
    __shared__ double sRed[VBQ_TOTALS][kMaxThreads / 32];
    __shared__ unsigned sGuard[VBQ_GROUP];
    __shared__ bool sLast;

    const int N = NT > 0 ? NT : a.N;
    const int lam = blockIdx.y;
    const unsigned outm = a.outm & 15u;
    const int col = threadIdx.x & (VBQ_GROUP - 1);
    const int rsub = threadIdx.x >> 4;
    const bool logvar = (a.flags & VBQ_FLAG_LOGVAR) != 0;
    const long long u0 = a.total_units * blockIdx.x / gridDim.x;
    const long long u1 = a.total_units * (blockIdx.x + 1) / gridDim.x;
    const int C = a.C;
    const int rows = (int)a.rows;                       // the host splits calls so that rows*C < 2^29
    const size_t lam_off = (size_t)lam * (size_t)a.lam_stride;
    // shared-memory byte address of padded entry (n, i) of this thread's channel = pbi + 64*K + 128*n, K = 2^n + i
    // (entry_of(n, i) = K + 2n): K is the 1-based heap index of the node, children 2K and 2K+1
    const int pbi = (int)__cvta_generic_to_shared(sT + col);
    const float *sTc = sT + col;

    double acc_dist = 0.0;
    int acc_level = 0;   // < 2^31: at most 2^29 coordinates per launch, depth <= 10

    long long unit = u0;
    while (unit < u1) {
        // ---- segment: a run of row passes inside one 16-channel group --------------------------------------
        const int g = (int)(unit / a.passes);
        const int p0 = (int)(unit - (long long)g * a.passes);
        const int p1 = (int)min(a.passes, (long long)p0 + (u1 - unit));
        unit += p1 - p0;

        __syncthreads();
        {
            const float4 *src = reinterpret_cast<const float4 *>(a.packed + (size_t)g * kPadEntries * VBQ_GROUP);
            float4 *dst = reinterpret_cast<float4 *>(sT);
            for (int k = threadIdx.x; k < kPadEntries * (VBQ_GROUP / 4); k += kThreads) dst[k] = __ldg(src + k);
            if (threadIdx.x < VBQ_GROUP) {
                const int j = threadIdx.x;
                const int cj = min(g * VBQ_GROUP + j, C - 1);
                const size_t po = ((size_t)lam * a.pen_channels + (a.pen_channels == 1 ? 0 : cj)) * (N + 1);
                float prev = 0.0f;
                bool mono = true;   // certified ranking needs 0 <= pen_0 <= pen_1 <= ... (false for NaN)
                for (int n = 0; n <= kSmemDepth; ++n) {
                    const float p = n <= N ? a.pen[po + n] : CUDART_INF_F;
                    mono = mono && (p >= prev);
                    prev = p;
                    sPen[n * VBQ_GROUP + j] = p;
                }
                sGuard[j] = mono ? kKeyGuard : 0xffffffffu;   // 0xffffffff: every coordinate takes the slow path
            }
        }
        __syncthreads();

        const int c = g * VBQ_GROUP + col;
        const bool c_ok = c < C;
        const int cc = min(c, C - 1);
        const char *mu_b = reinterpret_cast<const char *>(a.mu);
        const char *sg_b = reinterpret_cast<const char *>(a.sigma);
        float *zhat_c = a.zhat ? a.zhat + lam_off : nullptr;
        int *qidx_c = a.qidx ? a.qidx + lam_off : nullptr;
        int *level_c = a.level ? a.level + lam_off : nullptr;
        float *bits_c = a.bits ? a.bits + lam_off : nullptr;
        float2 pen2[kSmemDepth + 1];
#pragma unroll
        for (int n = 0; n <= kSmemDepth; ++n) {
            const float v = sPen[n * VBQ_GROUP + col];
            pen2[n] = make_float2(v, v);
        }
        const unsigned guard = sGuard[col];
        const float z0 = sTc[entry_of(0, 0) * VBQ_GROUP];

        const int row_end = c_ok ? min(p1 * RP, rows) : 0;   // threads of channels >= C never pass the row test
        int row = p0 * RP + rsub;
        unsigned off = (unsigned)row * (unsigned)C + (unsigned)cc;     // element offset of (row, channel)
        const unsigned off_step = (unsigned)(RP * C);

        auto stage_rows = [&](int it_row, unsigned it_off, int slot) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (it_row + u * RP < row_end) {
                    const unsigned bo = (it_off + u * off_step) * 4u;
                    cp_async_f32(myStage + ((slot * 2 + 0) * U + u) * kThreads,
                                 reinterpret_cast<const float *>(mu_b + bo));
                    cp_async_f32(myStage + ((slot * 2 + 1) * U + u) * kThreads,
                                 reinterpret_cast<const float *>(sg_b + bo));
                }
            }
            cp_async_commit();
        };
#pragma unroll
        for (int k = 0; k < kStages - 1; ++k) stage_rows(row + k * U * RP, off + k * U * off_step, k);
        int slot = 0;

        for (; row - rsub < p1 * RP; row += U * RP, off += U * off_step) {
            float mu[U], sg[U];
            float2 nmu2[P], r2[P];   // r2 = sqrt(1/2)/sigma
            cp_async_wait<kStages - 2>();
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool ok = row + u * RP < row_end;
                mu[u] = ok ? myStage[((slot * 2 + 0) * U + u) * kThreads] : 0.0f;
                float s = ok ? myStage[((slot * 2 + 1) * U + u) * kThreads] : 1.0f;
                if (logvar) s = sqrtf(expf(s));
                sg[u] = s;
            }
            {
                const int ps = slot == 0 ? kStages - 1 : slot - 1;
                stage_rows(row + (kStages - 1) * U * RP, off + (kStages - 1) * U * off_step, ps);
                slot = slot == kStages - 1 ? 0 : slot + 1;
            }
#pragma unroll
            for (int k = 0; k < P; ++k) {
                nmu2[k] = make_float2(-mu[2 * k], -mu[2 * k + 1]);
                r2[k] = __fmul2_rn(make_float2(rcp_approx(sg[2 * k]), rcp_approx(sg[2 * k + 1])),
                                   make_float2(0.70710678f, 0.70710678f));
            }

            unsigned key[U][kSmemDepth + 1];
            unsigned K[U];   // 1-based heap index of the path node at the current depth
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int n = 0; n <= kSmemDepth; ++n) key[u][n] = 0x7ffffff0u | (unsigned)n;

            // ---- depth 0: the median ---------------------------------------------------------------------
#pragma unroll
            for (int k = 0; k < P; ++k) {
                const float2 d = __fadd2_rn(make_float2(z0, z0), nmu2[k]);
                const float2 t = __fmul2_rn(d, r2[k]);
                const float2 A = __ffma2_rn(t, t, pen2[0]);
                key[2 * k][0] = make_key(A.x, 0u);
                key[2 * k + 1][0] = make_key(A.y, 0u);
                K[2 * k] = __funnelshift_l(__float_as_uint(d.x), 1u, 1);       // 2 + (mu > z0)
                K[2 * k + 1] = __funnelshift_l(__float_as_uint(d.y), 1u, 1);
            }
            int m_done = 0;   // deepest level scored (warp-uniform)

            // ---- depths 1..N, fully unrolled -------------------------------------------------------------
#pragma unroll
            for (int n = 1; n <= kSmemDepth; ++n) {
                if (NT == 0 && n > N) break;
                if (NT > 0 && n > NT) break;
                if (PRUNE && n % 3 == 0) {
                    // sound early exit: every deeper loss is >= pen_n, so once the best key plus the guard is below
                    // the key of pen_n no deeper candidate can win or come within the guard
                    const unsigned floor_key = __float_as_uint(pen2[n].x) & kKeyMask;
                    bool done = guard == kKeyGuard;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        unsigned m = key[u][0];
#pragma unroll
                        for (int j = 1; j < n; ++j) m = min(m, key[u][j]);
                        done = done && floor_key > kKeyGuard + 16u && m < floor_key - (kKeyGuard + 16u);
                    }
                    if (__all_sync(0xffffffffu, done)) break;
                }
                float z[U];
#pragma unroll
                for (int u = 0; u < U; ++u) z[u] = lds_u32((unsigned)(imad((int)K[u], kRowStrideBytes, pbi) + 2 * n * kRowStrideBytes));
#pragma unroll
                for (int k = 0; k < P; ++k) {
                    const float2 d = __fadd2_rn(make_float2(z[2 * k], z[2 * k + 1]), nmu2[k]);
                    if (!(NT > 0 && n == NT)) {   // the branch of the deepest compile-time depth is never used
                        K[2 * k] = __funnelshift_l(__float_as_uint(d.x), K[2 * k], 1);          // 2K + (mu > z)
                        K[2 * k + 1] = __funnelshift_l(__float_as_uint(d.y), K[2 * k + 1], 1);
                    }
                    const float2 t = __fmul2_rn(d, r2[k]);
                    const float2 A = __ffma2_rn(t, t, pen2[n]);
                    key[2 * k][n] = make_key(A.x, (unsigned)n);
                    key[2 * k + 1][n] = make_key(A.y, (unsigned)n);
                }
                m_done = n;
            }
            const int kd = (NT > 0 && m_done == NT) ? NT : m_done + 1;   // depth of the node K points at

            // ---- winner and certificate -----------------------------------------------------------------------
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned *k_ = key[u];
                unsigned m = __vimin3_u32(k_[0], k_[1], k_[2]);
                m = __vimin3_u32(m, k_[3], k_[4]);
                m = __vimin3_u32(m, k_[5], k_[6]);
                m = __vimin3_u32(m, k_[7], k_[8]);
                m = __vimin3_u32(m, k_[9], k_[10]);
                const unsigned nm = ~m;   // key + ~m = key - m - 1: 0xffffffff for the winner itself
                unsigned gap = 0xffffffffu;
#pragma unroll
                for (int n = 0; n <= kSmemDepth; ++n) gap = __viaddmin_u32(k_[n], nm, gap);
                int n = (int)(m & 15u);
                int Pn = (int)(K[u] >> (kd - n));   // heap index of the winning path node: 2^n + i
                if (gap <= guard) {   // not certified (or penalties not monotone): literal search
                    const int r = reference_search(sTc, sPen + col, mu[u], sg[u], N);
                    n = r >> 16;
                    Pn = (1 << n) + (r & 0xffff);
                }
                if (row + u * RP < row_end) {
                    const unsigned o = off + u * off_step;
                    const int i = Pn - (1 << n);
                    const int q = ((2 * i + 1) << (N - n)) - 1;
                    if (outm & 2u) qidx_c[o] = q;
                    if (outm & 4u) level_c[o] = n;
                    if (outm & 8u) bits_c[o] = (float)n;
                    if (TOTALS || (outm & 1u)) {
                        const float zh = lds_u32((unsigned)(imad(Pn, kRowStrideBytes, pbi) + 2 * n * kRowStrideBytes));
                        if (outm & 1u) zhat_c[o] = zh;
                        if (TOTALS) {
                            const float r1 = u & 1 ? r2[u / 2].y : r2[u / 2].x;
                            const float t = (zh - mu[u]) * r1;
                            acc_level += n;
                            acc_dist += (double)(t * t);
                        }
                    }
                }
            }
        }
        cp_async_wait<0>();
    }

    if (TOTALS) {
        // raw-length mode: the code length of depth n is n itself; no entropy model on this path
        double v[VBQ_TOTALS] = {(double)acc_level, (double)acc_level, 0.0, acc_dist};
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
        if (sLast && threadIdx.x < VBQ_TOTALS) {   // the last CTA of this lambda adds the partials in a fixed order
            __threadfence();
            const volatile double *p = a.partials + (size_t)lam * kMaxGrid * VBQ_TOTALS;
            double s = a.accumulate ? a.totals[lam * VBQ_TOTALS + threadIdx.x] : 0.0;
            for (unsigned b = 0; b < gridDim.x; ++b) s += p[b * VBQ_TOTALS + threadIdx.x];
            a.totals[lam * VBQ_TOTALS + threadIdx.x] = s;
            if (threadIdx.x == 0) a.ticket[lam] = 0u;
        }
    }
}

template <bool PRUNE, bool TOTALS, int NT, int U, int T>
static int launch_bisect(QArgs a, int dev, int sms, cudaStream_t st) {
    constexpr int rows_per_pass = T / VBQ_GROUP;
    a.passes = (a.rows + rows_per_pass - 1) / rows_per_pass;
    a.total_units = a.passes * a.n_groups;
    long long gx = (a.total_units + U - 1) / U;
    if (gx > sms) gx = sms;
    if (gx > kMaxGrid) gx = kMaxGrid;
    const size_t smem = ((size_t)kPadEntries * VBQ_GROUP + (size_t)(kSmemDepth + 1) * VBQ_GROUP +
                         (size_t)kStages * 2 * U * T) * sizeof(float);
    auto kern = vbq_bisect_kernel<PRUNE, TOTALS, NT, U, T>;
    VBQ_ENSURE_MAX_SMEM(kern, dev);
    kern<<<dim3((int)gx, a.n_lambda), T, smem, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    return VBQ_OK;
}

template <bool PRUNE, int U, int T>
static int launch_bisect2(const QArgs &a, int dev, int sms, cudaStream_t st) {
    const bool tot = a.totals != nullptr;
    if (a.N == kSmemDepth)
        return tot ? launch_bisect<PRUNE, true, kSmemDepth, U, T>(a, dev, sms, st)
                   : launch_bisect<PRUNE, false, kSmemDepth, U, T>(a, dev, sms, st);
    return tot ? launch_bisect<PRUNE, true, 0, U, T>(a, dev, sms, st) : launch_bisect<PRUNE, false, 0, U, T>(a, dev, sms, st);
}

// raw code lengths (no length table, no entropy model), max_bits_per_coord <= 10; returns -1 if not applicable
int vbq_launch_quantize_bisect(const QArgs &a, int dev, int sms, cudaStream_t st) {
    if (a.len || a.em || a.N > kSmemDepth) return -1;
    const bool prune = !(a.flags & VBQ_FLAG_NO_PRUNE);
    static const int tune = getenv("VBQ_TUNE") ? atoi(getenv("VBQ_TUNE")) : 0;   // development: threads/128*10 + U
    if (!prune) {
        if (tune == 42) return launch_bisect2<false, 2, 512>(a, dev, sms, st);
        if (tune == 62) return launch_bisect2<false, 2, 768>(a, dev, sms, st);
        if (tune == 44) return launch_bisect2<false, 4, 512>(a, dev, sms, st);
        if (tune == 34) return launch_bisect2<false, 4, 384>(a, dev, sms, st);
    }
    return prune ? launch_bisect2<true, 2, 640>(a, dev, sms, st) : launch_bisect2<false, 2, 640>(a, dev, sms, st);
}
