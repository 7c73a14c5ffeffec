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

#include <type_traits>

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

// (bits & mask) | n as ONE LOP3: the mask 0xfffffff0 sits in a register, the depth is an immediate
template <int DEPTH>
__device__ __forceinline__ unsigned make_key(float loss, unsigned mask) {
    unsigned k;
    asm("lop3.b32 %0, %1, %2, %3, 0xEC;" : "=r"(k) : "r"(__float_as_uint(loss)), "n"(DEPTH), "r"(mask));
    return k;
}

// NT > 0: max_bits_per_coord == NT at compile time; NT == 0: runtime depth (<= kSmemDepth).
// OUT >= 0: the set of requested outputs (bit 0 zhat, 1 qidx, 2 level, 3 bits) is known at compile time; OUT < 0: runtime.
// VEC: C % 4 == 0 and 16-byte aligned latents: every warp stages its own 2U rows x 16 channels of mu and sigma with ONE
// 16-byte cp.async per lane and iteration (a 512-byte tile per warp) instead of four 4-byte copies per thread.
template <bool PRUNE, bool TOTALS, int NT, int OUT, bool VEC, int U, int kThreads>
__global__ void __launch_bounds__(kThreads, 1) vbq_bisect_kernel(const QArgs a) {
    static_assert(U % 2 == 0, "coordinates are processed in f32x2 pairs");
    static_assert(!VEC || U == 2, "the per-warp tile of the vectorised staging is one 16-byte chunk per lane");
    constexpr int RP = kThreads / VBQ_GROUP;              // rows covered by one pass of the CTA
    constexpr int P = U / 2;
    extern __shared__ __align__(16) float smem[];
    float *sT = smem;                                   // [kPadEntries][16] code points of depths 0..10
    float *sPen = sT + kPadEntries * VBQ_GROUP;         // [kSmemDepth+1][16] penalties (+inf beyond N)
    // staging ring, kStages deep.  !VEC: [kStages][2][U][kThreads], thread-private slots.
    // VEC: per warp [kStages][2 arrays][2U rows][16 channels]; consumer (col, rsub&1) reads row u*2 + (rsub&1)
    float *sStage = sPen + (kSmemDepth + 1) * VBQ_GROUP;
    float *myStage = VEC ? sStage + (threadIdx.x >> 5) * (kStages * 4 * U * VBQ_GROUP) + (threadIdx.x & 31)
                         : sStage + threadIdx.x;
    constexpr int kSlotStride = VEC ? 4 * U * VBQ_GROUP : 2 * U * kThreads;   // floats per ring slot
    constexpr int kArrStride = VEC ? 2 * U * VBQ_GROUP : U * kThreads;        // mu -> sigma
    constexpr int kRowStride = VEC ? 2 * VBQ_GROUP : kThreads;               // u -> u + 1
    __shared__ double sRed[VBQ_TOTALS][kMaxThreads / 32];
    __shared__ unsigned sGuard[VBQ_GROUP];
    __shared__ bool sLast;

    const int N = NT > 0 ? NT : a.N;
    const int lam = blockIdx.y;
    const unsigned outm = OUT >= 0 ? (unsigned)OUT : (a.outm & 15u);
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
    const unsigned kmask = a.keymask;                   // 0xfffffff0 as a runtime value: stays in one register

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
        const float *mu_c = a.mu + cc, *sg_c = a.sigma + cc;      // element (row, channel) = base[row * C]
        float *zhat_c = a.zhat ? a.zhat + lam_off + cc : nullptr;
        int *qidx_c = a.qidx ? a.qidx + lam_off + cc : nullptr;
        int *level_c = a.level ? a.level + lam_off + cc : nullptr;
        float *bits_c = a.bits ? a.bits + lam_off + cc : nullptr;
        float pen[kSmemDepth + 1];
#pragma unroll
        for (int n = 0; n <= kSmemDepth; ++n) pen[n] = sPen[n * VBQ_GROUP + col];
        const unsigned guard = sGuard[col];
        const float z0 = sTc[entry_of(0, 0) * VBQ_GROUP];

        const int row_end = c_ok ? min(p1 * RP, rows) : 0;   // threads of channels >= C never pass the row test
        // a full iteration (all U row passes inside the matrix, all 16 channels of the group real) needs no predicates
        const int full_rows = (g * VBQ_GROUP + VBQ_GROUP <= C) ? min(p1 * RP, rows) : 0;
        int row = p0 * RP + rsub;
        unsigned off = (unsigned)row * (unsigned)C;          // element offset of row `row`
        const unsigned off_step = (unsigned)(RP * C);

        // stage kStages-1 iterations ahead; every iteration commits exactly one group (possibly empty)
        // VEC producer role of this lane: array (lane>>4), tile row ((lane>>2)&3) = u*2 + r, 16-byte chunk (lane&3)
        const int lane = threadIdx.x & 31;
        const int prod_row = 2 * (threadIdx.x >> 5) + ((lane >> 2) & 1) + ((lane >> 3) & 1) * RP;   // row inside a CTA iteration
        const int prod_col = g * VBQ_GROUP + (lane & 3) * 4;
        const float *prod_src = nullptr;   // advanced by one CTA iteration per stage_rows call
        float *prod_dst = nullptr;
        int prod_limit = 0;
        if (VEC) {
            prod_src = ((lane >> 4) ? a.sigma : a.mu) + ((size_t)(p0 * RP + prod_row) * C + prod_col);
            prod_dst = myStage - lane + (lane >> 4) * kArrStride + ((lane >> 2) & 3) * VBQ_GROUP + (lane & 3) * 4;
            prod_limit = prod_col < C ? min(p1 * RP, rows) : 0;
        }
        const size_t it_step = (size_t)U * off_step;
        auto stage_rows = [&](int it_row, unsigned it_off, int slot) {
            const bool full = it_row - rsub + U * RP <= full_rows;       // CTA-uniform
            if (VEC) {
                if (full || it_row - rsub + prod_row < prod_limit) cp_async_16(prod_dst + slot * kSlotStride, prod_src);
                prod_src += it_step;
            } else if (full) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    cp_async_f32_idx(myStage + slot * kSlotStride + u * kRowStride, mu_c, it_off + u * off_step);
                    cp_async_f32_idx(myStage + slot * kSlotStride + kArrStride + u * kRowStride, sg_c, it_off + u * off_step);
                }
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (it_row + u * RP < row_end) {
                        cp_async_f32_idx(myStage + slot * kSlotStride + u * kRowStride, mu_c, it_off + u * off_step);
                        cp_async_f32_idx(myStage + slot * kSlotStride + kArrStride + u * kRowStride, sg_c, it_off + u * off_step);
                    }
                }
            }
            cp_async_commit();
        };
#pragma unroll
        for (int k = 0; k < kStages - 1; ++k) stage_rows(row + k * U * RP, off + k * U * off_step, k);
        int slot = 0;

        // one iteration: U coordinates of this thread (rows row, row+RP, ...); CHECK = row bounds must be tested
        auto iteration = [&](auto check_tag) {
            constexpr bool CHECK = decltype(check_tag)::value;
            float mu[U], sg[U];
            float2 nmu2[P], r2[P];   // r2 = sqrt(1/2)/sigma
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool ok = !CHECK || row + u * RP < row_end;
                mu[u] = ok ? myStage[slot * kSlotStride + u * kRowStride] : 0.0f;
                float s = ok ? myStage[slot * kSlotStride + kArrStride + u * kRowStride] : 1.0f;
                if (logvar) s = sqrtf(expf(s));
                sg[u] = s;
            }
            {   // refill the slot consumed in the previous iteration
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
                const float2 A = __ffma2_rn(t, t, make_float2(pen[0], pen[0]));
                key[2 * k][0] = make_key<0>(A.x, kmask);
                key[2 * k + 1][0] = make_key<0>(A.y, kmask);
                K[2 * k] = __funnelshift_l(__float_as_uint(d.x), 1u, 1);       // 2 + (mu > z0)
                K[2 * k + 1] = __funnelshift_l(__float_as_uint(d.y), 1u, 1);
            }
            int m_done = 0;   // deepest level scored (warp-uniform)

            // ---- depths 1..N, fully unrolled -------------------------------------------------------------
            auto depth = [&](auto n_tag) {
                constexpr int n = decltype(n_tag)::value;
                float z[U];
#pragma unroll
                for (int u = 0; u < U; ++u)
                    z[u] = lds_pure((unsigned)(imad((int)K[u], kRowStrideBytes, pbi) + 2 * n * kRowStrideBytes));
#pragma unroll
                for (int k = 0; k < P; ++k) {
                    const float2 d = __fadd2_rn(make_float2(z[2 * k], z[2 * k + 1]), nmu2[k]);
                    if (!(NT > 0 && n == NT)) {   // the branch of the deepest compile-time depth is never used
                        K[2 * k] = __funnelshift_l(__float_as_uint(d.x), K[2 * k], 1);          // 2K + (mu > z)
                        K[2 * k + 1] = __funnelshift_l(__float_as_uint(d.y), K[2 * k + 1], 1);
                    }
                    const float2 t = __fmul2_rn(d, r2[k]);
                    const float2 A = __ffma2_rn(t, t, make_float2(pen[n], pen[n]));
                    key[2 * k][n] = make_key<n>(A.x, kmask);
                    key[2 * k + 1][n] = make_key<n>(A.y, kmask);
                }
                m_done = n;
            };
            auto prune_here = [&](int n) -> bool {
                // sound early exit: every deeper loss is >= pen_n, so once the best key plus the guard is below the
                // key of pen_n no deeper candidate can win or come within the guard
                const unsigned floor_key = __float_as_uint(pen[n]) & kKeyMask;
                bool done = guard == kKeyGuard && floor_key > kKeyGuard + 16u;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    unsigned m = key[u][0];
#pragma unroll
                    for (int j = 1; j <= kSmemDepth; ++j)
                        if (j < n) m = min(m, key[u][j]);
                    done = done && m < floor_key - (kKeyGuard + 16u);
                }
                return __all_sync(0xffffffffu, done);
            };
#define VBQ_DEPTH(n_)                                                        \
    if ((NT > 0 ? n_ <= NT : n_ <= N) && !stop) {                            \
        if (PRUNE && n_ % 3 == 0 && prune_here(n_)) stop = true;             \
        else depth(std::integral_constant<int, n_>{});                       \
    }
            bool stop = false;
            VBQ_DEPTH(1) VBQ_DEPTH(2) VBQ_DEPTH(3) VBQ_DEPTH(4) VBQ_DEPTH(5)
            VBQ_DEPTH(6) VBQ_DEPTH(7) VBQ_DEPTH(8) VBQ_DEPTH(9) VBQ_DEPTH(10)
#undef VBQ_DEPTH
            static_assert(kSmemDepth == 10, "the depth macro list above covers depths 1..10");
            const int kd = (NT > 0 && m_done == NT) ? NT : m_done + 1;   // depth of the node K points at

            // ---- winner and certificate -----------------------------------------------------------------------
            int wn[U], wP[U];          // winning depth and heap index 2^n + i of the winning path node
            unsigned gapmin = 0xffffffffu;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned *k_ = key[u];
                unsigned m = __vimin3_u32(k_[0], k_[1], k_[2]);
                m = __vimin3_u32(m, k_[3], k_[4]);
                m = __vimin3_u32(m, k_[5], k_[6]);
                m = __vimin3_u32(m, k_[7], k_[8]);
                m = __vimin3_u32(m, k_[9], k_[10]);
                const unsigned nm = ~m;   // key + ~m = key - m - 1: 0xffffffff for the winner itself
                unsigned g0 = 0xffffffffu, g1 = 0xffffffffu;   // two chains for instruction-level parallelism
#pragma unroll
                for (int n = 0; n <= kSmemDepth; n += 2) g0 = __viaddmin_u32(k_[n], nm, g0);
#pragma unroll
                for (int n = 1; n <= kSmemDepth; n += 2) g1 = __viaddmin_u32(k_[n], nm, g1);
                gapmin = __vimin3_u32(gapmin, g0, g1);
                wn[u] = (int)(m & 15u);
                wP[u] = (int)(K[u] >> (kd - wn[u]));
            }
            if (gapmin <= guard) {   // some coordinate is not certified (or penalties not monotone): literal search
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int r = reference_search(sTc, sPen + col, mu[u], sg[u], N);
                    wn[u] = r >> 16;
                    wP[u] = (1 << wn[u]) + (r & 0xffff);
                }
            }
            float dist[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int n = wn[u], Pn = wP[u];
                dist[u] = 0.0f;
                if (!CHECK || row + u * RP < row_end) {
                    const unsigned o = off + u * off_step;
                    // sorted index q = (2i+1) 2^(N-n) - 1 with i = Pn - 2^n:  (2 Pn + 1) 2^(N-n) - 2^(N+1) - 1
                    const int q = ((2 * Pn + 1) << (N - n)) - (2 << N) - 1;
                    if (outm & 2u) qidx_c[o] = q;
                    if (outm & 4u) level_c[o] = n;
                    if (outm & 8u) bits_c[o] = (float)n;
                    if (TOTALS || (outm & 1u)) {
                        const float zh = lds_pure((unsigned)(imad(n, 2 * kRowStrideBytes, imad(Pn, kRowStrideBytes, pbi))));
                        if (outm & 1u) zhat_c[o] = zh;
                        if (TOTALS) {
                            const float r1 = u & 1 ? r2[u / 2].y : r2[u / 2].x;
                            const float t = (zh - mu[u]) * r1;
                            acc_level += n;
                            dist[u] = t * t;
                        }
                    }
                }
            }
            if (TOTALS) {   // the float32 terms of one iteration are added in float32, then accumulated in float64
                float dsum = dist[0];
#pragma unroll
                for (int u = 1; u < U; ++u) dsum += dist[u];
                acc_dist += (double)dsum;
            }
        };

        for (; row - rsub < p1 * RP; row += U * RP, off += U * off_step) {
            cp_async_wait<kStages - 2>();        // this iteration's rows have landed
            if (VEC) __syncwarp();               // ... for every lane of the warp (the tile is staged cooperatively)
            if (row - rsub + U * RP <= full_rows) iteration(std::false_type{});
            else iteration(std::true_type{});
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

template <bool PRUNE, bool TOTALS, int NT, int OUT, bool VEC, int U, int T>
static int launch_bisect(QArgs a, int dev, int sms, cudaStream_t st) {
    constexpr int rows_per_pass = T / VBQ_GROUP;
    a.passes = (a.rows + rows_per_pass - 1) / rows_per_pass;
    a.total_units = a.passes * a.n_groups;
    long long gx = (a.total_units + U - 1) / U;
    if (gx > sms) gx = sms;
    if (gx > kMaxGrid) gx = kMaxGrid;
    const size_t smem = ((size_t)kPadEntries * VBQ_GROUP + (size_t)(kSmemDepth + 1) * VBQ_GROUP +
                         (size_t)kStages * 2 * U * T) * sizeof(float);   // the same ring size with and without VEC
    auto kern = vbq_bisect_kernel<PRUNE, TOTALS, NT, OUT, VEC, U, T>;
    VBQ_ENSURE_MAX_SMEM(kern, dev);
    kern<<<dim3((int)gx, a.n_lambda), T, smem, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    return VBQ_OK;
}

template <bool PRUNE, bool TOTALS, int NT, int U, int T>
static int launch_bisect3(const QArgs &a, int dev, int sms, cudaStream_t st) {
    // the two output sets the facade and the benchmark ask for are compiled in; anything else tests the mask at run time
    const bool vec = U == 2 && a.C % 4 == 0 && (((uintptr_t)a.mu | (uintptr_t)a.sigma) & 15) == 0;
    if (!vec) return launch_bisect<PRUNE, TOTALS, NT, -1, false, U, T>(a, dev, sms, st);
    switch (a.outm & 15u) {
        case 2u | 8u: return launch_bisect<PRUNE, TOTALS, NT, 2 | 8, true, U, T>(a, dev, sms, st);   // sorted index + code length
        case 1u | 4u: return launch_bisect<PRUNE, TOTALS, NT, 1 | 4, true, U, T>(a, dev, sms, st);   // z_hat + depth
        default: return launch_bisect<PRUNE, TOTALS, NT, -1, true, U, T>(a, dev, sms, st);
    }
}

template <bool PRUNE, int U, int T>
static int launch_bisect2(const QArgs &a, int dev, int sms, cudaStream_t st) {
    const bool tot = a.totals != nullptr;
    if (a.N == kSmemDepth)
        return tot ? launch_bisect3<PRUNE, true, kSmemDepth, U, T>(a, dev, sms, st)
                   : launch_bisect3<PRUNE, false, kSmemDepth, U, T>(a, dev, sms, st);
    return tot ? launch_bisect3<PRUNE, true, 0, U, T>(a, dev, sms, st) : launch_bisect3<PRUNE, false, 0, U, T>(a, dev, sms, st);
}

// raw code lengths (no length table, no entropy model), max_bits_per_coord <= 10; returns -1 if not applicable
int vbq_launch_quantize_bisect(const QArgs &a, int dev, int sms, cudaStream_t st) {
    if (a.len || a.em || a.N > kSmemDepth) return -1;
    const bool prune = !(a.flags & VBQ_FLAG_NO_PRUNE);
    return prune ? launch_bisect2<true, 2, 640>(a, dev, sms, st) : launch_bisect2<false, 2, 640>(a, dev, sms, st);
}
