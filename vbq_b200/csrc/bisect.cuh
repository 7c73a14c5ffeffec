// bisect.cuh — helpers shared by the certified-bisection kernels (quantize_bisect.cu: one lambda; sweep_bisect.cu: all
// lambdas of a call from one walk).  See quantize_bisect.cu for the argument why one candidate per depth suffices and
// for the certificate.
#pragma once
#include <type_traits>

#include "tree.cuh"

constexpr unsigned kKeyGuard = 192u;
constexpr unsigned kKeyMask = 0xfffffff0u;

// Literal restatement of the reference search for one coordinate (slow path): both bracket ends of every depth,
// IEEE float32 scores, first maximum in the order left_0..left_N, right_1..right_N.  Returns depth << 24 | index.
// sTc = this channel's column of the padded shared-memory tree, sPenc = its penalties (stride pen_stride floats).
static __device__ __noinline__ int reference_search(const float *sTc, const float *sPenc, int pen_stride, float mu,
                                                    float sg, int N) {
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
        const float npn = -sPenc[n * pen_stride];
        const float sl = score_exact(sTc[entry_of(n, il) * VBQ_GROUP], mu, sg, rs, npn);
        const float sr = score_exact(sTc[entry_of(n, ir) * VBQ_GROUP], mu, sg, rs, npn);
        if (sl > bestL) { bestL = sl; nL = n; iL = il; }
        if (sr > bestR) { bestR = sr; nR = n; iR = ir; }
        ip = 2 * ip + b;
    }
    return bestR > bestL ? (nR << 24 | iR) : (nL << 24 | iL);
}

// The same for max_bits_per_coord > 10: depths 0..10 from the shared-memory tree, deeper ones from the heap-order table
// of this channel in global memory (gT[(1 << n) - 1 + i] is code point (n, i)).
static __device__ __noinline__ int reference_search_deep(const float *sTc, const float *gT, const float *sPenc,
                                                         int pen_stride, float mu, float sg, int N) {
    const float rs = rcp_rn(sg);
    auto point = [&](int n, int i) -> float {
        return n <= kSmemDepth ? sTc[entry_of(n, i) * VBQ_GROUP] : __ldg(gT + ((1 << n) - 1 + i));
    };
    const float z0 = point(0, 0);
    float bestL = score_exact(z0, mu, sg, rs, -sPenc[0]), bestR = -CUDART_INF_F;
    int nL = 0, iL = 0, nR = 0, iR = 0;
    int ip = mu > z0 ? 1 : 0;
    for (int n = 1; n <= N; ++n) {
        const float zp = point(n, ip);
        const int b = mu > zp ? 1 : 0;
        const int fg = ip + b;
        const int il = clamp_index(fg, n, N, false), ir = clamp_index(fg, n, N, true);
        const float npn = -sPenc[n * pen_stride];
        const float sl = score_exact(il == ip ? zp : point(n, il), mu, sg, rs, npn);
        const float sr = score_exact(ir == ip ? zp : point(n, ir), mu, sg, rs, npn);
        if (sl > bestL) { bestL = sl; nL = n; iL = il; }
        if (sr > bestR) { bestR = sr; nR = n; iR = ir; }
        ip = 2 * ip + b;
    }
    return bestR > bestL ? (nR << 24 | iR) : (nL << 24 | iL);
}

// Next unclaimed tile of the CTA's current segment: lane 0 increments the shared counter, the warp gets the old value.
// atom.inc (wrap bound 2^31-1, i.e. a plain increment) on purpose: for atomicAdd / atom.add ptxas emits its
// warp-aggregation sequence (vote, find-leader, popc, lanemask: 17 instructions) around this single-lane atomic.
__device__ __forceinline__ int claim_tile(int *counter, int lane) {
    int j = 0;
    const unsigned addr = (unsigned)__cvta_generic_to_shared(counter);
    if (lane == 0) asm volatile("atom.shared.inc.u32 %0, [%1], 0x7fffffff;" : "=r"(j) : "r"(addr) : "memory");
    return __shfl_sync(0xffffffffu, j, 0);
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

// Work decomposition.  A TILE is 4 consecutive rows x the 16 channels of one group: one warp iteration (lane =
// (row parity, channel), two coordinates per thread: rows 2u + parity, u = 0, 1).  The tiles of a launch, ordered by
// (group, row), are cut into one contiguous span per CTA; inside a span the warps of the CTA CLAIM tiles from a
// shared-memory counter, kStages-1 tiles ahead of the one they compute (claim -> cp.async -> compute), so that all
// warps of a CTA finish within one tile of each other whatever the scheduler's warp priorities were.  A CTA whose
// span crosses a group boundary loads a second tree; the cut positions charge kSwitchTiles tiles for that.
constexpr int kTileRows = 4;
constexpr int kTileFloats = 2 * kTileRows * VBQ_GROUP;   // mu rows then sigma rows: [2][4][16]
constexpr int kSwitchTiles = 80;   // measured: a second segment costs a CTA about 2.8 us = 50-80 tiles

// first real tile (in group-major order) of virtual position v: every group is preceded by kSwitchTiles virtual tiles
__host__ __device__ __forceinline__ long long span_cut(long long v, long long tiles_per_group, int n_groups) {
    const long long vg = tiles_per_group + kSwitchTiles;
    long long g = v / vg;
    if (g > n_groups) g = n_groups;
    const long long o = v - g * vg;
    return g * tiles_per_group + (o > kSwitchTiles ? o - kSwitchTiles : 0);
}

