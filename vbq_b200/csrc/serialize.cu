// serialize.cu — the step after the search: the emitted symbols as a wire format for an external entropy coder
// (SURVEY §8 row f4).  The reference never serialises its indices: it counts them (quantizer.py:135-146, per-channel
// np.bincount of the sorted quantile index) and reports ideal code lengths (ipynb:452-455).  Here
//   * vbq_pack_indices / vbq_unpack_indices: fixed-width bit packing of sorted quantile indices, N+1 bits per symbol
//     (Q = 2^(N+1)-1 symbols), little-endian bit order inside little-endian 32-bit words: symbol k occupies bits
//     [k(N+1), (k+1)(N+1)) of the stream.  HBM-bound byte work: 4 B read + (N+1)/8 B written per symbol.
//   * vbq_symbol_histogram: the per-channel frequency tables of the symbols (what quantizer.py:135-146 builds with a
//     Python loop of np.bincount), privatised per 16-channel group in shared memory.
#include "common.h"

// A CTA of 256 threads packs tiles of 8192 symbols: the tile is read with coalesced 16-byte loads into shared memory
// (index s + s/32: the later stride-32 reads are conflict-free), then thread t packs symbols 32t .. 32t+31 into exactly
// B = N+1 words.
constexpr int kPackThreads = 256, kPackTile = kPackThreads * 32;
__global__ void __launch_bounds__(kPackThreads) pack_indices_kernel(const int *__restrict__ q, long long n, int B,
                                                                    unsigned *__restrict__ words, long long n_words) {
    __shared__ unsigned sh[kPackTile + kPackTile / 32];
    const long long n_tiles = (n + kPackTile - 1) / kPackTile;
    const bool aligned = ((uintptr_t)q & 15) == 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long base = tile * kPackTile;
        if (aligned && base + kPackTile <= n) {
            const int4 *src = reinterpret_cast<const int4 *>(q + base);
#pragma unroll
            for (int i = 0; i < kPackTile / (4 * kPackThreads); ++i) {
                const int v4 = i * kPackThreads + threadIdx.x;
                const int4 v = __ldg(src + v4);
                const int s = v4 * 4;           // 4 consecutive symbols never straddle a multiple of 32
                unsigned *d = sh + s + (s >> 5);
                d[0] = (unsigned)v.x; d[1] = (unsigned)v.y; d[2] = (unsigned)v.z; d[3] = (unsigned)v.w;
            }
        } else {
            for (int s = threadIdx.x; s < kPackTile; s += kPackThreads)
                sh[s + (s >> 5)] = base + s < n ? (unsigned)__ldg(q + base + s) : 0u;
        }
        __syncthreads();
        const unsigned *mine = sh + threadIdx.x * 33;
        unsigned long long acc = 0;
        int nb = 0;
        long long w = (base / 32 + threadIdx.x) * B;
#pragma unroll
        for (int s = 0; s < 32; ++s) {
            acc |= (unsigned long long)mine[s] << nb;
            nb += B;
            if (nb >= 32) {
                if (w < n_words) words[w] = (unsigned)acc;
                ++w;
                acc >>= 32;
                nb -= 32;
            }
        }
        __syncthreads();
    }
}

__global__ void unpack_indices_kernel(const unsigned *__restrict__ words, long long n_words, long long n, int B,
                                      int *__restrict__ q) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const unsigned mask = B >= 32 ? 0xffffffffu : ((1u << B) - 1u);
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const long long bit = k * B;
        const long long w = bit >> 5;
        const int sh = (int)(bit & 31);
        unsigned long long two = __ldg(words + w);
        if (sh + B > 32 && w + 1 < n_words) two |= (unsigned long long)__ldg(words + w + 1) << 32;
        q[k] = (int)((unsigned)(two >> sh) & mask);
    }
}

// counts[c][s] += #{rows r : q[r][c] == s}; a CTA owns one 16-channel group and a slice of the rows
constexpr int kHistGroup = 16;
__global__ void symbol_histogram_kernel(const int *__restrict__ q, long long rows, int C, int Q,
                                        unsigned long long *__restrict__ counts, long long rows_per_cta) {
    extern __shared__ unsigned sh_hist[];   // [16][Q]
    const int g = blockIdx.y;
    for (int k = threadIdx.x; k < kHistGroup * Q; k += blockDim.x) sh_hist[k] = 0u;
    __syncthreads();
    const long long r0 = (long long)blockIdx.x * rows_per_cta;
    const long long r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
    const int col = threadIdx.x & (kHistGroup - 1);
    const int c = g * kHistGroup + col;
    if (c < C) {
        for (long long r = r0 + (threadIdx.x >> 4); r < r1; r += blockDim.x >> 4) {
            const int s = __ldg(q + r * C + c);
            if (s >= 0 && s < Q) atomicAdd(&sh_hist[col * Q + s], 1u);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < kHistGroup * Q; k += blockDim.x) {
        const int cc = g * kHistGroup + k / Q;
        const unsigned v = sh_hist[k];
        if (v != 0u && cc < C) atomicAdd(counts + (size_t)cc * Q + (k % Q), (unsigned long long)v);
    }
}

extern "C" long long vbq_packed_index_words(long long n, int N) {
    if (n < 0 || N < 0 || N > VBQ_MAX_DEPTH) return -1;
    return (n * (N + 1) + 31) / 32;
}

extern "C" int vbq_pack_indices(const int *d_qidx, long long n, int N, unsigned *d_words, void *stream) {
    if (n < 0) return vbq_fail(VBQ_ERR_BAD_SHAPE, "vbq_pack_indices: n=%lld", n);
    RETURN_IF(vbq_check_depth(N));
    if (n == 0) return VBQ_OK;
    if (!d_qidx || !d_words) return vbq_fail(VBQ_ERR_NULL_POINTER, "vbq_pack_indices: null pointer");
    int dev = 0, sms = 0;
    RETURN_IF(vbq_current_device(&dev, &sms));
    long long grid = (n + kPackTile - 1) / kPackTile;
    if (grid > 8ll * sms) grid = 8ll * sms;
    pack_indices_kernel<<<(int)grid, kPackThreads, 0, (cudaStream_t)stream>>>(d_qidx, n, N + 1, d_words,
                                                                              vbq_packed_index_words(n, N));
    CUDA_TRY(cudaGetLastError());
    return VBQ_OK;
}

extern "C" int vbq_unpack_indices(const unsigned *d_words, long long n, int N, int *d_qidx, void *stream) {
    if (n < 0) return vbq_fail(VBQ_ERR_BAD_SHAPE, "vbq_unpack_indices: n=%lld", n);
    RETURN_IF(vbq_check_depth(N));
    if (n == 0) return VBQ_OK;
    if (!d_qidx || !d_words) return vbq_fail(VBQ_ERR_NULL_POINTER, "vbq_unpack_indices: null pointer");
    int grid;
    RETURN_IF(vbq_grid_for(n, 256, &grid));
    unpack_indices_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_words, vbq_packed_index_words(n, N), n, N + 1, d_qidx);
    CUDA_TRY(cudaGetLastError());
    return VBQ_OK;
}

extern "C" int vbq_symbol_histogram(const int *d_qidx, long long rows, int C, int N, unsigned long long *d_counts,
                                    void *stream) {
    if (rows < 0 || C < 1) return vbq_fail(VBQ_ERR_BAD_SHAPE, "vbq_symbol_histogram: rows=%lld C=%d", rows, C);
    RETURN_IF(vbq_check_depth(N));
    if (N > 10) return vbq_fail(VBQ_ERR_BAD_DEPTH, "vbq_symbol_histogram: N=%d (the shared-memory histogram holds N <= 10)", N);
    if (!d_counts || (rows > 0 && !d_qidx)) return vbq_fail(VBQ_ERR_NULL_POINTER, "vbq_symbol_histogram: null pointer");
    if (rows == 0) return VBQ_OK;
    const int Q = (1 << (N + 1)) - 1;
    int dev = 0, sms = 0;
    RETURN_IF(vbq_current_device(&dev, &sms));
    const int groups = (C + kHistGroup - 1) / kHistGroup;
    int slices = (sms + groups - 1) / groups;   // about one CTA per SM: every CTA flushes its 16 x Q bins once
    if (slices < 1) slices = 1;
    long long rows_per_cta = (rows + slices - 1) / slices;
    if (rows_per_cta < 64) rows_per_cta = 64;
    slices = (int)((rows + rows_per_cta - 1) / rows_per_cta);
    const size_t smem = (size_t)kHistGroup * Q * sizeof(unsigned);
    VBQ_ENSURE_MAX_SMEM(symbol_histogram_kernel, dev);
    symbol_histogram_kernel<<<dim3(slices, groups), 512, smem, (cudaStream_t)stream>>>(d_qidx, rows, C, Q, d_counts,
                                                                                       rows_per_cta);
    CUDA_TRY(cudaGetLastError());
    return VBQ_OK;
}
