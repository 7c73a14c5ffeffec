// tma.cuh — sm_100a bulk-copy plumbing of the search kernels: tensor maps for the channel-last latent / output
// matrices (host side) and the PTX wrappers for mbarriers, cp.async.bulk.tensor (TMA) loads / stores and plain bulk
// copies (device side).  A tile of a kernel is `box_rows` consecutive rows x the 16 channels of one group = one TMA
// box of 64-byte rows; ragged edges (rows beyond the matrix, channels beyond C) are zero-filled on load and clipped on
// store by the tensor map, so the kernels carry no per-element bounds tests for memory traffic.
#pragma once
#include <cuda.h>

#include "common.h"

// (planes, rows, C) float32 / int32 array, channel-last; plane stride `plane_stride` elements.  Box = 16 channels x
// box_rows rows x 1 plane.  Needs a 16-byte aligned base and C % 4 == 0 (row pitch multiple of 16 bytes).
int vbq_make_tensor_map(CUtensorMap *out, const void *base, int C, long long rows, long long planes,
                        long long plane_stride, int box_rows);

#ifdef __CUDACC__
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// make freshly initialised barriers visible to the async proxy (TMA completes transactions on them)
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// blocks until the phase with the given parity has completed (try_wait suspends in hardware; the loop covers time-outs)
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "VBQ_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra VBQ_DONE;\n"
        "bra VBQ_WAIT;\n"
        "VBQ_DONE:\n"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}

// the same with a suspend-time hint: a warp that finds the phase incomplete sleeps in hardware for up to `ns` nanoseconds
// (or until the phase completes) instead of competing for issue slots
__device__ __forceinline__ void mbar_wait_sleepy(unsigned bar, unsigned parity, unsigned ns) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "VBQ_WAITS:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra VBQ_DONES;\n"
        "bra VBQ_WAITS;\n"
        "VBQ_DONES:\n"
        "}" ::"r"(bar), "r"(parity), "r"(ns)
        : "memory");
}

// one bounded wait: returns true if the phase has completed, false after roughly `ns` nanoseconds at the latest
__device__ __forceinline__ bool mbar_try_wait_hint(unsigned bar, unsigned parity, unsigned ns) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0;
}

// non-blocking probe of the same condition
__device__ __forceinline__ bool mbar_test(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

// generic-proxy writes to shared memory (st.shared) -> visible to the async proxy (a following TMA store)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_map(const CUtensorMap *m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<unsigned long long>(m)) : "memory");
}
// global (c, row, plane) box -> shared; completes `box bytes` on the mbarrier
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap *m, int c, int row, int plane, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(reinterpret_cast<unsigned long long>(m)), "r"(c), "r"(row), "r"(plane), "r"(bar)
        : "memory");
}
// shared -> global (c, row, plane) box, bulk async-group completion
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *m, unsigned src, int c, int row, int plane) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<unsigned long long>(m)),
                 "r"(src), "r"(c), "r"(row), "r"(plane)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the N_ most recent bulk groups have finished READING shared memory (the source may be overwritten)
template <int N_>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N_) : "memory"); }
template <int N_>
__device__ __forceinline__ void tma_store_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N_) : "memory"); }

// contiguous global -> shared bulk copy (16-byte aligned, size multiple of 16); completes `bytes` on the mbarrier
__device__ __forceinline__ void bulk_load(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// named barrier over `count` threads (id 1..15; id 0 is __syncthreads)
__device__ __forceinline__ void named_barrier(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
#endif
