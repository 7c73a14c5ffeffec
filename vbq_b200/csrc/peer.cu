// peer.cu — the only exchange step of the data-parallel path: the all-reduce of the per-lambda
// rate / distortion totals, done over NVLink peer memory by the search kernel itself instead of a separate NCCL launch.
//
// Every rank owns an INBOX of kPeerSlots x world entries (n_lambda_max x VBQ_TOTALS doubles + a sequence number) that
// all ranks of the node map (cudaIpc).  The sums of the call with sequence number q are written into entry
// [q % kPeerSlots][own rank] of EVERY rank's inbox, followed by q itself (system-scope release) — by an idle lane of the
// NEXT search kernel while that kernel runs (vbq_quantize_peer), or by a one-CTA kernel (vbq_peer_push).
// vbq_peer_collect(q) then waits, on the device, until every entry of slot q % kPeerSlots carries q and adds the world's
// sums in rank order: deterministic, no collective library, no stream dependency between consecutive search kernels.
// The reference's consumer of the sums is utils.py:546-553.
#include <string.h>

#include <new>

#include "tree.cuh"

constexpr int kPeerSlots = 8;       // calls that may be in flight before the oldest must have been collected
constexpr int kMaxPeers = 64;

struct vbq_peer_ctx {
    int rank = 0, world = 1, n_lambda_max = 1;
    long long entry = 0;            // doubles per entry: n_lambda_max * VBQ_TOTALS + 1
    double *inbox = nullptr;        // this rank's inbox (device memory owned by the context)
    double *peers[kMaxPeers] = {};  // every rank's inbox as seen from this process
    bool opened[kMaxPeers] = {};
    double **d_peers = nullptr;     // the same table in device memory
};

static void free_peer(vbq_peer_ctx *c) {
    if (!c) return;
    for (int p = 0; p < c->world; ++p)
        if (c->opened[p]) cudaIpcCloseMemHandle(c->peers[p]);
    if (c->inbox) cudaFree(c->inbox);
    if (c->d_peers) cudaFree(c->d_peers);
    delete c;
}

static size_t inbox_bytes(const vbq_peer_ctx *c) { return (size_t)kPeerSlots * c->world * c->entry * sizeof(double); }

extern "C" int vbq_peer_ctx_create(int rank, int world, int n_lambda_max, vbq_peer_ctx **out) {
    if (!out) return vbq_fail(VBQ_ERR_NULL_POINTER, "vbq_peer_ctx_create: null output");
    *out = nullptr;
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world || n_lambda_max < 1)
        return vbq_fail(VBQ_ERR_BAD_SHAPE, "vbq_peer_ctx_create: rank=%d world=%d n_lambda_max=%d", rank, world, n_lambda_max);
    vbq_peer_ctx *c = new (std::nothrow) vbq_peer_ctx;
    if (!c) return vbq_fail(VBQ_ERR_CUDA, "vbq_peer_ctx_create: out of host memory");
    c->rank = rank; c->world = world; c->n_lambda_max = n_lambda_max;
    c->entry = (long long)n_lambda_max * VBQ_TOTALS + 1;
    cudaError_t e = cudaMalloc(&c->inbox, inbox_bytes(c));
    if (e == cudaSuccess) e = cudaMemset(c->inbox, 0, inbox_bytes(c));     // sequence numbers start at 1
    if (e == cudaSuccess) e = cudaMalloc(&c->d_peers, sizeof(double *) * world);
    if (e != cudaSuccess) {
        free_peer(c);
        return vbq_fail(VBQ_ERR_CUDA, "vbq_peer_ctx_create: %s", cudaGetErrorString(e));
    }
    c->peers[rank] = c->inbox;
    if (world == 1) CUDA_TRY(cudaMemcpy(c->d_peers, c->peers, sizeof(double *), cudaMemcpyHostToDevice));
    *out = c;
    return VBQ_OK;
}

extern "C" int vbq_peer_ctx_destroy(vbq_peer_ctx *c) {
    free_peer(c);
    return VBQ_OK;
}

extern "C" int vbq_peer_ctx_handle(vbq_peer_ctx *c, unsigned char *handle64) {
    if (!c || !handle64) return vbq_fail(VBQ_ERR_NULL_POINTER, "vbq_peer_ctx_handle: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the handle travels as 64 bytes");
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, c->inbox));
    memcpy(handle64, &h, 64);
    return VBQ_OK;
}

extern "C" int vbq_peer_ctx_connect(vbq_peer_ctx *c, const unsigned char *handles) {
    if (!c || !handles) return vbq_fail(VBQ_ERR_NULL_POINTER, "vbq_peer_ctx_connect: null pointer");
    for (int p = 0; p < c->world; ++p) {
        if (p == c->rank || c->opened[p]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * (size_t)p, 64);
        void *ptr = nullptr;
        CUDA_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        c->peers[p] = (double *)ptr;
        c->opened[p] = true;
    }
    CUDA_TRY(cudaMemcpy(c->d_peers, c->peers, sizeof(double *) * c->world, cudaMemcpyHostToDevice));
    return VBQ_OK;
}

// sums of a finished call -> every rank's inbox (used when the search kernel could not do it itself)
__global__ void peer_push_kernel(const double *__restrict__ totals, int n, double *const *inbox, int world, long long off,
                                 long long flag, unsigned long long seq) {
    for (int p = 0; p < world; ++p)
        for (int i = threadIdx.x; i < n; i += blockDim.x) inbox[p][off + i] = totals[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
        for (int p = 0; p < world; ++p)
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(inbox[p] + flag), "l"(seq) : "memory");
}

// lane r waits for rank r's entry of the slot, then the entries are added in rank order
__global__ void peer_collect_kernel(const double *inbox, int world, long long entry, long long slot_off, int n,
                                    unsigned long long seq, double *__restrict__ out) {
    for (int r = threadIdx.x; r < world; r += blockDim.x) {
        const double *f = inbox + slot_off + (long long)r * entry + (entry - 1);
        unsigned long long v;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
        } while (v < seq);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < world; ++r) s += __ldcv(inbox + slot_off + (long long)r * entry + i);
        out[i] = s;
    }
}

extern "C" int vbq_peer_push(vbq_peer_ctx *peer, unsigned long long seq, int n_lambda, const double *d_totals, void *stream) {
    if (!peer || !d_totals) return vbq_fail(VBQ_ERR_NULL_POINTER, "vbq_peer_push: null pointer");
    if (n_lambda < 1 || n_lambda > peer->n_lambda_max || seq == 0)
        return vbq_fail(VBQ_ERR_BAD_SHAPE, "vbq_peer_push: n_lambda=%d seq=%llu", n_lambda, seq);
    for (int p = 0; p < peer->world; ++p)
        if (!peer->peers[p]) return vbq_fail(VBQ_ERR_NULL_POINTER, "vbq_peer_push: context not connected (rank %d)", p);
    const long long off = ((long long)(seq % kPeerSlots) * peer->world + peer->rank) * peer->entry;
    peer_push_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(d_totals, n_lambda * VBQ_TOTALS, peer->d_peers, peer->world, off,
                                                        off + peer->entry - 1, seq);
    CUDA_TRY(cudaGetLastError());
    return VBQ_OK;
}

extern "C" int vbq_quantize_peer(const float *d_mu, const float *d_sigma, long long rows, int C, const float *d_table,
                                 const float *d_packed, int N, const float *d_penalty, const float *h_penalty,
                                 const float *d_length, int n_lambda, int pen_channels, const float *d_entropy_model,
                                 float *d_zhat, int *d_qidx, int *d_level, float *d_bits, float *d_em_bits,
                                 double *d_totals, void *d_workspace, long long workspace_bytes, unsigned flags,
                                 void *stream, vbq_peer_ctx *peer, unsigned long long push_seq, const double *d_push_totals,
                                 unsigned long long collect_seq, double *d_collected) {
    if (!peer) return vbq_fail(VBQ_ERR_NULL_POINTER, "vbq_quantize_peer: needs a peer context");
    if (n_lambda > peer->n_lambda_max || (push_seq != 0) != (d_push_totals != nullptr) ||
        (collect_seq != 0) != (d_collected != nullptr))
        return vbq_fail(VBQ_ERR_BAD_SHAPE, "vbq_quantize_peer: n_lambda=%d (context: %d), push %llu, collect %llu", n_lambda,
                        peer->n_lambda_max, push_seq, collect_seq);
    for (int p = 0; p < peer->world; ++p)
        if (!peer->peers[p]) return vbq_fail(VBQ_ERR_NULL_POINTER, "vbq_quantize_peer: context not connected (rank %d)", p);
    PeerPush push;
    push.inbox = peer->d_peers;
    push.world = peer->world;
    push.off = ((long long)(push_seq % kPeerSlots) * peer->world + peer->rank) * peer->entry;
    push.flag = push.off + peer->entry - 1;
    push.seq = push_seq;
    push.src = d_push_totals;
    push.own = peer->inbox;
    push.coff = (long long)(collect_seq % kPeerSlots) * peer->world * peer->entry;
    push.entry = peer->entry;
    push.cseq = collect_seq;
    push.cout = d_collected;
    push.fused = false;
    // n_lambda == 1 and a TMA kernel: the search kernel does both while it runs; otherwise two one-CTA kernels in front
    if (n_lambda != 1 || !(push_seq || collect_seq)) {
        if (push_seq) RETURN_IF(vbq_peer_push(peer, push_seq, n_lambda, d_push_totals, stream));
        if (collect_seq) RETURN_IF(vbq_peer_collect(peer, collect_seq, n_lambda, d_collected, stream));
        return vbq_quantize_impl(d_mu, d_sigma, rows, C, d_table, d_packed, N, d_penalty, h_penalty, d_length, n_lambda,
                                 pen_channels, d_entropy_model, d_zhat, d_qidx, d_level, d_bits, d_em_bits, d_totals,
                                 d_workspace, workspace_bytes, flags, stream, nullptr);
    }
    RETURN_IF(vbq_quantize_impl(d_mu, d_sigma, rows, C, d_table, d_packed, N, d_penalty, h_penalty, d_length, n_lambda,
                                pen_channels, d_entropy_model, d_zhat, d_qidx, d_level, d_bits, d_em_bits, d_totals,
                                d_workspace, workspace_bytes, flags, stream, &push));
    if (!push.fused) {   // another kernel ran: the exchange follows it
        if (push_seq) RETURN_IF(vbq_peer_push(peer, push_seq, n_lambda, d_push_totals, stream));
        if (collect_seq) RETURN_IF(vbq_peer_collect(peer, collect_seq, n_lambda, d_collected, stream));
    }
    return VBQ_OK;
}

extern "C" int vbq_peer_collect(vbq_peer_ctx *peer, unsigned long long seq, int n_lambda, double *d_totals, void *stream) {
    if (!peer || !d_totals) return vbq_fail(VBQ_ERR_NULL_POINTER, "vbq_peer_collect: null pointer");
    if (n_lambda < 1 || n_lambda > peer->n_lambda_max || seq == 0)
        return vbq_fail(VBQ_ERR_BAD_SHAPE, "vbq_peer_collect: n_lambda=%d seq=%llu", n_lambda, seq);
    peer_collect_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(peer->inbox, peer->world, peer->entry,
                                                            (long long)(seq % kPeerSlots) * peer->world * peer->entry,
                                                            n_lambda * VBQ_TOTALS, seq, d_totals);
    CUDA_TRY(cudaGetLastError());
    return VBQ_OK;
}
