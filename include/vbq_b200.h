/*
 * vbq_b200 — C ABI of the B200-native VBQ rate-distortion quantization path.
 *
 * The reference (mandt-lab/vbq) has no FFI layer: its boundary is the Python surface of
 * img-compression/quantizer.py, learned_prior.py, vae_models.py, utils.py and one notebook cell.  Each entry
 * point below names the reference interface it replaces (paths relative to the reference root).  All pointers
 * prefixed d_ are DEVICE pointers owned by the caller; the library never allocates, keeps no global state and
 * orders all work on the given stream (a cudaStream_t passed as void*, NULL = legacy default stream).  Every
 * function returns a VBQ_* status (0 = ok) and never throws; vbq_last_error() gives the thread's last message.
 *
 * Data layout
 *   latents        (rows, C) float32, channel-last, row-major  — quantizer.py:196-197 reshape(-1, C)
 *   code points    (C, Q) float32, Q = 2^(N+1)-1, HEAP order: entry h = 2^n-1+i is F_c^-1((i+1/2)2^-n)
 *                  — quantizer.py:30-36 `all_code_points`; the notebook's `codepoints` (ipynb:383-390) is one row
 *   packed table   ceil(C/16) groups x 2069 entries x 16 channels, the shared-memory image of a 16-channel group:
 *                  bit depths 0..10, each stored as [pad, 2^n points, pad] so bracket ends need no clamping;
 *                  followed by ceil(C/16) "walk trees" of 40992 floats (the same code points in heap order, scaled by
 *                  2^24, bit depths 1..7 twice: the shared-memory image of vbq_bisect_tma_kernel); made by
 *                  vbq_pack_code_points, vbq_packed_table_floats(C, N) floats in all
 *   prior params   (C, 43) float32: for layer k=0..3: matrix (d_{k+1} x d_k row-major), bias (d_{k+1}),
 *                  factor (d_{k+1}, k<3), dims (1,3,3,3,1), already softplus/tanh-transformed
 *                  — learned_prior.py:30-58 `_matrices`, `_biases`, `_factors`
 *   penalties      (n_lambda, pen_channels, N+1) float32 = fl(float32(lambda) * float32(len[c][n])),
 *                  pen_channels = 1 (raw lengths len = n, quantizer.py:166-169) or C (corrected lengths
 *                  n + R_lambda[c][n], quantizer.py:170-180); utils.py:393-396 forms exactly this product
 *   sorted index   q = (2i+1) 2^(N-n) - 1, the rank of code point (n,i) among the channel's Q ascending code
 *                  points — what quantizer.py:135,223 compute by searchsorted(code_points_by_channel, z_hat)
 */
#ifndef VBQ_B200_H
#define VBQ_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define VBQ_VERSION 100

enum {
    VBQ_OK = 0,
    VBQ_ERR_NULL_POINTER = 1,
    VBQ_ERR_BAD_SHAPE = 2,     /* rows < 0, C < 1, n_lambda < 1, pen_channels not in {1, C} */
    VBQ_ERR_BAD_DEPTH = 3,     /* N outside [0, VBQ_MAX_DEPTH] */
    VBQ_ERR_BAD_FLAGS = 4,
    VBQ_ERR_WORKSPACE = 5,     /* workspace missing or too small */
    VBQ_ERR_CUDA = 6,          /* a CUDA runtime call failed; see vbq_last_error() */
    VBQ_ERR_MISALIGNED = 7
};

#define VBQ_MAX_DEPTH 20
#define VBQ_PRIOR_PARAMS 43
#define VBQ_GROUP 16             /* channels interleaved per shared-memory group */
#define VBQ_SMEM_LEVELS 11       /* bit depths 0..10 of a group live in shared memory */
#define VBQ_TOTALS 4             /* per lambda: sum raw depth n, sum code length, sum entropy-model bits, sum (z_hat-mu)^2/(2 sigma^2) */

/* flags of vbq_quantize */
#define VBQ_FLAG_LOGVAR   1u     /* d_sigma holds log-variances; sigma = sqrt(exp(logvar)) (quantizer.py:193-198) */
#define VBQ_FLAG_NO_PRUNE 2u     /* visit every bit depth even when deeper levels provably cannot win */
#define VBQ_FLAG_RESERVE_SM 64u       /* launch one CTA fewer than there are SMs, so that a concurrent kernel on another
                                         stream (e.g. the NCCL all-reduce of the previous call's totals) finds a free SM */
#define VBQ_FLAG_WORKSPACE_ZEROED 256u /* the caller guarantees that the first 256 bytes per 64 lambdas of d_workspace (the
                                         ticket counters) are zero: true after a cudaMemset at allocation and after every
                                         completed call, which leaves them zero again.  Saves the per-call memset node
                                         (about 5 us of stream time on a B200). */
#define VBQ_FLAG_TABLE_STABLE 1024u     /* the caller vouches that d_packed was completely written before the PREVIOUS operation
                                         on `stream` was enqueued (e.g. it synchronized after vbq_pack_code_points): the search
                                         kernel then fetches its code points while the stream's previous kernel still drains
                                         (programmatic dependent launch) instead of after it */
#define VBQ_FLAG_NEIGHBOUR_EVERY_DEPTH 2048u /* arbitrary penalties (vbq_bisect_tma_kernel both-ends variant, vbq_both_sweep_kernel): score the second
                                          bracket end at every depth, not only where the penalties allow it to win (same
                                          results; diagnostics and tests) */
#define VBQ_FLAG_NO_TMA 512u           /* single lambda: stage the latents with per-warp cp.async (vbq_bisect_kernel) instead
                                         of the TMA pipeline (vbq_bisect_tma_kernel); same results, for comparison */
#define VBQ_FLAG_BRACKET_WALK 128u     /* single lambda: use the nearer-bracket-end walk (strict mode) even where the
                                         certified bisection kernel applies (same results; for comparison) */
#define VBQ_FLAG_REFERENCE_WALK 32u   /* score both bracket ends of every depth (the slower, literal formulation; same results) */
#define VBQ_FLAG_NO_SWEEP 16u         /* walk the tree once per lambda instead of once for all lambdas */
#define VBQ_FLAG_ACCUMULATE_TOTALS 8u /* add this call's sums to d_totals instead of overwriting them */
#define VBQ_FLAG_FAST     4u     /* score with d*d*(0.5/sigma^2) + pen on the nearer bracket end only (not bit-faithful
                                    to utils.py:318-320 rounding; differs only inside float32 rounding ties) */

int vbq_version(void);
const char *vbq_status_string(int status);
const char *vbq_last_error(void);

/* ---- prior models -------------------------------------------------------------------------------------- */

/* BMSHJ2018Prior.cdf (learned_prior.py:109-148): d_x, d_cdf (rows, C) float32 channel-last. */
int vbq_learned_cdf(const float *d_params, int C, const float *d_x, long long rows, float *d_cdf, void *stream);

/* BMSHJ2018Prior.inverse_cdf (learned_prior.py:173-218): d_xi (rows, C) float64 in (0,1) -> d_z (rows, C)
 * float32.  Solves logits_c(z) = logit(xi) in float64 by bracketed Newton and rounds to float32, so the
 * result is a pure function of (channel parameters, xi). */
int vbq_learned_inverse_cdf(const float *d_params, int C, const double *d_xi, long long rows, float *d_z,
                            void *stream);

/* StandardGaussianPrior / FactoredGaussianPrior.inverse_cdf (vae_models.py:23-25, :40-43) and the notebook's
 * norm.ppf(xi, scale=empirical_std) (ipynb:385): d_z = mean[c] + std[c] * ndtri(xi), float64.
 * d_mean / d_std may be NULL (0 / 1). */
int vbq_gaussian_inverse_cdf(const double *d_mean, const double *d_std, int C, const double *d_xi,
                             long long rows, double *d_z, void *stream);

/* ---- code-point tables (ChannelwisePriorCDFQuantizer.build_code_points, quantizer.py:25-63) ------------ */

int vbq_build_code_points_learned(const float *d_params, int C, int N, float *d_table, void *stream);
int vbq_build_code_points_gaussian(const double *d_mean, const double *d_std, int C, int N, float *d_table,
                                   void *stream);
long long vbq_packed_table_floats(int C, int N);
int vbq_pack_code_points(const float *d_table, int C, int N, float *d_packed, void *stream);

/* ---- the hot path -------------------------------------------------------------------------------------- */

long long vbq_quantize_workspace_bytes(int n_lambda);

/* Replaces ChannelwisePriorCDFQuantizer.get_all_N_bit_intervals + compress_batch_channel_latents
 * (quantizer.py:65-80, :156-188), utils.curry_normal_logpdf + utils.batch_quantize_indep_dims
 * (utils.py:307-327, :363-423) and the sorted-index / entropy-model gather of compress_latents
 * (quantizer.py:223-228) for n_lambda rate-distortion trade-offs at once.  For every coordinate it finds the
 * bracketing code points of mu at each bit depth and returns the first maximiser (candidate order left_0..left_N,
 * right_1..right_N) of  -0.5*((z-mu)/sigma)^2 - penalty[lambda][c][n]  in float32 arithmetic.
 *
 * Outputs (any may be NULL), each (n_lambda, rows, C):
 *   d_zhat      float32  chosen code point                       (Z_hat_dict[lamb])
 *   d_qidx      int32    its sorted quantile index               (I / qidx, quantizer.py:135,223)
 *   d_level     int32    its bit depth n                         (raw_num_bits in raw-length mode)
 *   d_bits      float32  d_length[lambda][c][n] (n if d_length is NULL)  (num_bits_dict[lamb])
 *   d_em_bits   float32  d_entropy_model[lambda][c][q]           (num_bits, quantizer.py:226-228)
 * d_totals (n_lambda, VBQ_TOTALS) float64 receives the per-lambda sums (deterministic reduction order).
 * d_length has the shape of d_penalty; d_entropy_model is (n_lambda, C, Q) float32.
 * While the call runs, an entropy-model output plane may temporarily hold integer heap indices (em_gather_kernel replaces
 * them before the call's work on the stream completes); code lengths above 512 bits and entropy-model entries above 512
 * bits per coordinate are outside the exact range of the integer sums of the multi-lambda kernels (lengths fall back to a
 * float32 sum, entropy-model sums saturate). */
int vbq_quantize(const float *d_mu, const float *d_sigma, long long rows, int C,
                 const float *d_table, const float *d_packed, int N,
                 const float *d_penalty, const float *d_length, int n_lambda, int pen_channels,
                 const float *d_entropy_model,
                 float *d_zhat, int *d_qidx, int *d_level, float *d_bits, float *d_em_bits,
                 double *d_totals, void *d_workspace, long long workspace_bytes,
                 unsigned flags, void *stream);

/* vbq_quantize with an additional HOST copy of the penalties (same shape and values as d_penalty; may be NULL).  The
 * reference forms lambda * code_length on the host (utils.py:393-396), so its caller always has them; when they do
 * not depend on the channel (raw code lengths, quantizer.py:166-169) and n_lambda == 1 the search kernel takes them
 * as launch constants instead of spending a register per bit depth on them.  Results are identical. */
int vbq_quantize_hp(const float *d_mu, const float *d_sigma, long long rows, int C,
                    const float *d_table, const float *d_packed, int N,
                    const float *d_penalty, const float *h_penalty, const float *d_length, int n_lambda,
                    int pen_channels, const float *d_entropy_model,
                    float *d_zhat, int *d_qidx, int *d_level, float *d_bits, float *d_em_bits,
                    double *d_totals, void *d_workspace, long long workspace_bytes,
                    unsigned flags, void *stream);

/* ---- the exchange step of the data-parallel path (the sums are consumed by utils.py:546-553) ------------ */

/* One context per rank (one process per GPU).  It owns this rank's INBOX in device memory; the 64-byte handle of
 * vbq_peer_ctx_handle travels to the other ranks of the node by any means (e.g. torch.distributed.all_gather), and
 * vbq_peer_ctx_connect(handles = world x 64 bytes, in rank order) maps their inboxes (cudaIpc, NVLink peer access). */
typedef struct vbq_peer_ctx vbq_peer_ctx;
int vbq_peer_ctx_create(int rank, int world, int n_lambda_max, vbq_peer_ctx **out);
int vbq_peer_ctx_handle(vbq_peer_ctx *ctx, unsigned char *handle64);
int vbq_peer_ctx_connect(vbq_peer_ctx *ctx, const unsigned char *handles);
int vbq_peer_ctx_destroy(vbq_peer_ctx *ctx);

/* The sums of the call with sequence number `seq` (>= 1, chosen by the caller: the same on every rank for the same call,
 * increasing by one per call) are DELIVERED by writing them into every rank's inbox (peer stores over NVLink + a
 * system-scope release of the sequence number) and COLLECTED by waiting, on the device, for every rank's entry and adding
 * the entries in rank order (deterministic) into d_totals (n_lambda, VBQ_TOTALS).  At most 8 calls may be delivered but
 * not yet collected.  vbq_peer_push / vbq_peer_collect are one-CTA kernels.  vbq_quantize_peer is vbq_quantize_hp that
 * also delivers the COMPLETED totals of an earlier call (`d_push_totals`, sequence number push_seq; 0 / NULL = none) and
 * collects a still earlier one (collect_seq -> d_collected; 0 / NULL = none): with n_lambda == 1 an idle lane of the
 * search kernel does both while the search runs, so a sequence of calls exchanges its totals once per call without any
 * other stream operation and without a collective library. */
int vbq_peer_push(vbq_peer_ctx *ctx, unsigned long long seq, int n_lambda, const double *d_totals, void *stream);
int vbq_peer_collect(vbq_peer_ctx *ctx, unsigned long long seq, int n_lambda, double *d_totals, void *stream);
int vbq_quantize_peer(const float *d_mu, const float *d_sigma, long long rows, int C,
                      const float *d_table, const float *d_packed, int N,
                      const float *d_penalty, const float *h_penalty, const float *d_length, int n_lambda,
                      int pen_channels, const float *d_entropy_model,
                      float *d_zhat, int *d_qidx, int *d_level, float *d_bits, float *d_em_bits,
                      double *d_totals, void *d_workspace, long long workspace_bytes,
                      unsigned flags, void *stream, vbq_peer_ctx *peer, unsigned long long push_seq,
                      const double *d_push_totals, unsigned long long collect_seq, double *d_collected);

/* ---- the hot path for host-resident latents ---------------------------------------------------------------- */

/* output selection bits of vbq_host_ctx_create */
#define VBQ_OUT_ZHAT    1u
#define VBQ_OUT_QIDX    2u
#define VBQ_OUT_LEVEL   4u
#define VBQ_OUT_BITS    8u
#define VBQ_OUT_EM_BITS 16u
#define VBQ_OUT_TOTALS  32u

/* A context owns three device staging slots of `chunk_rows` rows (inputs plus the selected outputs for n_lambda
 * trade-offs), three streams and the totals workspace.  It is the only object of this library that allocates;
 * one context serves one calling thread at a time. */
typedef struct vbq_host_ctx vbq_host_ctx;
int vbq_host_ctx_create(int C, int N, int n_lambda, long long chunk_rows, unsigned outputs, vbq_host_ctx **out);
int vbq_host_ctx_destroy(vbq_host_ctx *ctx);

/* vbq_quantize for HOST arrays (pinned memory recommended): the reference's host-side call
 * compress_batch_channel_latents(batch_means, batch_stds, lambs) on NumPy inputs (quantizer.py:156-188).
 * h_mu / h_sigma are (rows, C); outputs are (n_lambda, rows, C) host arrays (NULL to skip; must have been
 * selected at context creation); h_totals is (n_lambda, VBQ_TOTALS).  Tables, penalties and entropy models are
 * device pointers as in vbq_quantize.  Rows are processed in chunks whose upload, kernel and download overlap on
 * three streams of the context; the call returns when all results are in host memory.  The d_* tables must be
 * complete when the call is made (it does not order itself after other streams). */
int vbq_quantize_host(vbq_host_ctx *ctx, const float *h_mu, const float *h_sigma, long long rows,
                      const float *d_table, const float *d_packed, const float *d_penalty, const float *d_length,
                      int pen_channels, const float *d_entropy_model,
                      float *h_zhat, int *h_qidx, int *h_level, float *h_bits, float *h_em_bits, double *h_totals,
                      unsigned flags);

/* ---- word embeddings: the notebook's float64 search (ipynb:429-443) ------------------------------------------- */

/* compress_coordinates for one shared prior: d_mu, d_sigma (n) float32, d_codepoints (Q) float64 heap order
 * (ipynb:383-390), d_lengths (N+1) float64 code length of each bit depth, N <= 12.  Minimises
 * (c-mu)^2 + (2 beta) sigma^2 len(c) with the notebook's float64 roundings, first minimum in heap order.
 * pen_f32 != 0: (2 beta) sigma^2 is formed in float32 first (NumPy 1.17, or a Python-float beta under NumPy 2).
 * Outputs (n), any may be NULL: the optimum cast to float32, its heap index, its bit depth. */
int vbq_compress_coordinates_f64(const float *d_mu, const float *d_sigma, long long n, const double *d_codepoints,
                                 int N, const double *d_lengths, double beta, int pen_f32, float *d_optima,
                                 int *d_heap_index, int *d_level, void *stream);

/* ---- after the search: symbols for an external entropy coder ------------------------------------- */

/* The reference counts its sorted quantile indices per channel (quantizer.py:135-146, np.bincount in a Python loop)
 * and reports ideal code lengths (ipynb:452-455); it has no wire format.  vbq_symbol_histogram adds, for every
 * channel c and symbol s < Q = 2^(N+1)-1, the number of rows with d_qidx[r][c] == s to d_counts (C, Q) uint64
 * (the caller zeroes it; N <= 10).  vbq_pack_indices writes n symbols as a bit stream of N+1 bits per symbol
 * (symbol k = bits [k(N+1), (k+1)(N+1)), least significant bit first, in little-endian 32-bit words;
 * vbq_packed_index_words(n, N) words); vbq_unpack_indices is its inverse. */
long long vbq_packed_index_words(long long n, int N);
int vbq_pack_indices(const int *d_qidx, long long n, int N, unsigned *d_words, void *stream);
int vbq_unpack_indices(const unsigned *d_words, long long n, int N, int *d_qidx, void *stream);
int vbq_symbol_histogram(const int *d_qidx, long long rows, int C, int N, unsigned long long *d_counts,
                         void *stream);

/* ---- stand-alone operator forms ------------------------------------------------------------------------------- */

/* ChannelwisePriorCDFQuantizer.get_all_N_bit_intervals (quantizer.py:65-80): d_mu (rows, C) -> d_left, d_right
 * (C, N+1, rows): at every bit depth the largest code point below mu and the smallest one >= mu, with the edge
 * semantics of the reference's padded search grids. */
int vbq_intervals(const float *d_mu, long long rows, int C, const float *d_table, int N, float *d_left,
                  float *d_right, void *stream);

/* utils.batch_quantize_indep_dims (utils.py:363-423) on explicit candidates: d_P (M, BK) float32 code points,
 * d_L (M, BK) or (n_lambda, M, BK) code lengths (int32, or float32 when l_is_float), scores
 * fun_P - fl(lambda*L) with fun_P = d_funP (M, BK) if given, else -0.5*((P-loc)/scale)^2 (utils.py:318-320) from
 * d_loc, d_scale (BK).  Outputs (n_lambda, BK): chosen code point, its code length (dtype of d_L), optional index. */
int vbq_argmax_candidates(const float *d_P, const void *d_L, int l_is_float, int l_per_lambda, const float *d_funP,
                          const float *d_loc, const float *d_scale, const float *d_lambs, int n_lambda, int M,
                          long long BK, float *d_zhat, void *d_bits, int *d_index, void *stream);

/* Self-test helper: out[i] = the kernel's division a[i]/b[i] (reciprocal + FMA correction) so that tests can
 * compare it with IEEE division bit for bit. */
int vbq_selftest_divide(const float *d_a, const float *d_b, long long n, float *d_out, void *stream);

/* Self-test helper (host only): out[b] .. out[b+1] is the range of 4-row x 16-channel tiles, in (channel group, row)
 * order, that CTA b of a `grid`-CTA launch of the bisection kernels processes; out has grid+1 entries. */
int vbq_selftest_span_cuts(long long rows, int C, int grid, long long *out);

#ifdef __cplusplus
}
#endif
#endif /* VBQ_B200_H */
