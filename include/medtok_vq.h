/*
 * medtok_vq.h -- C ABI of the B200-native MedTok vector quantizer (libmedtok_vq.so).
 *
 * The reference (mims-harvard/MedTok) is pure PyTorch: its quantizer has no plugin / FFI layer,
 * the boundary is the nn.Module surface of MedTok/norm_ema_quantizer.py and
 * MedTok/vector_quantization_soft_one_new.py.  This header is the C ABI a maintainer binds
 * *underneath* those modules (see INTEGRATION.md for the ctypes stub); every entry point names
 * the reference lines whose ATen ops it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; tensors are row-major,
 *     contiguous, fp32 unless stated; code indices are int64 (torch.argmin / topk dtype)
 *   - the caller owns every buffer, including the workspace; the library never allocates, frees or
 *     retains device memory and keeps no global state beyond cached function attributes
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises the
 *     host; calls are re-entrant for distinct (stream, workspace) pairs
 *   - return value: 0 on success, a positive cudaError_t, or a negative MTVQ_ERR_* code;
 *     mtvq_last_error() returns a thread-local message.  No C++ exception crosses the ABI.
 *   - there is no CPU fallback and no alternate backend: sm_100a only.
 */
#ifndef MEDTOK_VQ_H
#define MEDTOK_VQ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTVQ_ABI_VERSION 2

#if defined(__GNUC__)
#define MTVQ_API __attribute__((visibility("default")))
#else
#define MTVQ_API
#endif

#define MTVQ_ERR_BAD_ARG      (-1)  /* shape / alignment / null-pointer violation detected on the host */
#define MTVQ_ERR_UNSUPPORTED  (-2)  /* e.g. D not a multiple of 16, D > 256, top-k > 8 */
#define MTVQ_ERR_WORKSPACE    (-3)  /* workspace smaller than mtvq_workspace_bytes() */
#define MTVQ_ERR_NO_DEVICE    (-4)  /* current device is not compute capability 10.x */

/* l2norm reduction order (how the REFERENCE's torch CPU path sums the squares, DESIGN.md "canonical
 * arithmetic"): unit-stride rows ([N,D,1,1] module inputs, codebooks) vs a channel-strided view
 * (b c h w inputs with h*w > 1, normalised through rearrange's permuted view). */
#define MTVQ_NORM_CONTIGUOUS 0
#define MTVQ_NORM_STRIDED    1
#define MTVQ_NORM_NONE       2      /* rows are used as given (already normalised by the caller): denom = 1 exactly */
/* or-ed into norm_mode of mtvq_hard_assign[_exact]: rank the codes by the inner product z.w alone, first index on
 * ties -- `dists = samples @ means.t(); dists.max(dim=-1).indices` of kmeans (norm_ema_quantizer.py:34,41) */
#define MTVQ_SCORE_DOT       4

/* flags for mtvq_hard_forward */
#define MTVQ_F_TRAINING        1u   /* module.training                          (norm_ema_quantizer.py:185,191) */
#define MTVQ_F_UPDATE          2u   /* module.embedding.update                  (:191) */
#define MTVQ_F_SKIP_FINALIZE   4u   /* multi-GPU: stop after the packed statistics; the caller all-reduces
                                       `stats` and then calls mtvq_ema_finalize (replaces :188,195,203) */

#define MTVQ_F_ACCUMULATE      8u   /* keep what `stats` already holds and add this call's rows to it: a pass over a
                                       vocabulary that arrives in batches (inference.py:106-115) ends in ONE
                                       all-reduce + mtvq_ema_finalize; use together with MTVQ_F_SKIP_FINALIZE */

MTVQ_API int mtvq_abi_version(void);
MTVQ_API const char * mtvq_last_error(void);

/* Bytes of caller-provided scratch needed by the calls below for up to n rows, dim d, k codes:
 * topk = 0 for the hard-path calls, topk = the soft path's top-k (1..8) for mtvq_soft_forward (a soft
 * workspace also serves the hard-path calls). */
MTVQ_API size_t mtvq_workspace_bytes(int64_t n, int d, int k, int topk);

/* Number of doubles in the packed statistics buffer: [ bins(k) | embed_sum(k*d) | sse | rows ].
 * This one buffer is what crosses NVLink (one all-reduce instead of the reference's two plus a
 * rank-local loss, norm_ema_quantizer.py:195,203,212). */
MTVQ_API size_t mtvq_stats_len(int d, int k);

/* ---- hard path: NormEMAVectorQuantizer.forward (norm_ema_quantizer.py:166-218) ------------- */

/* Whole training/eval forward of one rank's rows, stream-ordered:
 *   z        [n,d]  raw rows (b h w c order); normalised inside           (:169-171, l2norm :8-9)
 *   weight   [k,d]  codebook, updated IN PLACE when TRAINING|UPDATE       (:199-210, :136-138)
 *   cluster_size [k] module buffer, updated IN PLACE                      (:189,197, ema_inplace :11-12)
 *   z_q      [n,d]  out: zf + (w[idx] - zf), rows of the PRE-update codebook (:181,214)
 *   idx      [n]    out: int64 first-index argmin of ||zf||^2 + ||w||^2 - 2 zf.w  (:175-179)
 *   loss     [1]    out: beta * mean((z_q - zf)^2) over this call's rows   (:212)
 *   stats    [mtvq_stats_len]  out (fp64): this rank's bins | embed_sum | sse | rows  (:187,194,202)
 * With MTVQ_F_SKIP_FINALIZE the codebook / cluster_size are left untouched. */
MTVQ_API int mtvq_hard_forward(const float *z, int64_t n, int d, int norm_mode,
                      float *weight, int k, float *cluster_size,
                      double beta, double decay, unsigned flags,
                      float *z_q, int64_t *idx, float *loss, double *stats,
                      void *workspace, size_t workspace_bytes, void *stream);

/* EMA step from (all-reduced) packed statistics: cluster_size <- decay*cs + (1-decay)*bins and,
 * when update_weight, w_k <- l2norm(decay*w_k + (1-decay)*l2norm(embed_sum_k / bins_k)), rows with
 * bins_k == 0 keeping w_k (:197-210).  Also (re)computes loss = beta * sse / (rows * d) from the
 * packed totals when loss != NULL (global-batch loss; pass NULL to keep the rank-local one). */
MTVQ_API int mtvq_ema_finalize(float *weight, float *cluster_size, const double *stats, int k, int d,
                      double beta, double decay, int update_cluster_size, int update_weight,
                      float *loss, void *stream);

/* mtvq_ema_finalize fused with the cross-GPU exchange of the statistics (replaces the all-reduces of
 * norm_ema_quantizer.py:188,195,203 AND the update :197-210 with ONE kernel per rank, reduce-scatter shape).
 * peer_stats / peer_flags / peer_results are DEVICE arrays of `world` pointers into buffers mapped into this process
 * (CUDA IPC / symmetric memory over NVLink):
 *   peer_stats[r]    rank r's packed statistics of this step (mtvq_hard_forward with MTVQ_F_SKIP_FINALIZE)
 *   peer_flags[r]    rank r's flag array, >= 129 uint32, zero before the first step
 *   peer_results[r]  rank r's result buffer, k*d + k floats
 * `epoch` must grow by one per call and be the same on every rank; the caller alternates between two statistics buffers
 * (epoch parity) so that a rank running ahead never overwrites what a peer still reads.  Each rank owns a contiguous slice
 * of the codes: it waits until every peer has announced its statistics (one releasing system-scope store per peer), sums
 * the W partial statistics OF ITS SLICE in rank order in fp64 straight out of the peers' memory, applies the EMA step,
 * pushes the new rows / cluster sizes into every peer's result buffer, announces the slice, and copies the other
 * slices out of its own result buffer as their owners announce them.  Every row is computed once and copied, so all
 * ranks end up with bit-identical weight / cluster_size, and a rank moves ~(statistics + codebook) bytes over NVLink
 * instead of world x statistics.  loss (may be NULL) receives beta * sum(sse) / (sum(rows) * d).  A peer that never
 * arrives makes the kernel trap after MEDTOK_EXCHANGE_TIMEOUT_S (default 600) seconds instead of hanging. */
MTVQ_API int mtvq_ema_finalize_peers(float *weight, float *cluster_size, const void *const *peer_stats,
                            void *const *peer_flags, void *const *peer_results, int world, int rank, unsigned epoch,
                            int k, int d, double beta, double decay, int update_cluster_size, int update_weight,
                            float *loss, void *stream);

/* Backward of mtvq_hard_forward w.r.t. z (the codebook gets no gradient, norm_ema_quantizer.py:212-214):
 *   g = g_zq + g_loss * 2 beta / (n d) * (zf - z_q);   grad_z = (g - zf (zf . g)) / max(||z||, 1e-12)
 * z: the raw rows of the forward (contiguous [n,d]); z_q: its output; g_zq [n,d] may be NULL (no gradient through the
 * straight-through output); g_loss [1] (device) may be NULL.  fp32. */
MTVQ_API int mtvq_hard_backward(const float *z, const float *z_q, const float *g_zq, const float *g_loss,
                       int64_t n, int d, double beta, float *grad_z, void *stream);

/* Pieces of the forward, exposed for tests, k-means and the lookup-only VectorQuantizer
 * (norm_ema_quantizer.py:222-268). */

/* idx[n] = first-index argmin_k d(zf_n, w_k); also writes denom[n] = max(||z_n||, 1e-12)
 * (denom may be NULL).  Tensor-core coarse pass + exact fp32 re-scoring of every near-tie. */
MTVQ_API int mtvq_hard_assign(const float *z, int64_t n, int d, int norm_mode, const float *weight, int k,
                     int64_t *idx, float *denom, void *workspace, size_t workspace_bytes, void *stream);

/* Same result through the exact fp32 CUDA-core kernel only (reference order, no tensor cores);
 * the slow cross-check used by tests and by the near-tie overflow fallback. */
MTVQ_API int mtvq_hard_assign_exact(const float *z, int64_t n, int d, int norm_mode, const float *weight, int k,
                           int64_t *idx, float *denom, void *workspace, size_t workspace_bytes, void *stream);

/* z_q = zf + (w[idx] - zf) and sse += sum (w[idx]-zf)^2 (fp64) into `stats` (F.embedding :117-118,
 * mse/straight-through :212-214).  denom as written by mtvq_hard_assign. */
MTVQ_API int mtvq_gather_ste(const float *z, const float *denom, const float *weight, const int64_t *idx,
                    int64_t n, int d, int k, float *z_q, double *stats, void *stream);

/* bins[k] += #{n: idx_n = k} (one_hot/sum :183-194 without the N x K one-hot), rows += n, and
 * embed_sum[k,d] += sum_{n: idx_n = k} zf_n (replaces the second N x K GEMM, :202): warp-aggregated
 * histogram, counting sort of the rows by code, fp64 segmented sums. */
MTVQ_API int mtvq_code_sums(const float *z, const float *denom, const int64_t *idx, int64_t n, int d, int k,
                   double *stats, void *workspace, size_t workspace_bytes, void *stream);

/* k-means centre update (kmeans, norm_ema_quantizer.py:43-54, cosine branch) from packed statistics of
 * mtvq_code_sums: means_k <- l2norm(embed_sum_k / bins_k); a cluster without members keeps its centre. */
MTVQ_API int mtvq_kmeans_update(float *means, const double *stats, int k, int d, void *stream);

/* Lookup-only quantizer, VectorQuantizer.forward of norm_ema_quantizer.py:222-268: z_q[n,d] = weight[idx]
 * (F.embedding :117-118,253).  z (raw rows, b h w c order) may be NULL; otherwise *sse (fp64, may be NULL) +=
 * sum (z_q - z)^2 for the loss of :262-264.  counts (fp32 [k], may be NULL) receives the per-code row counts the
 * eval branch feeds to the cluster-size EMA (:257-261); it needs a workspace of mtvq_workspace_bytes(n, d, k, 0). */
MTVQ_API int mtvq_lookup(const float *z, const float *weight, const int64_t *idx, int64_t n, int d, int k,
                float *z_q, double *sse, float *counts, void *workspace, size_t workspace_bytes, void *stream);

/* ---- soft path: VectorQuantizer.specific_embedding / shared core
 *      (vector_quantization_soft_one_new.py:120-125,147-185,187-217) ------------------------- */

/* One soft quantisation of n projected rows x [n,d] against `k` RAW codebook rows `codes`
 * (the active slice; normalised inside, :148,198-200):
 *   idx [n,topk] int64 ascending-distance code indices (slice-relative)        (:157,159,203)
 *   vals[n,topk] the distances, wts[n,topk] = softmax(-vals)                    (:158,160,204)
 *   zq  [n,d]    sum_j wts_j * cn[idx_j]                                        (:164-165,205)
 *   out [n,d]    x + (zq - x)                                                   (:181-182,214)
 *   sse [1] fp64 += sum (zq - x)^2  (vq_loss = sse/(n d), commit = beta*vq_loss, :169-173,208-209) */
MTVQ_API int mtvq_soft_forward(const float *x, int64_t n, int d, const float *codes, int k, int topk,
                      int64_t *idx, float *vals, float *wts, float *zq, float *out, double *sse,
                      void *workspace, size_t workspace_bytes, void *stream);

/* Backward of mtvq_soft_forward w.r.t. x and the raw codebook rows (autograd of :196-214, maths in
 * DESIGN.md): g_zq is dL/dzq [n,d] (from the vq loss and anything else consuming zq); writes
 * grad_x [n,d] (through the softmax weights and normalize) and ACCUMULATES grad_codes [k,d]. */
MTVQ_API int mtvq_soft_backward(const float *x, int64_t n, int d, const float *codes, int k, int topk,
                       const int64_t *idx, const float *wts, const float *g_zq,
                       float *grad_x, float *grad_codes, void *workspace, size_t workspace_bytes,
                       void *stream);

/* The whole backward of one soft quantisation in one pass (what autograd does for :196-214 in the reference): g_out =
 * dL/d(returned z_q) [n,d] (may be NULL), g_vq / g_commit = DEVICE scalars dL/d(vq_loss), dL/d(commit_loss) (may be NULL);
 * writes grad_x [n,d] = g_out - g_commit*beta*diff + (soft path through the weights and normalize), diff = (zq - x)*2/(n d),
 * and ACCUMULATES the raw codebook rows' gradient into grad_codes [k,d].  zq is recomputed from idx / wts.  With
 * training == 0 the losses are constants (:211-212): grad_x = g_out.  workspace: 4*(k*d + pad128(k)) bytes. */
MTVQ_API int mtvq_soft_backward_fused(const float *x, int64_t n, int d, const float *codes, int k, int topk,
                             const int64_t *idx, const float *wts, const float *g_out, const float *g_vq,
                             const float *g_commit, double beta, int training, float *grad_x, float *grad_codes,
                             void *workspace, size_t workspace_bytes, void *stream);

/* VectorQuantizer.codebook_usage (vector_quantization_soft_one_new.py:219-236): next[cap] = the last `cap`
 * entries of concat(window[cap], float(idx[count])) and *distinct (device int) = number of distinct values in it
 * (the reference's len(torch.unique(window)); usage = distinct / n_e on the host).  window and next must not
 * overlap; scratch needs 4 * ((n_e + 31) / 32 + 1) bytes. */
MTVQ_API int mtvq_codebook_usage(const float *window, const int64_t *idx, int64_t count, int cap, int n_e, float *next,
                        void *scratch, size_t scratch_bytes, int *distinct, void *stream);

/* ---- diagnostics (not part of the drop-in surface) ---------------------------------------- */

/* counters of the last assignment run on this workspace (same n, d, k, topk as that call), copied to
 * HOST after a stream sync: out_host[0] = rows sent to exact re-scoring (near-ties), [1] = rows that
 * overflowed the candidate list and took the full exact scan, [2] = candidates re-scored in total,
 * [3] = kernels/memsets this thread enqueued since the previous call of this function. */
MTVQ_API int mtvq_debug_counters(const void *workspace, int64_t n, int d, int k, int topk, int64_t *out_host,
                        void *stream);

/* raw tensor-core scores of the first `rows` rows against all codes, [rows, kpad] fp32
 * (coarse score = zf.w - ||w||^2/2 from fp16 operands); used by tests to validate the MMA path. */
MTVQ_API int mtvq_debug_coarse_scores(const float *z, int64_t n, int d, int norm_mode, const float *weight, int k,
                             float *scores, int64_t rows, int *kpad_host,
                             void *workspace, size_t workspace_bytes, void *stream);

/* Profiling hook (bench.py): the next mtvq_hard_forward / mtvq_hard_assign issued by this host thread
 * records the two caller-owned cudaEvent_t handles on `stream` immediately before and after the launch
 * of the dominant kernel (the tcgen05 coarse-score pass over all rows).  One-shot: cleared after use;
 * pass NULLs to cancel.  Never synchronises. */
MTVQ_API int mtvq_debug_set_events(void *event_before, void *event_after);

/* Pipeline trace (DESIGN.md "slot loop"): while set (device buffer of 512*8 int64, or NULL to stop), the
 * coarse kernels of this host thread record clock64() stamps of accumulator slot 0 on CTA 0:
 * [tile*8 + 0] issuer got the slot, +1 issuer done issuing, +2 epilogue saw the tile, +3 slot released,
 * +4 tile reduced. */
MTVQ_API int mtvq_debug_set_trace(void *trace_dev);

/* route of mtvq_soft_forward on this host thread: 0 = production (one tensor pass + two-pass fallback for overflow
 * rows), 1 = two tensor passes for every row (k-th score bound, then collect), 2 = no tensor pass, every row scored
 * against every code by the exact CUDA-core scan (cross-check; slow).  All three give bit-identical results. */
MTVQ_API int mtvq_debug_set_soft_route(int route);

/* route of the hard assignment (mtvq_hard_assign / mtvq_hard_forward) on this host thread: -1 or 0 = automatic (the
 * MTVQ_HARD_ROUTE environment variable, else by shape), 1 = always the single tensor pass with score records, 2 = the
 * two-pass route (record-free pass, rows sorted by their best 128-code chunk, per-chunk re-scoring fused with the row
 * work) whenever the shape allows it (D <= 64, at least two and at most 8192 chunks of 128 codes, enough rows for
 * four-row-tile blocks).  Both routes give bit-identical indices; the per-code sums differ in summation order only. */
MTVQ_API int mtvq_debug_set_hard_route(int route);
MTVQ_API int mtvq_debug_get_hard_route(void);   /* the value last set on this host thread (-1 = automatic) */
/* the route mtvq_hard_assign / mtvq_hard_forward take for this shape on the current device and host thread: 1 = single
 * tensor pass, 2 = two-pass; 0 = the shape is not supported at all (see mtvq_last_error) */
MTVQ_API int mtvq_hard_route(int64_t n, int d, int k);

#ifdef __cplusplus
}
#endif
#endif /* MEDTOK_VQ_H */
