// rows.h -- launchers of the HBM-bound row kernels (rows.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "workspace.h"

namespace mtvq {

cudaError_t launch_hist_idx64(const int64_t* idx, int64_t n, int* bins_i32, int k, int num_sms, cudaStream_t s);
cudaError_t launch_lookup(const float* z, const float* w, const int64_t* idx, int64_t n, int d, int k, float* z_q,
                          double* sse, int num_sms, cudaStream_t s);
cudaError_t launch_counts_to_float(const int* bins, int k, float* counts, cudaStream_t s);
// window' = last `cap` of concat(window, idx) into `next`; *distinct = #distinct values in window'
// (bitmap: (n_e + 31) / 32 + 1 words of scratch)
cudaError_t launch_usage(const float* window, const int64_t* idx, int64_t cur, int cap, int n_e, float* next,
                         unsigned* bitmap, int* distinct, int num_sms, cudaStream_t s);
cudaError_t launch_gather_ste(const float* z, const float* denom, const float* w, const int64_t* idx, int64_t n, int d,
                              int k, float* z_q, double* stats, int num_sms, cudaStream_t s);
// gradient of (z_q, loss) w.r.t. the raw rows z (straight-through + commitment term through l2norm)
cudaError_t launch_hard_backward(const float* z, const float* z_q, const float* g_zq, const float* g_loss, int64_t n, int d,
                                 double beta, float* grad_z, int num_sms, cudaStream_t s);
cudaError_t launch_code_sums(const float* z, const float* denom, const int64_t* idx, int64_t n, int d, int k,
                             double* stats, const Workspace& ws, bool want_sums, int num_sms, cudaStream_t s);
cudaError_t launch_sorted_rows(const float* z, const float* denom, const float* w, const int64_t* idx, int64_t n, int d,
                               int k, float* z_q, double* stats, const Workspace& ws, int num_sms, cudaStream_t s);
cudaError_t launch_ema_finalize(float* weight, float* cluster_size, const double* stats, int k, int d, double beta,
                                double decay, int update_cs, int update_w, float* loss, cudaStream_t s);
// finalize fused with the cross-GPU exchange (reduce-scatter shape): peer_stats / peer_flags / peer_results are DEVICE
// arrays of `world` peer-mapped pointers
cudaError_t launch_ema_finalize_peers(float* weight, float* cluster_size, const void* const* peer_stats,
                                      void* const* peer_flags, void* const* peer_results, int world, int rank, unsigned epoch,
                                      int k, int d, double beta, double decay, int update_cs, int update_w, float* loss,
                                      int num_sms, cudaStream_t s);
// k-means centre update: means_k <- l2norm(embed_sum_k / bins_k) where bins_k > 0 (packed statistics as above)
cudaError_t launch_kmeans_update(float* means, const double* stats, int k, int d, cudaStream_t s);
cudaError_t launch_loss(const double* stats, int k, int d, double beta, float* loss, cudaStream_t s);

}  // namespace mtvq
