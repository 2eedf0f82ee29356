// soft.h -- launchers of the soft-one (top-k) row kernels (soft.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "workspace.h"

namespace mtvq {

cudaError_t launch_soft_resolve(const float* x, int64_t n, int d, int k, int topk, const Workspace& ws, int64_t* idx,
                                float* vals, float* wts, float* zq, float* out, double* sse, int num_sms,
                                cudaStream_t stream);
// fused backward: codes_n [k,d] and norms [k] are caller scratch (filled here from the raw codebook rows)
cudaError_t launch_soft_backward_fused(const float* x, int64_t n, int d, const float* codes, int k, int topk,
                                       const int64_t* idx, const float* wts, const float* g_out, const float* g_vq,
                                       const float* g_commit, float beta, int training, float* grad_x, float* grad_codes,
                                       float* codes_n, float* norms, int num_sms, cudaStream_t stream);
cudaError_t launch_norms(const float* x, int64_t n, int d, float* denom, int num_sms, cudaStream_t stream);
cudaError_t launch_soft_backward(const float* x, int64_t n, int d, const float* codes, int k, int topk,
                                 const int64_t* idx, const float* wts, const float* g_zq, float* grad_x,
                                 float* grad_codes, int num_sms, cudaStream_t stream);

}  // namespace mtvq
