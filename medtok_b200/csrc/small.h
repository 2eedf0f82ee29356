// small.h -- launchers of the small-batch route (small.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "workspace.h"

namespace mtvq {

// true when n rows x k codes x d dims is small enough that scoring every pair exactly on the CUDA cores beats the
// fixed cost of the tensor pass (the batch sizes MedTok's training loop uses)
bool   small_route_ok(int64_t n, int d, int k, int norm_mode);

// exact top-k (soft: cand / cand_cnt get the k winners per row; codes normalised when `normalize`, published to
// codes_n_out / cc_out) or exact argmin (hard: idx_out + bins; rows with NaN distances go to full_rows / CNT_FULL for
// the exact scan).  denom_out[n] receives the rows' norms.
cudaError_t launch_small_assign(const float* x, int64_t n, int d, int norm_mode, const float* codes, int k, int normalize,
                                int topk, float* codes_n_out, float* cc_out, float* denom_out, void* scratch, int* cand,
                                int* cand_cnt, int cand_cap, int64_t* idx_out, int* bins, int* full_rows, int* counters,
                                cudaStream_t stream);

}  // namespace mtvq
