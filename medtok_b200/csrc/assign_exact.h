// assign_exact.h -- launchers of the exact (reference-order fp32) assignment kernels
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "workspace.h"

namespace mtvq {

// fp32 codebook [k,d] -> ws.b_image (fp16 operand tiles + bias planes), ws.cc, ws.meta (margin);
// flags bit 0 (soft path): rows are l2-normalised first and also written to ws.codes_n; bit 1 (MTVQ_SCORE_DOT): the
// coarse score is the inner product alone (no -||w||^2/2 term)
cudaError_t launch_codebook_prep(const float* w, int k, int d, int flags, const Workspace& ws,
                                 cudaStream_t stream);

// near-tie rows (ws.flag_rows / counters[CNT_FLAGGED]) x their collected candidates -> idx[row] (int64) and
// ws.bins_i32; rows whose candidate list is empty or overflowed go to ws.full_rows / counters[CNT_FULL]
cudaError_t launch_resolve(const float* z, int d, int norm_mode, const float* w, int k, const Workspace& ws,
                           int cand_cap, int64_t* idx, int num_sms, cudaStream_t stream);
// the same for lists of single columns and known row norms (two-pass route): eight lanes per row, four rows per warp;
// z_q / stats non-null (fused rows): also z_q, squared error and per-code sums of the rows decided here
cudaError_t launch_resolve8(const float* z, const float* denom, int d, int norm_mode, const float* w, int k,
                            const Workspace& ws, int cand_cap, int64_t* idx, int num_sms, cudaStream_t stream,
                            float* z_q = nullptr, double* stats = nullptr);

// exact first-index argmin over ALL k codes for the rows in row_list (or rows [0,n) when null) -> idx[row]
// (int64) and, when bins != nullptr, the per-code counts of those rows
cudaError_t launch_exact_scan(const float* z, int64_t n, int d, int norm_mode, const float* w, int k,
                              const float* cc, const int* row_list, const int* n_rows_dev, int64_t* idx, int* bins,
                              float* denom_out, int num_sms, cudaStream_t stream);

}  // namespace mtvq
