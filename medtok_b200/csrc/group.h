// group.h -- glue kernels of the two-pass hard route (group.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "workspace.h"

namespace mtvq {

// ws.gtot (entries per chunk, from the GROUP pass) -> ws.goff / ws.gtile, the tile table and counters[CNT_TILES]; then ws.grp / ws.grank / ws.best_val / denom (the row norms the GROUP pass wrote) ->
// ws.entries sorted by chunk, placed without atomics
cudaError_t launch_group_sort(const Workspace& ws, const float* denom, int64_t n, int k, int num_sms, cudaStream_t s);

// z_q / squared error / per-code sums of the rows only the exact full scan could decide (ws.full_rows; resolve_kernel
// finishes the other near-tie rows itself), and the counts of the packed statistics (stats[c] += bins[c], rows += n)
cudaError_t launch_finish_flagged(const float* z, const float* denom, const float* w, const int64_t* idx, int64_t n, int d,
                                  int k, float* z_q, double* stats, const Workspace& ws, int num_sms, cudaStream_t s);

}  // namespace mtvq
