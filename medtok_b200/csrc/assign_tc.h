// assign_tc.h -- parameters / launcher of the tcgen05 coarse-score kernel (assign_tc.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "workspace.h"

namespace mtvq {

enum CoarseMode { MODE_ARGMAX = 0, MODE_THRESH = 1, MODE_COLLECT = 2, MODE_DUMP = 3, MODE_TOPK = 4 };

// MODE_TOPK keeps the fp16 scores of every thread's best granules in an L2-resident scratch: per CTA 8 slots x 4 x
// 1024 records x 16 B
constexpr size_t kTopkScratchPerCta = 8 * 4 * 1024 * 16;
constexpr int    kTopkMaxCtas       = 160;

struct CoarseParams {
    // rows
    const float* z;          // [*, d] raw rows (normalised by the A loader)
    int64_t      n;          // rows to process when n_rows_dev == nullptr
    int          d;
    int          norm_mode;  // MTVQ_NORM_*
    const int*   row_list;   // optional gather list (COLLECT over near-tie rows)
    const int*   n_rows_dev; // optional device-side row count (length of row_list)
    // codebook (prepared by codebook_prep_kernel)
    const void*  b_image;
    int          kpad;       // codes padded to a multiple of 128
    int          k;
    const float* meta;       // META_MARGIN ...
    // outputs / state
    float*       denom;      // [n] or nullptr
    int64_t*     idx_out;    // [n]  ARGMAX: winning code of every decided row (torch.argmin dtype); near-tie rows
                             //      are written by the resolve kernels
    int*         bins;       // [k]  ARGMAX: per-code row counts of the decided rows (zeroed by the caller)
    float*       best_val;   // [n]
    int*         flag_rows;  // [n]   ARGMAX: near-tie rows (exact re-scoring decides)
    int*         over_rows;  // [n]   ARGMAX: the near-tie rows that also need a COLLECT pass (counters[CNT_OVER])
    int*         counters;   // [CNT_COUNT]
    int*         cand;       // [rows * cand_cap]   indexed by position in row_list
    int*         cand_cnt;   // [rows]
    int          cand_cap;
    int          topk;
    void*        topk_scratch; // TOPK: kTopkScratchPerCta bytes per CTA
    float*       dump;       // DUMP: [dump_rows, kpad]
    int64_t      dump_rows;
    // shared-memory descriptor strides (kept as parameters so tests can probe the encoding)
    uint32_t     lbo_bytes;  // 2048
    uint32_t     sbo_bytes;  // 128
    long long*   trace;       // optional clock64 trace of slot 0 of CTA 0 (mtvq_debug_set_trace), 8 x int64 per tile
    int          debug_flags; // timing experiments only (MTVQ_DBG): 1 = epilogue does no work, 2 = no bias K-step,
                              // 4 = A tiles loaded once per CTA; results are wrong when set
};

// rows per CTA block = 128 * row_tiles; bounded by shared memory (row_tiles * d <= 256)
int         coarse_row_tiles(int d, int64_t nrows, int num_sms);
cudaError_t launch_coarse(const CoarseParams& p, int mode, int row_tiles, int num_sms, cudaStream_t stream);
int         coarse_smem_bytes();

}  // namespace mtvq
