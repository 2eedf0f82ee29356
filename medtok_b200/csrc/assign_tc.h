// assign_tc.h -- parameters / launcher of the tcgen05 coarse-score kernel (assign_tc.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "workspace.h"

namespace mtvq {

enum CoarseMode { MODE_ARGMAX = 0, MODE_THRESH = 1, MODE_COLLECT = 2, MODE_DUMP = 3, MODE_TOPK = 4, MODE_GROUP = 5, MODE_REFINE = 6 };

// Two-pass hard route (D <= 64, large batches).  MODE_GROUP is the tensor pass with a
// branch-free epilogue: per row only the best 128-code chunk (its maximum, and whether another chunk comes within the
// margin) -- no score records.  The rows are then counting-sorted by chunk (group.cu) and MODE_REFINE re-scores every
// 128-row tile of the sorted order against its ONE chunk: it decides the column, lists near-ties for the exact
// re-scoring, and -- fused -- writes z_q, the squared error and the per-code sums of the decided rows.
constexpr int kGroupMaxChunks = 8192;      // chunk ids are packed into 14 bits; the per-CTA histogram lives in 32 KB
constexpr int kGrpOtherShift  = 14;
constexpr int kGrpCrossBit    = 1 << 28;   // a second chunk is within the margin (the row is listed under both)
constexpr int kGrpOverBit     = 1 << 29;   // three or more chunks (or no finite score): COLLECT pass decides
constexpr unsigned kEntCross  = 1u << 31;  // sorted entry {row | flags, best coarse score, row norm, 0} (16 bytes)
constexpr unsigned kEntSecond = 1u << 30;
constexpr unsigned kEntRowMask = (1u << 30) - 1;
// cand_cnt[row]: bits 0-23 = columns listed from the front of the row's candidate list, bits 24-31 = columns listed from
// its back (the second tile of a row listed under two chunks, REFINE only)
constexpr int kCandBackShift = 24;

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
    // two-pass hard route
    int*         grp;         // GROUP out [n]: best chunk | other chunk << 14 | kGrpCrossBit | kGrpOverBit
    int*         ghist;       // GROUP out [CTAs][kpad / 128]: offset of the CTA's share within the chunk's range of entries
    int*         gtot;        // GROUP out [kpad / 128]: entries per chunk (zeroed by the caller)
    int*         grank;       // GROUP out [n]: rank of the row's entry within its (CTA, chunk) share
    int*         grank2;      //           [n]: the same for the second entry of a row listed under two chunks
    const int4*  entries;     // REFINE in: entries sorted by chunk {row | flags, best coarse score, row norm, 0}
    const int2*  tile_tab;    // REFINE in: per 128-entry tile {chunk | count << 16, first entry}
    const int*   n_tiles_dev; // REFINE in: number of tiles
    const float* w;           // REFINE, fused rows: fp32 codebook [k, d]
    float*       z_q;         //         straight-through output rows [n, d]
    double*      stats;       //         packed statistics (bins [k] | embed_sum [k, d] | sse | rows)
    int          fuse_rows;   //         0 = decide only, 1 = also gather / STE / SSE / per-code sums of the decided rows
    long long*   trace;       // optional clock64 trace of slot 0 of CTA 0 (mtvq_debug_set_trace), 8 x int64 per tile
    int          debug_flags; // timing experiments only (MTVQ_DBG): 1 = epilogue does no work, 2 = no bias K-step,
                              // 4 = A tiles loaded once per CTA; results are wrong when set
};

// rows per CTA block = 128 * row_tiles; bounded by shared memory (row_tiles * d <= 256)
int         coarse_row_tiles(int d, int64_t nrows, int num_sms);
cudaError_t launch_coarse(const CoarseParams& p, int mode, int row_tiles, int num_sms, cudaStream_t stream);
int         coarse_smem_bytes();

}  // namespace mtvq
