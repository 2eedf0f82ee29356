// workspace.h -- host-side carving of the caller-provided scratch blob (the library never allocates).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace mtvq {

constexpr int kMaxDim = 256;   // D must be a multiple of 16 and <= kMaxDim

// meta block written by the codebook prep kernel (floats)
enum MetaSlot { META_MARGIN = 0, META_EPS = 1, META_MAXNORM = 2, META_SPREAD = 3, META_COUNT = 8 };
// the -||w||^2/2 K-step is skipped when max_k | ||w_k||^2 - ||w_0||^2 | / 2 <= kBiasSkipTol (eps carries the slack)
constexpr float kBiasSkipTol = 1.0e-6f;
// counters (ints), zeroed at the start of every assign
enum Counter { CNT_FLAGGED = 0, CNT_FULL = 1, CNT_CANDS = 2, CNT_LAUNCHES = 3, CNT_OVER = 4, CNT_TILES = 5, CNT_COUNT = 8 };

// cudaFuncSetAttribute applies to the current device only: "configured once" flags are kept per device (a host may
// drive several GPUs from one process, even though the module layer uses one process per GPU)
struct PerDeviceOnce {
    bool done[64] = {};
    bool& flag() {
        int dev = 0;
        cudaGetDevice(&dev);
        return done[dev & 63];
    }
};

// kernels this thread has enqueued (every launcher adds its own; reported through mtvq_debug_counters)
extern thread_local long g_kernel_launches;

struct Workspace {
    // codebook side
    void*    b_image;      // fp16 operand image, kpad/128 tiles x (d/8+2) planes x 128 rows x 16 B
    float*   cc;           // [kpad]  canonical ||w_k||^2
    float*   codes_n;      // [k,d]   normalised codebook rows (soft path only)
    float*   meta;         // [META_COUNT]
    // row side
    float*   denom;        // [n]
    float*   best_val;     // [n]   best coarse score
    int*     flag_rows;    // [n]   rows needing exact re-scoring
    int*     cand;         // [n * kCandMax]  (soft: [n * cand_cap])
    int*     cand_cnt;     // [n]
    int*     full_rows;    // [n]
    int*     counters;     // [CNT_COUNT]
    void*    topk_scratch; // soft path: per-CTA record scratch of the single-pass top-k kernel (L2 resident)
    void*    small_part;   // small-batch route: per-(row, code slice) partial top-k lists (small.cu)
    // statistics side
    int*     bins_i32;     // [k]
    int*     offsets;      // [k+1]
    int*     cursor;       // [k]
    int*     perm;         // [n]
    // two-pass hard route (assign_tc.h MODE_GROUP / MODE_REFINE, group.cu)
    int*     grp;          // [n]            per-row chunk word of the GROUP pass
    int*     ghist;        // [160][kpad / 128]  entries per (CTA of the GROUP pass, chunk), then the first entry of each share
    int*     goff;         // [kpad / 128]   first entry | first tile of each chunk
    int*     gtile;
    int*     gtot;         // [kpad / 128]   entries per chunk
    int*     grank;        // [n]            rank of a row's entry within its (CTA, chunk) share (second entry: grank2)
    int*     grank2;
    void*    entries;      // [2 n] int4     rows sorted by chunk (a near-tie row spanning two chunks is listed twice)
    void*    tile_tab;     // [group_tile_capacity] int2
    size_t   total_bytes;
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int    pad_codes(int k) { return (int)align_up((size_t)k, 128); }
// near-tie candidates kept per row before the row falls back to the full exact scan
inline int    cand_capacity(bool soft) { return soft ? 24 : 16; }
constexpr int kGroupMaxCtas = 160;   // CTAs of the GROUP pass (one per SM)
inline size_t group_tile_capacity(int64_t n, int k) { return (size_t)(2 * n / 128) + (size_t)(pad_codes(k) / 128) + 2; }
inline size_t image_bytes(int d, int k) { return (size_t)(pad_codes(k) / 128) * (size_t)(d / 8 + 2) * 2048; }

// Small-batch route (small.cu): work bound below which scoring every (row, code) pair exactly on the CUDA cores beats the
// tensor pass's fixed cost, and its scratch: the partial top-k lists -- 8 bytes per (row, 64-code slice, list entry), a few
// MB under the second bound -- followed by the normalised rows stored d-major (xnT[d][n_pad]) and their squared norms.
inline bool   small_route_fits(int64_t n, int d, int k) {
    // measured crossover on B200 (tools_small_batch_probe.py, D = 64): N = 256 x K = 21000 -> 100 us here vs 164 us on the
    // tensor route; N = 1024 x K = 7000 -> 103 vs 104 us; N = 1024 x K = 21000 -> 230 vs 169 us
    return n > 0 && n <= 4096 && (double)n * (double)k * (double)d <= 6.0e8 && (double)n * (double)k <= 3.0e7;
}
inline size_t small_npad(int64_t n) { return align_up((size_t)n, 64); }
inline size_t small_part_bytes(int64_t n, int k, int topk) {
    return align_up((size_t)n * (size_t)((k + 63) / 64) * (size_t)(topk < 1 ? 1 : topk) * 8, 256);
}
inline size_t small_scratch_bytes(int64_t n, int d, int k, int topk) {
    return small_part_bytes(n, k, topk) + sizeof(float) * small_npad(n) * ((size_t)d + 1);
}

// topk == 0: hard path only; topk >= 1: the buffers of the soft path (normalised codebook, per-CTA record scratch, wider
// candidate lists) as well -- also for top-1, which VectorQuantizer(..., k=1) may ask for.  Everything whose size does
// not depend on the path comes first, so those offsets (counters, ...) are the same for both.
inline Workspace carve(void* base, int64_t n, int d, int k, int topk) {
    Workspace w{};
    size_t off = 0;
    char*  p   = static_cast<char*>(base);
    auto take  = [&](size_t bytes) {
        void* r = p ? p + off : nullptr;
        off += align_up(bytes, 256);
        return r;
    };
    const bool   soft     = topk >= 1;
    const int    kpad     = pad_codes(k);
    const size_t cand_cap = (size_t)cand_capacity(soft);
    w.b_image   = take(image_bytes(d, k));
    w.cc        = (float*)take(sizeof(float) * kpad);
    w.meta      = (float*)take(sizeof(float) * META_COUNT);
    w.counters  = (int*)take(sizeof(int) * CNT_COUNT);
    w.denom     = (float*)take(sizeof(float) * n);
    w.best_val  = (float*)take(sizeof(float) * n);
    w.flag_rows = (int*)take(sizeof(int) * n);
    w.cand_cnt  = (int*)take(sizeof(int) * n);
    w.full_rows = (int*)take(sizeof(int) * n);
    w.bins_i32  = (int*)take(sizeof(int) * k);
    w.offsets   = (int*)take(sizeof(int) * (k + 1));
    w.cursor    = (int*)take(sizeof(int) * k);
    w.perm      = (int*)take(sizeof(int) * n);
    w.cand      = (int*)take(sizeof(int) * n * cand_cap);
    w.codes_n   = (float*)take(soft ? sizeof(float) * (size_t)k * d : 0);
    w.topk_scratch = take(soft ? (size_t)160 * (8 * 4 * 1024 * 16) : 0);   // kTopkMaxCtas x kTopkScratchPerCta
    w.small_part = take(small_route_fits(n, d, k) ? small_scratch_bytes(n, d, k, soft ? topk : 1) : 0);
    const size_t nchunks = (size_t)kpad / 128;
    w.grp       = (int*)take(sizeof(int) * n);
    w.ghist     = (int*)take(sizeof(int) * nchunks * kGroupMaxCtas);
    w.goff      = (int*)take(sizeof(int) * nchunks);
    w.gtile     = (int*)take(sizeof(int) * (nchunks + 1));
    w.gtot      = (int*)take(sizeof(int) * nchunks);
    w.grank     = (int*)take(sizeof(int) * n);
    w.grank2    = (int*)take(sizeof(int) * n);
    w.entries   = take(16 * (size_t)(2 * n));
    w.tile_tab  = take(8 * group_tile_capacity(n, k));
    w.total_bytes = off;
    return w;
}

}  // namespace mtvq
