// api.cu -- extern "C" surface of libmedtok_vq.so (declared in include/medtok_vq.h).
// Host-side argument checks, workspace carving and kernel sequencing only; no device allocation,
// no host synchronisation (diagnostic entry points excepted), no exceptions across the ABI.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/medtok_vq.h"
#include "assign_exact.h"
#include "assign_tc.h"
#include "group.h"
#ifdef MTVQ_WITH_MICROBENCH
#include "microbench.h"
#endif
#include "rows.h"
#include "small.h"
#include "soft.h"
#include "workspace.h"

using namespace mtvq;

namespace mtvq { thread_local long g_kernel_launches = 0; }

namespace {

thread_local char g_err[512] = "";
thread_local long long* g_trace = nullptr;
thread_local int g_soft_route = 0;   // mtvq_debug_set_soft_route
thread_local bool g_no_small = false; // mtvq_debug_set_soft_route(3): force the tensor route for small batches too
thread_local cudaEvent_t g_ev_before = nullptr, g_ev_after = nullptr;

int fail(int code, const char* fmt, const char* a = "", long b = 0, long c = 0) {
    snprintf(g_err, sizeof(g_err), fmt, a, b, c);
    return code;
}
int cuda_fail(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return (int)e;
}
#define MTVQ_CUDA(call, where)                                 \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return cuda_fail(e__, where);  \
    } while (0)

struct DeviceInfo {
    int device = -1;
    int num_sms = 0;
    int major = 0;
};
int device_info(DeviceInfo* out) {
    static DeviceInfo cache[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    if (dev < 0 || dev >= 64) return fail(MTVQ_ERR_NO_DEVICE, "device ordinal out of range%s", "");
    if (cache[dev].device != dev) {
        cudaDeviceProp p;
        e = cudaGetDeviceProperties(&p, dev);
        if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
        cache[dev].device = dev;
        cache[dev].num_sms = p.multiProcessorCount;
        cache[dev].major = p.major;
    }
    if (cache[dev].major != 10)
        return fail(MTVQ_ERR_NO_DEVICE, "libmedtok_vq is sm_100a-only; device has compute capability %s%ld.x", "",
                    (long)cache[dev].major);
    *out = cache[dev];
    return 0;
}

int check_norm_mode(int norm_mode) {
    if (norm_mode < 0 || (norm_mode & ~7) || (norm_mode & 3) == 3)
        return fail(MTVQ_ERR_BAD_ARG, "norm_mode must be MTVQ_NORM_* optionally or-ed with MTVQ_SCORE_DOT%s (got %ld)", "", (long)norm_mode);
    return 0;
}
int check_dims(int64_t n, int d, int k) {
    if (n < 0 || k <= 0 || d <= 0) return fail(MTVQ_ERR_BAD_ARG, "bad extents%s n=%ld k=%ld", "", (long)n, (long)k);
    if (d % 16 != 0 || d > kMaxDim)
        return fail(MTVQ_ERR_UNSUPPORTED, "embedding dim must be a multiple of 16 and <= 256%s (got %ld)", "", (long)d);
    if (n > 0x7fffffffLL) return fail(MTVQ_ERR_UNSUPPORTED, "more than 2^31-1 rows per call%s", "");
    return 0;
}
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

bool desc_swap() {
    static int v = -1;
    if (v < 0) { const char* s = getenv("MTVQ_DESC_SWAP"); v = (s && s[0] == '1') ? 1 : 0; }
    return v == 1;
}

void fill_coarse(CoarseParams& p, const float* z, int64_t n, int d, int norm_mode, int k, const Workspace& ws,
                 float* denom, int topk, bool soft, int64_t* idx_out = nullptr) {
    memset(&p, 0, sizeof(p));
    p.z = z; p.n = n; p.d = d; p.norm_mode = norm_mode;
    p.b_image = ws.b_image; p.kpad = pad_codes(k); p.k = k; p.meta = ws.meta;
    p.denom = denom; p.idx_out = idx_out; p.bins = ws.bins_i32; p.best_val = ws.best_val;
    p.flag_rows = ws.flag_rows; p.over_rows = ws.full_rows; p.counters = ws.counters;
    p.cand = ws.cand; p.cand_cnt = ws.cand_cnt; p.cand_cap = cand_capacity(soft); p.topk = topk;
    p.topk_scratch = ws.topk_scratch;
    p.trace = g_trace;
    p.lbo_bytes = desc_swap() ? 128 : 2048;
    p.sbo_bytes = desc_swap() ? 2048 : 128;
    { static int dbg = -1; if (dbg < 0) { const char* e = getenv("MTVQ_DBG"); dbg = e ? atoi(e) : 0; } p.debug_flags = dbg; }
}

// Two-pass route of the hard assignment: record-free tensor pass -> rows sorted by their best
// 128-code chunk -> REFINE pass over 128-row tiles of one chunk each.  Worth it where the single pass is bound by its
// epilogue (D <= 64: four row tiles per block) and the batch is large enough to amortise the sort.
// MTVQ_HARD_ROUTE=1 forces the single-pass route (A/B timing, cross-check tests).
int hard_route_env() {
    static int v = -1;
    if (v < 0) { const char* s = getenv("MTVQ_HARD_ROUTE"); v = s ? atoi(s) : 0; }
    return v;
}
thread_local int g_hard_route = -1;      // mtvq_debug_set_hard_route: -1 = environment / automatic
bool two_pass_ok(int64_t n, int d, int k, int num_sms) {
    const int route = g_hard_route >= 0 ? g_hard_route : hard_route_env();
    if (route == 1) return false;
    const int nchunks = pad_codes(k) / 128;
    if (coarse_row_tiles(d, n, num_sms) != 4 || nchunks < 2 || nchunks > kGroupMaxChunks || n >= (1LL << 30) ||
        num_sms > kGroupMaxCtas)
        return false;
    // measured on B200 (tools_route_check.py): the record-free tensor pass saves ~16 % of the single pass, the sort + REFINE
    // tail costs ~0.16 us per 1000 rows more than the single pass's tail -- a net win from about K x D = 8192 x 64 upwards
    return route == 2 || (n >= 262144 && (int64_t)pad_codes(k) * d >= (1 << 19));
}

// fused row outputs of the two-pass route (training forward): z_q, squared error and per-code sums come out of the
// REFINE pass (decided rows) and finish_flagged (near-tie rows) instead of the sorted-rows pass
struct FusedRows {
    float*  z_q;
    double* stats;
    bool    done;      // out: the route produced them
};

// hard assignment: idx[n] (int64, first-index argmin) + ws.bins_i32 (per-code row counts) + denom;
// `codes` are the rows the distances are taken against
int assign_rows(const float* z, int64_t n, int d, int norm_mode, const float* codes, int k, float* denom, int64_t* idx,
                const Workspace& ws, const DeviceInfo& di, bool exact_only, cudaStream_t st, FusedRows* fused = nullptr) {
    MTVQ_CUDA(cudaMemsetAsync(ws.counters, 0, sizeof(int) * CNT_COUNT, st), "memset counters");
    MTVQ_CUDA(cudaMemsetAsync(ws.bins_i32, 0, sizeof(int) * k, st), "memset bins");
    if (n == 0) return 0;
    if (!exact_only && small_route_ok(n, d, k, norm_mode) && !g_no_small) {
        // small batch: every (row, code) distance directly in the reference's fp32 order (small.cu); no tensor pass
        MTVQ_CUDA(launch_small_assign(z, n, d, norm_mode, codes, k, 0, 1, nullptr, ws.cc, denom, ws.small_part, nullptr,
                                      nullptr, 0, idx, ws.bins_i32, ws.full_rows, ws.counters, st), "small_assign");
        MTVQ_CUDA(launch_exact_scan(z, n, d, norm_mode, codes, k, ws.cc, ws.full_rows, ws.counters + CNT_FULL, idx,
                                    ws.bins_i32, nullptr, di.num_sms, st), "exact_scan(NaN rows)");
        return 0;
    }
    MTVQ_CUDA(launch_codebook_prep(codes, k, d, (norm_mode & MTVQ_SCORE_DOT) ? 2 : 0, ws, st), "codebook_prep");
    if (exact_only) {
        MTVQ_CUDA(launch_exact_scan(z, n, d, norm_mode, codes, k, ws.cc, nullptr, nullptr, idx, ws.bins_i32, denom,
                                    di.num_sms, st), "exact_scan");
        return 0;
    }
    CoarseParams p;
    fill_coarse(p, z, n, d, norm_mode, k, ws, denom, 1, false, idx);
    const bool two_pass = two_pass_ok(n, d, k, di.num_sms);
    if (!two_pass) MTVQ_CUDA(cudaMemsetAsync(ws.cand_cnt, 0, sizeof(int) * n, st), "memset cand_cnt");   // GROUP zeroes it per row
    if (two_pass) {
        if (di.num_sms > kGroupMaxCtas) return fail(MTVQ_ERR_UNSUPPORTED, "more SMs than the group tables hold%s", "");
        p.grp = ws.grp;
        MTVQ_CUDA(cudaMemsetAsync(ws.gtot, 0, sizeof(int) * (pad_codes(k) / 128), st), "memset gtot");
        p.ghist = ws.ghist;
        p.gtot = ws.gtot;
        p.grank = ws.grank;
        p.grank2 = ws.grank2;
        static int trace_group = -1;
        if (trace_group < 0) { const char* e = getenv("MTVQ_TRACE_GROUP"); trace_group = (e && e[0] == '1') ? 1 : 0; }
        if (trace_group) p.debug_flags |= 0x100;         // the pipeline trace (if set) records the GROUP pass, not REFINE
        if (g_ev_before) cudaEventRecord(g_ev_before, st);
        MTVQ_CUDA(launch_coarse(p, MODE_GROUP, 4, di.num_sms, st), "coarse group");
        if (g_ev_after) cudaEventRecord(g_ev_after, st);
        g_ev_before = g_ev_after = nullptr;
        MTVQ_CUDA(launch_group_sort(ws, denom, n, k, di.num_sms, st), "group sort");
        CoarseParams r = p;
        r.denom = nullptr;
        if (trace_group) r.trace = nullptr;
        r.debug_flags = 0;
        r.entries = static_cast<const int4*>(ws.entries);
        r.tile_tab = static_cast<const int2*>(ws.tile_tab);
        r.n_tiles_dev = ws.counters + CNT_TILES;
        if (fused) { r.w = codes; r.z_q = fused->z_q; r.stats = fused->stats; r.fuse_rows = 1; }
        MTVQ_CUDA(launch_coarse(r, MODE_REFINE, 1, di.num_sms, st), "coarse refine");
    } else {
        if (g_ev_before) cudaEventRecord(g_ev_before, st);
        MTVQ_CUDA(launch_coarse(p, MODE_ARGMAX, coarse_row_tiles(d, n, di.num_sms), di.num_sms, st), "coarse argmax");
        if (g_ev_after) cudaEventRecord(g_ev_after, st);
        g_ev_before = g_ev_after = nullptr;
    }
    // near-tie rows (a few percent) carry a candidate list out of the tensor pass; the few whose list is incomplete
    // (a third granule within the margin) get it from a COLLECT pass -- the row count is only known on the device,
    // the kernel splits the codebook over CTAs to fill the SMs (assign_tc.cu "work items").  Exact fp32 decides.
    CoarseParams c = p;
    c.row_list = ws.full_rows;                    // the over list (consumed before resolve reuses the buffer)
    c.n_rows_dev = ws.counters + CNT_OVER;
    c.denom = nullptr;
    MTVQ_CUDA(launch_coarse(c, MODE_COLLECT, 1, di.num_sms, st), "coarse collect");   // few rows: 128-row items
    if (two_pass) {
        MTVQ_CUDA(launch_resolve8(z, denom, d, norm_mode, codes, k, ws, p.cand_cap, idx, di.num_sms, st,
                                  fused ? fused->z_q : nullptr, fused ? fused->stats : nullptr), "resolve8");
    } else {
        MTVQ_CUDA(launch_resolve(z, d, norm_mode, codes, k, ws, p.cand_cap, idx, di.num_sms, st), "resolve");
    }
    MTVQ_CUDA(launch_exact_scan(z, n, d, norm_mode, codes, k, ws.cc, ws.full_rows, ws.counters + CNT_FULL, idx, ws.bins_i32,
                                nullptr, di.num_sms, st), "exact_scan(overflow)");
    if (two_pass && fused) {
        MTVQ_CUDA(launch_finish_flagged(z, denom, codes, idx, n, d, k, fused->z_q, fused->stats, ws, di.num_sms, st),
                  "finish_flagged");
        fused->done = true;
    }
    return 0;
}

}  // namespace

extern "C" {

int mtvq_abi_version(void) { return MTVQ_ABI_VERSION; }
const char* mtvq_last_error(void) { return g_err; }

size_t mtvq_workspace_bytes(int64_t n, int d, int k, int topk) {
    if (n < 0 || d <= 0 || k <= 0) return 0;
    return carve(nullptr, n, d, k, topk < 0 ? 0 : topk).total_bytes;
}
size_t mtvq_stats_len(int d, int k) { return (size_t)k + (size_t)k * d + 2; }

static int hard_assign_impl(const float* z, int64_t n, int d, int norm_mode, const float* weight, int k, int64_t* idx,
                            float* denom, void* workspace, size_t workspace_bytes, void* stream, bool exact_only) {
    if (int r = check_dims(n, d, k)) return r;
    if (int r = check_norm_mode(norm_mode)) return r;
    if (!z || !weight || !idx || !workspace) return fail(MTVQ_ERR_BAD_ARG, "null pointer%s", "");
    if (!aligned16(z) || !aligned16(weight) || !aligned16(workspace))
        return fail(MTVQ_ERR_BAD_ARG, "z, weight and workspace must be 16-byte aligned%s", "");
    DeviceInfo di;
    if (int r = device_info(&di)) return r;
    Workspace ws = carve(workspace, n, d, k, 0);
    if (ws.total_bytes > workspace_bytes)
        return fail(MTVQ_ERR_WORKSPACE, "workspace too small%s: need %ld, got %ld", "", (long)ws.total_bytes,
                    (long)workspace_bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* den = denom ? denom : ws.denom;
    return assign_rows(z, n, d, norm_mode, weight, k, den, idx, ws, di, exact_only, st);
}

int mtvq_hard_assign(const float* z, int64_t n, int d, int norm_mode, const float* weight, int k, int64_t* idx,
                     float* denom, void* workspace, size_t workspace_bytes, void* stream) {
    return hard_assign_impl(z, n, d, norm_mode, weight, k, idx, denom, workspace, workspace_bytes, stream, false);
}
int mtvq_hard_assign_exact(const float* z, int64_t n, int d, int norm_mode, const float* weight, int k, int64_t* idx,
                           float* denom, void* workspace, size_t workspace_bytes, void* stream) {
    return hard_assign_impl(z, n, d, norm_mode, weight, k, idx, denom, workspace, workspace_bytes, stream, true);
}

int mtvq_gather_ste(const float* z, const float* denom, const float* weight, const int64_t* idx, int64_t n, int d,
                    int k, float* z_q, double* stats, void* stream) {
    if (int r = check_dims(n, d, k)) return r;
    if (!z || !denom || !weight || !idx || !z_q || !stats) return fail(MTVQ_ERR_BAD_ARG, "null pointer%s", "");
    if (!aligned16(z) || !aligned16(weight) || !aligned16(z_q)) return fail(MTVQ_ERR_BAD_ARG, "16-byte alignment%s", "");
    DeviceInfo di;
    if (int r = device_info(&di)) return r;
    if (n > 0)
        MTVQ_CUDA(launch_gather_ste(z, denom, weight, idx, n, d, k, z_q, stats, di.num_sms, static_cast<cudaStream_t>(stream)),
                  "gather_ste");
    return 0;
}

int mtvq_code_sums(const float* z, const float* denom, const int64_t* idx, int64_t n, int d, int k, double* stats,
                   void* workspace, size_t workspace_bytes, void* stream) {
    if (int r = check_dims(n, d, k)) return r;
    if (!z || !denom || !idx || !stats || !workspace) return fail(MTVQ_ERR_BAD_ARG, "null pointer%s", "");
    DeviceInfo di;
    if (int r = device_info(&di)) return r;
    Workspace ws = carve(workspace, n, d, k, 0);
    if (ws.total_bytes > workspace_bytes) return fail(MTVQ_ERR_WORKSPACE, "workspace too small%s", "");
    if (n > 0) {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        MTVQ_CUDA(launch_hist_idx64(idx, n, ws.bins_i32, k, di.num_sms, st), "hist");
        MTVQ_CUDA(launch_code_sums(z, denom, idx, n, d, k, stats, ws, true, di.num_sms, st), "code_sums");
    }
    return 0;
}

int mtvq_ema_finalize(float* weight, float* cluster_size, const double* stats, int k, int d, double beta, double decay,
                      int update_cluster_size, int update_weight, float* loss, void* stream) {
    if (int r = check_dims(0, d, k)) return r;
    if (!stats || (update_weight && !weight) || (update_cluster_size && !cluster_size))
        return fail(MTVQ_ERR_BAD_ARG, "null pointer%s", "");
    DeviceInfo di;
    if (int r = device_info(&di)) return r;
    MTVQ_CUDA(launch_ema_finalize(weight, cluster_size, stats, k, d, beta, decay, update_cluster_size, update_weight, loss,
                                  static_cast<cudaStream_t>(stream)), "ema_finalize");
    return 0;
}

int mtvq_kmeans_update(float* means, const double* stats, int k, int d, void* stream) {
    if (int r = check_dims(0, d, k)) return r;
    if (!means || !stats) return fail(MTVQ_ERR_BAD_ARG, "null pointer%s", "");
    DeviceInfo di;
    if (int r = device_info(&di)) return r;
    MTVQ_CUDA(launch_kmeans_update(means, stats, k, d, static_cast<cudaStream_t>(stream)), "kmeans_update");
    return 0;
}

int mtvq_ema_finalize_peers(float* weight, float* cluster_size, const void* const* peer_stats, void* const* peer_flags,
                            void* const* peer_results, int world, int rank, unsigned epoch, int k, int d, double beta,
                            double decay, int update_cluster_size, int update_weight, float* loss, void* stream) {
    if (int r = check_dims(0, d, k)) return r;
    if (!peer_stats || !peer_flags || !peer_results || (update_weight && !weight) || (update_cluster_size && !cluster_size))
        return fail(MTVQ_ERR_BAD_ARG, "null pointer%s", "");
    if (world < 1 || world > 64 || rank < 0 || rank >= world) return fail(MTVQ_ERR_BAD_ARG, "bad world / rank%s", "");
    DeviceInfo di;
    if (int r = device_info(&di)) return r;
    MTVQ_CUDA(launch_ema_finalize_peers(weight, cluster_size, peer_stats, peer_flags, peer_results, world, rank, epoch, k, d,
                                        beta, decay, update_cluster_size, update_weight, loss, di.num_sms,
                                        static_cast<cudaStream_t>(stream)),
              "ema_finalize_peers");
    return 0;
}

int mtvq_hard_forward(const float* z, int64_t n, int d, int norm_mode, float* weight, int k, float* cluster_size,
                      double beta, double decay, unsigned flags, float* z_q, int64_t* idx, float* loss, double* stats,
                      void* workspace, size_t workspace_bytes, void* stream) {
    if (int r = check_dims(n, d, k)) return r;
    if (int r = check_norm_mode(norm_mode)) return r;
    if (!weight || !loss || !stats || !workspace || (n > 0 && (!z || !z_q || !idx)))
        return fail(MTVQ_ERR_BAD_ARG, "null pointer%s", "");
    if (!aligned16(z) || !aligned16(weight) || !aligned16(z_q) || !aligned16(workspace))
        return fail(MTVQ_ERR_BAD_ARG, "z, weight, z_q and workspace must be 16-byte aligned%s", "");
    DeviceInfo di;
    if (int r = device_info(&di)) return r;
    Workspace ws = carve(workspace, n, d, k, 0);
    if (ws.total_bytes > workspace_bytes)
        return fail(MTVQ_ERR_WORKSPACE, "workspace too small%s: need %ld, got %ld", "", (long)ws.total_bytes,
                    (long)workspace_bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool training = flags & MTVQ_F_TRAINING, update = flags & MTVQ_F_UPDATE;
    const bool upd_w = training && update;            // norm_ema_quantizer.py:191
    const bool upd_cs = (!training) || update;        // :185 / :191 are the only branches touching cluster_size
    if (upd_cs && !cluster_size && !(flags & MTVQ_F_SKIP_FINALIZE))
        return fail(MTVQ_ERR_BAD_ARG, "cluster_size is null%s", "");

    if (!(flags & MTVQ_F_ACCUMULATE))
        MTVQ_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * mtvq_stats_len(d, k), st), "memset stats");
    FusedRows fused{z_q, stats, false};
    if (int r = assign_rows(z, n, d, norm_mode, weight, k, ws.denom, idx, ws, di, false, st, upd_w ? &fused : nullptr)) return r;
    if (n > 0 && !fused.done) {
        if (upd_w) {   // per-code sums wanted: one fused pass over the rows in code-sorted order
            MTVQ_CUDA(launch_sorted_rows(z, ws.denom, weight, idx, n, d, k, z_q, stats, ws, di.num_sms, st), "sorted_rows");
        } else {       // streaming gather/STE/SSE + bins only
            MTVQ_CUDA(launch_gather_ste(z, ws.denom, weight, idx, n, d, k, z_q, stats, di.num_sms, st), "gather_ste");
            MTVQ_CUDA(launch_code_sums(z, ws.denom, idx, n, d, k, stats, ws, false, di.num_sms, st), "code_sums");
        }
    }
    if (flags & MTVQ_F_SKIP_FINALIZE) {
        MTVQ_CUDA(launch_loss(stats, k, d, beta, loss, st), "loss");
    } else {
        MTVQ_CUDA(launch_ema_finalize(weight, cluster_size, stats, k, d, beta, decay, upd_cs, upd_w, loss, st),
                  "ema_finalize");
    }
    return 0;
}

int mtvq_hard_backward(const float* z, const float* z_q, const float* g_zq, const float* g_loss, int64_t n, int d,
                       double beta, float* grad_z, void* stream) {
    if (int r = check_dims(n, d, 1)) return r;
    if (n > 0 && (!z || !z_q || !grad_z)) return fail(MTVQ_ERR_BAD_ARG, "null pointer%s", "");
    if (!aligned16(z) || !aligned16(z_q) || !aligned16(grad_z) || (g_zq && !aligned16(g_zq)))
        return fail(MTVQ_ERR_BAD_ARG, "16-byte alignment%s", "");
    DeviceInfo di;
    if (int r = device_info(&di)) return r;
    if (n > 0)
        MTVQ_CUDA(launch_hard_backward(z, z_q, g_zq, g_loss, n, d, beta, grad_z, di.num_sms, static_cast<cudaStream_t>(stream)),
                  "hard_backward");
    return 0;
}

int mtvq_lookup(const float* z, const float* weight, const int64_t* idx, int64_t n, int d, int k, float* z_q,
                double* sse, float* counts, void* workspace, size_t workspace_bytes, void* stream) {
    if (int r = check_dims(n, d, k)) return r;
    if (!weight || !idx || !z_q) return fail(MTVQ_ERR_BAD_ARG, "null pointer%s", "");
    if ((z && !aligned16(z)) || !aligned16(weight) || !aligned16(z_q)) return fail(MTVQ_ERR_BAD_ARG, "16-byte alignment%s", "");
    DeviceInfo di;
    if (int r = device_info(&di)) return r;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n > 0) MTVQ_CUDA(launch_lookup(z, weight, idx, n, d, k, z_q, sse, di.num_sms, st), "lookup");
    if (counts) {
        if (!workspace) return fail(MTVQ_ERR_BAD_ARG, "counts need the workspace%s", "");
        Workspace ws = carve(workspace, n, d, k, 0);
        if (ws.total_bytes > workspace_bytes) return fail(MTVQ_ERR_WORKSPACE, "workspace too small%s", "");
        MTVQ_CUDA(launch_hist_idx64(idx, n, ws.bins_i32, k, di.num_sms, st), "hist");
        MTVQ_CUDA(launch_counts_to_float(ws.bins_i32, k, counts, st), "counts");
    }
    return 0;
}

int mtvq_codebook_usage(const float* window, const int64_t* idx, int64_t count, int cap, int n_e, float* next,
                        void* scratch, size_t scratch_bytes, int* distinct, void* stream) {
    if (!window || !idx || !next || !scratch || !distinct || count <= 0 || cap <= 0 || n_e <= 0)
        return fail(MTVQ_ERR_BAD_ARG, "bad argument%s", "");
    if (scratch_bytes < sizeof(unsigned) * ((size_t)(n_e + 31) / 32 + 1)) return fail(MTVQ_ERR_WORKSPACE, "scratch too small%s", "");
    DeviceInfo di;
    if (int r = device_info(&di)) return r;
    MTVQ_CUDA(launch_usage(window, idx, count, cap, n_e, next, static_cast<unsigned*>(scratch), distinct, di.num_sms,
                           static_cast<cudaStream_t>(stream)), "usage");
    return 0;
}

int mtvq_soft_forward(const float* x, int64_t n, int d, const float* codes, int k, int topk, int64_t* idx, float* vals,
                      float* wts, float* zq, float* out, double* sse, void* workspace, size_t workspace_bytes,
                      void* stream) {
    if (int r = check_dims(n, d, k)) return r;
    if (topk < 1 || topk > 8 || topk > k) return fail(MTVQ_ERR_UNSUPPORTED, "top-k must be in [1, min(8, k)]%s", "");
    if (!codes || !sse || !workspace || (n > 0 && (!x || !idx || !vals || !wts || !zq || !out)))
        return fail(MTVQ_ERR_BAD_ARG, "null pointer%s", "");
    if (!aligned16(x) || !aligned16(codes) || !aligned16(zq) || !aligned16(out) || !aligned16(workspace))
        return fail(MTVQ_ERR_BAD_ARG, "16-byte alignment%s", "");
    DeviceInfo di;
    if (int r = device_info(&di)) return r;
    Workspace ws = carve(workspace, n, d, k, topk);
    if (ws.total_bytes > workspace_bytes) return fail(MTVQ_ERR_WORKSPACE, "workspace too small%s", "");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MTVQ_CUDA(cudaMemsetAsync(ws.counters, 0, sizeof(int) * CNT_COUNT, st), "memset counters");
    if (n > 0 && g_soft_route == 0 && !g_no_small && small_route_ok(n, d, k, MTVQ_NORM_CONTIGUOUS)) {
        // small batch (MedTok's training batches): exact scores on the CUDA cores, k-way merge, then the usual row kernel
        MTVQ_CUDA(launch_small_assign(x, n, d, MTVQ_NORM_CONTIGUOUS, codes, k, 1, topk, ws.codes_n, ws.cc, ws.denom,
                                      ws.small_part, ws.cand, ws.cand_cnt, cand_capacity(true), nullptr, nullptr, nullptr,
                                      ws.counters, st), "small_assign");
        MTVQ_CUDA(launch_soft_resolve(x, n, d, k, topk, ws, idx, vals, wts, zq, out, sse, di.num_sms, st), "soft_resolve");
        return 0;
    }
    MTVQ_CUDA(launch_codebook_prep(codes, k, d, 1, ws, st), "codebook_prep(normalise)");
    if (n == 0) return 0;
    MTVQ_CUDA(cudaMemsetAsync(ws.cand_cnt, 0, sizeof(int) * n, st), "memset cand_cnt");
    CoarseParams p;
    fill_coarse(p, x, n, d, MTVQ_NORM_CONTIGUOUS, k, ws, ws.denom, topk, true);
    const int rt = coarse_row_tiles(d, n, di.num_sms);
    // ONE tensor pass: every thread tracks its best granules (scores kept as fp16 records in an L2-resident scratch)
    // and the row block's end turns them into the per-row candidate lists.  Rows whose tracking overflowed (more than
    // eight qualifying granules in one thread's share, or more than cand_cap candidates) are listed in flag_rows with
    // their k-th score bound in best_val and get their list from a COLLECT pass -- the row count is only known on the
    // device, the kernel splits the codebook over CTAs to fill the SMs.
    if (g_soft_route == 2) {
        // cross-check route: no tensor pass, every row is scored against every code by the exact CUDA-core scan
        MTVQ_CUDA(launch_norms(x, n, d, ws.denom, di.num_sms, st), "norms");
    } else if (g_soft_route == 1) {
        MTVQ_CUDA(launch_coarse(p, MODE_THRESH, rt, di.num_sms, st), "coarse thresh");
        CoarseParams c = p;
        c.denom = nullptr;
        MTVQ_CUDA(launch_coarse(c, MODE_COLLECT, rt, di.num_sms, st), "coarse collect");
    } else {
        p.over_rows = ws.flag_rows;
        if (g_ev_before) cudaEventRecord(g_ev_before, st);
        MTVQ_CUDA(launch_coarse(p, MODE_TOPK, rt, di.num_sms, st), "coarse topk");
        if (g_ev_after) cudaEventRecord(g_ev_after, st);
        g_ev_before = g_ev_after = nullptr;
        CoarseParams c = p;
        c.row_list = ws.flag_rows;
        c.n_rows_dev = ws.counters + CNT_OVER;
        c.denom = nullptr;
        MTVQ_CUDA(launch_coarse(c, MODE_COLLECT, 1, di.num_sms, st), "coarse collect(overflow)");
    }
    MTVQ_CUDA(launch_soft_resolve(x, n, d, k, topk, ws, idx, vals, wts, zq, out, sse, di.num_sms, st), "soft_resolve");
    return 0;
}

int mtvq_soft_backward(const float* x, int64_t n, int d, const float* codes, int k, int topk, const int64_t* idx,
                       const float* wts, const float* g_zq, float* grad_x, float* grad_codes, void* workspace,
                       size_t workspace_bytes, void* stream) {
    if (int r = check_dims(n, d, k)) return r;
    if (topk < 1 || topk > 8) return fail(MTVQ_ERR_UNSUPPORTED, "top-k must be in [1, 8]%s", "");
    if (!x || !codes || !idx || !wts || !g_zq || !grad_x || !grad_codes) return fail(MTVQ_ERR_BAD_ARG, "null pointer%s", "");
    DeviceInfo di;
    if (int r = device_info(&di)) return r;
    (void)workspace; (void)workspace_bytes;
    if (n > 0)
        MTVQ_CUDA(launch_soft_backward(x, n, d, codes, k, topk, idx, wts, g_zq, grad_x, grad_codes, di.num_sms,
                                       static_cast<cudaStream_t>(stream)), "soft_backward");
    return 0;
}

int mtvq_soft_backward_fused(const float* x, int64_t n, int d, const float* codes, int k, int topk, const int64_t* idx,
                             const float* wts, const float* g_out, const float* g_vq, const float* g_commit, double beta,
                             int training, float* grad_x, float* grad_codes, void* workspace, size_t workspace_bytes,
                             void* stream) {
    if (int r = check_dims(n, d, k)) return r;
    if (topk < 1 || topk > 8) return fail(MTVQ_ERR_UNSUPPORTED, "top-k must be in [1, 8]%s", "");
    if (!codes || !grad_codes || !workspace || (n > 0 && (!x || !idx || !wts || !grad_x)))
        return fail(MTVQ_ERR_BAD_ARG, "null pointer%s", "");
    if (!aligned16(x) || !aligned16(codes) || !aligned16(grad_x) || !aligned16(grad_codes) || !aligned16(workspace) ||
        (g_out && !aligned16(g_out)))
        return fail(MTVQ_ERR_BAD_ARG, "16-byte alignment%s", "");
    const size_t need = sizeof(float) * ((size_t)k * d + (size_t)pad_codes(k));
    if (workspace_bytes < need) return fail(MTVQ_ERR_WORKSPACE, "workspace too small%s: need %ld", "", (long)need);
    DeviceInfo di;
    if (int r = device_info(&di)) return r;
    float* codes_n = static_cast<float*>(workspace);
    float* norms = codes_n + (size_t)k * d;
    if (n > 0)
        MTVQ_CUDA(launch_soft_backward_fused(x, n, d, codes, k, topk, idx, wts, g_out, g_vq, g_commit, (float)beta, training,
                                             grad_x, grad_codes, codes_n, norms, di.num_sms,
                                             static_cast<cudaStream_t>(stream)), "soft_backward_fused");
    return 0;
}

int mtvq_debug_counters(const void* workspace, int64_t n, int d, int k, int topk, int64_t* out_host, void* stream) {
    if (!workspace || !out_host) return fail(MTVQ_ERR_BAD_ARG, "null pointer%s", "");
    Workspace ws = carve(const_cast<void*>(workspace), n, d, k, topk < 0 ? 0 : topk);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int h[CNT_COUNT] = {0};
    cudaError_t e = cudaMemcpyAsync(h, ws.counters, sizeof(h), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return cuda_fail(e, "memcpy counters");
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "sync");
    out_host[0] = h[CNT_FLAGGED];
    out_host[1] = h[CNT_FULL];
    out_host[2] = h[CNT_CANDS];
    out_host[3] = g_kernel_launches;
    g_kernel_launches = 0;
    return 0;
}

int mtvq_debug_coarse_scores(const float* z, int64_t n, int d, int norm_mode, const float* weight, int k, float* scores,
                             int64_t rows, int* kpad_host, void* workspace, size_t workspace_bytes, void* stream) {
    if (int r = check_dims(n, d, k)) return r;
    DeviceInfo di;
    if (int r = device_info(&di)) return r;
    Workspace ws = carve(workspace, n, d, k, 0);
    if (ws.total_bytes > workspace_bytes) return fail(MTVQ_ERR_WORKSPACE, "workspace too small%s", "");
    if (kpad_host) *kpad_host = pad_codes(k);
    if (!scores) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MTVQ_CUDA(cudaMemsetAsync(ws.counters, 0, sizeof(int) * CNT_COUNT, st), "memset counters");
    MTVQ_CUDA(launch_codebook_prep(weight, k, d, 0, ws, st), "codebook_prep");
    CoarseParams p;
    fill_coarse(p, z, n, d, norm_mode, k, ws, ws.denom, 1, false);
    p.dump = scores;
    p.dump_rows = rows < n ? rows : n;
    const char* rt_env = getenv("MTVQ_DEBUG_ROW_TILES");
    int rt = rt_env ? atoi(rt_env) : coarse_row_tiles(d, n, di.num_sms);
    if (rt != 1 && rt != 2 && rt != 4) rt = 1;
    if (rt * d > 256) rt = 256 / d;
    MTVQ_CUDA(launch_coarse(p, MODE_DUMP, rt, di.num_sms, st), "coarse dump");
    return 0;
}

int mtvq_debug_set_soft_route(int route) {
    if (route < 0 || route > 4) return fail(MTVQ_ERR_BAD_ARG, "route must be 0..4%s", "");
    if (route == 3) { g_no_small = true; return 0; }      // 3: small batches take the tensor route too (A/B, tests)
    if (route == 4) { g_no_small = false; return 0; }     // 4: back to automatic
    g_soft_route = route;
    return 0;
}

int mtvq_hard_route(int64_t n, int d, int k) {
    if (check_dims(n, d, k)) return 0;
    DeviceInfo di;
    if (device_info(&di)) return 0;
    return two_pass_ok(n, d, k, di.num_sms) ? 2 : 1;
}

int mtvq_debug_set_hard_route(int route) {
    if (route < -1 || route > 2) return fail(MTVQ_ERR_BAD_ARG, "route must be -1 (automatic), 0, 1 (single pass) or 2 (two-pass whenever it applies)%s", "");
    g_hard_route = route;
    return 0;
}

int mtvq_debug_get_hard_route(void) { return g_hard_route; }

int mtvq_debug_set_trace(void* trace_dev) {
    g_trace = static_cast<long long*>(trace_dev);
    return 0;
}

int mtvq_debug_set_events(void* event_before, void* event_after) {
    g_ev_before = static_cast<cudaEvent_t>(event_before);
    g_ev_after = static_cast<cudaEvent_t>(event_after);
    return 0;
}

#ifdef MTVQ_WITH_MICROBENCH
int mtvq_microbench(int which, int warps, int iters, double* result_host, void* stream) {
    DeviceInfo di;
    if (int r = device_info(&di)) return r;
    MTVQ_CUDA(run_microbench(which, warps, iters, di.num_sms, result_host, static_cast<cudaStream_t>(stream)), "microbench");
    return 0;
}
#endif

}  // extern "C"
