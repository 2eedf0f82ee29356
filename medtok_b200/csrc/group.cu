// group.cu -- glue kernels of the two-pass hard route (assign_tc.h: MODE_GROUP -> here -> MODE_REFINE -> here).
//
// The GROUP tensor pass leaves, per row, the 128-code chunk holding its best coarse score (`grp`, plus a second chunk
// for the few rows with a near-tie across chunks), that score (`best_val`), the rank of every entry within its (CTA,
// chunk) share and the table of share sizes (`ghist`).  Here the rows are counting-sorted by chunk into 128-entry tiles for the REFINE pass (norm_ema_quantizer.py
// :175-179 is still what is being computed: the argmin; the sort only makes the second look at a row a 128-code GEMM tile
// instead of a scan), and after the exact re-scoring of the near-tie rows their share of :181,202,212-214 is added.
#include "group.h"

#include "assign_tc.h"
#include "common.cuh"

namespace mtvq {

// one CTA: exclusive prefix sums of the entries and of the 128-entry tiles per chunk, then the tile table
__global__ void __launch_bounds__(1024) group_scan_kernel(const int* __restrict__ gtot, int nchunks, int* __restrict__ goff,
                                                           int* __restrict__ gtile, int2* __restrict__ tile_tab,
                                                           int* __restrict__ counters) {
    __shared__ int wtot_e[32], wtot_t[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int per = (nchunks + 1023) / 1024;
    const int lo = min(nchunks, (int)threadIdx.x * per), hi = min(nchunks, lo + per);
    int se = 0, stl = 0;
    for (int c = lo; c < hi; ++c) { const int h = gtot[c]; se += h; stl += (h + kTileRows - 1) / kTileRows; }
    int ie = se, it = stl;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int ve = __shfl_up_sync(0xffffffffu, ie, o), vt = __shfl_up_sync(0xffffffffu, it, o);
        if (lane >= o) { ie += ve; it += vt; }
    }
    if (lane == 31) { wtot_e[wid] = ie; wtot_t[wid] = it; }
    __syncthreads();
    if (wid == 0) {
        int te = wtot_e[lane], tt = wtot_t[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int ve = __shfl_up_sync(0xffffffffu, te, o), vt = __shfl_up_sync(0xffffffffu, tt, o);
            if (lane >= o) { te += ve; tt += vt; }
        }
        wtot_e[lane] = te; wtot_t[lane] = tt;
    }
    __syncthreads();
    int re = ie - se + (wid > 0 ? wtot_e[wid - 1] : 0);
    int rt = it - stl + (wid > 0 ? wtot_t[wid - 1] : 0);
    for (int c = lo; c < hi; ++c) {
        const int h = gtot[c];
        goff[c] = re; gtile[c] = rt;
        re += h; rt += (h + kTileRows - 1) / kTileRows;
    }
    if (threadIdx.x == 1023) { gtile[nchunks] = wtot_t[31]; counters[CNT_TILES] = wtot_t[31]; }
    __syncthreads();
    // tile table: warp w takes chunks w, w + 32, ...; its lanes write the chunk's tiles side by side
    for (int c = wid; c < nchunks; c += 32) {
        const int h = gtot[c], e0 = goff[c], t0 = gtile[c];
        const int nt = (h + kTileRows - 1) / kTileRows;
        for (int t = lane; t < nt; t += 32)
            tile_tab[t0 + t] = make_int2(c | (min(kTileRows, h - t * kTileRows) << 16), e0 + t * kTileRows);
    }
}

// rows -> 16-byte entries {row | flags, best coarse score, row norm, 0} at (first entry of the chunk) + (offset of the
// (CTA, chunk) share) + (rank within the share): the GROUP pass ranked every entry in shared memory and reserved its
// shares with one atomic per (CTA, chunk), so placing the entries needs no atomics at all
__global__ void __launch_bounds__(256) group_place_kernel(const int* __restrict__ grp, const int* __restrict__ grank,
                                                           const int* __restrict__ grank2, const float* __restrict__ best_val,
                                                           const float* __restrict__ denom, int64_t n, int nchunks, int nctas,
                                                           int rows_per_block, const int* __restrict__ share0,
                                                           const int* __restrict__ goff,
                                                           int4* __restrict__ entries) {
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += (int64_t)gridDim.x * blockDim.x) {
        const int g = grp[row];
        if (g & kGrpOverBit) continue;
        const int cta = (int)((row / rows_per_block) % nctas);       // the GROUP pass hands row block b to CTA b % nctas
        const int cb = g & ((1 << kGrpOtherShift) - 1);
        const int mbits = __float_as_int(best_val[row]), dbits = __float_as_int(denom[row]);
        const unsigned cross = (g & kGrpCrossBit) ? kEntCross : 0u;
        entries[goff[cb] + share0[(size_t)cta * nchunks + cb] + grank[row]] = make_int4((int)((unsigned)row | cross), mbits, dbits, 0);
        if (cross) {
            const int other = (g >> kGrpOtherShift) & ((1 << kGrpOtherShift) - 1);
            entries[goff[other] + share0[(size_t)cta * nchunks + other] + grank2[row]] =
                make_int4((int)((unsigned)row | kEntCross | kEntSecond), mbits, dbits, 0);
        }
    }
}

// After the exact scan: the row work (z_q, squared error, per-code sums) of the few rows only the full scan could decide
// (ws.full_rows; every other near-tie row was finished by resolve_kernel, every decided row by the REFINE pass), 16 lanes
// per row; and the packed statistics' counts (bins as doubles, rows), which need the histogram to be complete.
__global__ void __launch_bounds__(256) finish_flagged_kernel(const float* __restrict__ z, const float* __restrict__ denom,
                                                              const float* __restrict__ w, const int64_t* __restrict__ idx,
                                                              const int* __restrict__ rows, const int* __restrict__ n_rows_dev,
                                                              const int* __restrict__ bins, int64_t n, int d, int k,
                                                              float* __restrict__ z_q, double* __restrict__ stats) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < k; c += (int64_t)gridDim.x * blockDim.x)
        stats[c] += (double)bins[c];
    if (blockIdx.x == 0 && threadIdx.x == 0) stats[(size_t)k + (size_t)k * d + 1] += (double)n;
    const int nflag = *n_rows_dev;
    const int vpr = d >> 2, q = threadIdx.x & 15;
    const int ngroups = (gridDim.x * blockDim.x) >> 4;
    double* embed_sum = stats + k;
    double sse = 0.0;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 4; i < nflag; i += ngroups) {
        const int row = rows[i];
        const int code = (int)idx[row];
        if (q >= vpr || code < 0 || code >= k) continue;
        const float dn = denom[row];
        const float4 x = __ldg(reinterpret_cast<const float4*>(z + (size_t)row * d) + q);
        const float4 c = __ldg(reinterpret_cast<const float4*>(w + (size_t)code * d) + q);
        const float zf0 = __fdiv_rn(x.x, dn), zf1 = __fdiv_rn(x.y, dn), zf2 = __fdiv_rn(x.z, dn), zf3 = __fdiv_rn(x.w, dn);
        double* a = embed_sum + (size_t)code * d + 4 * q;
        atomicAdd(a + 0, (double)zf0); atomicAdd(a + 1, (double)zf1); atomicAdd(a + 2, (double)zf2); atomicAdd(a + 3, (double)zf3);
        const float d0 = __fsub_rn(c.x, zf0), d1 = __fsub_rn(c.y, zf1), d2 = __fsub_rn(c.z, zf2), d3 = __fsub_rn(c.w, zf3);
        sse += (double)__fmul_rn(d0, d0) + (double)__fmul_rn(d1, d1) + (double)__fmul_rn(d2, d2) + (double)__fmul_rn(d3, d3);
        float4 o;
        o.x = __fadd_rn(zf0, d0); o.y = __fadd_rn(zf1, d1); o.z = __fadd_rn(zf2, d2); o.w = __fadd_rn(zf3, d3);
        *(reinterpret_cast<float4*>(z_q + (size_t)row * d) + q) = o;
    }
    sse = warp_sum(sse);
    if ((threadIdx.x & 31) == 0 && sse != 0.0) atomicAdd(&stats[(size_t)k + (size_t)k * d], sse);
}

cudaError_t launch_group_sort(const Workspace& ws, const float* denom, int64_t n, int k, int num_sms, cudaStream_t s) {
    const int nchunks = pad_codes(k) / kTileCodes;
    group_scan_kernel<<<1, 1024, 0, s>>>(ws.gtot, nchunks, ws.goff, ws.gtile, static_cast<int2*>(ws.tile_tab), ws.counters);
    ++g_kernel_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int64_t want = (n + 255) / 256;
    const int blocks = (int)(want < (int64_t)num_sms * 16 ? want : (int64_t)num_sms * 16);
    group_place_kernel<<<blocks, 256, 0, s>>>(ws.grp, ws.grank, ws.grank2, ws.best_val, denom, n, nchunks, num_sms, 4 * kTileRows,
                                              ws.ghist, ws.goff, static_cast<int4*>(ws.entries));
    ++g_kernel_launches;
    return cudaGetLastError();
}

cudaError_t launch_finish_flagged(const float* z, const float* denom, const float* w, const int64_t* idx, int64_t n, int d,
                                  int k, float* z_q, double* stats, const Workspace& ws, int num_sms, cudaStream_t s) {
    finish_flagged_kernel<<<num_sms, 256, 0, s>>>(z, denom, w, idx, ws.full_rows, ws.counters + CNT_FULL, ws.bins_i32, n, d, k,
                                                  z_q, stats);
    ++g_kernel_launches;
    return cudaGetLastError();
}

}  // namespace mtvq
