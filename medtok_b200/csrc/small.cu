// small.cu -- the small-batch route (a few hundred to a few thousand rows per call: what MedTok's training loop
// actually feeds the quantizer -- batch 256-1024 per GPU, six quantisations per step, train_MedTok.py:387,
// tokenizer.py:199).
//
// At that size the persistent tensor-core pass is all fixed cost (tensor-memory allocation, pipeline fill, a
// candidate pass plus an exact re-scoring pass, a dozen dependent launches: 120-150 us whatever N).  The work itself
// is tiny (N K D <= ~1.5e9 multiply-adds), so this route scores every (row, code) pair DIRECTLY in the reference's
// arithmetic -- sequential fmaf chain over d, fl(fl(zz + cc) - 2 dot) -- on the CUDA cores, spread over the whole GPU
// by code slice; nothing is approximate, so there is no candidate list, no margin and no resolve step:
//
//   small_prep_kernel   one warp per row and per code: the canonical l2-normalisation / squared norms (same operation
//                       order as the loader and prep kernels of the tensor route), rows stored d-major for the tiles
//   small_slice_kernel  CTA = one slice of 64 codes x a group of 64-row tiles: 64 x 64 register-tiled distance blocks,
//                       per-row ascending top-k of the slice -> partial lists in an L2-resident scratch
//   small_merge_kernel  one warp per row: k-way merge of the slices' lists -> the row's final top-k / argmin
//
// The soft path hands the merged indices to soft_resolve_kernel (softmax, weighted sum, straight-through, losses) as
// an exactly-k candidate list; the hard path gets idx + per-code counts directly.
#include "small.h"
#include "common.cuh"

namespace mtvq {

namespace {

constexpr int SR = 64;        // rows per register-tiled block
constexpr int SC = 64;        // codes per slice (= per CTA)
constexpr int kMaxTopS = 8;
constexpr int kNanMark = -2;  // list entry: "this slice saw a NaN distance for the row" (hard path only)

struct PartEntry { float v; int i; };

// shared-memory layout (floats): As[d][SR] | Bs[d][SC] | raw[64][d + 1] | dist[SR][SC + 1] | zz[SR] | cc[SC]
__host__ __device__ inline size_t small_smem_floats(int d) {
    return (size_t)d * SR + (size_t)d * SC + (size_t)SR * (d + 1) + (size_t)SR * (SC + 1) + SR + SC;
}

__device__ __forceinline__ bool less_vi(float v, int i, float bv, int bi) {     // NaN never wins; ties -> lower index
    return v < bv || (v == bv && i < bi);
}

// One warp normalises one row held in shared memory IN PLACE exactly like canon_denom / canon_sumsq do with one thread:
// the eight lane-accumulators of canon_norm8 run on lanes 0-7 and are folded in lane order, every lane then divides its
// elements, and warp_canon_sumsq folds the 32 cascade accumulators.  Returns den (all lanes) and *sumsq.
// kind: 0 = contiguous norm, 2 = none (do_norm false).
__device__ __forceinline__ float warp_normalise_row(float* row, int d, bool do_norm, int lane, float* sumsq) {
    float den = 1.0f;
    if (do_norm) {
        const int v = lane & 7;
        float a = 0.f;
        for (int c = 0; c < d; c += 8) { const float x = row[c + v]; a = __fadd_rn(a, __fmul_rn(x, x)); }
        float sm = __shfl_sync(0xffffffffu, a, 0);
#pragma unroll
        for (int t = 1; t < 8; ++t) sm = __fadd_rn(sm, __shfl_sync(0xffffffffu, a, t));
        den = __fsqrt_rn(sm);
        den = den < kNormEps ? kNormEps : den;
        __syncwarp();
        for (int i = lane; i < d; i += 32) row[i] = __fdiv_rn(row[i], den);
        __syncwarp();
    }
    *sumsq = warp_canon_sumsq(row, d, lane);
    return den;
}

// raw[64][d + 1] (row-major) -> T[d][64] (d-major); rows >= nvalid become zero.  Conflict free on both sides.
__device__ __forceinline__ void transpose_tile(const float* raw, float* T, int d, int nvalid, int tid) {
    const int r = tid & 63;
    for (int i = tid >> 6; i < d; i += 4) T[(size_t)i * 64 + r] = r < nvalid ? raw[r * (d + 1) + i] : 0.0f;
}

// global [rows][d] -> raw[64][d + 1] for a 16-byte vector `e` of the tile (four scalar stores: the odd row stride)
__device__ __forceinline__ void stage_vec(float* raw, int e, int nv, int d, float4 v) {
    const int r = e / nv, i = 4 * (e - r * nv);
    float* dst = raw + r * (d + 1) + i;
    dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
}

}  // namespace

// Pre-pass: one WARP per row and per code, in the canonical orders (so every slice CTA below reads finished operands
// instead of re-deriving them): rows -> xnT[d][n_pad] (d-major: a 64-row tile of As is then a straight coalesced copy),
// zz[n], denom[n]; codes -> cn[k][d] (when normalising), cc[k].
__global__ void __launch_bounds__(256) small_prep_kernel(const float* __restrict__ x, int n, int n_pad, int d, int norm_mode,
                                                          const float* __restrict__ codes, int k, int normalize,
                                                          float* __restrict__ xnT, float* __restrict__ zz_out,
                                                          float* __restrict__ denom_out, float* __restrict__ cn_out,
                                                          float* __restrict__ cc_out) {
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float* row = sm + wib * d;
    const int item = blockIdx.x * 8 + wib;
    if (item >= n + k) return;
    const bool is_row = item < n;
    const float* src = is_row ? x + (size_t)item * d : codes + (size_t)(item - n) * d;
    for (int i = lane; i < d; i += 32) row[i] = __ldg(src + i);
    __syncwarp();
    float ss;
    const bool do_norm = is_row ? (norm_kind(norm_mode) != 2) : (normalize != 0);
    const float den = warp_normalise_row(row, d, do_norm, lane, &ss);
    if (is_row) {
        for (int i = lane; i < d; i += 32) xnT[(size_t)i * n_pad + item] = row[i];
        if (lane == 0) { zz_out[item] = ss; if (denom_out) denom_out[item] = den; }
    } else {
        const int c = item - n;
        if (cn_out && normalize) for (int i = lane; i < d; i += 32) cn_out[(size_t)c * d + i] = row[i];
        if (lane == 0 && cc_out) cc_out[c] = ss;
    }
}

template <int KT>   // list length kept per thread: 1 (argmin), 5 (MedTok's k) or 8
__global__ void __launch_bounds__(256) small_slice_kernel(const float* __restrict__ xnT, const float* __restrict__ zz_in, int n,
                                                           int n_pad, int d, int norm_mode, const float* __restrict__ cn,
                                                           const float* __restrict__ cc_in, int k, int topk,
                                                           PartEntry* __restrict__ part, int nslices, int nan_first,
                                                           int rows_per_cta) {
    extern __shared__ __align__(16) float sm[];
    float* As   = sm;                                   // [d][SR]   normalised rows, d-major
    float* Bs   = As + (size_t)d * SR;                  // [d][SC]   (normalised) codes of the slice, d-major
    float* raw  = Bs + (size_t)d * SC;                  // [64][d + 1] staging (row stride d + 1: conflict-free transpose)
    float* dist = raw + (size_t)SR * (d + 1);           // [SR][SC + 1]
    float* zzs  = dist + SR * (SC + 1);                 // [SR]
    float* ccs  = zzs + SR;                             // [SC]
    const int tid = threadIdx.x;
    const int slice = blockIdx.x;
    const int code0 = slice * SC;
    const int ncodes = min(SC, k - code0);
    const bool dots = score_dot(norm_mode);
    const int nv = d >> 2;                              // float4 per row

    // grid.y splits the rows: a small batch must still fill the GPU (a CTA is 8 warps; with one CTA per SM every
    // dependent chain's latency is exposed)
    const int row_lo = blockIdx.y * rows_per_cta, row_hi = min(n, row_lo + rows_per_cta);
    // As tile = rows [row0, row0 + 64) of xnT: for every dimension 64 consecutive floats -> 16 d float4, d / 16 per thread
    const int a_vecs = d * (SR / 4);
    auto load_a = [&](int row0) {
        for (int e0 = tid; e0 < a_vecs; e0 += 4 * 256) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = e0 + u * 256;
                if (e < a_vecs) {
                    const int i = e >> 4, r4 = e & 15;                       // dimension, float4 within the 64-row span
                    v[u] = __ldg(reinterpret_cast<const float4*>(xnT + (size_t)i * n_pad + row0) + r4);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = e0 + u * 256;
                if (e < a_vecs) reinterpret_cast<float4*>(As)[e] = v[u];
            }
        }
    };
    // ---- the slice's codes (finished by the pre-pass): coalesced 16-byte loads, transposed through `raw` ----
    {
        const float4* src = reinterpret_cast<const float4*>(cn + (size_t)code0 * d);
        const int tot = ncodes * nv;
        for (int e0 = tid; e0 < tot; e0 += 4 * 256) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) if (e0 + u * 256 < tot) v[u] = __ldg(src + e0 + u * 256);
#pragma unroll
            for (int u = 0; u < 4; ++u) if (e0 + u * 256 < tot) stage_vec(raw, e0 + u * 256, nv, d, v[u]);
        }
        if (tid < SC) ccs[tid] = (tid < ncodes && !dots) ? __ldg(cc_in + code0 + tid) : 0.0f;
    }
    load_a(row_lo);                                     // (xnT is padded to whole tiles: no bounds inside a tile)
    if (tid < SR) zzs[tid] = (row_lo + tid < n && !dots) ? __ldg(zz_in + row_lo + tid) : 0.0f;
    __syncthreads();
    transpose_tile(raw, Bs, d, ncodes, tid);
    __syncthreads();

    const int tx = tid & 15, ty = tid >> 4;             // 4 codes x 4 rows per thread
    for (int row0 = row_lo; row0 < row_hi; row0 += SR) {
        const int nrows = min(SR, n - row0);
        if (row0 != row_lo) {
            __syncthreads();                            // previous tile: As / zz / dist are free again
            load_a(row0);
            if (tid < SR) zzs[tid] = (row0 + tid < n && !dots) ? __ldg(zz_in + row0 + tid) : 0.0f;
            __syncthreads();
        }
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
#pragma unroll 4
        for (int dd = 0; dd < d; ++dd) {                // sequential in d: the reference's fmaf chain
            const float4 a4 = *reinterpret_cast<const float4*>(As + (size_t)dd * SR + ty * 4);
            const float4 b4 = *reinterpret_cast<const float4*>(Bs + (size_t)dd * SC + tx * 4);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int ai = 0; ai < 4; ++ai)
#pragma unroll
                for (int bi = 0; bi < 4; ++bi) acc[ai][bi] = __fmaf_rn(a[ai], b[bi], acc[ai][bi]);
        }
#pragma unroll
        for (int ai = 0; ai < 4; ++ai) {
            const float zz = zzs[ty * 4 + ai];
#pragma unroll
            for (int bi = 0; bi < 4; ++bi)
                dist[(ty * 4 + ai) * (SC + 1) + tx * 4 + bi] = canon_dist(zz, ccs[tx * 4 + bi], acc[ai][bi]);
        }
        __syncthreads();
        // per-row ascending top-k of the slice: four threads per row scan 16 codes each, then merge through shuffles
        {
            const int r = tid >> 2, q = tid & 3;        // row within the tile, quarter of the slice
            float lv[KT];
            int li[KT];
            bool saw_nan = false;
#pragma unroll
            for (int t = 0; t < KT; ++t) { lv[t] = INFINITY; li[t] = 0x7fffffff; }
            for (int c = q * 16; c < q * 16 + 16; ++c) {
                if (c >= ncodes) break;
                float v = dist[r * (SC + 1) + c];
                int ci = code0 + c;
                if (v != v) { saw_nan = true; continue; }   // NaN never enters a list (hard path: see below)
                if (!less_vi(v, ci, lv[KT - 1], li[KT - 1])) continue;      // not among the KT best so far
#pragma unroll
                for (int t = 0; t < KT; ++t) {
                    if (less_vi(v, ci, lv[t], li[t])) {
                        const float tv = lv[t]; const int ti = li[t];
                        lv[t] = v; li[t] = ci; v = tv; ci = ti;
                    }
                }
            }
            // k rounds of a 4-lane argmin over the list heads (the four threads of a row are adjacent lanes)
            PartEntry* dst = part + ((size_t)(row0 + r) * nslices + slice) * topk;
            for (int t = 0; t < topk; ++t) {
                float bv = lv[0]; int bi = li[0];
#pragma unroll
                for (int o = 2; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (less_vi(ov, oi, bv, bi) || (bi == 0x7fffffff && oi != 0x7fffffff)) { bv = ov; bi = oi; }
                }
                if (li[0] == bi && bi != 0x7fffffff) {  // owner pops its head
#pragma unroll
                    for (int u = 0; u + 1 < KT; ++u) { lv[u] = lv[u + 1]; li[u] = li[u + 1]; }
                    lv[KT - 1] = INFINITY; li[KT - 1] = 0x7fffffff;
                }
                if (q == 0 && r < nrows) { dst[t].v = bv; dst[t].i = bi; }
            }
            // torch.argmin returns the first NaN: the hard path re-scores such rows with the exact scan (kNanMark)
            const unsigned nan4 = (__ballot_sync(0xffffffffu, saw_nan) >> ((tid & 31) & ~3)) & 0xfu;
            if (nan_first && nan4 && q == 0 && r < nrows) dst[0].i = kNanMark;
        }
    }
}

// one warp per row: merge the slices' ascending lists.  Soft path: cand[row] = the k winners (soft_resolve_kernel
// re-derives their distances and does the rest); hard path (idx_out != nullptr): idx + per-code counts.
template <int KT>
__global__ void __launch_bounds__(256) small_merge_kernel(const PartEntry* __restrict__ part, int n, int nslices, int topk,
                                                           int* __restrict__ cand, int* __restrict__ cand_cnt, int cand_cap,
                                                           int64_t* __restrict__ idx_out, int* __restrict__ bins,
                                                           int* __restrict__ full_rows, int* __restrict__ counters) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    float lv[KT];
    int li[KT];
    bool nan_row = false;
#pragma unroll
    for (int t = 0; t < KT; ++t) { lv[t] = INFINITY; li[t] = 0x7fffffff; }
    const PartEntry* src = part + (size_t)row * nslices * topk;
    const int tot = nslices * topk;
    for (int e0 = lane; e0 < tot; e0 += 8 * 32) {       // eight independent 8-byte loads in flight per lane
        PartEntry p[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            p[u].v = INFINITY; p[u].i = 0x7fffffff;
            if (e0 + u * 32 < tot) p[u] = src[e0 + u * 32];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            float v = p[u].v; int ci = p[u].i;
            if (ci == kNanMark) { nan_row = true; continue; }
            if (ci == 0x7fffffff || !less_vi(v, ci, lv[KT - 1], li[KT - 1])) continue;   // empty / not among the best KT
#pragma unroll
            for (int t = 0; t < KT; ++t) {
                if (less_vi(v, ci, lv[t], li[t])) {
                    const float tv = lv[t]; const int ti = li[t];
                    lv[t] = v; li[t] = ci; v = tv; ci = ti;
                }
            }
        }
    }
    nan_row = __any_sync(0xffffffffu, nan_row);
    int found = 0;
    int first = 0x7fffffff;
    for (int t = 0; t < topk; ++t) {
        float bv = lv[0]; int bi = li[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (less_vi(ov, oi, bv, bi) || (bi == 0x7fffffff && oi != 0x7fffffff)) { bv = ov; bi = oi; }
        }
        if (li[0] == bi && bi != 0x7fffffff) {
#pragma unroll
            for (int u = 0; u + 1 < KT; ++u) { lv[u] = lv[u + 1]; li[u] = li[u + 1]; }
            lv[KT - 1] = INFINITY; li[KT - 1] = 0x7fffffff;
        }
        if (bi != 0x7fffffff) {
            ++found;
            if (t == 0) first = bi;
            if (lane == 0 && cand) cand[(size_t)row * cand_cap + t] = bi;
        }
    }
    if (lane != 0) return;
    // soft path: a list with fewer than k finite distances is unusable -> soft_resolve sends the row to its full scan
    if (cand_cnt) cand_cnt[row] = found;
    if (idx_out) {
        if (nan_row || found == 0) full_rows[atomicAdd(&counters[CNT_FULL], 1)] = row;   // exact scan decides (NaN first)
        else { idx_out[row] = (int64_t)first; if (bins) atomicAdd(&bins[first], 1); }
    }
}

bool small_route_ok(int64_t n, int d, int k, int norm_mode) {
    if ((norm_mode & 3) == 1) return false;   // the strided-view norm is a sequential sum: left to the tensor route
    return small_route_fits(n, d, k);
}

cudaError_t launch_small_assign(const float* x, int64_t n, int d, int norm_mode, const float* codes, int k, int normalize,
                                int topk, float* codes_n_out, float* cc_out, float* denom_out, void* scratch, int* cand,
                                int* cand_cnt, int cand_cap, int64_t* idx_out, int* bins, int* full_rows, int* counters,
                                cudaStream_t stream) {
    const int nslices = (k + SC - 1) / SC;
    const size_t smem = small_smem_floats(d) * sizeof(float);
    static PerDeviceOnce once;
    bool& configured = once.flag();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(small_slice_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(small_slice_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(small_slice_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    if (normalize && (!codes_n_out || !cc_out)) return cudaErrorInvalidValue;
    const int n_pad = (int)small_npad(n);
    PartEntry* part = static_cast<PartEntry*>(scratch);
    float* xnT = reinterpret_cast<float*>(static_cast<char*>(scratch) + small_part_bytes(n, k, topk));
    float* zz = xnT + (size_t)n_pad * d;
    // 1. canonical normalisation of rows and codes, one warp each
    small_prep_kernel<<<(int)((n + k + 7) / 8), 256, sizeof(float) * 8 * d, stream>>>(x, (int)n, n_pad, d, norm_mode, codes, k,
                                                                                      normalize, xnT, zz, denom_out,
                                                                                      codes_n_out, cc_out);
    ++g_kernel_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // 2. (slice, row group) grid: at least ~3 CTAs per SM when the problem allows, whole 64-row tiles per CTA
    auto kern = topk == 1 ? small_slice_kernel<1> : (topk <= 5 ? small_slice_kernel<5> : small_slice_kernel<8>);
    const int ntiles = (int)((n + SR - 1) / SR);
    int tiles_per_cta = (int)(((int64_t)ntiles * nslices) / 444);
    if (tiles_per_cta < 1) tiles_per_cta = 1;
    const int groups = (ntiles + tiles_per_cta - 1) / tiles_per_cta;
    kern<<<dim3(nslices, groups), 256, smem, stream>>>(xnT, zz, (int)n, n_pad, d, norm_mode, normalize ? codes_n_out : codes,
                                                      cc_out, k, topk, part, nslices, idx_out != nullptr, tiles_per_cta * SR);
    ++g_kernel_launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // 3. per-row merge of the slices' lists
    auto merge = topk == 1 ? small_merge_kernel<1> : (topk <= 5 ? small_merge_kernel<5> : small_merge_kernel<8>);
    merge<<<(int)((n * 32 + 255) / 256), 256, 0, stream>>>(part, (int)n, nslices, topk, cand, cand_cnt, cand_cap, idx_out, bins,
                                                            full_rows, counters);
    ++g_kernel_launches;
    return cudaGetLastError();
}

}  // namespace mtvq
