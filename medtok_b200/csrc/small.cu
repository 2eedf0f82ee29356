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
//   small_slice_kernel  CTA = one slice of 64 codes x ALL rows: (optionally normalises its codes exactly like the
//                       prep kernel and publishes them to codes_n / cc), 64 x 64 register-tiled distance blocks,
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

// shared-memory layout (floats): As[d][SR] | Bs[d][SC] | raw[SC or SR][d + 1] | dist[SR][SC + 1] | zz[SR] | cc[SC]
__host__ __device__ inline size_t small_smem_floats(int d) {
    return (size_t)d * SR + (size_t)d * SC + (size_t)SR * (d + 1) + (size_t)SR * (SC + 1) + SR + SC;
}

__device__ __forceinline__ bool less_vi(float v, int i, float bv, int bi) {     // NaN never wins; ties -> lower index
    return v < bv || (v == bv && i < bi);
}

}  // namespace

__global__ void __launch_bounds__(256) small_slice_kernel(const float* __restrict__ x, int n, int d, int norm_mode,
                                                           const float* __restrict__ codes, int k, int normalize,
                                                           int topk, float* __restrict__ codes_n_out,
                                                           float* __restrict__ cc_out, float* __restrict__ denom_out,
                                                           PartEntry* __restrict__ part, int nslices, int nan_first) {
    extern __shared__ __align__(16) float sm[];
    float* As   = sm;                                   // [d][SR]   normalised rows, d-major
    float* Bs   = As + (size_t)d * SR;                  // [d][SC]   (normalised) codes of the slice, d-major
    float* raw  = Bs + (size_t)d * SC;                  // [64][d + 1] staging for coalesced global loads
    float* dist = raw + (size_t)SR * (d + 1);           // [SR][SC + 1]
    float* zzs  = dist + SR * (SC + 1);                 // [SR]
    float* ccs  = zzs + SR;                             // [SC]
    const int tid = threadIdx.x;
    const int slice = blockIdx.x;
    const int code0 = slice * SC;
    const int ncodes = min(SC, k - code0);
    const bool dots = score_dot(norm_mode);

    // ---- the slice's codes: coalesced load, then one thread per code in the canonical orders of the prep kernel ----
    for (int e = tid; e < ncodes * d; e += 256) {
        const int c = e / d, i = e - c * d;
        raw[c * (d + 1) + i] = __ldg(codes + (size_t)code0 * d + e);
    }
    __syncthreads();
    if (tid < SC) {
        if (tid < ncodes) {
            const float* wr = raw + tid * (d + 1);
            float den = 1.0f;
            if (normalize) den = canon_denom([&](int i) { return wr[i]; }, d, 0);
            for (int i = 0; i < d; ++i) Bs[(size_t)i * SC + tid] = normalize ? __fdiv_rn(wr[i], den) : wr[i];
            const float c2 = canon_sumsq([&](int i) { return Bs[(size_t)i * SC + tid]; }, d);
            ccs[tid] = dots ? 0.0f : c2;
            if (cc_out) cc_out[code0 + tid] = c2;
        } else {
            for (int i = 0; i < d; ++i) Bs[(size_t)i * SC + tid] = 0.0f;
            ccs[tid] = 0.0f;
        }
    }
    __syncthreads();
    if (codes_n_out && normalize) {                     // publish the normalised rows (soft_resolve reads them)
        for (int e = tid; e < ncodes * d; e += 256) {
            const int c = e / d, i = e - c * d;
            codes_n_out[(size_t)code0 * d + e] = Bs[(size_t)i * SC + c];
        }
    }

    const int tx = tid & 15, ty = tid >> 4;             // 4 codes x 4 rows per thread
    for (int row0 = 0; row0 < n; row0 += SR) {
        const int nrows = min(SR, n - row0);
        __syncthreads();                                // previous tile: raw / As / dist are free again
        for (int e = tid; e < nrows * d; e += 256) {
            const int r = e / d, i = e - r * d;
            raw[r * (d + 1) + i] = __ldg(x + (size_t)row0 * d + e);
        }
        __syncthreads();
        if (tid < SR) {
            if (tid < nrows) {
                const float* xr = raw + tid * (d + 1);
                const float den = canon_denom([&](int i) { return xr[i]; }, d, norm_mode);
                if (denom_out && slice == 0) denom_out[row0 + tid] = den;
                for (int i = 0; i < d; ++i) As[(size_t)i * SR + tid] = __fdiv_rn(xr[i], den);
                zzs[tid] = dots ? 0.0f : canon_sumsq([&](int i) { return As[(size_t)i * SR + tid]; }, d);
            } else {
                for (int i = 0; i < d; ++i) As[(size_t)i * SR + tid] = 0.0f;
                zzs[tid] = 0.0f;
            }
        }
        __syncthreads();
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
            float lv[kMaxTopS];
            int li[kMaxTopS];
            bool saw_nan = false;
#pragma unroll
            for (int t = 0; t < kMaxTopS; ++t) { lv[t] = INFINITY; li[t] = 0x7fffffff; }
            for (int c = q * 16; c < q * 16 + 16; ++c) {
                if (c >= ncodes) break;
                float v = dist[r * (SC + 1) + c];
                int ci = code0 + c;
                if (v != v) { saw_nan = true; continue; }   // NaN never enters a list (hard path: see below)
#pragma unroll
                for (int t = 0; t < kMaxTopS; ++t) {
                    if (t < topk && less_vi(v, ci, lv[t], li[t])) {
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
                    for (int u = 0; u + 1 < kMaxTopS; ++u) { lv[u] = lv[u + 1]; li[u] = li[u + 1]; }
                    lv[kMaxTopS - 1] = INFINITY; li[kMaxTopS - 1] = 0x7fffffff;
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
__global__ void __launch_bounds__(256) small_merge_kernel(const PartEntry* __restrict__ part, int n, int nslices, int topk,
                                                           int* __restrict__ cand, int* __restrict__ cand_cnt, int cand_cap,
                                                           int64_t* __restrict__ idx_out, int* __restrict__ bins,
                                                           int* __restrict__ full_rows, int* __restrict__ counters) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    float lv[kMaxTopS];
    int li[kMaxTopS];
    bool nan_row = false;
#pragma unroll
    for (int t = 0; t < kMaxTopS; ++t) { lv[t] = INFINITY; li[t] = 0x7fffffff; }
    const PartEntry* src = part + (size_t)row * nslices * topk;
    for (int e = lane; e < nslices * topk; e += 32) {
        PartEntry p = src[e];
        float v = p.v; int ci = p.i;
        if (ci == kNanMark) { nan_row = true; continue; }
        if (ci == 0x7fffffff) continue;
#pragma unroll
        for (int t = 0; t < kMaxTopS; ++t) {
            if (t < topk && less_vi(v, ci, lv[t], li[t])) {
                const float tv = lv[t]; const int ti = li[t];
                lv[t] = v; li[t] = ci; v = tv; ci = ti;
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
            for (int u = 0; u + 1 < kMaxTopS; ++u) { lv[u] = lv[u + 1]; li[u] = li[u + 1]; }
            lv[kMaxTopS - 1] = INFINITY; li[kMaxTopS - 1] = 0x7fffffff;
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

size_t small_scratch_bytes(int64_t n, int k, int topk) {
    const size_t nslices = (size_t)(k + SC - 1) / SC;
    return (size_t)n * nslices * (size_t)(topk < 1 ? 1 : topk) * sizeof(PartEntry);
}

bool small_route_ok(int64_t n, int d, int k) {
    // work bound: below it the direct fp32 scan beats the tensor pass's fixed cost (tools_small_batch_probe.py); the
    // second bound keeps the partial-list scratch at a few MB
    return n > 0 && n <= 4096 && (double)n * (double)k * (double)d <= 1.6e9 && (double)n * (double)k <= 3.0e7;
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
        cudaError_t e = cudaFuncSetAttribute(small_slice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    PartEntry* part = static_cast<PartEntry*>(scratch);
    small_slice_kernel<<<nslices, 256, smem, stream>>>(x, (int)n, d, norm_mode, codes, k, normalize, topk, codes_n_out, cc_out,
                                                       denom_out, part, nslices, idx_out != nullptr);
    ++g_kernel_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    small_merge_kernel<<<(int)((n * 32 + 255) / 256), 256, 0, stream>>>(part, (int)n, nslices, topk, cand, cand_cnt, cand_cap,
                                                                         idx_out, bins, full_rows, counters);
    ++g_kernel_launches;
    return cudaGetLastError();
}

}  // namespace mtvq
