// soft.cu -- row kernels of the soft-one (top-k) quantizer
// (vector_quantization_soft_one_new.py:147-185 shared core, :187-217 specific_embedding).
//   soft_resolve_kernel : exact fp32 distances of the collected candidates (reference order),
//                         ascending top-k, softmax(-d), weighted sum of normalised codes, STE, SSE
//   soft_backward_kernel: gradient of zq w.r.t. x (through the softmax weights and normalize) and
//                         w.r.t. the raw codebook rows
// One warp per row; lanes split candidates (resolve) or dimensions (gather / backward).
#include "soft.h"
#include "common.cuh"

namespace mtvq {

namespace {

struct Pick { float v; int i; };

// lexicographic (value, index) minimum across the warp; NaN never wins
__device__ __forceinline__ Pick warp_argmin(Pick p) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, p.v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, p.i, o);
        if (ov < p.v || (ov == p.v && oi < p.i)) { p.v = ov; p.i = oi; }
    }
    return p;
}

constexpr int kMaxTop = 8;

}  // namespace

// Eight lanes per row, four rows per warp.  A row's candidate codes (a handful, listed by the tensor pass) are staged
// into the group's shared-memory tile with coalesced 16-byte loads (8 lanes = one 128-byte line), eight candidates per
// batch, and lane g then runs candidate g's sequential fmaf chain out of shared memory; merge, softmax and the weighted
// sum stay inside the 8-lane group.  (One warp per row spent ~1500 issue slots per row with mostly idle lanes and was
// issue-bound at 2.2 ms for 2^20 rows.)  Rows whose list is unusable (fewer than top-k entries or overflow) are
// appended to `full_rows` for soft_full_scan_kernel.
constexpr int kSoftWarps = 8;
constexpr int kLaneList  = 3;     // candidates per lane: cand_cap (24) / 8 lanes

// lexicographic (value, index) minimum across an aligned 8-lane group; NaN never wins
__device__ __forceinline__ Pick group_argmin(Pick p) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, p.v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, p.i, o);
        if (ov < p.v || (ov == p.v && oi < p.i)) { p.v = ov; p.i = oi; }
    }
    return p;
}

// canon_sumsq (common.cuh) of a row in shared memory by an 8-lane group: lane v runs the four accumulator chains
// acc[0..3][v] in the reference order, folds them 1,2,3 -> 0, and the eight lane sums are added in lane order.
__device__ __forceinline__ float group_canon_sumsq(const float* xs, int d, int g) {
    const int nvec = d >> 3, n4 = nvec >> 2;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int c = 0; c < n4; ++c) {
        const float e0 = xs[(c * 4 + 0) * 8 + g], e1 = xs[(c * 4 + 1) * 8 + g];
        const float e2 = xs[(c * 4 + 2) * 8 + g], e3 = xs[(c * 4 + 3) * 8 + g];
        a0 = __fadd_rn(a0, __fmul_rn(e0, e0)); a1 = __fadd_rn(a1, __fmul_rn(e1, e1));
        a2 = __fadd_rn(a2, __fmul_rn(e2, e2)); a3 = __fadd_rn(a3, __fmul_rn(e3, e3));
    }
    for (int i = n4 * 4; i < nvec; ++i) { const float e = xs[i * 8 + g]; a0 = __fadd_rn(a0, __fmul_rn(e, e)); }
    a0 = __fadd_rn(a0, a1); a0 = __fadd_rn(a0, a2); a0 = __fadd_rn(a0, a3);
    float fin = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) fin = __fadd_rn(fin, __shfl_sync(0xffffffffu, a0, t, 8));
    return fin;
}

template <int NQ, int KT>   // NQ: 16-byte units of a row per lane of the group, ceil(d / 32); KT: top-k bound (KT == 5: k is 5)
__global__ void __launch_bounds__(kSoftWarps * 32) soft_resolve_kernel(
    const float* __restrict__ x, const float* __restrict__ denom, int64_t n, int d, int k, int topk_rt,
    const float* __restrict__ cn, const float* __restrict__ cc, const int* __restrict__ cand,
    const int* __restrict__ cand_cnt, int cand_cap, int64_t* __restrict__ idx_out, float* __restrict__ val_out,
    float* __restrict__ wts_out, float* __restrict__ zq_out, float* __restrict__ out, double* __restrict__ sse_out,
    int* __restrict__ full_rows, int* __restrict__ counters, int split) {
    extern __shared__ __align__(16) float soft_smem[];
    const int topk = (KT == 5) ? 5 : topk_rt;    // compile-time for MedTok's k = 5: the per-row loops below fully unroll
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int grp = lane >> 3, g = lane & 7;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nv = d >> 2;                       // float4 per row (<= 64)
    const int ds = d + 4;                        // tile row stride: 16-byte aligned, conflict free for 8 lanes x float4
    float* xs = soft_smem + (size_t)(wib * 4 + grp) * (d + 8 * ds);   // per group: normalised row | 8-candidate tile
    float* tile = xs + d;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* cn4 = reinterpret_cast<const float4*>(cn);
    double sse = 0.0;

    // A warp takes 32 consecutive rows at a time (candidate counts and norms: one coalesced load per lane), four rows
    // per step; while a step is being scored the raw rows and candidate lists of the next step are already in flight,
    // so no HBM round trip sits on the critical path (L2 gather of the candidates -> fmaf chain -> merge -> gather of
    // the top-k rows, which the staging has just pulled into L1).
    // (`split`: few rows -- a warp takes ONE four-row step of a 32-row block instead of all eight, so that a small batch
    //  spreads over the GPU instead of serialising eight dependent gather chains in a handful of warps)
    const int64_t nblk = (n + 31) >> 5;
    const int64_t nitems = split ? nblk * 8 : nblk;
    for (int64_t item = warp; item < nitems; item += nwarps) {
      const int64_t blk = split ? (item >> 3) : item;
      const int s_lo = split ? (int)(item & 7) : 0, s_hi = split ? s_lo + 1 : 8;
      const int64_t base = blk << 5;
      if (base + 4 * s_lo >= n) continue;
      int cnt_l = 0;
      float den_l = 1.0f;
      if (base + lane < n) { cnt_l = cand_cnt[base + lane]; den_l = denom[base + lane]; }
      float4 nx[NQ];
      int ncl[kLaneList];
      auto fetch = [&](int step) {
          const int64_t row = base + 4 * step + grp;
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
              const int f = g + 8 * q;
              nx[q] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (row < n && f < nv) nx[q] = __ldcs(x4 + row * nv + f);
          }
#pragma unroll
          for (int b = 0; b < kLaneList; ++b)
              ncl[b] = (row < n && g + 8 * b < cand_cap) ? __ldcs(cand + row * cand_cap + g + 8 * b) : -1;
      };
      fetch(s_lo);
      for (int step = s_lo; step < s_hi && base + 4 * step < n; ++step) {
        const int64_t row = base + 4 * step + grp;
        const int cnt = __shfl_sync(0xffffffffu, cnt_l, 4 * step + grp);
        const float den = __shfl_sync(0xffffffffu, den_l, 4 * step + grp);
        float4 xr[NQ];
        int cl[kLaneList];
#pragma unroll
        for (int q = 0; q < NQ; ++q) xr[q] = nx[q];
#pragma unroll
        for (int b = 0; b < kLaneList; ++b) cl[b] = ncl[b];
        __syncwarp();                                         // the previous step is done with xs / tile
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int f = g + 8 * q;
            if (f < nv)
                *reinterpret_cast<float4*>(xs + 4 * f) = make_float4(__fdiv_rn(xr[q].x, den), __fdiv_rn(xr[q].y, den),
                                                                     __fdiv_rn(xr[q].z, den), __fdiv_rn(xr[q].w, den));
        }
        if (step + 1 < s_hi && base + 4 * (step + 1) < n) fetch(step + 1);
        const bool active = row < n && cnt >= topk && cnt <= cand_cap;
        if (row < n && !active && g == 0) full_rows[atomicAdd(&counters[CNT_FULL], 1)] = (int)row;
        __syncwarp();
        const float xx = group_canon_sumsq(xs, d, g);        // ||xn||^2 in the reference's cascade order

        // each lane keeps an ascending (value, index) list of the candidates it scored
        float lv[kLaneList];
        int li[kLaneList];
#pragma unroll
        for (int t = 0; t < kLaneList; ++t) { lv[t] = INFINITY; li[t] = 0x7fffffff; }
        const int nbat = __reduce_max_sync(0xffffffffu, active ? (cnt + 7) >> 3 : 0);
#pragma unroll
        for (int b = 0; b < kLaneList; ++b) {
            if (b >= nbat) break;                             // warp-uniform
            const int nb = active ? min(8, max(0, cnt - 8 * b)) : 0;
            const int mine = g < nb ? cl[b] : -1;
            const int maxnb = __reduce_max_sync(0xffffffffu, nb);
            __syncwarp();                                     // the previous batch's chains are done with the tile
            for (int j0 = 0; j0 < maxnb; j0 += 2) {           // two candidates = 2 NQ independent 16-byte loads per lane
                float4 v[2][NQ];
                bool ok[2];
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const int col = __shfl_sync(0xffffffffu, mine, j0 + jj, 8);
                    ok[jj] = col >= 0 && col < k;
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const int f = g + 8 * q;
                        if (ok[jj] && f < nv) v[jj][q] = __ldg(cn4 + (size_t)col * nv + f);
                    }
                }
#pragma unroll
                for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const int f = g + 8 * q;
                        if (ok[jj] && f < nv) *reinterpret_cast<float4*>(tile + (j0 + jj) * ds + 4 * f) = v[jj][q];
                    }
            }
            __syncwarp();
            if (mine >= 0 && mine < k) {
                const float4* tr = reinterpret_cast<const float4*>(tile + g * ds);
                float dot = 0.0f;
#pragma unroll 8
                for (int i = 0; i < nv; ++i) {                // sequential fmaf chain over d (reference order)
                    const float4 xv = *reinterpret_cast<const float4*>(xs + 4 * i);
                    const float4 cv = tr[i];
                    dot = __fmaf_rn(xv.x, cv.x, dot);
                    dot = __fmaf_rn(xv.y, cv.y, dot);
                    dot = __fmaf_rn(xv.z, cv.z, dot);
                    dot = __fmaf_rn(xv.w, cv.w, dot);
                }
                float v = canon_dist(xx, __ldg(cc + mine), dot);
                int ci = mine;
#pragma unroll
                for (int t = 0; t < kLaneList; ++t) {
                    if (v < lv[t] || (v == lv[t] && ci < li[t])) {
                        const float tv = lv[t]; const int ti = li[t];
                        lv[t] = v; li[t] = ci; v = tv; ci = ti;
                    }
                }
            }
        }

        // merge: topk rounds of a group-wide argmin over the list heads
        float sv[KT];
        int si[KT], sl[KT];                                   // value, code, tile slot (= owner lane of the group)
#pragma unroll
        for (int t = 0; t < KT; ++t) {
            sl[t] = 0;
            if (t < topk) {
                Pick best = group_argmin(Pick{lv[0], li[0]});
                sv[t] = best.v; si[t] = best.i;
                const bool owner = li[0] == best.i && lv[0] == best.v;
                sl[t] = __ffs((__ballot_sync(0xffffffffu, owner) >> (8 * grp)) & 0xffu) - 1;
                if (owner) {                                  // owner pops its head
#pragma unroll
                    for (int u = 0; u + 1 < kLaneList; ++u) { lv[u] = lv[u + 1]; li[u] = li[u + 1]; }
                    lv[kLaneList - 1] = INFINITY; li[kLaneList - 1] = 0x7fffffff;
                }
            } else { sv[t] = INFINITY; si[t] = 0; }
        }
        // single batch everywhere in the warp (the normal case): the top-k rows are still in the groups' tiles
        const bool all_in_tile = __all_sync(0xffffffffu, !active || cnt <= 8);
        if (!active) continue;                                // (no warp-level primitive below this point)
        // softmax(-vals): exp(x - max) / sum with max = -vals[0]
        float wgt[KT];
        float s = 0.0f;
#pragma unroll
        for (int t = 0; t < KT; ++t) {
            wgt[t] = 0.0f;
            if (t < topk) { wgt[t] = expf(__fsub_rn(sv[0], sv[t])); s = __fadd_rn(s, wgt[t]); }
        }
#pragma unroll
        for (int t = 0; t < KT; ++t) if (t < topk) wgt[t] = __fdiv_rn(wgt[t], s);
        if (g < topk) {
            float v = 0.f, wv = 0.f; int ii = 0;
#pragma unroll
            for (int t = 0; t < KT; ++t) if (t == g) { v = sv[t]; wv = wgt[t]; ii = si[t]; }
            idx_out[row * topk + g] = ii;
            val_out[row * topk + g] = v;
            wts_out[row * topk + g] = wv;
        }
        // zq = sum_j w_j * cn[idx_j] (products, then a sequential sum over j); out = x + (zq - x)
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int f = g + 8 * q;
            if (f < nv) {
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                if (all_in_tile) {                            // warp-uniform
#pragma unroll
                    for (int t = 0; t < KT; ++t)
                        if (t < topk) {
                            const float4 c = *reinterpret_cast<const float4*>(tile + sl[t] * ds + 4 * f);
                            a.x = __fadd_rn(a.x, __fmul_rn(wgt[t], c.x)); a.y = __fadd_rn(a.y, __fmul_rn(wgt[t], c.y));
                            a.z = __fadd_rn(a.z, __fmul_rn(wgt[t], c.z)); a.w = __fadd_rn(a.w, __fmul_rn(wgt[t], c.w));
                        }
                } else {
#pragma unroll
                    for (int t = 0; t < KT; ++t)
                        if (t < topk) {
                            const float4 c = __ldg(cn4 + (size_t)min(si[t], k - 1) * nv + f);
                            a.x = __fadd_rn(a.x, __fmul_rn(wgt[t], c.x)); a.y = __fadd_rn(a.y, __fmul_rn(wgt[t], c.y));
                            a.z = __fadd_rn(a.z, __fmul_rn(wgt[t], c.z)); a.w = __fadd_rn(a.w, __fmul_rn(wgt[t], c.w));
                        }
                }
                const float4 xv = xr[q];
                const float4 df = make_float4(__fsub_rn(a.x, xv.x), __fsub_rn(a.y, xv.y), __fsub_rn(a.z, xv.z),
                                              __fsub_rn(a.w, xv.w));
                __stcs(reinterpret_cast<float4*>(zq_out) + row * nv + f, a);
                __stcs(reinterpret_cast<float4*>(out) + row * nv + f,
                       make_float4(__fadd_rn(xv.x, df.x), __fadd_rn(xv.y, df.y), __fadd_rn(xv.z, df.z),
                                   __fadd_rn(xv.w, df.w)));
                sse += (double)__fmul_rn(df.x, df.x) + (double)__fmul_rn(df.y, df.y) + (double)__fmul_rn(df.z, df.z) +
                       (double)__fmul_rn(df.w, df.w);
            }
        }
      }
    }
    sse = warp_sum(sse);
    __shared__ double part[kSoftWarps];
    if (lane == 0) part[wib] = sse;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part[i];
        if (t != 0.0) atomicAdd(sse_out, t);
    }
}

// Fallback for the rows listed in full_rows: every code is scored (one warp per row, lanes split the codes).
__global__ void __launch_bounds__(256) soft_full_scan_kernel(
    const float* __restrict__ x, const float* __restrict__ denom, int64_t n, int d, int k, int topk,
    const float* __restrict__ cn, const float* __restrict__ cc, const int* __restrict__ full_rows,
    int64_t* __restrict__ idx_out, float* __restrict__ val_out, float* __restrict__ wts_out, float* __restrict__ zq_out,
    float* __restrict__ out, double* __restrict__ sse_out, const int* __restrict__ counters) {
    __shared__ __align__(16) float xn_s[8][kMaxDim];          // one normalised row per warp
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float* xs = xn_s[wib];
    double sse = 0.0;

    const int nfull = counters[CNT_FULL];
    for (int64_t p = warp; p < nfull; p += nwarps) {
        const int64_t row = full_rows[p];
        const float* xr = x + row * d;
        const float den = denom[row];
        __syncwarp();
        for (int i = lane; i < d; i += 32) xs[i] = __fdiv_rn(__ldg(xr + i), den);   // xn, divided once
        __syncwarp();
        // ||xn||^2 in the reference's cascade order (every lane the same value; shared-memory broadcasts)
        const float xx = canon_sumsq([&](int i) { return xs[i]; }, d);

        // each lane keeps an ascending list of its own best `topk` (value, index) pairs
        float lv[kMaxTop];
        int li[kMaxTop];
#pragma unroll
        for (int t = 0; t < kMaxTop; ++t) { lv[t] = INFINITY; li[t] = 0x7fffffff; }
        for (int col = lane; col < k; col += 32) {
            const float4* cr = reinterpret_cast<const float4*>(cn + (size_t)col * d);
            float dot = 0.0f;
            for (int i = 0; i < d; i += 4) {                       // sequential fmaf chain over d (reference order)
                const float4 cv = __ldg(cr + (i >> 2));
                const float4 xv = *reinterpret_cast<const float4*>(xs + i);
                dot = __fmaf_rn(xv.x, cv.x, dot);
                dot = __fmaf_rn(xv.y, cv.y, dot);
                dot = __fmaf_rn(xv.z, cv.z, dot);
                dot = __fmaf_rn(xv.w, cv.w, dot);
            }
            float v = canon_dist(xx, cc[col], dot);
            int ci = col;
            // insertion keeping (value, index) ascending
#pragma unroll
            for (int t = 0; t < kMaxTop; ++t) {
                if (t < topk && (v < lv[t] || (v == lv[t] && ci < li[t]))) {
                    const float tv = lv[t]; const int ti = li[t];
                    lv[t] = v; li[t] = ci; v = tv; ci = ti;
                }
            }
        }

        // merge: topk rounds of a warp-wide argmin over the list heads
        float sv[kMaxTop];
        int si[kMaxTop];
#pragma unroll
        for (int t = 0; t < kMaxTop; ++t) {
            if (t < topk) {
                Pick best = warp_argmin(Pick{lv[0], li[0]});
                sv[t] = best.v; si[t] = best.i;
                if (li[0] == best.i && lv[0] == best.v) {   // owner pops its head
#pragma unroll
                    for (int u = 0; u + 1 < kMaxTop; ++u) { lv[u] = lv[u + 1]; li[u] = li[u + 1]; }
                    lv[kMaxTop - 1] = INFINITY; li[kMaxTop - 1] = 0x7fffffff;
                }
            } else { sv[t] = INFINITY; si[t] = 0; }
        }
        // softmax(-vals): exp(x - max) / sum with max = -vals[0]
        float wgt[kMaxTop];
        float s = 0.0f;
#pragma unroll
        for (int t = 0; t < kMaxTop; ++t) {
            wgt[t] = 0.0f;
            if (t < topk) { wgt[t] = expf(__fsub_rn(sv[0], sv[t])); s = __fadd_rn(s, wgt[t]); }
        }
#pragma unroll
        for (int t = 0; t < kMaxTop; ++t) if (t < topk) wgt[t] = __fdiv_rn(wgt[t], s);
        if (lane < topk) {
            float v = 0.f, wv = 0.f; int ii = 0;
#pragma unroll
            for (int t = 0; t < kMaxTop; ++t) if (t == lane) { v = sv[t]; wv = wgt[t]; ii = si[t]; }
            idx_out[row * topk + lane] = ii;
            val_out[row * topk + lane] = v;
            wts_out[row * topk + lane] = wv;
        }
        // zq = sum_j w_j * cn[idx_j] (products, then a sequential sum over j); out = x + (zq - x)
        for (int c = lane; c < d; c += 32) {
            float a = 0.0f;
#pragma unroll
            for (int t = 0; t < kMaxTop; ++t)
                if (t < topk) a = __fadd_rn(a, __fmul_rn(wgt[t], __ldg(cn + (size_t)min(si[t], k - 1) * d + c)));
            const float xv = __ldg(xr + c);
            const float diff = __fsub_rn(a, xv);
            zq_out[row * d + c] = a;
            out[row * d + c] = __fadd_rn(xv, diff);
            sse += (double)__fmul_rn(diff, diff);
        }
    }
    sse = warp_sum(sse);
    __shared__ double part[8];
    if (lane == 0) part[threadIdx.x >> 5] = sse;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part[i];
        atomicAdd(sse_out, t);
    }
}

// ------------------------------------------------------------------------------------------------
// backward.  zq = sum_j w_j c_j, w = softmax(-v), v_j = ||xn||^2 + ||c_j||^2 - 2 xn.c_j,
// xn = x/||x||, c_j = W_r/||W_r|| (r = idx_j):
//   a_j = g.c_j ; s = sum_m w_m a_m ; h_j = dL/dv_j = -w_j (a_j - s)
//   dL/dxn  = sum_j h_j (2 xn - 2 c_j)
//   dL/dc_j = w_j g + h_j (2 c_j - 2 xn)
//   dL/dx   = (I - xn xn^T)/||x|| dL/dxn ;  dL/dW_r += (I - c_j c_j^T)/||W_r|| dL/dc_j
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) soft_backward_kernel(const float* __restrict__ x, int64_t n, int d,
                                                             const float* __restrict__ codes, int k, int topk,
                                                             const int64_t* __restrict__ idx,
                                                             const float* __restrict__ wts,
                                                             const float* __restrict__ g_zq, float* __restrict__ grad_x,
                                                             float* __restrict__ grad_codes) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    constexpr int S = kMaxDim / 32;   // dims per lane
    for (int64_t row = warp; row < n; row += nwarps) {
        float xv[S], gv[S], gxn[S];
        float nx2 = 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int c = s * 32 + lane;
            xv[s] = c < d ? x[row * d + c] : 0.f;
            gv[s] = c < d ? g_zq[row * d + c] : 0.f;
            gxn[s] = 0.f;
            nx2 += xv[s] * xv[s];
        }
        nx2 = (float)warp_sum((double)nx2);
        const float nx = fmaxf(sqrtf(nx2), kNormEps);
#pragma unroll
        for (int s = 0; s < S; ++s) xv[s] /= nx;   // xn

        float a[kMaxTop], wj[kMaxTop], nc[kMaxTop];
        int rj[kMaxTop];
        float sdot = 0.f;
#pragma unroll
        for (int t = 0; t < kMaxTop; ++t) {
            a[t] = 0.f; wj[t] = 0.f; nc[t] = 1.f; rj[t] = 0;
            if (t < topk) {
                rj[t] = (int)idx[row * topk + t];
                wj[t] = wts[row * topk + t];
                float n2 = 0.f, dt = 0.f;
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const int c = s * 32 + lane;
                    const float cv = c < d ? codes[(size_t)rj[t] * d + c] : 0.f;
                    n2 += cv * cv;
                    dt += cv * gv[s];
                }
                n2 = (float)warp_sum((double)n2);
                dt = (float)warp_sum((double)dt);
                nc[t] = fmaxf(sqrtf(n2), kNormEps);
                a[t] = dt / nc[t];
                sdot += wj[t] * a[t];
            }
        }
#pragma unroll
        for (int t = 0; t < kMaxTop; ++t) {
            if (t < topk) {
                const float h = -wj[t] * (a[t] - sdot);
                float gc[S];
                float proj = 0.f;
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const int c = s * 32 + lane;
                    const float cj = c < d ? codes[(size_t)rj[t] * d + c] / nc[t] : 0.f;
                    gxn[s] += h * (2.f * xv[s] - 2.f * cj);
                    gc[s] = wj[t] * gv[s] + h * (2.f * cj - 2.f * xv[s]);
                    proj += gc[s] * cj;
                }
                proj = (float)warp_sum((double)proj);
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const int c = s * 32 + lane;
                    if (c < d) {
                        const float cj = codes[(size_t)rj[t] * d + c] / nc[t];
                        atomicAdd(&grad_codes[(size_t)rj[t] * d + c], (gc[s] - cj * proj) / nc[t]);
                    }
                }
            }
        }
        float px = 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s) px += gxn[s] * xv[s];
        px = (float)warp_sum((double)px);
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int c = s * 32 + lane;
            if (c < d) grad_x[row * d + c] = (gxn[s] - xv[s] * px) / nx;
        }
    }
}

// canonical row norms alone (cross-check route of the soft forward: no tensor pass has written `denom`)
__global__ void norms_kernel(const float* __restrict__ x, int64_t n, int d, float* __restrict__ denom) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const float* xr = x + row * d;
    denom[row] = canon_denom([&](int i) { return __ldg(xr + i); }, d, 0);
}

cudaError_t launch_norms(const float* x, int64_t n, int d, float* denom, int num_sms, cudaStream_t stream) {
    (void)num_sms;
    norms_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(x, n, d, denom);
    ++g_kernel_launches;
    return cudaGetLastError();
}

cudaError_t launch_soft_resolve(const float* x, int64_t n, int d, int k, int topk, const Workspace& ws, int64_t* idx,
                                float* vals, float* wts, float* zq, float* out, double* sse, int num_sms,
                                cudaStream_t stream) {
    const int cap = cand_capacity(true);
    if (cap > 8 * kLaneList) return cudaErrorInvalidValue;
    // shared memory: per 8-lane group the normalised row and an 8-candidate tile; d > 128 runs 4 warps per CTA
    const int warps = d <= 128 ? kSoftWarps : kSoftWarps / 2;
    const size_t smem = sizeof(float) * warps * 4 * ((size_t)d + 8 * (size_t)(d + 4));
    static PerDeviceOnce once;
    bool& configured = once.flag();
    if (!configured) {
        cudaError_t e = cudaSuccess;
#define MTVQ_SOFT_ATTR(NQ_, KT_)                                                                                   \
        if (e == cudaSuccess)                                                                                      \
            e = cudaFuncSetAttribute(soft_resolve_kernel<NQ_, KT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        MTVQ_SOFT_ATTR(2, 5) MTVQ_SOFT_ATTR(4, 5) MTVQ_SOFT_ATTR(8, 5) MTVQ_SOFT_ATTR(2, 8) MTVQ_SOFT_ATTR(4, 8) MTVQ_SOFT_ATTR(8, 8)
#undef MTVQ_SOFT_ATTR
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int per_sm = (int)((220 * 1024) / (smem + 1024)) < 1 ? 1 : (int)((220 * 1024) / (smem + 1024));
    int64_t blocks = (n + 32 * warps - 1) / (32 * warps);
    // small batches: one four-row step per warp when that still fits the resident warps
    const int split = ((n + 3) / 4 <= (int64_t)num_sms * per_sm * warps) ? 1 : 0;
    if (split) blocks = ((n + 31) / 32 * 8 + warps - 1) / warps;
    if (blocks > (int64_t)num_sms * per_sm) blocks = (int64_t)num_sms * per_sm;   // all resident; warps stride over 32-row blocks
#define MTVQ_SOFT_ARGS x, ws.denom, n, d, k, topk, ws.codes_n, ws.cc, ws.cand, ws.cand_cnt, cap, idx, vals, wts, zq, out, sse, \
                       ws.full_rows, ws.counters, split
#define MTVQ_SOFT_LAUNCH(NQ_)                                                                                      \
    do {                                                                                                           \
        if (topk == 5) soft_resolve_kernel<NQ_, 5><<<(int)blocks, warps * 32, smem, stream>>>(MTVQ_SOFT_ARGS);     \
        else           soft_resolve_kernel<NQ_, 8><<<(int)blocks, warps * 32, smem, stream>>>(MTVQ_SOFT_ARGS);     \
    } while (0)
    if (d <= 64)       MTVQ_SOFT_LAUNCH(2);
    else if (d <= 128) MTVQ_SOFT_LAUNCH(4);
    else               MTVQ_SOFT_LAUNCH(8);
#undef MTVQ_SOFT_LAUNCH
#undef MTVQ_SOFT_ARGS
    ++g_kernel_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // rows without a usable candidate list (device-side count, normally zero): every code is scored
    soft_full_scan_kernel<<<num_sms * 2, 256, 0, stream>>>(x, ws.denom, n, d, k, topk, ws.codes_n, ws.cc, ws.full_rows, idx,
                                                          vals, wts, zq, out, sse, ws.counters);
    ++g_kernel_launches;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// fused backward of one soft quantisation (training step of the quantisation core): everything the autograd node owes
//   x      <- dL/dout (straight-through)  - g_commit * beta * diff  + (I - xn xn^T)/||x|| dL/dxn
//   W_r    <- (I - c c^T)/||W_r|| dL/dc        (atomic float4 adds into the raw codebook's gradient)
// with diff = (zq - x) * 2/(n d), dL/dzq = g_vq * diff, zq = sum_j w_j c_j recomputed from the gathered rows (nothing of
// the forward but idx / weights is re-read).  Eight lanes per row, four rows per warp, fp32 group reductions.
// ------------------------------------------------------------------------------------------------
__global__ void soft_backward_prep_kernel(const float* __restrict__ codes, int k, int d, float* __restrict__ codes_n,
                                          float* __restrict__ norms) {
    // one warp per code: ||W_r|| (clamped like F.normalize) and the normalised row
    const int lane = threadIdx.x & 31;
    const int code = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (code >= k) return;
    float s = 0.f;
    for (int i = lane; i < d; i += 32) { const float v = codes[(size_t)code * d + i]; s = fmaf(v, v, s); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float nrm = fmaxf(sqrtf(s), kNormEps);
    if (lane == 0) norms[code] = nrm;
    for (int i = lane; i < d; i += 32) codes_n[(size_t)code * d + i] = codes[(size_t)code * d + i] / nrm;
}

__device__ __forceinline__ float group_sum8(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

template <int NQ, int KT>
__global__ void __launch_bounds__(256) soft_backward_fused_kernel(
    const float* __restrict__ x, int64_t n, int d, const float* __restrict__ codes_n, const float* __restrict__ norms,
    int k, int topk_rt, const int64_t* __restrict__ idx, const float* __restrict__ wts, const float* __restrict__ g_out,
    const float* __restrict__ g_vq_p, const float* __restrict__ g_commit_p, float beta, int training,
    float* __restrict__ grad_x, float* __restrict__ grad_codes) {
    const int topk = (KT == 5) ? 5 : topk_rt;
    const int lane = threadIdx.x & 31, grp = lane >> 3, g = lane & 7;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nv = d >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* go4 = reinterpret_cast<const float4*>(g_out);
    const float4* cn4 = reinterpret_cast<const float4*>(codes_n);
    const float scale = 2.0f / ((float)n * (float)d);
    const float gvq = (training && g_vq_p) ? __ldg(g_vq_p) : 0.0f;
    const float gcm = (training && g_commit_p) ? __ldg(g_commit_p) * beta : 0.0f;
    const int64_t nsteps = (n + 3) >> 2;
    for (int64_t step = warp; step < nsteps; step += nwarps) {
        const int64_t row = step * 4 + grp;
        const bool valid = row < n;
        float4 xr[NQ], gx[NQ];
        float nx2 = 0.f;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int f = g + 8 * q;
            const bool ok = valid && f < nv;
            xr[q] = ok ? __ldcs(x4 + row * nv + f) : make_float4(0.f, 0.f, 0.f, 0.f);
            gx[q] = (ok && g_out) ? __ldcs(go4 + row * nv + f) : make_float4(0.f, 0.f, 0.f, 0.f);
            nx2 += dot4(xr[q], xr[q]);
        }
        // top-k ids / weights of the row: lane t of the group loads entry t, the group shares them
        int my_r = 0;
        float my_w = 0.f;
        if (valid && g < topk) { my_r = (int)idx[row * topk + g]; my_w = wts[row * topk + g]; }
        nx2 = group_sum8(nx2);
        const float nx = fmaxf(sqrtf(nx2), kNormEps);
        const float rnx = 1.0f / nx;
        int rj[KT];
        float wj[KT];
        // the gathered normalised rows stay in registers (D <= 128) or are re-read from L1 / L2 in each phase (D = 256)
        constexpr bool kKeep = NQ <= 4;
        float4 creg[kKeep ? KT : 1][kKeep ? NQ : 1];
#pragma unroll
        for (int t = 0; t < KT; ++t) {
            rj[t] = __shfl_sync(0xffffffffu, my_r, t, 8);
            wj[t] = __shfl_sync(0xffffffffu, my_w, t, 8);
            if (t >= topk) wj[t] = 0.f;
            rj[t] = min(max(rj[t], 0), k - 1);
        }
        auto ldc = [&](int t, int q) -> float4 {
            const int f = g + 8 * q;
            return (valid && t < topk && f < nv) ? __ldg(cn4 + (size_t)rj[t] * nv + f) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        if constexpr (kKeep) {
#pragma unroll
            for (int t = 0; t < KT; ++t)
#pragma unroll
                for (int q = 0; q < NQ; ++q) creg[t][q] = ldc(t, q);
        }
#define MTVQ_C(t, q) (kKeep ? creg[kKeep ? (t) : 0][kKeep ? (q) : 0] : ldc((t), (q)))
        // zq, diff, dL/dzq; a_j = dL/dzq . c_j
        float4 gz[NQ];
        float a[KT];
#pragma unroll
        for (int t = 0; t < KT; ++t) a[t] = 0.f;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            float4 zq = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int t = 0; t < KT; ++t) {
                const float4 cq = MTVQ_C(t, q);
                zq.x += wj[t] * cq.x; zq.y += wj[t] * cq.y; zq.z += wj[t] * cq.z; zq.w += wj[t] * cq.w;
            }
            const float4 df = make_float4((zq.x - xr[q].x) * scale, (zq.y - xr[q].y) * scale, (zq.z - xr[q].z) * scale,
                                          (zq.w - xr[q].w) * scale);
            gz[q] = make_float4(gvq * df.x, gvq * df.y, gvq * df.z, gvq * df.w);
            gx[q].x -= gcm * df.x; gx[q].y -= gcm * df.y; gx[q].z -= gcm * df.z; gx[q].w -= gcm * df.w;
#pragma unroll
            for (int t = 0; t < KT; ++t) a[t] += dot4(gz[q], MTVQ_C(t, q));
        }
        float sdot = 0.f;
#pragma unroll
        for (int t = 0; t < KT; ++t) { a[t] = group_sum8(a[t]); sdot += wj[t] * a[t]; }
        float h[KT];
#pragma unroll
        for (int t = 0; t < KT; ++t) h[t] = -wj[t] * (a[t] - sdot);
        // dL/dxn and the codebook rows' gradients
        float px = 0.f;
        float4 gxn[NQ];
        float proj[KT];
#pragma unroll
        for (int t = 0; t < KT; ++t) proj[t] = 0.f;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const float4 xn = make_float4(xr[q].x * rnx, xr[q].y * rnx, xr[q].z * rnx, xr[q].w * rnx);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int t = 0; t < KT; ++t) {
                const float4 cq = MTVQ_C(t, q);
                acc.x += h[t] * (2.f * xn.x - 2.f * cq.x); acc.y += h[t] * (2.f * xn.y - 2.f * cq.y);
                acc.z += h[t] * (2.f * xn.z - 2.f * cq.z); acc.w += h[t] * (2.f * xn.w - 2.f * cq.w);
                // dL/dc_j (kept in place of c_j . dL/dc_j accumulation first)
                const float4 gc = make_float4(wj[t] * gz[q].x + h[t] * (2.f * cq.x - 2.f * xn.x),
                                              wj[t] * gz[q].y + h[t] * (2.f * cq.y - 2.f * xn.y),
                                              wj[t] * gz[q].z + h[t] * (2.f * cq.z - 2.f * xn.z),
                                              wj[t] * gz[q].w + h[t] * (2.f * cq.w - 2.f * xn.w));
                proj[t] += dot4(gc, cq);
            }
            gxn[q] = acc;
            px += dot4(acc, xn);
        }
        px = group_sum8(px);
#pragma unroll
        for (int t = 0; t < KT; ++t) proj[t] = group_sum8(proj[t]);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int f = g + 8 * q;
            const bool live = valid && f < nv;
            const float4 xn = make_float4(xr[q].x * rnx, xr[q].y * rnx, xr[q].z * rnx, xr[q].w * rnx);
            if (live) __stcs(reinterpret_cast<float4*>(grad_x) + row * nv + f,
                   make_float4(gx[q].x + (gxn[q].x - xn.x * px) * rnx, gx[q].y + (gxn[q].y - xn.y * px) * rnx,
                               gx[q].z + (gxn[q].z - xn.z * px) * rnx, gx[q].w + (gxn[q].w - xn.w * px) * rnx));
            if (training && live) {
#pragma unroll
                for (int t = 0; t < KT; ++t) {
                    if (t < topk) {
                    const float4 cq = MTVQ_C(t, q);
                    const float rn = 1.0f / __ldg(norms + rj[t]);
                    const float4 gc = make_float4(wj[t] * gz[q].x + h[t] * (2.f * cq.x - 2.f * xn.x),
                                                  wj[t] * gz[q].y + h[t] * (2.f * cq.y - 2.f * xn.y),
                                                  wj[t] * gz[q].z + h[t] * (2.f * cq.z - 2.f * xn.z),
                                                  wj[t] * gz[q].w + h[t] * (2.f * cq.w - 2.f * xn.w));
                    atomicAdd(reinterpret_cast<float4*>(grad_codes) + (size_t)rj[t] * nv + f,
                              make_float4((gc.x - cq.x * proj[t]) * rn, (gc.y - cq.y * proj[t]) * rn,
                                          (gc.z - cq.z * proj[t]) * rn, (gc.w - cq.w * proj[t]) * rn));
                    }
                }
            }
        }
    }
#undef MTVQ_C
}

cudaError_t launch_soft_backward_fused(const float* x, int64_t n, int d, const float* codes, int k, int topk,
                                       const int64_t* idx, const float* wts, const float* g_out, const float* g_vq,
                                       const float* g_commit, float beta, int training, float* grad_x, float* grad_codes,
                                       float* codes_n, float* norms, int num_sms, cudaStream_t stream) {
    soft_backward_prep_kernel<<<(k * 32 + 255) / 256, 256, 0, stream>>>(codes, k, d, codes_n, norms);
    ++g_kernel_launches;
    int64_t blocks = ((n + 3) / 4 + 7) / 8;
    if (blocks > (int64_t)num_sms * 8) blocks = (int64_t)num_sms * 8;
    if (blocks < 1) blocks = 1;
#define MTVQ_BWD_ARGS x, n, d, codes_n, norms, k, topk, idx, wts, g_out, g_vq, g_commit, beta, training, grad_x, grad_codes
#define MTVQ_BWD_LAUNCH(NQ_)                                                                                  \
    do {                                                                                                      \
        if (topk == 5) soft_backward_fused_kernel<NQ_, 5><<<(int)blocks, 256, 0, stream>>>(MTVQ_BWD_ARGS);    \
        else           soft_backward_fused_kernel<NQ_, 8><<<(int)blocks, 256, 0, stream>>>(MTVQ_BWD_ARGS);    \
    } while (0)
    if (d <= 64)       MTVQ_BWD_LAUNCH(2);
    else if (d <= 128) MTVQ_BWD_LAUNCH(4);
    else               MTVQ_BWD_LAUNCH(8);
#undef MTVQ_BWD_LAUNCH
#undef MTVQ_BWD_ARGS
    ++g_kernel_launches;
    return cudaGetLastError();
}

cudaError_t launch_soft_backward(const float* x, int64_t n, int d, const float* codes, int k, int topk,
                                 const int64_t* idx, const float* wts, const float* g_zq, float* grad_x,
                                 float* grad_codes, int num_sms, cudaStream_t stream) {
    int64_t blocks = (n + 7) / 8;
    if (blocks > (int64_t)num_sms * 16) blocks = (int64_t)num_sms * 16;
    soft_backward_kernel<<<(int)blocks, 256, 0, stream>>>(x, n, d, codes, k, topk, idx, wts, g_zq, grad_x, grad_codes);
    ++g_kernel_launches;
    return cudaGetLastError();
}

}  // namespace mtvq
