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

__global__ void __launch_bounds__(256) soft_resolve_kernel(
    const float* __restrict__ x, const float* __restrict__ denom, int64_t n, int d, int k, int topk,
    const float* __restrict__ cn, const float* __restrict__ cc, const int* __restrict__ cand,
    const int* __restrict__ cand_cnt, int cand_cap, int64_t* __restrict__ idx_out, float* __restrict__ val_out,
    float* __restrict__ wts_out, float* __restrict__ zq_out, float* __restrict__ out, double* __restrict__ sse_out,
    int* __restrict__ counters) {
    __shared__ __align__(16) float xn_s[8][kMaxDim];          // one normalised row per warp
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float* xs = xn_s[wib];
    double sse = 0.0;

    for (int64_t row = warp; row < n; row += nwarps) {
        const float* xr = x + row * d;
        const float den = denom[row];
        __syncwarp();
        for (int i = lane; i < d; i += 32) xs[i] = __fdiv_rn(__ldg(xr + i), den);   // xn, divided once
        __syncwarp();
        // ||xn||^2 in the reference's cascade order (every lane the same value; shared-memory broadcasts)
        const float xx = canon_sumsq([&](int i) { return xs[i]; }, d);
        const int cnt = cand_cnt[row];
        const bool full = (cnt < topk) || (cnt > cand_cap);

        // each lane keeps an ascending list of its own best `topk` (value, index) pairs
        float lv[kMaxTop];
        int li[kMaxTop];
#pragma unroll
        for (int t = 0; t < kMaxTop; ++t) { lv[t] = INFINITY; li[t] = 0x7fffffff; }
        const int total = full ? k : cnt;
        for (int j = lane; j < total; j += 32) {
            const int col = full ? j : cand[row * cand_cap + j];
            if (col >= k) continue;
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
        if (full && lane == 0) atomicAdd(&counters[CNT_FULL], 1);

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
                if (t < topk) a = __fadd_rn(a, __fmul_rn(wgt[t], __ldg(cn + (size_t)si[t] * d + c)));
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

cudaError_t launch_soft_resolve(const float* x, int64_t n, int d, int k, int topk, const Workspace& ws, int64_t* idx,
                                float* vals, float* wts, float* zq, float* out, double* sse, int num_sms,
                                cudaStream_t stream) {
    int64_t blocks = (n + 7) / 8;
    if (blocks > (int64_t)num_sms * 16) blocks = (int64_t)num_sms * 16;
    soft_resolve_kernel<<<(int)blocks, 256, 0, stream>>>(x, ws.denom, n, d, k, topk, ws.codes_n, ws.cc, ws.cand, ws.cand_cnt,
                                                         cand_capacity(topk), idx, vals, wts, zq, out, sse, ws.counters);
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
