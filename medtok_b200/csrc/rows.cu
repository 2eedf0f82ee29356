// rows.cu -- the HBM-bound row kernels around the assignment (all coalesced / vectorised, no N x K
// intermediate anywhere):
//   hist_kernel         per-code row counts of caller-provided indices
//   gather_ste_kernel   z_q = zf + (w[idx] - zf), SSE (fp64), per-code counts         (:181,183-194,212-214)
//   scan / scatter      counting sort of the rows by code
//   segsum_kernel       embed_sum[k,:] = sum of the rows of code k, fp64 accumulation      (:202)
//   ema_finalize_kernel cluster_size EMA + normalised-EMA codebook update + loss scalar    (:189,197-212)
// Line numbers refer to /root/reference/MedTok/norm_ema_quantizer.py.
#include "rows.h"

#include <stdio.h>
#include <stdlib.h>
#include "common.cuh"

namespace mtvq {

// per-code row counts of a given index vector (F.one_hot(...).sum(0) of the reference, :183-194; the assign kernels
// count their own winners, this serves mtvq_code_sums on caller-provided indices): equal codes inside a warp are
// merged with match.any so a hot code costs one 32-bit RED per warp instead of 32.
__global__ void __launch_bounds__(256) hist_kernel(const int64_t* __restrict__ src, int64_t n, int* __restrict__ bins) {
    const int lane = threadIdx.x & 31;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x - lane; base < n; base += nthreads) {
        const int64_t row = base + lane;
        const bool valid = row < n;
        const int code = valid ? (int)src[row] : -1;
        const unsigned act = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const unsigned same = __match_any_sync(act, code);
            if (lane == __ffs(same) - 1) atomicAdd(&bins[code], __popc(same));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// gather + straight-through + SSE
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_ste_kernel(const float* __restrict__ z, const float* __restrict__ denom,
                                                          const float* __restrict__ w, const int64_t* __restrict__ idx,
                                                          int64_t n, int d, int k, float* __restrict__ z_q,
                                                          double* __restrict__ stats) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;

    // one float4 per thread per stream, kU independent streams in flight; a row's d/4 vectors sit in adjacent
    // lanes (coalesced).  z and z_q are touched once -> streaming loads/stores keep the codebook in L2.
    constexpr int kU = 4;
    const int vpr = d >> 2;
    const int64_t nvec = n * vpr;
    const float4* z4 = reinterpret_cast<const float4*>(z);
    float4* q4 = reinterpret_cast<float4*>(z_q);
    double sse = 0.0;
    for (int64_t v0 = tid; v0 < nvec; v0 += nthreads * kU) {
        float4 x[kU], c[kU];
        float den[kU];
        int64_t code[kU];
        int c4[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int64_t v = v0 + (int64_t)u * nthreads;
            const bool ok = v < nvec;
            const int64_t row = ok ? v / vpr : 0;
            c4[u] = ok ? (int)(v - row * vpr) : 0;
            den[u] = __ldg(denom + row);
            code[u] = __ldg(idx + row);
            x[u] = ok ? __ldcs(z4 + v) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) c[u] = __ldg(reinterpret_cast<const float4*>(w + code[u] * d) + c4[u]);
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int64_t v = v0 + (int64_t)u * nthreads;
            if (v < nvec) {
                const float zf0 = __fdiv_rn(x[u].x, den[u]), zf1 = __fdiv_rn(x[u].y, den[u]),
                            zf2 = __fdiv_rn(x[u].z, den[u]), zf3 = __fdiv_rn(x[u].w, den[u]);
                const float d0 = __fsub_rn(c[u].x, zf0), d1 = __fsub_rn(c[u].y, zf1), d2 = __fsub_rn(c[u].z, zf2),
                            d3 = __fsub_rn(c[u].w, zf3);
                sse += (double)__fmul_rn(d0, d0) + (double)__fmul_rn(d1, d1) + (double)__fmul_rn(d2, d2) +
                       (double)__fmul_rn(d3, d3);
                float4 o;
                o.x = __fadd_rn(zf0, d0); o.y = __fadd_rn(zf1, d1); o.z = __fadd_rn(zf2, d2); o.w = __fadd_rn(zf3, d3);
                __stcs(q4 + v, o);
            }
        }
    }
    sse = warp_sum(sse);
    __shared__ double part[8];
    if (lane == 0) part[threadIdx.x >> 5] = sse;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += part[i];
        atomicAdd(&stats[(size_t)k + (size_t)k * d], s);
    }
}

// ------------------------------------------------------------------------------------------------
// backward of the hard forward (norm_ema_quantizer.py:212-214 through l2norm :8-9): the only
// gradient goes to z --  g = g_zq + g_loss * 2 beta / (N D) * (zf - z_q),  grad_z = (g - zf (zf . g)) / max(|z|, eps).
// One warp per row, fp32 (a gradient, not part of the bit-exact contract).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hard_backward_kernel(const float* __restrict__ z, const float* __restrict__ z_q,
                                                             const float* __restrict__ g_zq, const float* __restrict__ g_loss,
                                                             int64_t n, int d, float coef, float* __restrict__ grad_z) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float gl = g_loss ? __ldg(g_loss) * coef : 0.0f;
    const int nv = d >> 2;
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n; row += nwarps) {
        const float4* zr = reinterpret_cast<const float4*>(z + row * d);
        float ss = 0.f;
        for (int v = lane; v < nv; v += 32) {
            const float4 x = __ldg(zr + v);
            ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float den = fmaxf(sqrtf(ss), kNormEps), rden = 1.0f / den;
        float dotg = 0.f;
        for (int v = lane; v < nv; v += 32) {
            const float4 x = __ldg(zr + v);
            const float4 q = __ldg(reinterpret_cast<const float4*>(z_q + row * d) + v);
            const float4 go = g_zq ? __ldg(reinterpret_cast<const float4*>(g_zq + row * d) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float f0 = x.x * rden, f1 = x.y * rden, f2 = x.z * rden, f3 = x.w * rden;
            dotg += f0 * (go.x + gl * (f0 - q.x)) + f1 * (go.y + gl * (f1 - q.y)) + f2 * (go.z + gl * (f2 - q.z)) +
                    f3 * (go.w + gl * (f3 - q.w));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dotg += __shfl_xor_sync(0xffffffffu, dotg, o);
        for (int v = lane; v < nv; v += 32) {
            const float4 x = __ldg(zr + v);
            const float4 q = __ldg(reinterpret_cast<const float4*>(z_q + row * d) + v);
            const float4 go = g_zq ? __ldg(reinterpret_cast<const float4*>(g_zq + row * d) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float f0 = x.x * rden, f1 = x.y * rden, f2 = x.z * rden, f3 = x.w * rden;
            float4 o;
            o.x = (go.x + gl * (f0 - q.x) - f0 * dotg) * rden;
            o.y = (go.y + gl * (f1 - q.y) - f1 * dotg) * rden;
            o.z = (go.z + gl * (f2 - q.z) - f2 * dotg) * rden;
            o.w = (go.w + gl * (f3 - q.w) - f3 * dotg) * rden;
            __stcs(reinterpret_cast<float4*>(grad_z + row * d) + v, o);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// lookup-only quantizer (norm_ema_quantizer.py:222-268): z_q = w[idx]; optional SSE against raw z (:262-264)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lookup_kernel(const float* __restrict__ z, const float* __restrict__ w,
                                                      const int64_t* __restrict__ idx, int64_t n, int d, int k,
                                                      float* __restrict__ z_q, double* __restrict__ sse_out) {
    const int vpr = d >> 2;
    const int64_t nvec = n * vpr;
    const int lane = threadIdx.x & 31;
    double sse = 0.0;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = v / vpr;
        const int c4 = (int)(v - row * vpr);
        const int64_t code = __ldg(idx + row);
        if (code < 0 || code >= k) __trap();                    // F.embedding device-asserts on the same input
        const float4 c = __ldg(reinterpret_cast<const float4*>(w + code * d) + c4);
        __stcs(reinterpret_cast<float4*>(z_q) + v, c);
        if (z != nullptr) {
            const float4 x = __ldcs(reinterpret_cast<const float4*>(z) + v);
            const float d0 = __fsub_rn(c.x, x.x), d1 = __fsub_rn(c.y, x.y), d2 = __fsub_rn(c.z, x.z), d3 = __fsub_rn(c.w, x.w);
            sse += (double)__fmul_rn(d0, d0) + (double)__fmul_rn(d1, d1) + (double)__fmul_rn(d2, d2) + (double)__fmul_rn(d3, d3);
        }
    }
    if (sse_out == nullptr) return;
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

__global__ void counts_to_float_kernel(const int* __restrict__ bins, int k, float* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) counts[i] = (float)bins[i];
}

// ------------------------------------------------------------------------------------------------
// codebook usage (vector_quantization_soft_one_new.py:219-236): sliding float32 window of the last `cap` code
// indices, usage = #distinct values in the window / n_e.  The reference shifts the window with a clone and counts with
// sort + unique; here the shifted window is written to a second buffer while a bitmap collects the values.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) usage_shift_kernel(const float* __restrict__ window, const int64_t* __restrict__ idx,
                                                           int64_t cur, int cap, float* __restrict__ next,
                                                           unsigned* __restrict__ bitmap, int nbits) {
    // window' = concat(window[cur:], idx[-min(cur, cap):]) (reference :224-232 keeps the last `cap` entries)
    const int64_t keep = cur >= cap ? 0 : cap - cur;          // old entries that survive
    const int64_t skip = cur >= cap ? cur - cap : 0;           // new entries that fall off immediately
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = i < keep ? window[i + (cap - keep)] : (float)idx[skip + (i - keep)];
        next[i] = v;
        const int b = (int)v;
        // out-of-range / non-integral values share one extra bit (counted as one distinct value)
        const bool in_range = b >= 0 && b < nbits && (float)b == v;
        unsigned* word = &bitmap[in_range ? (b >> 5) : ((nbits + 31) >> 5)];
        const unsigned bit = in_range ? 1u << (b & 31) : 1u;
        if (!(*reinterpret_cast<volatile unsigned*>(word) & bit)) atomicOr(word, bit);
    }
}

__global__ void __launch_bounds__(256) usage_count_kernel(const unsigned* __restrict__ bitmap, int nwords,
                                                           int* __restrict__ distinct) {
    int c = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += gridDim.x * blockDim.x) c += __popc(bitmap[i]);
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(distinct, c);
}

// ------------------------------------------------------------------------------------------------
// counting sort of rows by code
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) scan_bins_kernel(const int* __restrict__ bins, int k, int64_t n, int d,
                                                          int* __restrict__ offsets, int* __restrict__ cursor,
                                                          double* __restrict__ stats) {
    // `stats` may be null: the fused sorted-rows pass then adds the counts to the packed statistics itself, and this
    // single-CTA kernel touches one page of small arrays only (its run time is address-translation latency)
    // exclusive scan of the k counts by one CTA: contiguous slice per thread -> warp shuffles -> 32 warp totals
    __shared__ int warp_tot[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int per = (k + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(k, lo + per);
    int s = 0;
    for (int i0 = lo; i0 < hi; i0 += 8) {         // 8 independent loads in flight
        int v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (i0 + u < hi) ? bins[i0 + u] : 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[u];
    }
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int t = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, t, o);
            if (lane >= o) t += v;
        }
        warp_tot[lane] = t;                       // inclusive totals of warps 0..lane
    }
    __syncthreads();
    int run = incl - s + (wid > 0 ? warp_tot[wid - 1] : 0);
    for (int i0 = lo; i0 < hi; i0 += 8) {
        int v[8];
        double acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            v[u] = (i0 + u < hi) ? bins[i0 + u] : 0;   // L1 hits
            acc[u] = (stats != nullptr && i0 + u < hi) ? stats[i0 + u] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (i0 + u < hi) {
                offsets[i0 + u] = run;
                cursor[i0 + u] = run;
                if (stats != nullptr) stats[i0 + u] = acc[u] + (double)v[u];   // packed statistics accumulate: bins as doubles (exact)
                run += v[u];
            }
        }
    }
    if (threadIdx.x == 1023) offsets[k] = warp_tot[31];
    if (threadIdx.x == 0 && stats != nullptr) stats[(size_t)k + (size_t)k * d + 1] += (double)n;
}

__global__ void scatter_rows_kernel(const int64_t* __restrict__ idx, int64_t n, int* __restrict__ cursor,
                                    int* __restrict__ perm) {
    const int lane = threadIdx.x & 31;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x - lane; base < n; base += nthreads) {
        const int64_t row = base + lane;
        const bool valid = row < n;
        const int code = valid ? (int)idx[row] : -1;
        const unsigned act = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const unsigned same = __match_any_sync(act, code);
            const int leader = __ffs(same) - 1;
            int start = 0;
            if (lane == leader) start = atomicAdd(&cursor[code], __popc(same));
            start = __shfl_sync(same, start, leader);
            perm[start + __popc(same & ((1u << lane) - 1))] = (int)row;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// segmented fp64 sums over the code-sorted rows.  Work item = (range of kSegRows consecutive sorted rows,
// float4 column c4): the d/4 threads of a group read one 4*d-byte row together (coalesced) and walk the
// range 4 rows at a time (independent loads in flight); a group flushes with fp64 REDs when the code
// changes, so each code costs ~ (rows/kSegRows + 1) * d atomics whatever the skew.
// ------------------------------------------------------------------------------------------------
constexpr int kSegRows = 32;

__global__ void __launch_bounds__(256) segsum_kernel(const float* __restrict__ z, const float* __restrict__ denom,
                                                      const int64_t* __restrict__ idx, const int* __restrict__ perm,
                                                      int64_t n, int d, double* __restrict__ embed_sum) {
    const int vpr = d >> 2;                                   // float4 per row
    const int64_t nranges = (n + kSegRows - 1) / kSegRows;
    const int64_t items = nranges * vpr;
    for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
        const int64_t range = it / vpr;
        const int c4 = (int)(it - range * vpr);
        const int64_t p0 = range * kSegRows;
        const int cnt = (int)min((int64_t)kSegRows, n - p0);
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int cur = -1;
        auto flush = [&]() {
            if (cur >= 0) {
                double* dst = embed_sum + (size_t)cur * d + 4 * c4;
                atomicAdd(dst + 0, a0); atomicAdd(dst + 1, a1); atomicAdd(dst + 2, a2); atomicAdd(dst + 3, a3);
            }
            a0 = a1 = a2 = a3 = 0.0;
        };
        for (int j0 = 0; j0 < cnt; j0 += 4) {
            int rows[4], codes[4];
            float dens[4];
            float4 x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool ok = j0 + u < cnt;
                rows[u] = ok ? perm[p0 + j0 + u] : 0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool ok = j0 + u < cnt;
                codes[u] = ok ? (int)__ldg(idx + rows[u]) : -1;
                dens[u] = ok ? __ldg(denom + rows[u]) : 1.0f;
                x[u] = ok ? __ldg(reinterpret_cast<const float4*>(z + (size_t)rows[u] * d) + c4) : make_float4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (j0 + u < cnt) {
                    if (codes[u] != cur) { flush(); cur = codes[u]; }
                    a0 += (double)__fdiv_rn(x[u].x, dens[u]);
                    a1 += (double)__fdiv_rn(x[u].y, dens[u]);
                    a2 += (double)__fdiv_rn(x[u].z, dens[u]);
                    a3 += (double)__fdiv_rn(x[u].w, dens[u]);
                }
            }
        }
        flush();
    }
}

// ------------------------------------------------------------------------------------------------
// training forward, fused: ONE pass over the rows in code-sorted order does gather + straight-through + SSE
// (:181,212-214) AND the per-code sums (:202), so z is read once instead of twice.  Same work split as
// segsum_kernel; a row's d/4 threads write the 4*d-byte z_q row as whole 128-byte lines.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 3) sorted_rows_kernel(const float* __restrict__ z, const float* __restrict__ denom,
                                                           const float* __restrict__ w, const int64_t* __restrict__ idx,
                                                           const int* __restrict__ perm, const int* __restrict__ bins,
                                                           int64_t n, int d, int k,
                                                           float* __restrict__ z_q, double* __restrict__ stats) {
    const int vpr = d >> 2;
    const int lane = threadIdx.x & 31;
    double* embed_sum = stats + k;
    // packed statistics: per-code counts as doubles (exact) and the row count, accumulated like the sums
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < k; c += (int64_t)gridDim.x * blockDim.x)
        stats[c] += (double)bins[c];
    if (blockIdx.x == 0 && threadIdx.x == 0) stats[(size_t)k + (size_t)k * d + 1] += (double)n;
    const int64_t nranges = (n + kSegRows - 1) / kSegRows;
    const int64_t items = nranges * vpr;
    double sse = 0.0;
    for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
        const int64_t range = it / vpr;
        const int c4 = (int)(it - range * vpr);
        const int64_t p0 = range * kSegRows;
        const int cnt = (int)min((int64_t)kSegRows, n - p0);
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int cur = -1;
        auto flush = [&]() {
            if (cur >= 0) {
                double* dst = embed_sum + (size_t)cur * d + 4 * c4;
                atomicAdd(dst + 0, a0); atomicAdd(dst + 1, a1); atomicAdd(dst + 2, a2); atomicAdd(dst + 3, a3);
            }
            a0 = a1 = a2 = a3 = 0.0;
        };
        // software pipeline: the (row, code, norm) triples of the NEXT four sorted positions are fetched (two dependent
        // round trips: perm -> idx / denom) while the rows and codebook rows of the current four are in flight
        int rows[4], codes[4];
        float dens[4];
        auto fetch_meta = [&](int j) {
#pragma unroll
            for (int u = 0; u < 4; ++u) rows[u] = (j + u < cnt) ? perm[p0 + j + u] : 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool ok = j + u < cnt;
                codes[u] = ok ? (int)__ldg(idx + rows[u]) : 0;
                dens[u] = ok ? __ldg(denom + rows[u]) : 1.0f;
            }
        };
        fetch_meta(0);
        for (int j0 = 0; j0 < cnt; j0 += 4) {
            float4 x[4], c[4];
            int crow[4], ccode[4];
            float cden[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                crow[u] = rows[u]; ccode[u] = codes[u]; cden[u] = dens[u];
                const bool ok = j0 + u < cnt;
                x[u] = ok ? __ldcs(reinterpret_cast<const float4*>(z + (size_t)crow[u] * d) + c4) : make_float4(0, 0, 0, 0);
                c[u] = __ldg(reinterpret_cast<const float4*>(w + (size_t)ccode[u] * d) + c4);
            }
            if (j0 + 4 < cnt) fetch_meta(j0 + 4);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (j0 + u < cnt) {
                    if (ccode[u] != cur) { flush(); cur = ccode[u]; }
                    const float zf0 = __fdiv_rn(x[u].x, cden[u]), zf1 = __fdiv_rn(x[u].y, cden[u]),
                                zf2 = __fdiv_rn(x[u].z, cden[u]), zf3 = __fdiv_rn(x[u].w, cden[u]);
                    a0 += (double)zf0; a1 += (double)zf1; a2 += (double)zf2; a3 += (double)zf3;
                    const float d0 = __fsub_rn(c[u].x, zf0), d1 = __fsub_rn(c[u].y, zf1), d2 = __fsub_rn(c[u].z, zf2),
                                d3 = __fsub_rn(c[u].w, zf3);
                    sse += (double)__fmul_rn(d0, d0) + (double)__fmul_rn(d1, d1) + (double)__fmul_rn(d2, d2) +
                           (double)__fmul_rn(d3, d3);
                    float4 o;
                    o.x = __fadd_rn(zf0, d0); o.y = __fadd_rn(zf1, d1); o.z = __fadd_rn(zf2, d2); o.w = __fadd_rn(zf3, d3);
                    __stcs(reinterpret_cast<float4*>(z_q + (size_t)crow[u] * d) + c4, o);
                }
            }
        }
        flush();
    }
    sse = warp_sum(sse);
    __shared__ double part[8];
    if (lane == 0) part[threadIdx.x >> 5] = sse;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part[i];
        atomicAdd(&stats[(size_t)k + (size_t)k * d], t);
    }
}

// ------------------------------------------------------------------------------------------------
// EMA: one thread per code, every reduction in the reference's order (oracle: vqo_ema_*)
// ------------------------------------------------------------------------------------------------
constexpr int kEmaCodes = 32;   // codes per CTA

// Cross-GPU exchange (multi-GPU data parallel, norm_ema_quantizer.py:188,195,203): peer-mapped buffers of
// ema_finalize_peers_kernel below.
struct PeerTable {
    const double* const* stats;   // [world] peer-mapped statistics buffers of this epoch's parity
    unsigned* const*     flags;   // [world] peer-mapped flag arrays, flags[r][q] = last epoch rank q announced to rank r
    int world, rank;
    unsigned epoch;
    unsigned long long timeout_ns;   // upper bound on the wait for the peers' flags
    float* const*        results;    // [world] peer-mapped result buffers: updated rows [k*d] | cluster_size [k]  (fp32)
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) ema_finalize_kernel(float* __restrict__ weight, float* __restrict__ cluster_size,
                                                            const double* __restrict__ stats, int k, int d,
                                                            float beta, float decay, float alpha, int update_cs,
                                                            int update_w, float* __restrict__ loss) {
    // tile[c][i] (row stride d + 1: conflict-free when 32 threads walk 32 different rows) holds, per code,
    // first the per-code mean e = fl(fl32(sum) / bins) (or the old row of a dead code), then the updated row
    extern __shared__ float tile[];
    float* wold = tile + kEmaCodes * (d + 1);
    __shared__ double bins_s[kEmaCodes];
    const int code0 = blockIdx.x * kEmaCodes;
    auto stat = [&](size_t i) -> double { return stats[i]; };     // the local or all-reduced packed statistics
    if (blockIdx.x == 0 && threadIdx.x == 0 && loss) {
        const double sse = stat((size_t)k + (size_t)k * d), rows = stat((size_t)k + (size_t)k * d + 1);
        const float mse = (float)(sse / (rows * (double)d));
        *loss = __fmul_rn(beta, mse);
    }
    if (threadIdx.x < kEmaCodes) bins_s[threadIdx.x] = code0 + threadIdx.x < k ? stat(code0 + threadIdx.x) : 0.0;
    __syncthreads();
    if (update_cs) {   // ema_inplace: fl(cs*decay) then fma(new, alpha, .)
        const int code = code0 + threadIdx.x;
        if (threadIdx.x < kEmaCodes && code < k) {
            const float t = __fmul_rn(cluster_size[code], decay);
            cluster_size[code] = __fmaf_rn((float)bins_s[threadIdx.x], alpha, t);
        }
    }
    if (!update_w) return;
    // elementwise steps run on all threads (coalesced); only the two norms, whose summation order is part of
    // the contract (oracle: vqo_ema_*), run one thread per code
    __shared__ float den_s[kEmaCodes];
    const int ncodes = min(kEmaCodes, k - code0);
    // (latency-bound: 256 threads and 8 independent element loads per thread keep the whole tile's loads in flight)
#pragma unroll 8
    for (int e = threadIdx.x; e < ncodes * d; e += blockDim.x) {
        const int c = e / d, i = e - c * d;
        const float wv = weight[(size_t)(code0 + c) * d + i];
        const double bins = bins_s[c];
        wold[c * (d + 1) + i] = wv;
        // per-code mean; a dead code keeps its row (still re-normalised below, like the reference)
        tile[c * (d + 1) + i] = (bins == 0.0) ? wv : __fdiv_rn((float)stat((size_t)k + (size_t)(code0 + c) * d + i), (float)bins);
    }
    __syncthreads();
    if (threadIdx.x < ncodes) {
        float den = 1.0f;
        if (bins_s[threadIdx.x] != 0.0) {
            const float* e = tile + threadIdx.x * (d + 1);
            den = canon_norm_seq([&](int i) { return e[i]; }, d);   // strided (.t()) view in the reference
            den = den < kNormEps ? kNormEps : den;
        }
        den_s[threadIdx.x] = den;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < ncodes * d; e += blockDim.x) {
        const int c = e / d, i = e - c * d;
        float v = tile[c * (d + 1) + i];
        if (bins_s[c] != 0.0) v = __fdiv_rn(v, den_s[c]);
        tile[c * (d + 1) + i] = __fmaf_rn(v, alpha, __fmul_rn(wold[c * (d + 1) + i], decay));
    }
    __syncthreads();
    if (threadIdx.x < ncodes) {
        const float* e = tile + threadIdx.x * (d + 1);
        float den = canon_norm8([&](int i) { return e[i]; }, d);
        den_s[threadIdx.x] = den < kNormEps ? kNormEps : den;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < ncodes * d; e += blockDim.x) {
        const int c = e / d, i = e - c * d;
        weight[(size_t)(code0 + c) * d + i] = __fdiv_rn(tile[c * (d + 1) + i], den_s[c]);
    }
}

// k-means centre update (norm_ema_quantizer.py:43-54, cosine branch): means_k <- l2norm(sum_k / bins_k), a cluster
// without members keeps its centre.  The row is contiguous in the reference -> canon_norm8 order.
__global__ void __launch_bounds__(256) kmeans_update_kernel(float* __restrict__ means, const double* __restrict__ stats,
                                                             int k, int d) {
    extern __shared__ float tile[];
    __shared__ double bins_s[kEmaCodes];
    __shared__ float den_s[kEmaCodes];
    const int code0 = blockIdx.x * kEmaCodes;
    const int ncodes = min(kEmaCodes, k - code0);
    if (threadIdx.x < kEmaCodes) bins_s[threadIdx.x] = threadIdx.x < ncodes ? stats[code0 + threadIdx.x] : 0.0;
    __syncthreads();
    for (int e = threadIdx.x; e < ncodes * d; e += blockDim.x) {
        const int c = e / d, i = e - c * d;
        const double bins = bins_s[c];
        if (bins != 0.0) tile[c * (d + 1) + i] = __fdiv_rn((float)stats[(size_t)k + (size_t)(code0 + c) * d + i], (float)bins);
    }
    __syncthreads();
    if (threadIdx.x < ncodes && bins_s[threadIdx.x] != 0.0) {
        const float* e = tile + threadIdx.x * (d + 1);
        const float den = canon_norm8([&](int i) { return e[i]; }, d);
        den_s[threadIdx.x] = den < kNormEps ? kNormEps : den;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < ncodes * d; e += blockDim.x) {
        const int c = e / d, i = e - c * d;
        if (bins_s[c] != 0.0) means[(size_t)(code0 + c) * d + i] = __fdiv_rn(tile[c * (d + 1) + i], den_s[c]);
    }
}

// ------------------------------------------------------------------------------------------------
// Exchange fused into the finalize, reduce-scatter shape (norm_ema_quantizer.py:188,195,203 + :197-210 in ONE kernel per
// rank).  The codes are split into `world` contiguous slices; rank r
//   1. announces "my statistics of step `epoch` are complete" (one releasing system-scope store per peer) and waits for
//      every peer's announcement;
//   2. for the codes of ITS slice only: reads the W partial statistics straight out of the peers' memory over NVLink
//      (16-byte loads, several in flight per thread), adds them in rank order in fp64, runs the EMA / re-normalisation
//      and writes the new rows + cluster_size both into its own codebook and into every peer's result buffer;
//   3. announces "my slice is final" to every peer (second flag bank), and copies the other slices, as their owners
//      announce them, from its local result buffer into its codebook.
// Every row is computed by exactly one rank and copied by the others, so the replicas are bit-identical by construction;
// a rank moves ~ (stats + codebook) bytes over NVLink per step instead of world x stats.
// The grid is persistent (all CTAs co-resident) and walks the CTA's own-slice blocks before any foreign block, so no CTA
// can spin on a peer while one of its own blocks -- which that peer may be waiting for -- is still unscheduled.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 ld_cv_f64x2(const double* p) {
    double2 v;
    asm volatile("ld.global.cv.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void wait_flag(const unsigned* flag, unsigned epoch, const PeerTable& pt, int who) {
    const unsigned long long t0 = global_timer_ns();
    while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
        __nanosleep(100);
        if (global_timer_ns() - t0 > pt.timeout_ns) {
            printf("medtok_vq: rank %d waited more than %llu s for rank %d (step %u)\n", pt.rank,
                   pt.timeout_ns / 1000000000ull, who, epoch);
            __trap();
        }
    }
}

constexpr int kPeerFlagBank = 64;     // flags[0..63]: statistics ready; flags[64..127]: slice final; flags[128]: local counter

__global__ void __launch_bounds__(256) ema_finalize_peers_kernel(float* __restrict__ weight, float* __restrict__ cluster_size,
                                                                  PeerTable pt, int k, int d, float beta, float decay,
                                                                  float alpha, int update_cs, int update_w,
                                                                  float* __restrict__ loss) {
    extern __shared__ float tile[];
    float* wold = tile + kEmaCodes * (d + 1);
    __shared__ double bins_s[kEmaCodes];
    __shared__ float den_s[kEmaCodes];
    __shared__ float cs_s[kEmaCodes];
    const int W = pt.world, me = pt.rank;
    const int nblocks = (k + kEmaCodes - 1) / kEmaCodes;
    const int per_rank = (nblocks + W - 1) / W;                       // code blocks per slice
    const int own_lo = min(nblocks, me * per_rank), own_hi = min(nblocks, own_lo + per_rank);
    const int n_own = own_hi - own_lo;
    unsigned* my_flags = pt.flags[me];

    // ---- 1. statistics of this step are complete on every rank ----
    if (threadIdx.x < W) {
        if (blockIdx.x == 0) {
            __threadfence_system();
            st_release_sys(pt.flags[threadIdx.x] + me, pt.epoch);
        }
        wait_flag(my_flags + threadIdx.x, pt.epoch, pt, threadIdx.x);
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0 && loss) {                // global-batch loss: rank-ordered totals
        double sse = 0.0, rows = 0.0;
        for (int r = 0; r < W; ++r) {
            sse += __ldcv(pt.stats[r] + (size_t)k + (size_t)k * d);
            rows += __ldcv(pt.stats[r] + (size_t)k + (size_t)k * d + 1);
        }
        *loss = __fmul_rn(beta, (float)(sse / (rows * (double)d)));
    }
    const size_t kd = (size_t)k * d;
    for (int pb = blockIdx.x; pb < nblocks; pb += gridDim.x) {
        // physical order: own slice first, then the other slices starting after mine (spreads the waits)
        const int lb = (own_lo + pb) % nblocks;
        const int owner = min(W - 1, lb / per_rank);
        const int code0 = lb * kEmaCodes;
        const int ncodes = min(kEmaCodes, k - code0);
        if (owner != me) {
            // ---- 3b. a foreign slice: wait for its owner, then copy from MY result buffer ----
            if (threadIdx.x == 0) wait_flag(my_flags + kPeerFlagBank + owner, pt.epoch, pt, owner);
            __syncthreads();
            const float* res = pt.results[me];
            if (update_w)
                for (int e = threadIdx.x; e < ncodes * d; e += blockDim.x)
                    weight[(size_t)code0 * d + e] = __ldcv(res + (size_t)code0 * d + e);
            if (update_cs && threadIdx.x < ncodes) cluster_size[code0 + threadIdx.x] = __ldcv(res + kd + code0 + threadIdx.x);
            __syncthreads();
            continue;
        }
        // ---- 2. my slice: rank-ordered fp64 sums of the peers' partial statistics, EMA, push ----
        if (threadIdx.x < kEmaCodes) {
            double b = 0.0;
            if (threadIdx.x < ncodes)
                for (int r = 0; r < W; ++r) b += __ldcv(pt.stats[r] + code0 + threadIdx.x);
            bins_s[threadIdx.x] = b;
        }
        __syncthreads();
        if (threadIdx.x < ncodes) {
            float cs = cluster_size ? cluster_size[code0 + threadIdx.x] : 0.0f;
            if (update_cs) cs = __fmaf_rn((float)bins_s[threadIdx.x], alpha, __fmul_rn(cs, decay));   // ema_inplace
            cs_s[threadIdx.x] = cs;
        }
        if (update_w) {
            // element pairs: one 16-byte load per peer and pair, all W loads of a pair independent (in flight together)
            const int npair = (ncodes * d) >> 1;                      // d is even
            for (int e2 = threadIdx.x; e2 < npair; e2 += blockDim.x) {
                const int e = 2 * e2, c = e / d, i = e - c * d;
                double s0 = 0.0, s1 = 0.0;
                const size_t off = (size_t)k + (size_t)(code0 + c) * d + i;   // even offset from an even base? k may be odd:
                if ((off & 1) == 0) {
                    for (int r = 0; r < W; ++r) { const double2 v = ld_cv_f64x2(pt.stats[r] + off); s0 += v.x; s1 += v.y; }
                } else {
                    for (int r = 0; r < W; ++r) { s0 += __ldcv(pt.stats[r] + off); s1 += __ldcv(pt.stats[r] + off + 1); }
                }
                const float2 wv = *reinterpret_cast<const float2*>(weight + (size_t)(code0 + c) * d + i);
                const double bins = bins_s[c];
                wold[c * (d + 1) + i] = wv.x; wold[c * (d + 1) + i + 1] = wv.y;
                tile[c * (d + 1) + i]     = (bins == 0.0) ? wv.x : __fdiv_rn((float)s0, (float)bins);
                tile[c * (d + 1) + i + 1] = (bins == 0.0) ? wv.y : __fdiv_rn((float)s1, (float)bins);
            }
            __syncthreads();
            if (threadIdx.x < ncodes) {
                float den = 1.0f;
                if (bins_s[threadIdx.x] != 0.0) {
                    const float* e = tile + threadIdx.x * (d + 1);
                    den = canon_norm_seq([&](int i) { return e[i]; }, d);   // strided (.t()) view in the reference
                    den = den < kNormEps ? kNormEps : den;
                }
                den_s[threadIdx.x] = den;
            }
            __syncthreads();
            for (int e = threadIdx.x; e < ncodes * d; e += blockDim.x) {
                const int c = e / d, i = e - c * d;
                float v = tile[c * (d + 1) + i];
                if (bins_s[c] != 0.0) v = __fdiv_rn(v, den_s[c]);
                tile[c * (d + 1) + i] = __fmaf_rn(v, alpha, __fmul_rn(wold[c * (d + 1) + i], decay));
            }
            __syncthreads();
            if (threadIdx.x < ncodes) {
                const float* e = tile + threadIdx.x * (d + 1);
                const float den = canon_norm8([&](int i) { return e[i]; }, d);
                den_s[threadIdx.x] = den < kNormEps ? kNormEps : den;
            }
            __syncthreads();
            for (int e = threadIdx.x; e < ncodes * d; e += blockDim.x) {
                const int c = e / d, i = e - c * d;
                const float v = __fdiv_rn(tile[c * (d + 1) + i], den_s[c]);
                weight[(size_t)code0 * d + e] = v;
                for (int r = 0; r < W; ++r)
                    if (r != me) pt.results[r][(size_t)code0 * d + e] = v;     // push over NVLink
            }
        }
        if (threadIdx.x < ncodes) {
            if (update_cs) cluster_size[code0 + threadIdx.x] = cs_s[threadIdx.x];
            for (int r = 0; r < W; ++r)
                if (r != me) pt.results[r][kd + code0 + threadIdx.x] = cs_s[threadIdx.x];
        }
        // ---- 3a. last block of my slice: the slice is final everywhere -> tell the peers ----
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned done = atomicAdd(my_flags + 2 * kPeerFlagBank, 1u) + 1u;
            if (done == (unsigned)n_own) {
                my_flags[2 * kPeerFlagBank] = 0u;                       // ready for the next step (stream-ordered)
                __threadfence_system();
                for (int r = 0; r < W; ++r)
                    if (r != me) st_release_sys(pt.flags[r] + kPeerFlagBank + me, pt.epoch);
            }
        }
        __syncthreads();
    }
}

__global__ void loss_kernel(const double* __restrict__ stats, int k, int d, float beta, float* __restrict__ loss) {
    const double sse = stats[(size_t)k + (size_t)k * d], rows = stats[(size_t)k + (size_t)k * d + 1];
    *loss = __fmul_rn(beta, (float)(sse / (rows * (double)d)));
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static int grid_for(int64_t work, int per_block, int num_sms, int max_waves) {
    int64_t b = (work + per_block - 1) / per_block;
    const int64_t cap = (int64_t)num_sms * max_waves;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

cudaError_t launch_hist_idx64(const int64_t* idx, int64_t n, int* bins_i32, int k, int num_sms, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(bins_i32, 0, sizeof(int) * k, s);
    if (e != cudaSuccess) return e;
    hist_kernel<<<grid_for(n, 256, num_sms, 8), 256, 0, s>>>(idx, n, bins_i32);
    ++g_kernel_launches;
    return cudaGetLastError();
}

cudaError_t launch_lookup(const float* z, const float* w, const int64_t* idx, int64_t n, int d, int k, float* z_q,
                          double* sse, int num_sms, cudaStream_t s) {
    lookup_kernel<<<grid_for(n * (d / 4), 256, num_sms, 16), 256, 0, s>>>(z, w, idx, n, d, k, z_q, sse);
    ++g_kernel_launches;
    return cudaGetLastError();
}

cudaError_t launch_counts_to_float(const int* bins, int k, float* counts, cudaStream_t s) {
    counts_to_float_kernel<<<(k + 255) / 256, 256, 0, s>>>(bins, k, counts);
    ++g_kernel_launches;
    return cudaGetLastError();
}

cudaError_t launch_usage(const float* window, const int64_t* idx, int64_t cur, int cap, int n_e, float* next,
                         unsigned* bitmap, int* distinct, int num_sms, cudaStream_t s) {
    const int nwords = (n_e + 31) / 32 + 1;
    cudaError_t e = cudaMemsetAsync(bitmap, 0, sizeof(unsigned) * nwords, s);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(distinct, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    usage_shift_kernel<<<grid_for(cap, 256, num_sms, 4), 256, 0, s>>>(window, idx, cur, cap, next, bitmap, n_e);
    ++g_kernel_launches;
    usage_count_kernel<<<grid_for(nwords, 256, num_sms, 1), 256, 0, s>>>(bitmap, nwords, distinct);
    ++g_kernel_launches;
    return cudaGetLastError();
}

cudaError_t launch_gather_ste(const float* z, const float* denom, const float* w, const int64_t* idx, int64_t n, int d,
                              int k, float* z_q, double* stats, int num_sms, cudaStream_t s) {
    gather_ste_kernel<<<grid_for(n * (d / 4), 256 * 4, num_sms, 6), 256, 0, s>>>(z, denom, w, idx, n, d, k, z_q, stats);
    ++g_kernel_launches;
    return cudaGetLastError();
}

cudaError_t launch_hard_backward(const float* z, const float* z_q, const float* g_zq, const float* g_loss, int64_t n, int d,
                                 double beta, float* grad_z, int num_sms, cudaStream_t s) {
    const float coef = (float)(2.0 * beta / ((double)n * (double)d));
    hard_backward_kernel<<<grid_for(n * 32, 256, num_sms, 16), 256, 0, s>>>(z, z_q, g_zq, g_loss, n, d, coef, grad_z);
    ++g_kernel_launches;
    return cudaGetLastError();
}

// bins (int32, from launch_emit_idx / launch_hist_idx64) -> stats bins + row count; when want_sums also the
// counting sort and the per-code sums
cudaError_t launch_code_sums(const float* z, const float* denom, const int64_t* idx, int64_t n, int d, int k,
                             double* stats, const Workspace& ws, bool want_sums, int num_sms, cudaStream_t s) {
    scan_bins_kernel<<<1, 1024, 0, s>>>(ws.bins_i32, k, n, d, ws.offsets, ws.cursor, stats);
    ++g_kernel_launches;
    if (!want_sums) return cudaGetLastError();
    scatter_rows_kernel<<<grid_for(n, 256, num_sms, 8), 256, 0, s>>>(idx, n, ws.cursor, ws.perm);
    ++g_kernel_launches;
    const int64_t items = ((n + kSegRows - 1) / kSegRows) * (d / 4);
    segsum_kernel<<<grid_for(items, 256, num_sms, 32), 256, 0, s>>>(z, denom, idx, ws.perm, n, d, stats + k);
    ++g_kernel_launches;
    return cudaGetLastError();
}

// training forward: counting sort by code, then the fused sorted-order pass (z_q, SSE, per-code sums)
cudaError_t launch_sorted_rows(const float* z, const float* denom, const float* w, const int64_t* idx, int64_t n, int d,
                               int k, float* z_q, double* stats, const Workspace& ws, int num_sms, cudaStream_t s) {
    scan_bins_kernel<<<1, 1024, 0, s>>>(ws.bins_i32, k, n, d, ws.offsets, ws.cursor, nullptr);
    ++g_kernel_launches;
    scatter_rows_kernel<<<grid_for(n, 256, num_sms, 8), 256, 0, s>>>(idx, n, ws.cursor, ws.perm);
    ++g_kernel_launches;
    const int64_t items = ((n + kSegRows - 1) / kSegRows) * (d / 4);
    sorted_rows_kernel<<<grid_for(items, 256, num_sms, 32), 256, 0, s>>>(z, denom, w, idx, ws.perm, ws.bins_i32, n, d, k, z_q,
                                                                         stats);
    ++g_kernel_launches;
    return cudaGetLastError();
}

static cudaError_t ema_finalize_configure() {
    static PerDeviceOnce once;
    bool& configured = once.flag();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(ema_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    return cudaSuccess;
}

cudaError_t launch_ema_finalize(float* weight, float* cluster_size, const double* stats, int k, int d, double beta,
                                double decay, int update_cs, int update_w, float* loss, cudaStream_t s) {
    // python: mul_(decay).add_(new, alpha=1-decay): both scalars are doubles cast to fp32
    const size_t smem = sizeof(float) * 2 * kEmaCodes * (d + 1);   // <= 65.8 KB at d = 256
    if (cudaError_t e = ema_finalize_configure()) return e;
    ema_finalize_kernel<<<(k + kEmaCodes - 1) / kEmaCodes, 256, smem, s>>>(
        weight, cluster_size, stats, k, d, (float)beta, (float)decay, (float)(1.0 - decay), update_cs, update_w,
        loss);
    ++g_kernel_launches;
    return cudaGetLastError();
}

cudaError_t launch_ema_finalize_peers(float* weight, float* cluster_size, const void* const* peer_stats,
                                      void* const* peer_flags, void* const* peer_results, int world, int rank, unsigned epoch,
                                      int k, int d, double beta, double decay, int update_cs, int update_w, float* loss,
                                      int num_sms, cudaStream_t s) {
    const size_t smem = sizeof(float) * 2 * kEmaCodes * (d + 1);
    static PerDeviceOnce once;
    bool& configured = once.flag();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(ema_finalize_peers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    PeerTable pt;
    pt.stats = reinterpret_cast<const double* const*>(peer_stats);
    pt.flags = reinterpret_cast<unsigned* const*>(peer_flags);
    pt.results = reinterpret_cast<float* const*>(peer_results);
    pt.world = world; pt.rank = rank; pt.epoch = epoch;
    {   // MEDTOK_EXCHANGE_TIMEOUT_S (default 600 s, the NCCL watchdog's order of magnitude)
        static long long limit_s = -1;
        if (limit_s < 0) { const char* e = getenv("MEDTOK_EXCHANGE_TIMEOUT_S"); limit_s = (e && atoll(e) > 0) ? atoll(e) : 600; }
        pt.timeout_ns = (unsigned long long)limit_s * 1000000000ull;
    }
    // persistent grid: every CTA resident at once (<= 2 per SM at 66 KB of shared memory for d = 256)
    const int nblocks = (k + kEmaCodes - 1) / kEmaCodes;
    int grid = nblocks < 2 * num_sms ? nblocks : 2 * num_sms;
    ema_finalize_peers_kernel<<<grid, 256, smem, s>>>(weight, cluster_size, pt, k, d, (float)beta, (float)decay,
                                                      (float)(1.0 - decay), update_cs, update_w, loss);
    ++g_kernel_launches;
    return cudaGetLastError();
}

cudaError_t launch_kmeans_update(float* means, const double* stats, int k, int d, cudaStream_t s) {
    const size_t smem = sizeof(float) * kEmaCodes * (d + 1);
    kmeans_update_kernel<<<(k + kEmaCodes - 1) / kEmaCodes, 256, smem, s>>>(means, stats, k, d);
    ++g_kernel_launches;
    return cudaGetLastError();
}

cudaError_t launch_loss(const double* stats, int k, int d, double beta, float* loss, cudaStream_t s) {
    loss_kernel<<<1, 1, 0, s>>>(stats, k, d, (float)beta, loss);
    ++g_kernel_launches;
    return cudaGetLastError();
}

}  // namespace mtvq
