// rows.cu -- the HBM-bound row kernels around the assignment (all coalesced / vectorised, no N x K
// intermediate anywhere):
//   emit_idx_kernel     int32 winners -> int64 indices (torch.argmin dtype)
//   gather_ste_kernel   z_q = zf + (w[idx] - zf), SSE (fp64), per-code counts         (:181,183-194,212-214)
//   scan / scatter      counting sort of the rows by code
//   segsum_kernel       embed_sum[k,:] = sum of the rows of code k, fp64 accumulation      (:202)
//   ema_finalize_kernel cluster_size EMA + normalised-EMA codebook update + loss scalar    (:189,197-212)
// Line numbers refer to /root/reference/MedTok/norm_ema_quantizer.py.
#include "rows.h"
#include "common.cuh"

namespace mtvq {

__global__ void emit_idx_kernel(const int* __restrict__ best_col, int64_t n, int64_t* __restrict__ idx) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        idx[i] = (int64_t)best_col[i];
}

// ------------------------------------------------------------------------------------------------
// gather + straight-through + SSE + histogram
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_ste_kernel(const float* __restrict__ z, const float* __restrict__ denom,
                                                          const float* __restrict__ w, const int64_t* __restrict__ idx,
                                                          int64_t n, int d, int k, float* __restrict__ z_q,
                                                          double* __restrict__ stats) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;

    // phase A: per-code counts, one thread per row, equal codes inside a warp aggregated
    for (int64_t base = tid - lane; base < n; base += nthreads) {
        const int64_t row = base + lane;
        const bool valid = row < n;
        const int code = valid ? (int)idx[row] : -1;
        const unsigned act = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const unsigned same = __match_any_sync(act, code);
            if (lane == __ffs(same) - 1) atomicAdd(&stats[code], (double)__popc(same));
        }
    }
    if (tid == 0) atomicAdd(&stats[(size_t)k + (size_t)k * d + 1], (double)n);

    // phase B: one float4 per thread per step; a row's d/4 vectors sit in adjacent lanes
    const int vec_per_row = d >> 2;
    const int64_t nvec = n * vec_per_row;
    double sse = 0.0;
    for (int64_t v = tid; v < nvec; v += nthreads) {
        const int64_t row = v / vec_per_row;
        const int c4 = (int)(v - row * vec_per_row);
        const float den = __ldg(denom + row);
        const int64_t code = __ldg(idx + row);
        const float4 x = __ldg(reinterpret_cast<const float4*>(z) + v);
        const float4 q = __ldg(reinterpret_cast<const float4*>(w + code * d) + c4);
        const float zf0 = __fdiv_rn(x.x, den), zf1 = __fdiv_rn(x.y, den), zf2 = __fdiv_rn(x.z, den),
                    zf3 = __fdiv_rn(x.w, den);
        const float d0 = __fsub_rn(q.x, zf0), d1 = __fsub_rn(q.y, zf1), d2 = __fsub_rn(q.z, zf2),
                    d3 = __fsub_rn(q.w, zf3);
        sse += (double)__fmul_rn(d0, d0) + (double)__fmul_rn(d1, d1) + (double)__fmul_rn(d2, d2) +
               (double)__fmul_rn(d3, d3);
        float4 o;
        o.x = __fadd_rn(zf0, d0); o.y = __fadd_rn(zf1, d1); o.z = __fadd_rn(zf2, d2); o.w = __fadd_rn(zf3, d3);
        reinterpret_cast<float4*>(z_q)[v] = o;
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
// counting sort of rows by code
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) scan_bins_kernel(const double* __restrict__ bins, int k,
                                                          int* __restrict__ offsets, int* __restrict__ cursor) {
    __shared__ int part[1024];
    const int per = (k + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(k, lo + per);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += (int)bins[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {   // Hillis-Steele inclusive scan
        int v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = part[threadIdx.x] - s;
    for (int i = lo; i < hi; ++i) {
        offsets[i] = run;
        cursor[i] = run;
        run += (int)bins[i];
    }
    if (threadIdx.x == 1023) offsets[k] = part[1023];
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
// segmented fp64 sums over the code-sorted rows: one warp per 32 consecutive sorted rows
// ------------------------------------------------------------------------------------------------
template <int SLOTS>   // d <= 32 * SLOTS
__global__ void __launch_bounds__(256) segsum_kernel(const float* __restrict__ z, const float* __restrict__ denom,
                                                      const int64_t* __restrict__ idx, const int* __restrict__ perm,
                                                      int64_t n, int d, double* __restrict__ embed_sum) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p0 = warp * 32; p0 < n; p0 += nwarps * 32) {
        const int cnt = (int)min((int64_t)32, n - p0);
        int my_row = 0, my_code = -1;
        float my_den = 1.0f;
        if (lane < cnt) {
            my_row = perm[p0 + lane];
            my_code = (int)idx[my_row];
            my_den = denom[my_row];
        }
        double acc[SLOTS];
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) acc[s] = 0.0;
        int cur = __shfl_sync(0xffffffffu, my_code, 0);
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
            const int row = __shfl_sync(0xffffffffu, my_row, j);
            const int code = __shfl_sync(0xffffffffu, my_code, j);
            const float den = __shfl_sync(0xffffffffu, my_den, j);
            if (code != cur) {
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const int c = s * 32 + lane;
                    if (c < d) atomicAdd(&embed_sum[(size_t)cur * d + c], acc[s]);
                    acc[s] = 0.0;
                }
                cur = code;
            }
            const float* zr = z + (size_t)row * d;
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const int c = s * 32 + lane;
                if (c < d) acc[s] += (double)__fdiv_rn(__ldg(zr + c), den);
            }
        }
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const int c = s * 32 + lane;
            if (c < d && cnt > 0) atomicAdd(&embed_sum[(size_t)cur * d + c], acc[s]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// EMA: one thread per code, every reduction in the reference's order (oracle: vqo_ema_*)
// ------------------------------------------------------------------------------------------------
__global__ void ema_finalize_kernel(float* __restrict__ weight, float* __restrict__ cluster_size,
                                    const double* __restrict__ stats, int k, int d, float beta, float decay,
                                    float alpha, int update_cs, int update_w, float* __restrict__ loss) {
    const int code = blockIdx.x * blockDim.x + threadIdx.x;
    if (code == 0 && loss) {
        const double sse = stats[(size_t)k + (size_t)k * d], rows = stats[(size_t)k + (size_t)k * d + 1];
        const float mse = (float)(sse / (rows * (double)d));
        *loss = __fmul_rn(beta, mse);
    }
    if (code >= k) return;
    const double bins = stats[code];
    if (update_cs) {   // ema_inplace: fl(cs*decay) then fma(new, alpha, .)
        const float t = __fmul_rn(cluster_size[code], decay);
        cluster_size[code] = __fmaf_rn((float)bins, alpha, t);
    }
    if (!update_w) return;
    float* wr = weight + (size_t)code * d;
    const double* sr = stats + k + (size_t)code * d;
    float e[kMaxDim];
    if (bins == 0.0) {   // dead code keeps its row (still re-normalised below, like the reference)
        for (int c = 0; c < d; ++c) e[c] = wr[c];
    } else {
        const float b = (float)bins;
        for (int c = 0; c < d; ++c) e[c] = __fdiv_rn((float)sr[c], b);
        float den = canon_norm_seq([&](int i) { return e[i]; }, d);   // strided (.t()) view in the reference
        den = den < kNormEps ? kNormEps : den;
        for (int c = 0; c < d; ++c) e[c] = __fdiv_rn(e[c], den);
    }
    for (int c = 0; c < d; ++c) e[c] = __fmaf_rn(e[c], alpha, __fmul_rn(wr[c], decay));
    float den = canon_norm8([&](int i) { return e[i]; }, d);
    den = den < kNormEps ? kNormEps : den;
    for (int c = 0; c < d; ++c) wr[c] = __fdiv_rn(e[c], den);
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

cudaError_t launch_emit_idx(const int* best_col, int64_t n, int64_t* idx, int num_sms, cudaStream_t s) {
    emit_idx_kernel<<<grid_for(n, 256, num_sms, 8), 256, 0, s>>>(best_col, n, idx);
    return cudaGetLastError();
}

cudaError_t launch_gather_ste(const float* z, const float* denom, const float* w, const int64_t* idx, int64_t n, int d,
                              int k, float* z_q, double* stats, int num_sms, cudaStream_t s) {
    gather_ste_kernel<<<grid_for(n * (d / 4), 256 * 4, num_sms, 8), 256, 0, s>>>(z, denom, w, idx, n, d, k, z_q, stats);
    return cudaGetLastError();
}

cudaError_t launch_code_sums(const float* z, const float* denom, const int64_t* idx, int64_t n, int d, int k,
                             double* stats, const Workspace& ws, int num_sms, cudaStream_t s) {
    scan_bins_kernel<<<1, 1024, 0, s>>>(stats, k, ws.offsets, ws.cursor);
    scatter_rows_kernel<<<grid_for(n, 256, num_sms, 8), 256, 0, s>>>(idx, n, ws.cursor, ws.perm);
    const int g = grid_for(n, 8 * 32, num_sms, 16);
    double* es = stats + k;
    if (d <= 32) segsum_kernel<1><<<g, 256, 0, s>>>(z, denom, idx, ws.perm, n, d, es);
    else if (d <= 64) segsum_kernel<2><<<g, 256, 0, s>>>(z, denom, idx, ws.perm, n, d, es);
    else if (d <= 128) segsum_kernel<4><<<g, 256, 0, s>>>(z, denom, idx, ws.perm, n, d, es);
    else segsum_kernel<8><<<g, 256, 0, s>>>(z, denom, idx, ws.perm, n, d, es);
    return cudaGetLastError();
}

cudaError_t launch_ema_finalize(float* weight, float* cluster_size, const double* stats, int k, int d, double beta,
                                double decay, int update_cs, int update_w, float* loss, cudaStream_t s) {
    // python: mul_(decay).add_(new, alpha=1-decay): both scalars are doubles cast to fp32
    ema_finalize_kernel<<<(k + 127) / 128, 128, 0, s>>>(weight, cluster_size, stats, k, d, (float)beta, (float)decay,
                                                        (float)(1.0 - decay), update_cs, update_w, loss);
    return cudaGetLastError();
}

cudaError_t launch_loss(const double* stats, int k, int d, double beta, float* loss, cudaStream_t s) {
    loss_kernel<<<1, 1, 0, s>>>(stats, k, d, (float)beta, loss);
    return cudaGetLastError();
}

}  // namespace mtvq
