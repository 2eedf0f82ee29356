// assign_exact.cu -- the fp32 "reference order" side of the assignment:
//   * codebook_prep_kernel : fp32 codebook -> fp16 tensor-core operand image + canonical ||w||^2 +
//                            the rigorous coarse-score error bound (margin)
//   * resolve_kernel       : near-tie rows: exact distances of the collected candidates, first-index argmin
//   * exact_scan_kernel    : exact distances against ALL codes (cross-check + candidate-overflow fallback)
// Every distance here is fl(fl(zz + cc) - 2*dot) with dot a sequential fmaf chain over d, i.e. bit-for-bit
// what the reference's CPU path computes (norm_ema_quantizer.py:175-179).
#include "assign_exact.h"
#include "common.cuh"

namespace mtvq {

// ------------------------------------------------------------------------------------------------
// codebook prep (one thread per code; K*D is tiny next to N*K*D)
// ------------------------------------------------------------------------------------------------
__global__ void codebook_prep_kernel(const float* __restrict__ w, int k, int kpad, int d, int normalize,
                                     unsigned char* __restrict__ image, float* __restrict__ cc,
                                     float* __restrict__ codes_n, float* __restrict__ meta) {
    const int code = blockIdx.x * blockDim.x + threadIdx.x;
    if (code >= kpad) return;
    const int nplanes = d >> 3;
    const int tile = code >> 7, r = code & 127;
    unsigned char* tile_base = image + (size_t)tile * (nplanes + 2) * 2048 + r * 16;
    const __half hz = __float2half(0.0f);

    if (code >= k) {   // padding code: can never win (bias -60000), contributes nothing
        uint4 z4 = make_uint4(0, 0, 0, 0);
        for (int j = 0; j < nplanes + 2; ++j) *reinterpret_cast<uint4*>(tile_base + j * 2048) = z4;
        __half hb[8] = {__float2half(-60000.0f), hz, hz, hz, hz, hz, hz, hz};
        *reinterpret_cast<uint4*>(tile_base + nplanes * 2048) = *reinterpret_cast<uint4*>(hb);
        cc[code] = 3.0e38f;
        return;
    }
    const float* wr = w + (size_t)code * d;
    float den = 1.0f;
    if (normalize) den = canon_denom([&](int i) { return __ldg(wr + i); }, d, 0);
    auto elem = [&](int i) { return normalize ? __fdiv_rn(__ldg(wr + i), den) : __ldg(wr + i); };
    const float c2 = canon_sumsq(elem, d);
    cc[code] = c2;
    {   // spread of the squared norms around code 0's: if it is negligible the kernels drop the bias K-step
        const float* w0 = w;
        float den0 = 1.0f;
        if (normalize) den0 = canon_denom([&](int i) { return __ldg(w0 + i); }, d, 0);
        const float c0 = canon_sumsq([&](int i) { return normalize ? __fdiv_rn(__ldg(w0 + i), den0) : __ldg(w0 + i); }, d);
        const float spread = 0.5f * fabsf(c2 - c0);
        if (spread == spread) atomicMax(reinterpret_cast<int*>(meta + META_SPREAD), __float_as_int(spread));
        else meta[META_SPREAD] = INFINITY;
    }
    for (int j = 0; j < nplanes; ++j) {
        __half h[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float x = elem(8 * j + i);
            h[i] = __float2half_rn(x);
            if (normalize) codes_n[(size_t)code * d + 8 * j + i] = x;
        }
        *reinterpret_cast<uint4*>(tile_base + j * 2048) = *reinterpret_cast<uint4*>(h);
    }
    // bias = -||w||^2 / 2 split into three fp16 terms (exact to ~2^-33 relative)
    const float b = -0.5f * c2;
    const __half h0 = __float2half_rn(b);
    const float r1 = b - __half2float(h0);
    const __half h1 = __float2half_rn(r1);
    const __half h2 = __float2half_rn(r1 - __half2float(h1));
    __half hb[8] = {h0, h1, h2, hz, hz, hz, hz, hz};
    *reinterpret_cast<uint4*>(tile_base + nplanes * 2048) = *reinterpret_cast<uint4*>(hb);
    *reinterpret_cast<uint4*>(tile_base + (nplanes + 1) * 2048) = make_uint4(0, 0, 0, 0);

    // |coarse - exact| <= eps for every row with ||zf|| <= 1:
    //   fp16 rounding of both operands: (2^-10 + 2^-22) * ||w||     (Cauchy-Schwarz)
    //   fp16 subnormal flush + fp32 accumulation in the tensor core: 2^-18 * (1 + ||w|| + ||w||^2/2)
    const float nw = sqrtf(c2) * 1.001f;
    const float eps = 9.7704e-4f * nw + 3.8147e-6f * (1.0f + nw + 0.5f * c2) + 1e-6f + kBiasSkipTol;
    if (eps == eps) {   // positive floats order like ints
        atomicMax(reinterpret_cast<int*>(meta + META_EPS), __float_as_int(eps));
        atomicMax(reinterpret_cast<int*>(meta + META_MARGIN), __float_as_int(2.0f * eps));
        atomicMax(reinterpret_cast<int*>(meta + META_MAXNORM), __float_as_int(nw));
    } else {
        meta[META_MARGIN] = INFINITY;   // NaN in the codebook: send everything to the exact path
    }
}

// ------------------------------------------------------------------------------------------------
// near-tie resolve: one thread per flagged row, exact distances of <= cand_cap candidates
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) resolve_kernel(const float* __restrict__ z, int d, int norm_mode,
                                                      const float* __restrict__ w, int k, const float* __restrict__ cc,
                                                      const int* __restrict__ flag_rows, const int* __restrict__ cand,
                                                      const int* __restrict__ cand_cnt, int cand_cap,
                                                      int64_t* __restrict__ idx, int* __restrict__ bins,
                                                      int* __restrict__ full_rows, int* __restrict__ counters) {
    const int nflag = counters[CNT_FLAGGED];
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nflag; p += gridDim.x * blockDim.x) {
        const int row = flag_rows[p];
        const int cnt = cand_cnt[p];
        if (cnt <= 0 || cnt > cand_cap) {
            full_rows[atomicAdd(&counters[CNT_FULL], 1)] = row;
            continue;
        }
        atomicAdd(&counters[CNT_CANDS], cnt);
        // the normalised row is re-formed element by element (x / den is a pure function of x and den), so
        // nothing is staged in local memory; the row itself stays in L1 across the candidates
        const float* zr = z + (size_t)row * d;
        const float den = canon_denom([&](int i) { return __ldg(zr + i); }, d, norm_mode);
        const float zz = canon_sumsq([&](int i) { return __fdiv_rn(__ldg(zr + i), den); }, d);
        float best = INFINITY;
        int bcol = 0x7fffffff;
        bool any = false;
        for (int j = 0; j < cnt; ++j) {
            const int col = cand[(size_t)p * cand_cap + j];
            if (col >= k) continue;
            const float4* wr = reinterpret_cast<const float4*>(w + (size_t)col * d);
            const float4* zr4 = reinterpret_cast<const float4*>(zr);
            float dot = 0.0f;
            for (int i = 0; i < (d >> 2); ++i) {
                const float4 wv = __ldg(wr + i);
                const float4 xv = __ldg(zr4 + i);
                dot = __fmaf_rn(__fdiv_rn(xv.x, den), wv.x, dot);
                dot = __fmaf_rn(__fdiv_rn(xv.y, den), wv.y, dot);
                dot = __fmaf_rn(__fdiv_rn(xv.z, den), wv.z, dot);
                dot = __fmaf_rn(__fdiv_rn(xv.w, den), wv.w, dot);
            }
            const float dist = canon_dist(zz, cc[col], dot);
            // torch.argmin: smallest value, first index on ties; NaN counts as the minimum
            if (!any || dist < best || (dist == best && col < bcol) || (dist != dist && !(best != best))) {
                best = dist; bcol = col; any = true;
            }
        }
        if (any) {
            idx[row] = (int64_t)bcol;
            atomicAdd(&bins[bcol], 1);
        } else {
            full_rows[atomicAdd(&counters[CNT_FULL], 1)] = row;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// exact scan against ALL codes: 64 rows x 64 codes per step, 4x4 register tile, sequential-in-d fmaf
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int XR = 64;   // rows per block
constexpr int XC = 64;   // codes per step
constexpr int XD = 16;   // d-chunk staged for the codebook
}

__global__ void __launch_bounds__(256) exact_scan_kernel(const float* __restrict__ z, int64_t n, int d, int norm_mode,
                                                          const float* __restrict__ w, int k,
                                                          const float* __restrict__ cc,
                                                          const int* __restrict__ row_list,
                                                          const int* __restrict__ n_rows_dev,
                                                          int64_t* __restrict__ idx, int* __restrict__ bins,
                                                          float* __restrict__ denom_out) {
    extern __shared__ float xs[];
    float* As  = xs;                       // [d][XR]   normalised rows, d-major
    float* Bs  = As + (size_t)d * XR;      // [XD][XC+1]
    float* zzs = Bs + XD * (XC + 1);       // [XR]
    float* red_v = zzs + XR;               // [XR][16]
    int*   red_i = reinterpret_cast<int*>(red_v + XR * 16);

    const int64_t nrows = n_rows_dev ? (int64_t)(*n_rows_dev) : n;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

    for (int64_t blk0 = (int64_t)blockIdx.x * XR; blk0 < nrows; blk0 += (int64_t)gridDim.x * XR) {
        __syncthreads();
        // stage normalised rows
        if (threadIdx.x < XR) {
            const int64_t pos = blk0 + threadIdx.x;
            if (pos < nrows) {
                const int64_t row = row_list ? (int64_t)row_list[pos] : pos;
                const float* zr = z + row * d;
                const float den = canon_denom([&](int i) { return __ldg(zr + i); }, d, norm_mode);
                if (denom_out) denom_out[row] = den;
                for (int i = 0; i < d; ++i) As[(size_t)i * XR + threadIdx.x] = __fdiv_rn(__ldg(zr + i), den);
                zzs[threadIdx.x] = canon_sumsq([&](int i) { return As[(size_t)i * XR + threadIdx.x]; }, d);
            } else {
                for (int i = 0; i < d; ++i) As[(size_t)i * XR + threadIdx.x] = 0.0f;
                zzs[threadIdx.x] = 0.0f;
            }
        }
        __syncthreads();

        float bestv[4];
        int besti[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) { bestv[a] = INFINITY; besti[a] = 0x7fffffff; }
        bool seen[4] = {false, false, false, false};

        for (int c0 = 0; c0 < k; c0 += XC) {
            float acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
            for (int d0 = 0; d0 < d; d0 += XD) {
                __syncthreads();
                // codebook chunk: XC codes x XD dims, 1024 floats / 256 threads
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int e = threadIdx.x + t * 256;
                    const int code = e >> 4, dd = e & 15;
                    float v = 0.0f;
                    if (c0 + code < k && d0 + dd < d) v = __ldg(w + (size_t)(c0 + code) * d + d0 + dd);
                    Bs[dd * (XC + 1) + code] = v;
                }
                __syncthreads();
                const int dlim = min(XD, d - d0);
                for (int dd = 0; dd < dlim; ++dd) {
                    float a[4], b[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) a[t] = As[(size_t)(d0 + dd) * XR + ty * 4 + t];
#pragma unroll
                    for (int t = 0; t < 4; ++t) b[t] = Bs[dd * (XC + 1) + tx * 4 + t];
#pragma unroll
                    for (int ai = 0; ai < 4; ++ai)
#pragma unroll
                        for (int bi = 0; bi < 4; ++bi) acc[ai][bi] = __fmaf_rn(a[ai], b[bi], acc[ai][bi]);
                }
            }
#pragma unroll
            for (int ai = 0; ai < 4; ++ai) {
                const float zz = zzs[ty * 4 + ai];
#pragma unroll
                for (int bi = 0; bi < 4; ++bi) {
                    const int code = c0 + tx * 4 + bi;
                    if (code < k) {
                        const float dist = canon_dist(zz, __ldg(cc + code), acc[ai][bi]);
                        const bool nan_new = dist != dist, nan_old = seen[ai] && (bestv[ai] != bestv[ai]);
                        if (!seen[ai] || (!nan_old && (nan_new || dist < bestv[ai]))) {
                            bestv[ai] = dist; besti[ai] = code; seen[ai] = true;
                        }
                    }
                }
            }
        }
        // reduce the 16 threads that share each row: min value, then min index (NaN first, like torch)
#pragma unroll
        for (int ai = 0; ai < 4; ++ai) {
            red_v[(ty * 4 + ai) * 16 + tx] = bestv[ai];
            red_i[(ty * 4 + ai) * 16 + tx] = besti[ai];
        }
        __syncthreads();
        if (threadIdx.x < XR) {
            const int64_t pos = blk0 + threadIdx.x;
            if (pos < nrows) {
                float bv = red_v[threadIdx.x * 16];
                int bi = red_i[threadIdx.x * 16];
                for (int t = 1; t < 16; ++t) {
                    const float v = red_v[threadIdx.x * 16 + t];
                    const int ii = red_i[threadIdx.x * 16 + t];
                    if (ii == 0x7fffffff) continue;
                    const bool nan_new = v != v, nan_old = bv != bv;
                    bool take;
                    if (bi == 0x7fffffff) take = true;
                    else if (nan_old && nan_new) take = ii < bi;
                    else if (nan_old) take = false;
                    else if (nan_new) take = true;
                    else take = (v < bv) || (v == bv && ii < bi);
                    if (take) { bv = v; bi = ii; }
                }
                const int64_t row = row_list ? (int64_t)row_list[pos] : pos;
                idx[row] = (int64_t)bi;
                if (bins != nullptr && bi >= 0 && bi < k) atomicAdd(&bins[bi], 1);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
cudaError_t launch_codebook_prep(const float* w, int k, int d, int normalize, const Workspace& ws,
                                 cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(ws.meta, 0, sizeof(float) * META_COUNT, stream);
    if (e != cudaSuccess) return e;
    const int kpad = pad_codes(k);
    codebook_prep_kernel<<<(kpad + 31) / 32, 32, 0, stream>>>(w, k, kpad, d, normalize,
                                                                  static_cast<unsigned char*>(ws.b_image), ws.cc,
                                                                  ws.codes_n, ws.meta);
    return cudaGetLastError();
}

cudaError_t launch_resolve(const float* z, int d, int norm_mode, const float* w, int k, const Workspace& ws,
                           int cand_cap, int64_t* idx, int num_sms, cudaStream_t stream) {
    resolve_kernel<<<num_sms * 4, 128, 0, stream>>>(z, d, norm_mode, w, k, ws.cc, ws.flag_rows, ws.cand, ws.cand_cnt,
                                                     cand_cap, idx, ws.bins_i32, ws.full_rows, ws.counters);
    return cudaGetLastError();
}

cudaError_t launch_exact_scan(const float* z, int64_t n, int d, int norm_mode, const float* w, int k,
                              const float* cc, const int* row_list, const int* n_rows_dev, int64_t* idx, int* bins,
                              float* denom_out, int num_sms, cudaStream_t stream) {
    const size_t smem = sizeof(float) * ((size_t)d * XR + XD * (XC + 1) + XR + XR * 16) + sizeof(int) * XR * 16;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(exact_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    int64_t blocks = n_rows_dev ? (int64_t)num_sms * 2 : (n + XR - 1) / XR;
    if (blocks > (int64_t)num_sms * 8) blocks = (int64_t)num_sms * 8;
    if (blocks < 1) blocks = 1;
    exact_scan_kernel<<<(int)blocks, 256, smem, stream>>>(z, n, d, norm_mode, w, k, cc, row_list, n_rows_dev, idx, bins,
                                                          denom_out);
    return cudaGetLastError();
}

}  // namespace mtvq
