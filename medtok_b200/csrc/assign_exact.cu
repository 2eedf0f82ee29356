// assign_exact.cu -- the fp32 "reference order" side of the assignment:
//   * codebook_prep_kernel : fp32 codebook -> fp16 tensor-core operand image + canonical ||w||^2 +
//                            the rigorous coarse-score error bound (margin)
//   * resolve_kernel       : near-tie rows: exact distances of the collected candidates, first-index argmin
//   * exact_scan_kernel    : exact distances against ALL codes (cross-check + candidate-overflow fallback)
// Every distance here is fl(fl(zz + cc) - 2*dot) with dot a sequential fmaf chain over d, i.e. bit-for-bit
// what the reference's CPU path computes (norm_ema_quantizer.py:175-179).
#include "assign_exact.h"
#include "assign_tc.h"
#include "common.cuh"

namespace mtvq {

// ------------------------------------------------------------------------------------------------
// codebook prep (K*D is tiny next to N*K*D, but it sits on the critical path of every step)
// ------------------------------------------------------------------------------------------------
// One WARP per code (a single thread per code ran every reduction as one dependent chain: 15 us for 8192 codes).  The
// canonical orders are kept exactly: canon_norm8's eight lane-accumulators run on lanes 0-7 and are folded in lane order,
// warp_canon_sumsq folds the 32 cascade accumulators (common.cuh).
__global__ void __launch_bounds__(256) codebook_prep_kernel(const float* __restrict__ w, int k, int kpad, int d, int flags,
                                                             unsigned char* __restrict__ image, float* __restrict__ cc,
                                                             float* __restrict__ codes_n, float* __restrict__ meta) {
    __shared__ __align__(16) float rows_s[8][kMaxDim];
    __shared__ float red_s[8][4];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int normalize = flags & 1;
    const bool dot_only = (flags & 2) != 0;      // MTVQ_SCORE_DOT: codes are ranked by z.w alone -> no -||w||^2/2 term
    const int nplanes = d >> 3;
    const __half hz = __float2half(0.0f);
    float* row = rows_s[wib];
    // running maxima of this warp's codes (lane 0): spread, eps, max norm; NaN anywhere -> everything to the exact path.
    // One set of atomics per CTA at the end: thousands of atomics on four addresses serialise (they were most of the time).
    float m_spread = 0.0f, m_eps = 0.0f, m_norm = 0.0f;
    bool bad = false;
    for (int code = blockIdx.x * 8 + wib; code < kpad; code += gridDim.x * 8) {
        const int tile = code >> 7, r = code & 127;
        unsigned char* tile_base = image + (size_t)tile * (nplanes + 2) * 2048 + r * 16;
        if (code >= k) {   // padding code: can never win (bias -60000), contributes nothing
            for (int j = lane; j < nplanes + 2; j += 32) {
                uint4 z4 = make_uint4(0, 0, 0, 0);
                if (j == nplanes) {
                    __half hb[8] = {__float2half(-60000.0f), hz, hz, hz, hz, hz, hz, hz};
                    z4 = *reinterpret_cast<uint4*>(hb);
                }
                *reinterpret_cast<uint4*>(tile_base + j * 2048) = z4;
            }
            if (lane == 0) cc[code] = 3.0e38f;
            continue;
        }
        __syncwarp();
        for (int i = lane; i < d; i += 32) row[i] = __ldg(w + (size_t)code * d + i);
        __syncwarp();
        if (normalize) {
            // canon_norm8: eight lane-accumulators on lanes 0-7, folded in lane order
            const int v = lane & 7;
            float a = 0.f;
            for (int c = 0; c < d; c += 8) { const float e = row[c + v]; a = __fadd_rn(a, __fmul_rn(e, e)); }
            float sm = __shfl_sync(0xffffffffu, a, 0);
#pragma unroll
            for (int t = 1; t < 8; ++t) sm = __fadd_rn(sm, __shfl_sync(0xffffffffu, a, t));
            float den = __fsqrt_rn(sm);
            den = den < kNormEps ? kNormEps : den;
            __syncwarp();
            for (int i = lane; i < d; i += 32) row[i] = __fdiv_rn(row[i], den);
            __syncwarp();
        }
        const float c2 = warp_canon_sumsq(row, d, lane);
        // the fp16 operand image: plane j holds elements [8j, 8j + 8) of the code as one 16-byte unit
        for (int j = lane; j < nplanes; j += 32) {
            __half h[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) h[i] = __float2half_rn(row[8 * j + i]);
            *reinterpret_cast<uint4*>(tile_base + j * 2048) = *reinterpret_cast<uint4*>(h);
        }
        if (normalize) for (int i = lane; i < d; i += 32) codes_n[(size_t)code * d + i] = row[i];
        if (lane != 0) continue;
        cc[code] = c2;
        // spread of the squared norms around 1 (the unit-norm codebooks of the EMA quantizer): if it is negligible the
        // kernels drop the bias K-step; two codes then differ by at most 2 x spread, covered by the slack in eps below
        const float spread = dot_only ? 0.0f : 0.5f * fabsf(c2 - 1.0f);
        // bias = -||w||^2 / 2 split into three fp16 terms (exact to ~2^-33 relative)
        const float b = dot_only ? 0.0f : -0.5f * c2;
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
        //   skipped bias step: scores shifted by up to kBiasSkipTol each way
        const float nw = sqrtf(c2) * 1.001f;
        const float eps = 9.7704e-4f * nw + 3.8147e-6f * (1.0f + nw + 0.5f * c2) + 1e-6f + 2.0f * kBiasSkipTol;
        if (eps == eps && spread == spread) {
            m_spread = fmaxf(m_spread, spread); m_eps = fmaxf(m_eps, eps); m_norm = fmaxf(m_norm, nw);
        } else {
            bad = true;
        }
    }
    if (lane == 0) { red_s[wib][0] = m_spread; red_s[wib][1] = m_eps; red_s[wib][2] = m_norm; red_s[wib][3] = bad ? 1.0f : 0.0f; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int t = 1; t < 8; ++t) {
            red_s[0][0] = fmaxf(red_s[0][0], red_s[t][0]); red_s[0][1] = fmaxf(red_s[0][1], red_s[t][1]);
            red_s[0][2] = fmaxf(red_s[0][2], red_s[t][2]); red_s[0][3] = fmaxf(red_s[0][3], red_s[t][3]);
        }
        // positive floats order like ints
        atomicMax(reinterpret_cast<int*>(meta + META_SPREAD), __float_as_int(red_s[0][0]));
        atomicMax(reinterpret_cast<int*>(meta + META_EPS), __float_as_int(red_s[0][1]));
        atomicMax(reinterpret_cast<int*>(meta + META_MARGIN), __float_as_int(2.0f * red_s[0][1]));
        atomicMax(reinterpret_cast<int*>(meta + META_MAXNORM), __float_as_int(red_s[0][2]));
        if (red_s[0][3] != 0.0f) {   // NaN in the codebook: send everything to the exact path, keep the bias step
            atomicMax(reinterpret_cast<int*>(meta + META_MARGIN), __float_as_int(INFINITY));
            atomicMax(reinterpret_cast<int*>(meta + META_SPREAD), __float_as_int(INFINITY));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// near-tie resolve: one WARP per flagged row.  The row's candidate list holds single columns (>= 0) and whole
// 32-column granules (-(granule + 1)).  The normalised row sits in shared memory (formed once, coalesced); a
// granule's 32 codes are staged through shared memory in 32-dimension slabs with coalesced loads, lane l then
// runs the sequential fmaf chain of code l.  Single columns are spread over the lanes and read their code directly.
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int kResWarps = 8;     // warps (rows in flight) per CTA
constexpr int kResSlab  = 32;    // dimensions staged per step

// lexicographic (NaN first, then value, then index) minimum: torch.argmin semantics
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) {
    const bool vn = v != v, bn = bv != bv;
    if (bi == 0x7fffffff) return i != 0x7fffffff;
    if (i == 0x7fffffff) return false;
    if (vn != bn) return vn;
    if (vn) return i < bi;
    return v < bv || (v == bv && i < bi);
}
}  // namespace

__global__ void __launch_bounds__(kResWarps * 32, 4) resolve_kernel(
    const float* __restrict__ z, int d, int norm_mode, const float* __restrict__ w, int k, const float* __restrict__ cc,
    const int* __restrict__ flag_rows, const int* __restrict__ cand, const int* __restrict__ cand_cnt, int cand_cap,
    int64_t* __restrict__ idx, int* __restrict__ bins, int* __restrict__ full_rows, int* __restrict__ counters) {
    __shared__ __align__(16) float xs[kResWarps][kMaxDim];                  // normalised row per warp
    // granule slab per warp; row stride 36 floats keeps 16-byte accesses aligned and conflict free: within a
    // quarter-warp phase the 8 lanes' rows (or 8 consecutive float4 of a row) fall into 8 different bank quads
    __shared__ __align__(16) float ws_[kResWarps][32][kResSlab + 4];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int nflag = counters[CNT_FLAGGED];
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    int ncands = 0;
    for (int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < nflag; p += nwarps) {
        const int row = flag_rows[p];
        // candidate lists are indexed by row; the count word has two fields (assign_tc.h kCandBackShift): columns listed
        // from the front of the list and -- second tile of a row listed under two chunks by the REFINE pass -- from its back
        const unsigned cw = (unsigned)cand_cnt[row];
        const int nfront = (int)(cw & ((1u << kCandBackShift) - 1u)), nback = (int)(cw >> kCandBackShift);
        const int cnt = nfront + nback;
        if (cnt <= 0 || cnt > cand_cap) {
            if (lane == 0) full_rows[atomicAdd(&counters[CNT_FULL], 1)] = row;
            continue;
        }
        const float* zr = z + (size_t)row * d;
        __syncwarp();
        for (int i = lane; i < d; i += 32) xs[wib][i] = __ldg(zr + i);      // raw row, one coalesced round trip
        __syncwarp();
        // canonical norm and squared norm with their independent accumulator chains spread over the lanes: the
        // additions inside each chain and the order in which the chains are folded are exactly those of
        // canon_norm8 / canon_sumsq (common.cuh), only the chains run side by side
        float den;
        if (norm_kind(norm_mode) == 0) {
            const int v = lane & 7;
            float a = 0.f;
            for (int c = 0; c < d; c += 8) { const float x = xs[wib][c + v]; a = __fadd_rn(a, __fmul_rn(x, x)); }
            float sm = __shfl_sync(0xffffffffu, a, 0);
#pragma unroll
            for (int t = 1; t < 8; ++t) sm = __fadd_rn(sm, __shfl_sync(0xffffffffu, a, t));
            den = __fsqrt_rn(sm);
            den = den < kNormEps ? kNormEps : den;
        } else {
            den = canon_denom([&](int i) { return xs[wib][i]; }, d, norm_mode);
        }
        __syncwarp();
        for (int i = lane; i < d; i += 32) xs[wib][i] = __fdiv_rn(xs[wib][i], den);
        __syncwarp();
        float zz;
        {
            const int u = lane >> 3, v = lane & 7;
            const int nvec = d >> 3, n4 = nvec >> 2;
            float a = 0.f;
            for (int c = 0; c < n4; ++c) { const float e = xs[wib][(c * 4 + u) * 8 + v]; a = __fadd_rn(a, __fmul_rn(e, e)); }
            if (u == 0)
                for (int i = n4 * 4; i < nvec; ++i) { const float e = xs[wib][i * 8 + v]; a = __fadd_rn(a, __fmul_rn(e, e)); }
            float f = __shfl_sync(0xffffffffu, a, v);                                 // acc[0][v]
            f = __fadd_rn(f, __shfl_sync(0xffffffffu, a, 8 + v));
            f = __fadd_rn(f, __shfl_sync(0xffffffffu, a, 16 + v));
            f = __fadd_rn(f, __shfl_sync(0xffffffffu, a, 24 + v));
            float fin = 0.f;
#pragma unroll
            for (int t = 0; t < 8; ++t) fin = __fadd_rn(fin, __shfl_sync(0xffffffffu, f, t));
            zz = fin;
        }
        const bool dots = score_dot(norm_mode);      // argmax z.w == argmin fl(fl(0 + 0) - 2 dot), first index on ties
        if (dots) zz = 0.0f;
        float best = 0.0f;
        int bcol = 0x7fffffff;
        const int mine = lane < cnt ? cand[(size_t)row * cand_cap + (lane < nfront ? lane : cand_cap - 1 - (lane - nfront))]
                                    : 0x7fffffff;                                          // cand_cap <= 32
        // single columns: lane j scores entry j
        if (lane < cnt && mine >= 0 && mine < k) {
            const float4* wr = reinterpret_cast<const float4*>(w + (size_t)mine * d);
            float dot = 0.0f;
            for (int i0 = 0; i0 < (d >> 2); i0 += 4) {     // d % 16 == 0: four 16-byte loads in flight per step
                float4 wv[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) wv[t] = __ldg(wr + i0 + t);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float4 xv = *reinterpret_cast<const float4*>(&xs[wib][4 * (i0 + t)]);
                    dot = __fmaf_rn(xv.x, wv[t].x, dot);
                    dot = __fmaf_rn(xv.y, wv[t].y, dot);
                    dot = __fmaf_rn(xv.z, wv[t].z, dot);
                    dot = __fmaf_rn(xv.w, wv[t].w, dot);
                }
            }
            best = canon_dist(zz, dots ? 0.0f : cc[mine], dot);
            bcol = mine;
            ++ncands;
        }
        // whole granules: lane l scores code 32 * granule + l.  A slab = 32 codes x 32 dims: lane t loads float4
        // (t & 7) of codes (t >> 3) + 4 r; the next slab's loads are in flight while this one's chain runs.
        for (int j = 0; j < cnt; ++j) {
            const int e = __shfl_sync(0xffffffffu, mine, j);
            if (e >= 0) continue;
            const int code0 = (-e - 1) * 32;
            const int code = code0 + lane;
            float4 pre[8];
            auto issue = [&](int d0) {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int cl = (lane >> 3) + 4 * r;
                    pre[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (code0 + cl < k && d0 + 4 * (lane & 7) < d)
                        pre[r] = __ldg(reinterpret_cast<const float4*>(w + (size_t)(code0 + cl) * d + d0) + (lane & 7));
                }
            };
            issue(0);
            float dot = 0.0f;
            for (int d0 = 0; d0 < d; d0 += kResSlab) {
                __syncwarp();
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    *reinterpret_cast<float4*>(&ws_[wib][(lane >> 3) + 4 * r][4 * (lane & 7)]) = pre[r];
                __syncwarp();
                if (d0 + kResSlab < d) issue(d0 + kResSlab);
                const int dl4 = min(kResSlab, d - d0) >> 2;   // d is a multiple of 16, not necessarily of 32
#pragma unroll 4
                for (int i = 0; i < dl4; ++i) {
                    const float4 xv = *reinterpret_cast<const float4*>(&xs[wib][d0 + 4 * i]);
                    const float4 wv = *reinterpret_cast<const float4*>(&ws_[wib][lane][4 * i]);
                    dot = __fmaf_rn(xv.x, wv.x, dot);
                    dot = __fmaf_rn(xv.y, wv.y, dot);
                    dot = __fmaf_rn(xv.z, wv.z, dot);
                    dot = __fmaf_rn(xv.w, wv.w, dot);
                }
            }
            if (code < k) {
                const float dist = canon_dist(zz, dots ? 0.0f : cc[code], dot);
                if (better(dist, code, best, bcol)) { best = dist; bcol = code; }
                ++ncands;
            }
        }
        // warp-wide lexicographic minimum
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bcol, o);
            if (better(ov, oi, best, bcol)) { best = ov; bcol = oi; }
        }
        if (lane == 0) {
            if (bcol != 0x7fffffff) {
                idx[row] = (int64_t)bcol;
                atomicAdd(&bins[bcol], 1);
            } else {
                full_rows[atomicAdd(&counters[CNT_FULL], 1)] = row;
            }
        }
    }
    __syncwarp();
    ncands = __reduce_add_sync(0xffffffffu, ncands);
    if (lane == 0 && ncands) atomicAdd(&counters[CNT_CANDS], ncands);
}

// ------------------------------------------------------------------------------------------------
// resolve for the two-pass route: every candidate is a single column (at most cand_cap <= 16 per row, 2-3 on average)
// and the row norms are known (the GROUP pass wrote them), so a whole warp per row leaves most lanes idle.  Eight lanes
// per row, four rows per warp: lane v of a group scores candidates v and v + 8 (the same sequential fmaf chain, the same
// canon_sumsq chains -- four per lane -- and fold order), three shuffle rounds pick the lexicographic minimum, and the
// group finishes the row (z_q, squared error, per-code sums) when the fused outputs are wanted.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 4) resolve8_kernel(
    const float* __restrict__ z, const float* __restrict__ denom, int d, int norm_mode, const float* __restrict__ w, int k,
    const float* __restrict__ cc, const int* __restrict__ flag_rows, const int* __restrict__ cand,
    const int* __restrict__ cand_cnt, int cand_cap, int64_t* __restrict__ idx, int* __restrict__ bins,
    int* __restrict__ full_rows, int* __restrict__ counters, float* __restrict__ z_q, double* __restrict__ stats) {
    __shared__ __align__(16) float xs_all[8][4][kMaxDim];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, g = lane >> 3, v = lane & 7;
    float* xs = xs_all[wib][g];
    const int nflag = counters[CNT_FLAGGED];
    const int ngroups = (gridDim.x * blockDim.x) >> 3;
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int rounds = (nflag + ngroups - 1) / ngroups;          // warp-uniform trip count (the shuffles below are full-warp)
    const bool dots = score_dot(norm_mode);
    const int nv4 = d >> 2, nvec = d >> 3, n4 = nvec >> 2;
    int ncands = 0;
    double sse = 0.0;
    for (int it = 0; it < rounds; ++it) {
        const int p = gid + it * ngroups;
        const bool active = p < nflag;
        const int row = active ? flag_rows[p] : 0;
        const unsigned cw = active ? (unsigned)cand_cnt[row] : 0u;
        const int nfront = (int)(cw & ((1u << kCandBackShift) - 1u)), nback = (int)(cw >> kCandBackShift);
        const int cnt = nfront + nback;
        const bool ok = active && cnt > 0 && cnt <= cand_cap;
        __syncwarp();
        if (ok) {
            const float4* zr = reinterpret_cast<const float4*>(z + (size_t)row * d);
            for (int j = v; j < nv4; j += 8) reinterpret_cast<float4*>(xs)[j] = __ldg(zr + j);
        }
        const float den = ok ? denom[row] : 1.0f;
        __syncwarp();
        if (ok)
            for (int i = v; i < d; i += 8) xs[i] = __fdiv_rn(xs[i], den);
        __syncwarp();
        // canon_sumsq: lane v runs the chains acc[0..3][v], folds them in order, the eight lanes are then added in order
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (ok && !dots) {
            for (int c = 0; c < n4; ++c) {
                const float e0 = xs[(c * 4 + 0) * 8 + v], e1 = xs[(c * 4 + 1) * 8 + v], e2 = xs[(c * 4 + 2) * 8 + v],
                            e3 = xs[(c * 4 + 3) * 8 + v];
                a0 = __fadd_rn(a0, __fmul_rn(e0, e0)); a1 = __fadd_rn(a1, __fmul_rn(e1, e1));
                a2 = __fadd_rn(a2, __fmul_rn(e2, e2)); a3 = __fadd_rn(a3, __fmul_rn(e3, e3));
            }
            for (int i = n4 * 4; i < nvec; ++i) { const float e = xs[i * 8 + v]; a0 = __fadd_rn(a0, __fmul_rn(e, e)); }
            a0 = __fadd_rn(__fadd_rn(__fadd_rn(a0, a1), a2), a3);
        }
        float zz = 0.0f;
#pragma unroll
        for (int t = 0; t < 8; ++t) zz = __fadd_rn(zz, __shfl_sync(0xffffffffu, a0, (lane & 24) + t));
        if (dots) zz = 0.0f;
        float best = 0.0f;
        int bcol = 0x7fffffff;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int ci = v + 8 * r;
            if (ok && ci < cnt) {
                const int code = cand[(size_t)row * cand_cap + (ci < nfront ? ci : cand_cap - 1 - (ci - nfront))];
                if (code >= 0 && code < k) {
                    const float4* wr = reinterpret_cast<const float4*>(w + (size_t)code * d);
                    float dot = 0.0f;
                    for (int i0 = 0; i0 < nv4; i0 += 4) {        // d % 16 == 0: four 16-byte loads in flight per step
                        float4 wv[4];
#pragma unroll
                        for (int t = 0; t < 4; ++t) wv[t] = __ldg(wr + i0 + t);
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const float4 xv = *reinterpret_cast<const float4*>(&xs[4 * (i0 + t)]);
                            dot = __fmaf_rn(xv.x, wv[t].x, dot);
                            dot = __fmaf_rn(xv.y, wv[t].y, dot);
                            dot = __fmaf_rn(xv.z, wv[t].z, dot);
                            dot = __fmaf_rn(xv.w, wv[t].w, dot);
                        }
                    }
                    const float dist = canon_dist(zz, dots ? 0.0f : cc[code], dot);
                    if (better(dist, code, best, bcol)) { best = dist; bcol = code; }
                    ++ncands;
                }
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {                        // stays inside the eight-lane group
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bcol, o);
            if (better(ov, oi, best, bcol)) { best = ov; bcol = oi; }
        }
        if (active && v == 0) {
            if (bcol != 0x7fffffff) {
                idx[row] = (int64_t)bcol;
                atomicAdd(&bins[bcol], 1);
            } else {
                full_rows[atomicAdd(&counters[CNT_FULL], 1)] = row;      // empty / overflowing list, NaN distances
            }
        }
        if (z_q != nullptr && ok && bcol != 0x7fffffff) {
            double* esum = stats + k + (size_t)bcol * d;
            const float4* wr = reinterpret_cast<const float4*>(w + (size_t)bcol * d);
            for (int j = v; j < nv4; j += 8) {
                const float4 zf = reinterpret_cast<const float4*>(xs)[j];
                const float4 c = __ldg(wr + j);
                const float d0 = __fsub_rn(c.x, zf.x), d1 = __fsub_rn(c.y, zf.y), d2 = __fsub_rn(c.z, zf.z), d3 = __fsub_rn(c.w, zf.w);
                sse += (double)__fmul_rn(d0, d0) + (double)__fmul_rn(d1, d1) + (double)__fmul_rn(d2, d2) + (double)__fmul_rn(d3, d3);
                float4 o;
                o.x = __fadd_rn(zf.x, d0); o.y = __fadd_rn(zf.y, d1); o.z = __fadd_rn(zf.z, d2); o.w = __fadd_rn(zf.w, d3);
                reinterpret_cast<float4*>(z_q + (size_t)row * d)[j] = o;
                atomicAdd(esum + 4 * j + 0, (double)zf.x); atomicAdd(esum + 4 * j + 1, (double)zf.y);
                atomicAdd(esum + 4 * j + 2, (double)zf.z); atomicAdd(esum + 4 * j + 3, (double)zf.w);
            }
        }
    }
    __syncwarp();
    ncands = __reduce_add_sync(0xffffffffu, ncands);
    if (lane == 0 && ncands) atomicAdd(&counters[CNT_CANDS], ncands);
    if (z_q != nullptr) {
        sse = warp_sum(sse);
        if (lane == 0 && sse != 0.0) atomicAdd(&stats[(size_t)k + (size_t)k * d], sse);
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
                zzs[threadIdx.x] = score_dot(norm_mode) ? 0.0f : canon_sumsq([&](int i) { return As[(size_t)i * XR + threadIdx.x]; }, d);
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
                        const float dist = canon_dist(zz, score_dot(norm_mode) ? 0.0f : __ldg(cc + code), acc[ai][bi]);
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
cudaError_t launch_codebook_prep(const float* w, int k, int d, int flags, const Workspace& ws,
                                 cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(ws.meta, 0, sizeof(float) * META_COUNT, stream);
    if (e != cudaSuccess) return e;
    const int kpad = pad_codes(k);
    int grid = (kpad + 7) / 8;
    if (grid > 592) grid = 592;                      // <= 4 CTAs per SM, each walking its share of the codes
    codebook_prep_kernel<<<grid, 256, 0, stream>>>(w, k, kpad, d, flags,
                                                                  static_cast<unsigned char*>(ws.b_image), ws.cc,
                                                                  ws.codes_n, ws.meta);
    ++g_kernel_launches;
    return cudaGetLastError();
}

cudaError_t launch_resolve8(const float* z, const float* denom, int d, int norm_mode, const float* w, int k,
                            const Workspace& ws, int cand_cap, int64_t* idx, int num_sms, cudaStream_t stream, float* z_q,
                            double* stats) {
    resolve8_kernel<<<num_sms * 4, 256, 0, stream>>>(z, denom, d, norm_mode, w, k, ws.cc, ws.flag_rows, ws.cand, ws.cand_cnt,
                                                     cand_cap, idx, ws.bins_i32, ws.full_rows, ws.counters, z_q, stats);
    ++g_kernel_launches;
    return cudaGetLastError();
}

cudaError_t launch_resolve(const float* z, int d, int norm_mode, const float* w, int k, const Workspace& ws,
                           int cand_cap, int64_t* idx, int num_sms, cudaStream_t stream) {
    resolve_kernel<<<num_sms * 4, kResWarps * 32, 0, stream>>>(z, d, norm_mode, w, k, ws.cc, ws.flag_rows, ws.cand,
                                                                ws.cand_cnt, cand_cap, idx, ws.bins_i32, ws.full_rows,
                                                                ws.counters);
    ++g_kernel_launches;
    return cudaGetLastError();
}

cudaError_t launch_exact_scan(const float* z, int64_t n, int d, int norm_mode, const float* w, int k,
                              const float* cc, const int* row_list, const int* n_rows_dev, int64_t* idx, int* bins,
                              float* denom_out, int num_sms, cudaStream_t stream) {
    const size_t smem = sizeof(float) * ((size_t)d * XR + XD * (XC + 1) + XR + XR * 16) + sizeof(int) * XR * 16;
    static PerDeviceOnce once;
    bool& configured = once.flag();
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
    ++g_kernel_launches;
    return cudaGetLastError();
}

}  // namespace mtvq
