// assign_tc.cu -- the distance hot loop on tcgen05 tensor cores (sm_100a).
//
// Replaces, for one pass of rows against a whole codebook (norm_ema_quantizer.py:175-179 and
// vector_quantization_soft_one_new.py:120-125,157,203):
//     d = ||z||^2 + ||w||^2 - 2 z.w^T   (N x K fp32, materialised by the reference)   + argmin / topk
// with a persistent, warp-specialised kernel that never writes a score to HBM:
//
//   coarse score  s[n,k] = zf_n . w_k - ||w_k||^2 / 2      (argmax s == argmin d)
//     fp16 operands, fp32 accumulation in tensor memory; the -||w||^2/2 term rides in one extra
//     UMMA K-step as three fp16 terms (A side = 1,1,1), so the epilogue is a pure max.
//   epilogue      one thread per row streams its 32-column granules out of TMEM (tcgen05.ld), reduces
//     them with FMNMX3 trees and keeps per-row running state in registers.
//
// The coarse score is only used to SELECT candidates: |s - exact| <= eps (rigorous bound computed by
// the codebook prep kernel), so every code within margin = 2 eps of the best coarse score is
// re-scored in exact fp32 reference order by resolve kernels (assign_exact.cu).  Modes:
//   ARGMAX   best column + "is anything else within margin" flag           (hard path, pass 1)
//   THRESH   per-row lower bound on the top-k-th score (k-th largest granule max)   (soft, pass A)
//   COLLECT  append every column with s >= best_val[row] - margin          (near-tie rows / soft pass B)
//   DUMP     raw scores to global memory (tests only)
//
// Roles (800 threads, 1 CTA per SM, persistent over row blocks of RT x 128 rows):
//   warps 0-15  epilogue: warpgroup e owns accumulator slot e (128 TMEM columns)
//   warps 16-19 A loader: rows -> canonical l2norm -> fp16 -> shared memory operand planes
//   warp  20    B producer: cp.async.bulk of pre-laid-out codebook tiles (L2 resident) into a ring
//   warps 21-24 MMA issuers, one elected thread each, issuer j feeds accumulator slot j (a single
//               issuing thread is instruction-latency bound at ~140 SASS instructions per tile);
//               warp 21 also owns the TMEM allocation
//
// Shared-memory operand layout (both A and B): K-major, no swizzle; "plane" j holds, for every row
// of the tile, the 8 fp16 elements [8j, 8j+8) as one 16-byte unit => LBO = 2048 B (next plane),
// SBO = 128 B (next 8-row core matrix).  One UMMA K-step (16 elements) = 2 planes = 4096 B.
#include "assign_tc.h"
#include "common.cuh"

namespace mtvq {

namespace {

constexpr int kEpiWarps   = 16;
constexpr int kLoadWarp0  = 16;
constexpr int kProdWarp   = 20;
constexpr int kMmaWarp    = 21;         // 21..24: one MMA issuer per accumulator slot
constexpr int kThreads    = 25 * 32;

constexpr int kPlaneBytes = kTileRows * 16;          // 2048
constexpr int kKblkBytes  = 8 * kPlaneBytes;         // 16384: 64 fp16 per row
constexpr int kBiasBytes  = 2 * kPlaneBytes;         // 4096
constexpr int kDataStages = 4;
constexpr int kBiasStages = 3;

// dynamic shared memory map
constexpr int kOffA      = 0;                                    // RT tiles, <= 64 KB in total
constexpr int kOffOnes   = 65536;
constexpr int kOffBData  = kOffOnes + kBiasBytes;
constexpr int kOffBBias  = kOffBData + kDataStages * kKblkBytes;
constexpr int kOffSlots  = kOffBBias + kBiasStages * kBiasBytes;  // 4 WG x 128 rows x 128 B
constexpr int kOffState  = kOffSlots + 4 * 128 * 128;             // 4 WG x 128 rows x 16 B
constexpr int kOffBars   = kOffState + 4 * 128 * 16;
constexpr int kNumBars   = 2 * kDataStages + 2 * kBiasStages + 4 + 4 + 4 + 4;
constexpr int kOffTmem   = kOffBars + kNumBars * 8;
constexpr int kSmemBytes = kOffTmem + 16;

struct Bars {
    uint64_t* bd_full;   // [kDataStages]
    uint64_t* bd_empty;  // [kDataStages]
    uint64_t* bb_full;   // [kBiasStages]
    uint64_t* bb_empty;  // [kBiasStages]
    uint64_t* a_full;    // [4]
    uint64_t* a_empty;   // [4]
    uint64_t* acc_full;  // [4]
    uint64_t* acc_empty; // [4]
};

__device__ __forceinline__ Bars carve_bars(unsigned char* smem) {
    uint64_t* b = reinterpret_cast<uint64_t*>(smem + kOffBars);
    Bars r;
    r.bd_full = b;                 b += kDataStages;
    r.bd_empty = b;                b += kDataStages;
    r.bb_full = b;                 b += kBiasStages;
    r.bb_empty = b;                b += kBiasStages;
    r.a_full = b;                  b += 4;
    r.a_empty = b;                 b += 4;
    r.acc_full = b;                b += 4;
    r.acc_empty = b;
    return r;
}

__device__ __forceinline__ float max32(const float (&v)[32]) {
    // written as a chain so ptxas fuses pairs into FMNMX3 (16 instructions for 32 values)
    float a = fmaxf(fmaxf(v[0], v[1]), v[2]);
    float b = fmaxf(fmaxf(v[3], v[4]), v[5]);
    float c = fmaxf(fmaxf(v[6], v[7]), v[8]);
    float d = fmaxf(fmaxf(v[9], v[10]), v[11]);
    float e = fmaxf(fmaxf(v[12], v[13]), v[14]);
    float f = fmaxf(fmaxf(v[15], v[16]), v[17]);
    float g = fmaxf(fmaxf(v[18], v[19]), v[20]);
    float h = fmaxf(fmaxf(v[21], v[22]), v[23]);
    float i = fmaxf(fmaxf(v[24], v[25]), v[26]);
    float j = fmaxf(fmaxf(v[27], v[28]), v[29]);
    float k = fmaxf(v[30], v[31]);
    a = fmaxf(fmaxf(a, b), c);
    d = fmaxf(fmaxf(d, e), f);
    g = fmaxf(fmaxf(g, h), i);
    j = fmaxf(j, k);
    return fmaxf(fmaxf(fmaxf(a, d), g), j);
}

template <int NTOP>
__device__ __forceinline__ void topn_insert(float (&t)[NTOP], float x) {
#pragma unroll
    for (int i = NTOP - 1; i >= 1; --i) t[i] = fmaxf(t[i], fminf(t[i - 1], x));
    t[0] = fmaxf(t[0], x);
}

// slot rows are XOR-swizzled in 16-byte units so that 8 consecutive rows hit 8 different bank groups
__device__ __forceinline__ void slot_store(unsigned char* slot_row, int r, const float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float4 q = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        *reinterpret_cast<float4*>(slot_row + ((i ^ (r & 7)) << 4)) = q;
    }
}
__device__ __forceinline__ float slot_load(const unsigned char* slot_row, int r, int col) {
    return *reinterpret_cast<const float*>(slot_row + (((col >> 2) ^ (r & 7)) << 4) + ((col & 3) << 2));
}

}  // namespace

template <int RT, int MODE, int NTOP>
__global__ void __launch_bounds__(kThreads, 1) coarse_kernel(const CoarseParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int d         = p.d;
    const int nplanes   = d >> 3;
    const int nkblk     = (nplanes + 7) >> 3;
    const int a_tile_b  = nplanes * kPlaneBytes;
    const int tile_b    = (nplanes + 2) * kPlaneBytes;   // one codebook tile in the global image
    const int nchunks   = p.kpad / kTileCodes;
    const int64_t nrows = p.n_rows_dev ? (int64_t)(*p.n_rows_dev) : p.n;
    const int rows_blk  = RT * kTileRows;
    const int nblocks   = (int)((nrows + rows_blk - 1) / rows_blk);

    Bars bars = carve_bars(smem);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kOffTmem);

    // ---- one-time setup ------------------------------------------------------------------------
    if (threadIdx.x == 0) {
        // a codebook stage is read by the RT tiles (= RT different issuers) of one chunk
        for (int i = 0; i < kDataStages; ++i) { mbar_init(&bars.bd_full[i], 1); mbar_init(&bars.bd_empty[i], RT); }
        for (int i = 0; i < kBiasStages; ++i) { mbar_init(&bars.bb_full[i], 1); mbar_init(&bars.bb_empty[i], RT); }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&bars.a_full[i], 128);
            mbar_init(&bars.a_empty[i], 4 / RT);   // one commit from each issuer that reads the tile
            mbar_init(&bars.acc_full[i], 1);
            mbar_init(&bars.acc_empty[i], 128);
        }
        fence_mbar_init();
    }
    if (warp >= kLoadWarp0 && warp < kLoadWarp0 + 4) {
        // constant A-side planes of the bias K-step: row = (1, 1, 1, 0, 0, 0, 0, 0 | 0 x 8)
        const int r = threadIdx.x - kLoadWarp0 * 32;
        const __half one = __float2half(1.0f), zero = __float2half(0.0f);
        __half2 h01 = __halves2half2(one, one), h23 = __halves2half2(one, zero), hz = __halves2half2(zero, zero);
        uint4 q0, qz;
        q0.x = *reinterpret_cast<uint32_t*>(&h01); q0.y = *reinterpret_cast<uint32_t*>(&h23);
        q0.z = *reinterpret_cast<uint32_t*>(&hz);  q0.w = q0.z;
        qz.x = qz.y = qz.z = qz.w = q0.z;
        *reinterpret_cast<uint4*>(smem + kOffOnes + r * 16) = q0;
        *reinterpret_cast<uint4*>(smem + kOffOnes + kPlaneBytes + r * 16) = qz;
        fence_proxy_async_smem();
    }
    if (warp == kMmaWarp) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    // ============================================================================================
    if (warp < kEpiWarps) {
        // ---------------------------------- EPILOGUE --------------------------------------------
        const int e  = warp >> 2;                 // warpgroup == accumulator slot
        const int g  = e % RT;                    // row tile served by this warpgroup
        const int r  = (warp & 3) * 32 + lane;    // row within the tile == TMEM lane
        constexpr int G = 4 / RT;                 // warpgroups sharing one row tile
        const float margin = p.meta[META_MARGIN];
        const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(e * kTileCodes);
        unsigned char* my_slot = smem + kOffSlots + (e * 128 + r) * 128;
        float4* my_state = reinterpret_cast<float4*>(smem + kOffState) + (e * 128 + r);
        uint32_t uses = 0;

        for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
            const int64_t pos = (int64_t)blk * rows_blk + g * kTileRows + r;
            const bool valid = pos < nrows;
            const int64_t row = valid ? (p.row_list ? (int64_t)p.row_list[pos] : pos) : 0;

            float m1 = -INFINITY, m2 = -INFINITY, thr = -INFINITY;   // ARGMAX
            int best_chunk = 0;
            float top[NTOP];                                          // THRESH
#pragma unroll
            for (int i = 0; i < NTOP; ++i) top[i] = -INFINITY;
            float T = INFINITY;                                       // COLLECT
            if (MODE == MODE_COLLECT && valid) T = p.best_val[row] - margin;

            for (int c = e / RT; c < nchunks; c += G) {
                mbar_wait(&bars.acc_full[e], uses & 1);
                ++uses;
                tc_fence_after();
#pragma unroll 1
                for (int q = 0; q < 4; ++q) {
                    float v[32];
                    tmem_ld32(t_lane + q * kChunk, v);
                    tmem_ld_wait();
                    const int chunk = c * 4 + q;
                    if (MODE == MODE_DUMP) {
                        if (valid && pos < p.dump_rows) {
                            float4* dst = reinterpret_cast<float4*>(p.dump + pos * p.kpad + chunk * kChunk);
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                        }
                        continue;
                    }
                    const float cmax = max32(v);
                    if (MODE == MODE_ARGMAX) {
                        const bool trig = cmax >= thr;
                        if (__any_sync(0xffffffffu, trig)) {
                            if (trig) {
                                if (cmax > m1) {          // new record: remember the granule
                                    m2 = fmaxf(m2, m1);
                                    m1 = cmax;
                                    thr = m1 - margin;
                                    best_chunk = chunk;
                                    slot_store(my_slot, r, v);
                                } else {                  // another granule within the margin
                                    m2 = fmaxf(m2, cmax);
                                }
                            }
                        }
                    } else if (MODE == MODE_THRESH) {
                        topn_insert<NTOP>(top, cmax);
                    } else if (MODE == MODE_COLLECT) {
                        const bool trig = cmax >= T;
                        if (__any_sync(0xffffffffu, trig)) {
                            if (trig) {
                                // static indices only: a runtime index would push v[] into local memory
                                uint32_t hits = 0;
#pragma unroll
                                for (int i = 0; i < 32; ++i) hits |= (v[i] >= T) ? (1u << i) : 0u;
                                while (hits) {
                                    const int i = __ffs(hits) - 1;
                                    hits &= hits - 1;
                                    const int at = atomicAdd(&p.cand_cnt[pos], 1);
                                    if (at < p.cand_cap) p.cand[pos * p.cand_cap + at] = chunk * kChunk + i;
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&bars.acc_empty[e]);
            }

            // ---- end of row block: merge the warpgroups that shared this row tile, emit ----------
            if (MODE == MODE_ARGMAX) {
                if (G > 1) {
                    *my_state = make_float4(m1, m2, __int_as_float(best_chunk), 0.f);
                    named_bar_sync(1 + g, G * 128);
                }
                if (e < RT) {
                    float M = m1, m2all = m2;
                    if (G > 1) {
#pragma unroll
                        for (int j = 1; j < G; ++j) {
                            const float4 s = reinterpret_cast<const float4*>(smem + kOffState)[(e + j * RT) * 128 + r];
                            M = fmaxf(M, s.x);
                            m2all = fmaxf(m2all, s.y);
                        }
                    }
                    const float thrF = M - margin;
                    int cnt = 0, col = -1;
#pragma unroll
                    for (int j = 0; j < G; ++j) {
                        const int ee = e + j * RT;
                        float mj = m1;
                        int cj = best_chunk;
                        if (j > 0) {
                            const float4 s = reinterpret_cast<const float4*>(smem + kOffState)[ee * 128 + r];
                            mj = s.x;
                            cj = __float_as_int(s.z);
                        }
                        if (mj >= thrF) {
                            const unsigned char* sl = smem + kOffSlots + (ee * 128 + r) * 128;
#pragma unroll 1
                            for (int i = 0; i < 32; ++i) {
                                if (slot_load(sl, r, i) >= thrF) { ++cnt; col = cj * kChunk + i; }
                            }
                        }
                    }
                    const bool flagged = valid && ((cnt != 1) || (m2all >= thrF) || (col >= p.k));
                    if (valid) {
                        p.best_val[row] = M;
                        p.best_col[row] = flagged ? -1 : col;
                    }
                    const unsigned ball = __ballot_sync(0xffffffffu, flagged);
                    if (ball) {
                        int base = 0;
                        if (lane == 0) base = atomicAdd(&p.counters[CNT_FLAGGED], __popc(ball));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        if (flagged) p.flag_rows[base + __popc(ball & ((1u << lane) - 1))] = (int)row;
                    }
                }
                if (G > 1) named_bar_sync(1 + g, G * 128);
            } else if (MODE == MODE_THRESH) {
                float* st = reinterpret_cast<float*>(smem + kOffSlots);   // [4][128][8]
                if (G > 1) {
#pragma unroll
                    for (int i = 0; i < NTOP; ++i) st[(e * 128 + r) * 8 + i] = top[i];
                    named_bar_sync(1 + g, G * 128);
                }
                if (e < RT) {
                    if (G > 1) {
#pragma unroll
                        for (int j = 1; j < G; ++j)
#pragma unroll
                            for (int i = 0; i < NTOP; ++i) topn_insert<NTOP>(top, st[((e + j * RT) * 128 + r) * 8 + i]);
                    }
                    if (valid) {
                        float kth = top[0];
#pragma unroll
                        for (int i = 1; i < NTOP; ++i)
                            if (i < p.topk) kth = top[i];
                        p.best_val[row] = kth;
                    }
                }
                if (G > 1) named_bar_sync(1 + g, G * 128);
            }
        }
    } else if (warp < kLoadWarp0 + 4) {
        // ---------------------------------- A LOADER --------------------------------------------
        const int r = threadIdx.x - kLoadWarp0 * 32;
        uint32_t bi = 0;
        for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x, ++bi) {
            for (int g = 0; g < RT; ++g) {
                const int64_t pos = (int64_t)blk * rows_blk + g * kTileRows + r;
                const bool valid = pos < nrows;
                const int64_t row = valid ? (p.row_list ? (int64_t)p.row_list[pos] : pos) : 0;
                const float* zr = p.z + row * d;
                float den = 1.0f;
                if (valid) {
                    den = canon_denom([&](int i) { return __ldg(zr + i); }, d, p.norm_mode);
                    if (p.denom) p.denom[row] = den;
                }
                mbar_wait(&bars.a_empty[g], (bi & 1) ^ 1);
                unsigned char* tile = smem + kOffA + g * a_tile_b + r * 16;
                for (int j = 0; j < nplanes; ++j) {
                    uint4 q = make_uint4(0, 0, 0, 0);
                    if (valid) {
                        const float4 x0 = __ldg(reinterpret_cast<const float4*>(zr + 8 * j));
                        const float4 x1 = __ldg(reinterpret_cast<const float4*>(zr + 8 * j + 4));
                        __half2 h0 = __floats2half2_rn(__fdiv_rn(x0.x, den), __fdiv_rn(x0.y, den));
                        __half2 h1 = __floats2half2_rn(__fdiv_rn(x0.z, den), __fdiv_rn(x0.w, den));
                        __half2 h2 = __floats2half2_rn(__fdiv_rn(x1.x, den), __fdiv_rn(x1.y, den));
                        __half2 h3 = __floats2half2_rn(__fdiv_rn(x1.z, den), __fdiv_rn(x1.w, den));
                        q.x = *reinterpret_cast<uint32_t*>(&h0);
                        q.y = *reinterpret_cast<uint32_t*>(&h1);
                        q.z = *reinterpret_cast<uint32_t*>(&h2);
                        q.w = *reinterpret_cast<uint32_t*>(&h3);
                    }
                    *reinterpret_cast<uint4*>(tile + j * kPlaneBytes) = q;
                }
                fence_proxy_async_smem();
                mbar_arrive(&bars.a_full[g]);
            }
        }
    } else if (warp == kProdWarp) {
        // ---------------------------------- B PRODUCER ------------------------------------------
        if (elect_one()) {
            const unsigned char* img = static_cast<const unsigned char*>(p.b_image);
            uint32_t itD = 0, itB = 0;
            for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
                for (int c = 0; c < nchunks; ++c) {
                    const unsigned char* tile = img + (size_t)c * tile_b;
                    for (int kb = 0; kb < nkblk; ++kb, ++itD) {
                        const int st = itD % kDataStages;
                        const uint32_t bytes = (uint32_t)min(8, nplanes - 8 * kb) * kPlaneBytes;
                        mbar_wait(&bars.bd_empty[st], ((itD / kDataStages) & 1) ^ 1);
                        mbar_arrive_expect_tx(&bars.bd_full[st], bytes);
                        bulk_g2s(smem + kOffBData + st * kKblkBytes, tile + (size_t)kb * kKblkBytes, bytes,
                                 &bars.bd_full[st]);
                    }
                    const int sb = itB % kBiasStages;
                    mbar_wait(&bars.bb_empty[sb], ((itB / kBiasStages) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bars.bb_full[sb], kBiasBytes);
                    bulk_g2s(smem + kOffBBias + sb * kBiasBytes, tile + (size_t)nplanes * kPlaneBytes, kBiasBytes,
                             &bars.bb_full[sb]);
                    ++itB;
                }
            }
        }
    } else if (warp >= kMmaWarp) {
        // ---------------------------------- MMA ISSUERS -----------------------------------------
        // issuer j owns accumulator slot j, i.e. the tiles (c, g) with (c*RT + g) % 4 == j
        if (elect_one()) {
            const int j = warp - kMmaWarp;
            const int g = j % RT;
            constexpr int G = 4 / RT;
            const uint32_t idesc = umma_idesc_f16(kTileRows, kTileCodes);
            const uint32_t lbo = p.lbo_bytes, sbo = p.sbo_bytes;
            // descriptors are built once; per MMA only the 14-bit start-address field moves
            // (units of 16 B: one K-step = 2 planes = 256, one K-block / data stage = 1024)
            const uint64_t dA = umma_smem_desc(smem_u32(smem + kOffA + g * a_tile_b), lbo, sbo);
            const uint64_t dOnes = umma_smem_desc(smem_u32(smem + kOffOnes), lbo, sbo);
            const uint64_t dBD0 = umma_smem_desc(smem_u32(smem + kOffBData), lbo, sbo);
            const uint64_t dBB0 = umma_smem_desc(smem_u32(smem + kOffBBias), lbo, sbo);
            const uint32_t tacc = tmem_base + (uint32_t)(j * kTileCodes);
            uint32_t uses = 0, bi = 0;
            uint32_t chunk0 = 0;                     // chunks issued by the CTA before this block
            for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x, ++bi, chunk0 += nchunks) {
                if (j / RT >= nchunks) {
                    // fewer chunks than issuers: nothing to issue, but the row tile's release barrier
                    // still expects this issuer once per block (paced by the tile's own arrival)
                    mbar_wait(&bars.a_full[g], bi & 1);
                    mbar_arrive(&bars.a_empty[g]);
                    continue;
                }
                bool first = true;
                for (int c = j / RT; c < nchunks; c += G) {
                    const uint32_t gc = chunk0 + c;                       // running chunk number
                    mbar_wait(&bars.acc_empty[j], (uses & 1u) ^ 1u);
                    ++uses;
                    if (first) { mbar_wait(&bars.a_full[g], bi & 1); first = false; }
                    tc_fence_after();
                    const uint32_t itD = gc * nkblk;
                    for (int kb = 0; kb < nkblk; ++kb) {
                        const uint32_t it = itD + kb;
                        const uint32_t st = it % kDataStages;
                        mbar_wait(&bars.bd_full[st], (it / kDataStages) & 1);
                        tc_fence_after();
                        const int nsteps = min(8, nplanes - 8 * kb) >> 1;
                        uint64_t da = dA + (uint64_t)(kb * (kKblkBytes >> 4));
                        uint64_t db = dBD0 + (uint64_t)(st * (kKblkBytes >> 4));
                        for (int ks = 0; ks < nsteps; ++ks) {
                            umma_f16(tacc, da, db, idesc, (kb | ks) != 0);
                            da += (2 * kPlaneBytes) >> 4;
                            db += (2 * kPlaneBytes) >> 4;
                        }
                        umma_commit(&bars.bd_empty[st]);
                    }
                    const uint32_t sb = gc % kBiasStages;
                    mbar_wait(&bars.bb_full[sb], (gc / kBiasStages) & 1);
                    tc_fence_after();
                    umma_f16(tacc, dOnes, dBB0 + (uint64_t)(sb * (kBiasBytes >> 4)), idesc, 1u);
                    umma_commit(&bars.bb_empty[sb]);
                    umma_commit(&bars.acc_full[j]);
                    if (c + G >= nchunks) umma_commit(&bars.a_empty[g]);   // this issuer's last tile of the block
                }
            }
        }
    }

    // ---- teardown ----------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------------
template <int RT, int MODE, int NTOP>
static cudaError_t launch_one(const CoarseParams& p, int grid, cudaStream_t stream) {
    auto kern = coarse_kernel<RT, MODE, NTOP>;
    static bool configured = false;   // one per instantiation
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    kern<<<grid, kThreads, kSmemBytes, stream>>>(p);
    return cudaGetLastError();
}

int coarse_row_tiles(int d, int64_t nrows, int num_sms) {
    int rt = d <= 64 ? 4 : (d <= 128 ? 2 : 1);
    // keep every SM busy when there are few rows: shrink the block until there are >= 2 waves
    while (rt > 1 && (nrows + rt * 128 - 1) / (rt * 128) < 2 * (int64_t)num_sms) rt >>= 1;
    return rt;
}

cudaError_t launch_coarse(const CoarseParams& p, int mode, int row_tiles, int num_sms, cudaStream_t stream) {
    // NTOP only matters in THRESH mode (size of the per-row top list): 5 covers MedTok's k = 5
    const int ntop = (mode == MODE_THRESH && p.topk > 5) ? 8 : 5;
#define MTVQ_CASE(RT_, MODE_, NTOP_) \
    if (row_tiles == RT_ && mode == MODE_ && ntop == NTOP_) return launch_one<RT_, MODE_, NTOP_>(p, num_sms, stream);
#define MTVQ_MODES(RT_)             \
    MTVQ_CASE(RT_, MODE_ARGMAX, 5)  \
    MTVQ_CASE(RT_, MODE_COLLECT, 5) \
    MTVQ_CASE(RT_, MODE_DUMP, 5)    \
    MTVQ_CASE(RT_, MODE_THRESH, 5)  \
    MTVQ_CASE(RT_, MODE_THRESH, 8)
    MTVQ_MODES(1)
    MTVQ_MODES(2)
    MTVQ_MODES(4)
#undef MTVQ_CASE
#undef MTVQ_MODES
    return cudaErrorInvalidValue;
}

int coarse_smem_bytes() { return kSmemBytes; }

}  // namespace mtvq
