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
//   epilogue      each thread owns one row x 64 columns of a 128x128 accumulator tile: it pulls its two
//     32-column granules out of TMEM with two tcgen05.ld issued back to back, hands the accumulator slot
//     back to the MMA issuer at once, and only then reduces the 64 scores (FMNMX3 trees) into per-row
//     running state kept in registers.
//
// The coarse score is only used to SELECT candidates: |s - exact| <= eps (rigorous bound computed by
// the codebook prep kernel), so every code within margin = 2 eps of the best coarse score is
// re-scored in exact fp32 reference order by resolve kernels (assign_exact.cu).  Modes:
//   ARGMAX   best column + "is anything else within margin" flag           (hard path, pass 1)
//   THRESH   per-row lower bound on the top-k-th score (k-th largest granule max)   (soft, pass A)
//   COLLECT  append every column with s >= best_val[row] - margin          (near-tie rows / soft pass B)
//   DUMP     raw scores to global memory (tests only)
//
// Why the slot is released early (DESIGN.md "slot loop"): tensor memory holds only four 128x128 fp32
// accumulators, so tensor-pipe utilisation = 4 * T_mma / (latency of one MMA -> epilogue -> MMA round trip
// of a slot).  Everything on that round trip is kept minimal; reductions happen after the release.
//
// Roles (896 threads = 7 warpgroups, 1 CTA per SM, persistent over row blocks of RT x 128 rows):
//   warps  0-15 epilogue: warp w -> TMEM lane quarter w&3, column half (w>>2)&1, slot pair w>>3 (slots p, p+2)
//   warps 16-19 A loader: rows -> canonical l2norm -> fp16 -> shared-memory operand planes
//   warp  20    B producer: cp.async.bulk of pre-laid-out codebook tiles (L2 resident) into a ring
//   warps 24-27 MMA issuers, one elected thread each, issuer j feeds accumulator slot j; warp 24 owns TMEM
//   (warps 21-23 only pad the producer's warpgroup so that setmaxnreg can rebalance registers:
//    epilogue 96, loaders 56, producer 24, issuers 40 per thread)
//
// Shared-memory operand layout (both A and B): K-major, no swizzle; "plane" j holds, for every row
// of the tile, the 8 fp16 elements [8j, 8j+8) as one 16-byte unit => LBO = 2048 B (next plane),
// SBO = 128 B (next 8-row core matrix).  One UMMA K-step (16 elements) = 2 planes = 4096 B.
#include "assign_tc.h"
#include "common.cuh"

#include <stdlib.h>

namespace mtvq {

namespace {

constexpr int kEpiWarps   = 16;
constexpr int kLoadWarp0  = 16;
constexpr int kLoadWarps  = 4;
constexpr int kProdWarp   = 20;
constexpr int kMmaWarp    = 24;         // 24..27: one MMA issuer per accumulator slot
constexpr int kThreads    = 28 * 32;

constexpr int kPlaneBytes = kTileRows * 16;          // 2048
constexpr int kKblkBytes  = 8 * kPlaneBytes;         // 16384: 64 fp16 per row
constexpr int kBiasBytes  = 2 * kPlaneBytes;         // 4096
constexpr int kBiasStages = 3;
constexpr int kMinSplitChunks = 2;                   // COLLECT: fewest 128-code chunks per (row block, range) item
constexpr int kItemLoadCost   = 6;                   // COLLECT: an item's A-tile load + pipeline fill, in chunk-times
constexpr int kRecords    = 1024;                    // records per CTA in the TOPK global scratch (fixed layout)

// Dynamic shared memory map.  It depends on the row tiles per block: with RT = 4 (D <= 64) 512 epilogue threads keep two
// state sets each (1024 records); with RT = 1 (D > 128) one -- the 40 KB that frees go to two more codebook stages.
// The codebook ring is what hides the L2 latency of the stream: a D = 256 chunk fills four stages, so with a 4-stage
// ring the tensor pipe waits for every reload (measured 0.60 of peak, 0.67 with six stages).
// RT = 2 (D = 128) keeps the 4-stage map: six stages measured 2-12 % SLOWER there.
template <int RT>
struct SmemMap {
    static constexpr int kStages   = (RT == 1) ? 6 : 4;                  // codebook K-block stages of 16 KB
    static constexpr int kRec      = (RT == 1) ? 512 : 1024;             // epilogue threads x state sets
    static constexpr int kOffA     = 0;                                  // RT tiles, <= 64 KB in total
    static constexpr int kOffOnes  = 65536;
    static constexpr int kOffBData = kOffOnes + kBiasBytes;
    static constexpr int kOffBBias = kOffBData + kStages * kKblkBytes;
    static constexpr int kOffSlots = kOffBBias + kBiasStages * kBiasBytes;   // 4 chunks x kRec records x 16 B (fp16 granules)
    static constexpr int kOffState = kOffSlots + 4 * kRec * 16;              // kRec records x 16 B
    static constexpr int kOffBars  = kOffState + kRec * 16;
    static constexpr int kNumBars  = 2 * kStages + 2 * kBiasStages + 4 + 4 + 4 + 4;
    static constexpr int kOffTmem  = kOffBars + kNumBars * 8;
    static constexpr int kSmemBytes = kOffTmem + 16;
    static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA can opt in to");
};

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

template <int RT>
__device__ __forceinline__ Bars carve_bars(unsigned char* smem) {
    uint64_t* b = reinterpret_cast<uint64_t*>(smem + SmemMap<RT>::kOffBars);
    Bars r;
    r.bd_full = b;                 b += SmemMap<RT>::kStages;
    r.bd_empty = b;                b += SmemMap<RT>::kStages;
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

// record = the 32 scores of a row's best granule, kept as fp16 (64 B) in chunk-major order so that a warp's
// 16-byte stores are contiguous (conflict free).  fp16 is relative, the final scan widens its threshold.
template <int REC>
__device__ __forceinline__ void record_store(unsigned char* slots, int rec, const float (&v)[32]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        __half2 h0 = __floats2half2_rn(v[8 * c + 0], v[8 * c + 1]);
        __half2 h1 = __floats2half2_rn(v[8 * c + 2], v[8 * c + 3]);
        __half2 h2 = __floats2half2_rn(v[8 * c + 4], v[8 * c + 5]);
        __half2 h3 = __floats2half2_rn(v[8 * c + 6], v[8 * c + 7]);
        uint4 q;
        q.x = *reinterpret_cast<uint32_t*>(&h0); q.y = *reinterpret_cast<uint32_t*>(&h1);
        q.z = *reinterpret_cast<uint32_t*>(&h2); q.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(slots + (c * REC + rec) * 16) = q;
    }
}

// MODE_TOPK keeps records in a per-CTA global scratch (L2 resident): 32-byte unit ((slot * 2 + c2) * kRecords + rec),
// written with two 256-bit stores (STG.256 exists on sm_100a; each store instruction in the slot loop costs ~0.1 ms of a
// 4 ms pass, see profiles/r01_soft_path.md)
__device__ __forceinline__ void record_store_g(uint4* dst, const float (&v)[32]) {
#pragma unroll
    for (int c2 = 0; c2 < 2; ++c2) {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            __half2 h = __floats2half2_rn(v[16 * c2 + 2 * i], v[16 * c2 + 2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        asm volatile("st.global.cg.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 2 * c2 * kRecords), "r"(w[0]),
                     "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                     : "memory");
    }
}

constexpr int kTopSlots = 8;    // granules tracked per thread in MODE_TOPK (slot id rides in the low 3 mantissa bits)

__device__ __forceinline__ float select_kth(const float (&t)[kTopSlots], int kidx) {
    float r = t[0];
#pragma unroll
    for (int i = 1; i < kTopSlots; ++i) r = (i <= kidx) ? t[i] : r;
    return r;
}

// per-row running state of one row tile
template <int NLIST>
struct RowState {
    float m1, m2, m3, thr;  // ARGMAX: best granule max (its scores are in the record), second-best granule max,
                            //         max over every other granule seen, m1 - margin
    int   best_chunk, chunk2;   //     granule ids of m1 and m2
    float top[NLIST];       // THRESH / TOPK: largest granule maxima, descending
    float T;                // COLLECT: threshold
    int64_t pos, row;
    bool  valid;
};

template <int MODE, int NTOP, int NLIST>
__device__ __forceinline__ void consume(const float (&v)[32], int chunk, RowState<NLIST>& st, float margin,
                                        const CoarseParams& p, unsigned char* slots, int rec) {
    if (MODE == MODE_DUMP) {
        if (st.valid && st.pos < p.dump_rows) {
            float4* dst = reinterpret_cast<float4*>(p.dump + st.pos * p.kpad + chunk * kChunk);
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
        return;
    }
    const float cmax = max32(v);
    if (MODE == MODE_THRESH) {
        if constexpr (NLIST == NTOP) topn_insert<NTOP>(st.top, cmax);
    } else if (MODE == MODE_COLLECT) {
        const bool trig = cmax >= st.T;
        if (__any_sync(0xffffffffu, trig)) {
            if (trig) {
                // static indices only: a runtime index would push v[] into local memory
                uint32_t hits = 0;
#pragma unroll
                for (int i = 0; i < 32; ++i) hits |= (v[i] >= st.T) ? (1u << i) : 0u;
                while (hits) {
                    const int i = __ffs(hits) - 1;
                    hits &= hits - 1;
                    const int at = atomicAdd(&p.cand_cnt[st.row], 1);         // candidate lists are indexed by row
                    if (at < p.cand_cap) p.cand[st.row * p.cand_cap + at] = chunk * kChunk + i;
                }
            }
        }
    }
}

// ARGMAX on a thread's whole 64-column half-tile: one trigger test per 64 scores.  `va` holds granule `chunk`,
// `vb` granule `chunk + 1`.  Per thread the best granule is kept by value (the record), the second-best by id
// (chunk2; listed wholesale for the exact re-scoring if it ends up within the margin) and everything else only as
// a maximum (m3; if THAT ends up within the margin the row falls back to a COLLECT pass).
template <int NTOP, int REC>
__device__ __forceinline__ void consume_argmax64(const float (&va)[32], const float (&vb)[32], int chunk,
                                                 RowState<NTOP>& st, float margin, unsigned char* slots, int rec,
                                                 float* pub, const float* partner) {
    const float ma = max32(va), mb = max32(vb);
    const float cmax = fmaxf(ma, mb);
    const bool trig = cmax >= st.thr;
    if (__any_sync(0xffffffffu, trig)) {
        if (trig) {
            const bool b_wins = mb > ma;
            const float lo = fminf(ma, mb);
            // the ROW's best so far is at least what the thread owning the other column half has published (its value
            // only grows, so a stale read is still a bound): a granule below that minus the margin cannot matter
            const float other = ld_relaxed_smem(partner);
            if (cmax > st.m1) {          // new record
                const float nthr = cmax - margin;
                // Only granules within the margin of the CURRENT best can matter at the end (the best only grows), so
                // second place is tracked only in the rare case that the old record or the losing granule is that close;
                // otherwise m2 / m3 keep stale values below nthr, which the final test ignores.
                if (fmaxf(st.m1, lo) >= nthr) {
                    const int lo_c = b_wins ? chunk : chunk + 1;
                    if (lo > st.m1) { st.m3 = fmaxf(fmaxf(st.m3, st.m2), st.m1); st.m2 = lo;    st.chunk2 = lo_c; }
                    else            { st.m3 = fmaxf(fmaxf(st.m3, st.m2), lo);    st.m2 = st.m1; st.chunk2 = st.best_chunk; }
                }
                st.m1 = cmax;
                st_relaxed_smem(pub, cmax);
                st.thr = fmaxf(nthr, other - margin);
                if (b_wins) { st.best_chunk = chunk + 1; record_store<REC>(slots, rec, vb); }
                else        { st.best_chunk = chunk;     record_store<REC>(slots, rec, va); }
            } else {                      // rare: granule(s) within the margin of the record (or of the other half's best)
                if (cmax > st.m2) { st.m3 = fmaxf(fmaxf(st.m3, st.m2), lo); st.m2 = cmax; st.chunk2 = b_wins ? chunk + 1 : chunk; }
                else              { st.m3 = fmaxf(st.m3, cmax); }
                st.thr = fmaxf(st.thr, other - margin);
            }
        }
    }
}

// TOPK (single-pass soft top-k): a thread keeps its kTopSlots best granules of the row block -- the granule maximum
// (sorted list `top`, slot id packed into the low 3 mantissa bits: relative error 2^-20, covered by the margin), the
// granule id (shared memory) and the 32 scores as an fp16 record (global scratch).  A granule can only matter if its
// maximum is within the margin of the row's final k-th largest granule maximum, and that is at least the thread's own
// k-th best so far: the trigger `cmax >= thr`, thr = top[k-1] - margin, fires ~k/i of the time at the i-th granule.
template <int KSTATIC, int REC>
__device__ __forceinline__ void consume_topk(const float (&v)[32], int gran, float (&top)[kTopSlots], float& thr, int kidx,
                                             float margin, uint4* rec_base, uint16_t* gid, float4* pub, const float4* partner) {
    const float cmax = max32(v);
    const bool trig = cmax >= thr;
    if (__any_sync(0xffffffffu, trig)) {
        if (trig) {
            // thr >= top[7] always, so the new entry never falls off itself: it evicts top[7] and takes over its slot
            // (an evicted granule that still matters shows up as top[7] within the margin at the block end: overflow)
            const int slot = __float_as_int(top[kTopSlots - 1]) & 7;
            const float x = __int_as_float((__float_as_int(cmax) & ~7) | slot);
            record_store_g(rec_base + slot * (4 * kRecords), v);
            gid[slot * REC] = (uint16_t)gran;
            topn_insert<kTopSlots>(top, x);
            // Lower bound on the ROW's k-th largest granule maximum: the thread's own k-th, or what the thread that
            // owns the other column half has published (its values only grow, so a stale read is still a bound);
            // for k = 5 also the 3 + 2 splits.  Halves the trigger rate of the second half of a row block.
            float kth;
            const float4 o = ld_relaxed_smem4(partner);
            if (KSTATIC == 5) {
                kth = fmaxf(fmaxf(top[4], o.z), fmaxf(fminf(top[2], o.x), fminf(top[1], o.y)));
                st_relaxed_smem4(pub, make_float4(top[1], top[2], top[4], 0.f));
            } else {
                kth = fmaxf(select_kth(top, kidx), o.z);
                st_relaxed_smem4(pub, make_float4(0.f, 0.f, select_kth(top, kidx), 0.f));
            }
            thr = fmaxf(kth - margin, top[kTopSlots - 1]);
        }
    }
}

// scan_record on the global scratch (read through L2: other threads of the CTA wrote it)
__device__ __forceinline__ int scan_record_g(const uint4* src, float ts, int col0, int* dst, int n0, int cap) {
    uint4 q[4];     // 16-byte unit cc of the record: 32-byte unit cc / 2, half cc % 2
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) q[cc] = __ldcg(src + 2 * (cc >> 1) * kRecords + (cc & 1));
    int n = 0;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
        const uint32_t wds[4] = {q[cc].x, q[cc].y, q[cc].z, q[cc].w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&wds[t]));
            if (f.x >= ts) { if (n0 + n < cap) dst[n0 + n] = col0 + cc * 8 + 2 * t;     ++n; }
            if (f.y >= ts) { if (n0 + n < cap) dst[n0 + n] = col0 + cc * 8 + 2 * t + 1; ++n; }
        }
    }
    return n;
}

// scan one fp16 record for columns >= ts; optionally write them to dst[n...] (capacity cap)
template <int REC>
__device__ __forceinline__ int scan_record(const unsigned char* slots, int rec, float ts, int col0, int* last_col,
                                           int* dst, int n0, int cap) {
    int n = 0;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
        const uint4 q = *reinterpret_cast<const uint4*>(slots + (cc * REC + rec) * 16);
        const uint32_t wds[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&wds[t]));
            if (f.x >= ts) { const int col = col0 + cc * 8 + 2 * t;     *last_col = col; if (dst && n0 + n < cap) dst[n0 + n] = col; ++n; }
            if (f.y >= ts) { const int col = col0 + cc * 8 + 2 * t + 1; *last_col = col; if (dst && n0 + n < cap) dst[n0 + n] = col; ++n; }
        }
    }
    return n;
}

}  // namespace

template <int RT, int MODE, int NTOP, bool DBG>
__global__ void __launch_bounds__(kThreads, 1) coarse_kernel(const CoarseParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    using SM = SmemMap<RT>;
    constexpr int kDataStages = SM::kStages;
    constexpr int kOffA = SM::kOffA, kOffOnes = SM::kOffOnes, kOffBData = SM::kOffBData, kOffBBias = SM::kOffBBias;
    constexpr int kOffSlots = SM::kOffSlots, kOffState = SM::kOffState, kOffTmem = SM::kOffTmem;
    // thread id through a volatile read: ptxas otherwise re-derives warp/lane (S2R + shifts) inside the hot loops
    uint32_t tid_x;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid_x));
    const int warp = (int)(tid_x >> 5);
    const int lane = (int)(tid_x & 31);

    const int d         = p.d;
    const int nplanes   = d >> 3;
    const int nkblk     = (nplanes + 7) >> 3;
    const int a_tile_b  = nplanes * kPlaneBytes;
    const int tile_b    = (nplanes + 2) * kPlaneBytes;   // one codebook tile in the global image
    const int nchunks   = p.kpad / kTileCodes;
    const int64_t nrows = p.n_rows_dev ? (int64_t)(*p.n_rows_dev) : p.n;
    if (nrows <= 0) return;                               // (whole grid: a device-side row list is usually empty)
    const int rows_blk  = RT * kTileRows;
    const int nblocks   = (int)((nrows + rows_blk - 1) / rows_blk);
    // the 64 KB A area holds nreg = 4, 2 or 1 row tiles (D <= 64, <= 128, <= 256); they form a ring over
    // (block, row tile) so that with RT < nreg the loader runs ahead of the MMAs by whole blocks
    const uint32_t nreg = (uint32_t)(65536 / a_tile_b) >= 4u ? 4u : ((uint32_t)(65536 / a_tile_b) >= 2u ? 2u : 1u);

    // Work items.  ARGMAX / THRESH / DUMP reduce over a row's whole codebook, so an item is a row block.  COLLECT
    // compares every column with a per-row threshold known up front, so columns are independent too: when the
    // (device-side) row count leaves SMs idle -- the near-tie rows of the hard path are a few percent of N -- the
    // codebook is split into `ksplit` chunk ranges and an item is (row block, chunk range).
    int ksplit = 1;
    if (MODE == MODE_COLLECT) {
        // cost of a CTA's share in chunk-times: rounds x (chunks per item + the A-tile load of an item)
        const int grid = (int)gridDim.x;
        int best_cost = ((nblocks + grid - 1) / grid) * (nchunks + kItemLoadCost);
        for (int ks = 2; ks * kMinSplitChunks <= nchunks && ks <= 64; ++ks) {
            const int cps_ = (nchunks + ks - 1) / ks;
            if ((ks - 1) * cps_ >= nchunks) continue;            // would leave an empty range
            const int cost = ((nblocks * ks + grid - 1) / grid) * (cps_ + kItemLoadCost);
            if (cost < best_cost) { best_cost = cost; ksplit = ks; }
        }
    }
    const int cps    = (nchunks + ksplit - 1) / ksplit;          // chunks per item (the last range may be shorter)
    const int nitems = nblocks * ksplit;

    // bias K-step (-||w||^2/2): needed unless every code has the same squared norm (then it is a constant
    // shift of all scores); padding codes rely on it to lose, so it stays on when K is not a tile multiple
    const bool use_bias = !(DBG && (p.debug_flags & 2)) &&
                          (MODE == MODE_DUMP || p.k != p.kpad || !(p.meta[META_SPREAD] <= kBiasSkipTol));

    Bars bars = carve_bars<RT>(smem);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kOffTmem);

    // ---- one-time setup ------------------------------------------------------------------------
    if (tid_x == 0) {
        // every codebook stage is released by all four issuers (owners via tcgen05.commit, the others arrive)
        for (int i = 0; i < kDataStages; ++i) { mbar_init(&bars.bd_full[i], 1); mbar_init(&bars.bd_empty[i], 4); }
        for (int i = 0; i < kBiasStages; ++i) { mbar_init(&bars.bb_full[i], 1); mbar_init(&bars.bb_empty[i], 4); }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&bars.a_full[i], kLoadWarps);  // one elected lane per loader warp
            mbar_init(&bars.a_empty[i], 4 / RT);     // one commit from each issuer that reads the tile
            mbar_init(&bars.acc_full[i], 1);
            mbar_init(&bars.acc_empty[i], 8);        // one elected lane per epilogue warp serving the slot
        }
        fence_mbar_init();
    }
    if (warp >= kLoadWarp0 && warp < kLoadWarp0 + kLoadWarps) {
        // constant A-side planes of the bias K-step: row = (1, 1, 1, 0, 0, 0, 0, 0 | 0 x 8)
        const int r = (int)tid_x - kLoadWarp0 * 32;
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

    // register rebalancing between warpgroups (all 4 warps of a warpgroup execute the same instruction)
    if (warp < kEpiWarps) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
    } else if (warp < kLoadWarp0 + kLoadWarps) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    } else if (warp < kMmaWarp) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    }

    // ============================================================================================
    if (warp < kEpiWarps) {
        // ---------------------------------- EPILOGUE --------------------------------------------
        constexpr int NS = (RT == 4) ? 2 : 1;       // row tiles (state sets) per thread
        constexpr int NLIST = (MODE == MODE_TOPK) ? kTopSlots : NTOP;   // TOPK: NTOP is the static k (5) or 8 = runtime k
        const int quarter = warp & 3;               // TMEM lane quarter
        const int half    = (warp >> 2) & 1;        // columns [64*half, 64*half + 64) of every tile
        const int pair    = warp >> 3;              // serves accumulator slots pair and pair + 2
        const int rq      = quarter * 32 + lane;    // row within the tile == TMEM lane
        const float margin = p.meta[META_MARGIN];
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 64);
        unsigned char* slots = smem + kOffSlots;
        float4* state = reinterpret_cast<float4*>(smem + kOffState);
        int rec0 = warp * 32 + lane;                // record of state set ss = ss * 512 + rec0
        asm volatile("" : "+r"(rec0));
        uint32_t uses[2] = {0, 0};
        const bool tracing = DBG && (p.trace != nullptr) && blockIdx.x == 0 && warp == 0 && lane == 0;
        // everything the per-tile critical path needs, computed once
        uint32_t full_addr[2]  = {smem_u32(&bars.acc_full[pair]), smem_u32(&bars.acc_full[pair + 2])};
        uint32_t empty_addr[2] = {smem_u32(&bars.acc_empty[pair]), smem_u32(&bars.acc_empty[pair + 2])};
        uint32_t tld[2] = {t_lane + (uint32_t)(pair * kTileCodes), t_lane + (uint32_t)((pair + 2) * kTileCodes)};
        uint32_t is_lane0 = (lane == 0) ? 1u : 0u;
        int half2 = 2 * half;                       // first granule of the thread's column half within a chunk
        asm volatile("" : "+r"(half2));
        // make the values opaque so ptxas keeps them in registers instead of re-deriving them from the thread id
        asm volatile("" : "+r"(full_addr[0]), "+r"(full_addr[1]), "+r"(empty_addr[0]), "+r"(empty_addr[1]),
                          "+r"(tld[0]), "+r"(tld[1]), "+r"(is_lane0));

        // TOPK: margin widened by the slot-id packing error (2^-20 relative on a score bounded by |w| + |w|^2 / 2)
        const float nwmax = p.meta[META_MAXNORM];
        const float margin_k = margin + 4.0e-6f * (1.0f + nwmax + 0.5f * nwmax * nwmax);
        const int kidx = min(max(p.topk, 1), kTopSlots) - 1;
        uint4* topk_rec = static_cast<uint4*>(p.topk_scratch) + (size_t)blockIdx.x * (kTopkScratchPerCta / 16) + 2 * rec0;
        uint16_t* topk_gid = reinterpret_cast<uint16_t*>(smem + kOffSlots) + rec0;    // [slot][1024 records]
        if (MODE == MODE_TOPK || MODE == MODE_ARGMAX) {      // published bounds (kOffState): nothing known yet
#pragma unroll
            for (int ss = 0; ss < NS; ++ss)
                state[ss * 512 + rec0] = (MODE == MODE_TOPK) ? make_float4(-3.0e38f, -3.0e38f, -3.0e38f, 0.f)
                                                             : make_float4(-INFINITY, 0.f, 0.f, 0.f);
            named_bar_sync((RT == 1) ? 1 : 1 + pair, (RT == 1) ? 512 : 256);
        }
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int blk = (MODE == MODE_COLLECT) ? item / ksplit : item;
            const int cbeg = (MODE == MODE_COLLECT) ? (item - blk * ksplit) * cps : 0;
            const int cend = min(nchunks, cbeg + cps);
            RowState<NLIST> st[NS];
            // which row a state set belongs to: needed inside the chunk loop only by COLLECT / DUMP; ARGMAX / THRESH
            // work it out after the loop so that it does not hold registers through it
            auto locate = [&](int ss) {
                const int g = (RT == 4) ? pair + 2 * ss : (RT == 2 ? pair : 0);
                st[ss].pos = (int64_t)blk * rows_blk + g * kTileRows + rq;
                st[ss].valid = st[ss].pos < nrows;
                st[ss].row = st[ss].valid ? (p.row_list ? (int64_t)p.row_list[st[ss].pos] : st[ss].pos) : 0;
            };
            constexpr bool kLocateEarly = (MODE == MODE_COLLECT || MODE == MODE_DUMP);
#pragma unroll
            for (int ss = 0; ss < NS; ++ss) {
                if (kLocateEarly) locate(ss);
                st[ss].m1 = st[ss].m2 = st[ss].m3 = st[ss].thr = -INFINITY;
                st[ss].best_chunk = st[ss].chunk2 = 0;
#pragma unroll
                for (int i = 0; i < NLIST; ++i)  // TOPK: finite sentinels carrying the slot ids 0..7 in their low bits
                    st[ss].top[i] = (MODE == MODE_TOPK) ? __int_as_float((int)((0xFF61B1E0u & ~7u) | (unsigned)i)) : -INFINITY;
                st[ss].T = INFINITY;
                if (MODE == MODE_COLLECT && st[ss].valid) st[ss].T = p.best_val[st[ss].row] - margin;
            }

            for (int c = cbeg; c < cend; ++c) {
#pragma unroll
                for (int si = 0; si < 2; ++si) {
                    const int s = pair + 2 * si;                 // accumulator slot
                    if (((s - c * RT) & 3) >= RT) continue;      // no tile of this chunk lives in slot s
                    constexpr int kZero = 0;
                    const int ss = (RT == 4) ? si : kZero;
                    mbar_wait_addr(full_addr[si], uses[si] & 1);
                    if (DBG && tracing && si == 0 && uses[0] < 512) p.trace[uses[0] * 8 + 2] = clock64();
                    ++uses[si];
                    tc_fence_after();
                    float va[32], vb[32];
                    if (!(DBG && (p.debug_flags & 1))) {
                        tmem_ld32x2_wait(tld[si], va, vb);
                    }
                    // both granules are in registers: hand the accumulator back before reducing them
                    tc_fence_before();
                    __syncwarp();
                    if (is_lane0) mbar_arrive_relaxed_addr(empty_addr[si]);
                    if (DBG && tracing && si == 0 && uses[0] <= 512) p.trace[(uses[0] - 1) * 8 + 3] = clock64();
                    if (DBG && (p.debug_flags & 1)) continue;    // timing experiment: no epilogue work
                    if (MODE == MODE_ARGMAX) {
                        consume_argmax64<NLIST, SM::kRec>(va, vb, c * 4 + half2, st[ss], margin, slots, ss * 512 + rec0,
                                                &state[ss * 512 + rec0].x, &state[ss * 512 + (rec0 ^ 128)].x);
                    } else if (MODE == MODE_TOPK) {
                        if constexpr (NLIST == kTopSlots) {
                            constexpr int KS = (NTOP == 5) ? 5 : 0;
                            consume_topk<KS, SM::kRec>(va, c * 4 + half2, st[ss].top, st[ss].thr, kidx, margin_k,
                                             topk_rec + ss * 1024, topk_gid + ss * 512, state + ss * 512 + rec0,
                                             state + ss * 512 + (rec0 ^ 128));
                            consume_topk<KS, SM::kRec>(vb, c * 4 + half2 + 1, st[ss].top, st[ss].thr, kidx, margin_k,
                                             topk_rec + ss * 1024, topk_gid + ss * 512, state + ss * 512 + rec0,
                                             state + ss * 512 + (rec0 ^ 128));
                        }
                    } else {
                        consume<MODE, NTOP, NLIST>(va, c * 4 + half2, st[ss], margin, p, slots, ss * 512 + rec0);
                        consume<MODE, NTOP, NLIST>(vb, c * 4 + half2 + 1, st[ss], margin, p, slots, ss * 512 + rec0);
                    }
                    if (DBG && tracing && si == 0 && uses[0] <= 512) p.trace[(uses[0] - 1) * 8 + 4] = clock64();
                }
            }

            if (!kLocateEarly) {
#pragma unroll
                for (int ss = 0; ss < NS; ++ss) locate(ss);
            }
            // ---- end of row block: merge the threads that shared each row, emit ----------------------
            // a row's columns were split over 2 threads (column halves); with RT == 1 also over the 2 pairs
            constexpr int NP = (RT == 1) ? 2 : 1;
            const int bar_id = (RT == 1) ? 1 : 1 + pair;
            const int bar_n  = (RT == 1) ? 512 : 256;
            const bool leader = (half == 0) && (RT != 1 || pair == 0);
            if (MODE == MODE_ARGMAX) {
#pragma unroll
                for (int ss = 0; ss < NS; ++ss)
                    state[ss * 512 + rec0] = make_float4(st[ss].m1, st[ss].m2,
                                                         __int_as_float(st[ss].best_chunk | (st[ss].chunk2 << 16)), st[ss].m3);
                named_bar_sync(bar_id, bar_n);
                if (leader) {
#pragma unroll
                    for (int ss = 0; ss < NS; ++ss) {
                        float M = -INFINITY, m3all = -INFINITY;
#pragma unroll
                        for (int pp = 0; pp < NP; ++pp)
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                const int w2 = quarter + 4 * hh + 8 * ((RT == 1) ? pp : pair);
                                const float4 sv = state[ss * 512 + w2 * 32 + lane];
                                M = fmaxf(M, sv.x);
                                m3all = fmaxf(m3all, sv.w);
                            }
                        const float thrF = M - margin;
                        // records are fp16: accept anything that could be >= thrF before rounding
                        const float thrS = thrF - (fmaxf(fabsf(M), fabsf(thrF)) * 4.8829e-4f + 1e-7f);
                        int cnt = 0, col = -1, nsec = 0;
#pragma unroll
                        for (int pp = 0; pp < NP; ++pp)
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                const int w2 = quarter + 4 * hh + 8 * ((RT == 1) ? pp : pair);
                                const int rec = ss * 512 + w2 * 32 + lane;
                                const float4 sv = state[rec];
                                if (sv.x >= thrF) cnt += scan_record<SM::kRec>(slots, rec, thrS, (__float_as_int(sv.z) & 0xffff) * kChunk, &col, nullptr, 0, 0);
                                if (sv.y >= thrF) ++nsec;
                            }
                        // near-tie row: more than one column (or an unrecorded granule) within the margin of the best
                        // score.  Its candidates -- recorded columns one by one, second-best granules wholesale as
                        // -(granule + 1) -- go to the row's list for the exact fp32 re-scoring; if a granule known only
                        // through m3 is within the margin too (or the list would overflow) a COLLECT pass finds them.
                        const bool flagged = st[ss].valid && ((cnt != 1) || (nsec > 0) || (m3all >= thrF) || (col >= p.k));
                        const bool over = flagged && ((m3all >= thrF) || (cnt + nsec > p.cand_cap) || !(M == M));
                        if (flagged && !over) {
                            int* dst = p.cand + st[ss].row * p.cand_cap;
                            int nw = 0;
#pragma unroll
                            for (int pp = 0; pp < NP; ++pp)
#pragma unroll
                                for (int hh = 0; hh < 2; ++hh) {
                                    const int w2 = quarter + 4 * hh + 8 * ((RT == 1) ? pp : pair);
                                    const int rec = ss * 512 + w2 * 32 + lane;
                                    const float4 sv = state[rec];
                                    int dummy;
                                    if (sv.x >= thrF) nw += scan_record<SM::kRec>(slots, rec, thrS, (__float_as_int(sv.z) & 0xffff) * kChunk, &dummy, dst, nw, p.cand_cap);
                                    if (sv.y >= thrF) { dst[nw] = -((__float_as_int(sv.z) >> 16) + 1); ++nw; }
                                }
                            p.cand_cnt[st[ss].row] = nw;
                        }
                        if (over) p.best_val[st[ss].row] = M;      // COLLECT threshold; the row's list stays empty until then
                        // decided rows leave the kernel as final int64 indices + histogram counts (one RED per
                        // distinct code in the warp), spread over the whole tensor pass instead of a row kernel
                        const bool decided = st[ss].valid && !flagged;
                        const unsigned dall = __ballot_sync(0xffffffffu, decided);
                        if (decided) {
                            p.idx_out[st[ss].row] = (int64_t)col;
                            const unsigned same = __match_any_sync(dall, col);
                            if (lane == __ffs(same) - 1) atomicAdd(&p.bins[col], __popc(same));
                        }
                        const unsigned ball = __ballot_sync(0xffffffffu, flagged);
                        if (ball) {
                            int base = 0;
                            if (lane == 0) base = atomicAdd(&p.counters[CNT_FLAGGED], __popc(ball));
                            base = __shfl_sync(0xffffffffu, base, 0);
                            if (flagged) p.flag_rows[base + __popc(ball & ((1u << lane) - 1))] = (int)st[ss].row;
                        }
                        const unsigned oall = __ballot_sync(0xffffffffu, over);
                        if (oall) {
                            int base = 0;
                            if (lane == 0) base = atomicAdd(&p.counters[CNT_OVER], __popc(oall));
                            base = __shfl_sync(0xffffffffu, base, 0);
                            if (over) p.over_rows[base + __popc(oall & ((1u << lane) - 1))] = (int)st[ss].row;
                        }
                        // the next row block starts with no published best (the leader has read every entry of the row)
#pragma unroll
                        for (int pp = 0; pp < NP; ++pp)
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                const int w2 = quarter + 4 * hh + 8 * ((RT == 1) ? pp : pair);
                                state[ss * 512 + w2 * 32 + lane].x = -INFINITY;
                            }
                    }
                }
                named_bar_sync(bar_id, bar_n);
            } else if (MODE == MODE_TOPK) {
                if constexpr (NLIST == kTopSlots) {
                float* tl = reinterpret_cast<float*>(smem + SM::kOffSlots + 2 * kTopSlots * SM::kRec);   // [record][8], after the ids
#pragma unroll
                for (int ss = 0; ss < NS; ++ss) {
                    float4* t4 = reinterpret_cast<float4*>(tl + (ss * 512 + rec0) * 8);
                    t4[0] = make_float4(st[ss].top[0], st[ss].top[1], st[ss].top[2], st[ss].top[3]);
                    t4[1] = make_float4(st[ss].top[4], st[ss].top[5], st[ss].top[6], st[ss].top[7]);
                }
                named_bar_sync(bar_id, bar_n);
                if (leader) {
                    const uint4* rec_cta = static_cast<const uint4*>(p.topk_scratch) + (size_t)blockIdx.x * (kTopkScratchPerCta / 16);
                    const uint16_t* gid_all = reinterpret_cast<const uint16_t*>(smem + kOffSlots);
#pragma unroll
                    for (int ss = 0; ss < NS; ++ss) {
                        // k-th largest granule maximum over the threads that shared the row
#pragma unroll
                        for (int pp = 0; pp < NP; ++pp)
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                const int w2 = quarter + 4 * hh + 8 * ((RT == 1) ? pp : pair);
                                if (w2 == warp) continue;
                                const float4* o4 = reinterpret_cast<const float4*>(tl + (ss * 512 + w2 * 32 + lane) * 8);
                                const float4 oa = o4[0], ob = o4[1];
                                topn_insert<NLIST>(st[ss].top, oa.x); topn_insert<NLIST>(st[ss].top, oa.y);
                                topn_insert<NLIST>(st[ss].top, oa.z); topn_insert<NLIST>(st[ss].top, oa.w);
                                topn_insert<NLIST>(st[ss].top, ob.x); topn_insert<NLIST>(st[ss].top, ob.y);
                                topn_insert<NLIST>(st[ss].top, ob.z); topn_insert<NLIST>(st[ss].top, ob.w);
                            }
                        const float kth = select_kth(st[ss].top, kidx);
                        const float thrG = kth - margin_k;
                        // records are fp16: accept anything that could be >= thrG before rounding
                        const float thrS = thrG - (fmaxf(fabsf(kth), fabsf(thrG)) * 4.8829e-4f + 1e-7f);
                        bool over = !(kth > -1.0e37f);           // fewer than k granules seen, or NaN scores
                        int nw = 0;
                        int* dst = p.cand + st[ss].row * p.cand_cap;
                        if (st[ss].valid) {
#pragma unroll
                            for (int pp = 0; pp < NP; ++pp)
#pragma unroll
                                for (int hh = 0; hh < 2; ++hh) {
                                    const int w2 = quarter + 4 * hh + 8 * ((RT == 1) ? pp : pair);
                                    const int rec = ss * 512 + w2 * 32 + lane;
                                    for (int i = 0; i < kTopSlots; ++i) {
                                        const float e = tl[rec * 8 + i];
                                        if (!(e >= thrG)) break;
                                        if (i == kTopSlots - 1) over = true;     // an untracked granule may qualify too
                                        const int slot = __float_as_int(e) & 7;
                                        const int gran = gid_all[slot * SM::kRec + rec];
                                        nw += scan_record_g(rec_cta + (slot * 4) * kRecords + 2 * rec, thrS, gran * kChunk, dst, nw,
                                                            p.cand_cap);
                                    }
                                }
                            if (nw > p.cand_cap || nw < p.topk) over = true;
                            if (!over) p.cand_cnt[st[ss].row] = nw;
                            // overflow rows (rare) get their list from a COLLECT pass over `over_rows`: the k-th largest
                            // TRACKED granule maximum is still a lower bound on the row's k-th largest score
                            else p.best_val[st[ss].row] = (kth > -1.0e37f) ? kth - (margin_k - margin) : -INFINITY;
                        }
                        const bool ov = st[ss].valid && over;
                        const unsigned oall = __ballot_sync(0xffffffffu, ov);
                        if (oall) {
                            int base = 0;
                            if (lane == 0) base = atomicAdd(&p.counters[CNT_OVER], __popc(oall));
                            base = __shfl_sync(0xffffffffu, base, 0);
                            if (ov) p.over_rows[base + __popc(oall & ((1u << lane) - 1))] = (int)st[ss].row;
                        }
                    }
                }
                // the next row block starts with no published bounds (reset before the barrier: the thread of the other
                // column half must not see this block's values)
#pragma unroll
                for (int ss = 0; ss < NS; ++ss) state[ss * 512 + rec0] = make_float4(-3.0e38f, -3.0e38f, -3.0e38f, 0.f);
                named_bar_sync(bar_id, bar_n);
                }
            } else if (MODE == MODE_THRESH) {
                float* tl = reinterpret_cast<float*>(smem + kOffSlots);   // [record][8]
#pragma unroll
                for (int ss = 0; ss < NS; ++ss)
#pragma unroll
                    for (int i = 0; i < NTOP; ++i) tl[(ss * 512 + rec0) * 8 + i] = st[ss].top[i];
                named_bar_sync(bar_id, bar_n);
                if (leader) {
#pragma unroll
                    for (int ss = 0; ss < NS; ++ss) {
#pragma unroll
                        for (int pp = 0; pp < NP; ++pp)
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                const int w2 = quarter + 4 * hh + 8 * ((RT == 1) ? pp : pair);
                                if (w2 == warp) continue;
#pragma unroll
                                for (int i = 0; i < NTOP; ++i)
                                    topn_insert<NLIST>(st[ss].top, tl[(ss * 512 + w2 * 32 + lane) * 8 + i]);
                            }
                        if (st[ss].valid) {
                            float kth = st[ss].top[0];
#pragma unroll
                            for (int i = 1; i < NTOP; ++i)
                                if (i < p.topk) kth = st[ss].top[i];
                            p.best_val[st[ss].row] = kth;
                        }
                    }
                }
                named_bar_sync(bar_id, bar_n);
            }
        }
    } else if (warp < kLoadWarp0 + kLoadWarps) {
        // ---------------------------------- A LOADER --------------------------------------------
        // one thread per row.  The exact reference-order norm goes to `denom` (later kernels divide by
        // it); the fp16 MMA operand itself only needs to be inside the coarse error bound, so it is
        // produced with one reciprocal and a multiply per element instead of an IEEE division.
        const int r = (int)tid_x - kLoadWarp0 * 32;
        uint32_t bi = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++bi) {
            const int blk = (MODE == MODE_COLLECT) ? item / ksplit : item;
            const bool skip = DBG && (p.debug_flags & 4) && bi > 0;   // timing experiment only
            for (int g = 0; g < RT; ++g) {
                const int64_t pos = (int64_t)blk * rows_blk + g * kTileRows + r;
                const bool valid = pos < nrows && !skip;
                const int64_t row = valid ? (p.row_list ? (int64_t)p.row_list[pos] : pos) : 0;
                const float4* zr4 = reinterpret_cast<const float4*>(p.z + row * d);
                float rden = 0.0f;
                if (valid) {
                    float den;
                    if (norm_kind(p.norm_mode) == 2) {
                        den = 1.0f;                                   // MTVQ_NORM_NONE: rows are used as given
                    } else if (norm_kind(p.norm_mode) == 0) {
                        // canon_norm8 on float4 loads: lane v accumulates x[8c+v]^2 in order of c
                        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f, a6 = 0.f, a7 = 0.f;
                        for (int j = 0; j < nplanes; ++j) {
                            const float4 x0 = __ldg(zr4 + 2 * j), x1 = __ldg(zr4 + 2 * j + 1);
                            a0 = __fadd_rn(a0, __fmul_rn(x0.x, x0.x)); a1 = __fadd_rn(a1, __fmul_rn(x0.y, x0.y));
                            a2 = __fadd_rn(a2, __fmul_rn(x0.z, x0.z)); a3 = __fadd_rn(a3, __fmul_rn(x0.w, x0.w));
                            a4 = __fadd_rn(a4, __fmul_rn(x1.x, x1.x)); a5 = __fadd_rn(a5, __fmul_rn(x1.y, x1.y));
                            a6 = __fadd_rn(a6, __fmul_rn(x1.z, x1.z)); a7 = __fadd_rn(a7, __fmul_rn(x1.w, x1.w));
                        }
                        float s = a0;
                        s = __fadd_rn(s, a1); s = __fadd_rn(s, a2); s = __fadd_rn(s, a3);
                        s = __fadd_rn(s, a4); s = __fadd_rn(s, a5); s = __fadd_rn(s, a6); s = __fadd_rn(s, a7);
                        den = __fsqrt_rn(s);
                    } else {
                        const float* zp = p.z + row * d;
                        den = canon_norm_seq([&](int i) { return __ldg(zp + i); }, d);
                    }
                    den = den < kNormEps ? kNormEps : den;
                    if (p.denom) p.denom[row] = den;
                    rden = __frcp_rn(den);
                }
                auto pack8 = [&](int j) {
                    uint4 q = make_uint4(0, 0, 0, 0);
                    if (valid) {
                        const float4 x0 = __ldg(zr4 + 2 * j), x1 = __ldg(zr4 + 2 * j + 1);
                        __half2 h0 = __floats2half2_rn(x0.x * rden, x0.y * rden);
                        __half2 h1 = __floats2half2_rn(x0.z * rden, x0.w * rden);
                        __half2 h2 = __floats2half2_rn(x1.x * rden, x1.y * rden);
                        __half2 h3 = __floats2half2_rn(x1.z * rden, x1.w * rden);
                        q.x = *reinterpret_cast<uint32_t*>(&h0);
                        q.y = *reinterpret_cast<uint32_t*>(&h1);
                        q.z = *reinterpret_cast<uint32_t*>(&h2);
                        q.w = *reinterpret_cast<uint32_t*>(&h3);
                    }
                    return q;
                };
                const uint32_t aq = bi * RT + g, areg = aq % nreg, apar = (aq / nreg) & 1;   // A region ring
                unsigned char* tile = smem + kOffA + areg * a_tile_b + r * 16;
                if (nplanes <= 8) {
                    // D <= 64: the row is converted into registers BEFORE waiting for the tile to be
                    // released, so the global-load latency overlaps the previous block's MMAs
                    uint4 q[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) q[j] = (j < nplanes) ? pack8(j) : make_uint4(0, 0, 0, 0);
                    mbar_wait_backoff(&bars.a_empty[areg], apar ^ 1, 400);
                    if (!skip) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j < nplanes) *reinterpret_cast<uint4*>(tile + j * kPlaneBytes) = q[j];
                    }
                } else {
                    mbar_wait_backoff(&bars.a_empty[areg], apar ^ 1, 400);
                    if (!skip)
                        for (int j = 0; j < nplanes; ++j) *reinterpret_cast<uint4*>(tile + j * kPlaneBytes) = pack8(j);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars.a_full[areg]);
            }
        }
    } else if (warp == kProdWarp) {
        // ---------------------------------- B PRODUCER ------------------------------------------
        if (elect_one()) {
            const unsigned char* img = static_cast<const unsigned char*>(p.b_image);
            uint32_t itB = 0;
            uint32_t st = 0, ph = 0;      // ring position and phase, advanced per K-block (the ring need not be a power of 2)
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const int cbeg = (MODE == MODE_COLLECT) ? (item % ksplit) * cps : 0;
                const int cend = min(nchunks, cbeg + cps);
                for (int c = cbeg; c < cend; ++c) {
                    const unsigned char* tile = img + (size_t)c * tile_b;
                    for (int kb = 0; kb < nkblk; ++kb) {
                        const uint32_t bytes = (uint32_t)min(8, nplanes - 8 * kb) * kPlaneBytes;
                        mbar_wait(&bars.bd_empty[st], ph ^ 1);
                        mbar_arrive_expect_tx(&bars.bd_full[st], bytes);
                        bulk_g2s(smem + kOffBData + st * kKblkBytes, tile + (size_t)kb * kKblkBytes, bytes,
                                 &bars.bd_full[st]);
                        if (++st == (uint32_t)kDataStages) { st = 0; ph ^= 1; }
                    }
                    if (use_bias) {
                        const int sb = itB % kBiasStages;
                        mbar_wait(&bars.bb_empty[sb], ((itB / kBiasStages) & 1) ^ 1);
                        mbar_arrive_expect_tx(&bars.bb_full[sb], kBiasBytes);
                        bulk_g2s(smem + kOffBBias + sb * kBiasBytes, tile + (size_t)nplanes * kPlaneBytes, kBiasBytes,
                                 &bars.bb_full[sb]);
                        ++itB;
                    }
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
            const uint64_t dA0 = umma_smem_desc(smem_u32(smem + kOffA), lbo, sbo);
            const uint64_t dOnes = umma_smem_desc(smem_u32(smem + kOffOnes), lbo, sbo);
            const uint64_t dBD0 = umma_smem_desc(smem_u32(smem + kOffBData), lbo, sbo);
            const uint64_t dBB0 = umma_smem_desc(smem_u32(smem + kOffBBias), lbo, sbo);
            const uint32_t tacc = tmem_base + (uint32_t)(j * kTileCodes);
            uint32_t uses = 0, bi = 0;
            uint32_t dst0 = 0, dph0 = 0;             // codebook ring position / phase of the next chunk's first K-block
            uint32_t chunk0 = 0;                     // chunks issued by the CTA before this block
            const bool tracing = DBG && (p.trace != nullptr) && blockIdx.x == 0 && j == 0;
            int cbeg = 0, cend = nchunks;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++bi, chunk0 += (uint32_t)(cend - cbeg)) {
                if (MODE == MODE_COLLECT) { cbeg = (item % ksplit) * cps; cend = min(nchunks, cbeg + cps); }
                // Every issuer walks EVERY chunk so that each shared barrier sees all four issuers once per
                // phase (an mbarrier parity wait is only meaningful within one phase of the barrier);
                // it issues MMAs only for the tiles it owns: c % G == j / RT.
                bool first = true;
                const uint32_t aq = bi * RT + g, areg = aq % nreg, apar = (aq / nreg) & 1;   // A region ring
                const uint64_t dA = dA0 + (uint64_t)(areg * ((uint32_t)a_tile_b >> 4));
                for (int c = cbeg; c < cend; ++c) {
                    const bool own = (c % G) == (j / RT);
                    const uint32_t gc = chunk0 + (uint32_t)(c - cbeg);    // running chunk number
                    const uint32_t sb = gc % kBiasStages;
                    // this chunk's K-blocks occupy ring positions dst0, dst0 + 1, ... (wrapping, phase flips at the wrap)
                    const uint32_t cst = dst0, cph = dph0;
                    auto next_stage = [&](uint32_t& st_, uint32_t& ph_) { if (++st_ == (uint32_t)kDataStages) { st_ = 0; ph_ ^= 1; } };
                    // operands first, accumulator last: when the epilogue hands the slot back, every MMA of
                    // the tile is issued back to back
                    {
                        uint32_t st_ = cst, ph_ = cph;
                        for (int kb = 0; kb < nkblk; ++kb) { mbar_wait(&bars.bd_full[st_], ph_); next_stage(st_, ph_); }
                        dst0 = st_; dph0 = ph_;
                    }
                    if (use_bias) mbar_wait(&bars.bb_full[sb], (gc / kBiasStages) & 1);
                    if (!own) {
                        uint32_t st_ = cst, ph_ = cph;
                        for (int kb = 0; kb < nkblk; ++kb) { mbar_arrive(&bars.bd_empty[st_]); next_stage(st_, ph_); }
                        if (use_bias) mbar_arrive(&bars.bb_empty[sb]);
                        continue;
                    }
                    if (first) { mbar_wait(&bars.a_full[areg], apar); first = false; }
                    mbar_wait(&bars.acc_empty[j], (uses & 1u) ^ 1u);
                    if (DBG && tracing && uses < 512) p.trace[uses * 8 + 0] = clock64();
                    ++uses;
                    tc_fence_after();
                    uint32_t st = cst, sph = cph;
                    for (int kb = 0; kb < nkblk; ++kb) {
                        const int nsteps = min(8, nplanes - 8 * kb) >> 1;
                        uint64_t da = dA + (uint64_t)(kb * (kKblkBytes >> 4));
                        uint64_t db = dBD0 + (uint64_t)(st * (kKblkBytes >> 4));
                        for (int ks = 0; ks < nsteps; ++ks) {
                            umma_f16(tacc, da, db, idesc, (kb | ks) != 0);
                            da += (2 * kPlaneBytes) >> 4;
                            db += (2 * kPlaneBytes) >> 4;
                        }
                        umma_commit(&bars.bd_empty[st]);
                        next_stage(st, sph);
                    }
                    if (use_bias) umma_f16(tacc, dOnes, dBB0 + (uint64_t)(sb * (kBiasBytes >> 4)), idesc, 1u);
                    umma_commit(&bars.acc_full[j]);
                    if (use_bias) umma_commit(&bars.bb_empty[sb]);
                    if (c + G >= cend) umma_commit(&bars.a_empty[areg]);   // this issuer's last tile of the block
                    if (DBG && tracing && uses <= 512) p.trace[(uses - 1) * 8 + 1] = clock64();
                }
                if (first) {
                    // fewer chunks than issuers: this issuer owned no tile, but the row tile's release
                    // barrier still expects it once per block (paced by the tile's own arrival)
                    mbar_wait(&bars.a_full[areg], apar);
                    mbar_arrive(&bars.a_empty[areg]);
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
template <int RT, int MODE, int NTOP, bool DBG>
static cudaError_t launch_one(const CoarseParams& p, int grid, cudaStream_t stream) {
    auto kern = coarse_kernel<RT, MODE, NTOP, DBG>;
    static PerDeviceOnce once;         // one per instantiation and device
    bool& configured = once.flag();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemMap<RT>::kSmemBytes);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    kern<<<grid, kThreads, SmemMap<RT>::kSmemBytes, stream>>>(p);
    ++g_kernel_launches;
    return cudaGetLastError();
}

int coarse_row_tiles(int d, int64_t nrows, int num_sms) {
    int rt = d <= 64 ? 4 : (d <= 128 ? 2 : 1);
    {   // tuning experiments only: MTVQ_ROW_TILES caps the rows per block
        static int cap = -1;
        if (cap < 0) { const char* e = getenv("MTVQ_ROW_TILES"); cap = e ? atoi(e) : 0; }
        if (cap == 1 || cap == 2) rt = rt < cap ? rt : cap;
    }
    // keep every SM busy when there are few rows: shrink the block until there are >= 2 waves
    while (rt > 1 && (nrows + rt * 128 - 1) / (rt * 128) < 2 * (int64_t)num_sms) rt >>= 1;
    return rt;
}

cudaError_t launch_coarse(const CoarseParams& p, int mode, int row_tiles, int num_sms, cudaStream_t stream) {
    // NTOP only matters in THRESH mode (size of the per-row top list): 5 covers MedTok's k = 5
    // THRESH: size of the per-row top list (5 covers MedTok's k = 5).  TOPK: 5 = k is exactly 5 (static), 8 = runtime k.
    const int ntop = mode == MODE_TOPK ? (p.topk == 5 ? 5 : 8) : ((mode == MODE_THRESH && p.topk > 5) ? 8 : 5);
    if (mode == MODE_TOPK && (num_sms > kTopkMaxCtas || p.topk_scratch == nullptr)) return cudaErrorInvalidValue;
    // timing knobs / pipeline trace live in a separate instantiation so that the production kernels carry none of it
    const bool dbg = (mode == MODE_ARGMAX) && (p.trace != nullptr || p.debug_flags != 0);
#define MTVQ_CASE(RT_, MODE_, NTOP_)                                                                   \
    if (row_tiles == RT_ && mode == MODE_ && ntop == NTOP_) {                                           \
        if (MODE_ == MODE_ARGMAX && dbg) return launch_one<RT_, MODE_ARGMAX, 5, true>(p, num_sms, stream); \
        return launch_one<RT_, MODE_, NTOP_, false>(p, num_sms, stream);                                \
    }
#define MTVQ_MODES(RT_)             \
    MTVQ_CASE(RT_, MODE_ARGMAX, 5)  \
    MTVQ_CASE(RT_, MODE_COLLECT, 5) \
    MTVQ_CASE(RT_, MODE_DUMP, 5)    \
    MTVQ_CASE(RT_, MODE_THRESH, 5)  \
    MTVQ_CASE(RT_, MODE_THRESH, 8)  \
    MTVQ_CASE(RT_, MODE_TOPK, 5)    \
    MTVQ_CASE(RT_, MODE_TOPK, 8)
    MTVQ_MODES(1)
    MTVQ_MODES(2)
    MTVQ_MODES(4)
#undef MTVQ_CASE
#undef MTVQ_MODES
    return cudaErrorInvalidValue;
}

int coarse_smem_bytes() { return SmemMap<4>::kSmemBytes > SmemMap<1>::kSmemBytes ? SmemMap<4>::kSmemBytes : SmemMap<1>::kSmemBytes; }

}  // namespace mtvq
