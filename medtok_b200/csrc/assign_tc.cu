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
template <int RT, bool REFINE = false>
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
    static constexpr int kNumBars  = 2 * kStages + 2 * kBiasStages + 4 + 4 + 4 + 4 + 8 + 4 + 4 + 4 + 4;
    static constexpr int kOffTmem  = kOffBars + kNumBars * 8;
    static constexpr int kSmemBytes = kOffTmem + 16;
    static constexpr int kOffAcc = 0, kOffStage = 0;                      // REFINE only
    static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA can opt in to");
};

// MODE_REFINE (D <= 64, one 128-code chunk per 128-row tile): ONE fp16 row tile and ONE codebook stage (consecutive tiles
// mostly share their chunk: it is reloaded only when the chunk changes), and instead of score records the fp64 per-code
// accumulators of the fused row work (128 codes x d) and THREE staging tiles into which the gathered fp32 rows land by
// cp.async (row stride 4 d + 16 bytes: the row-per-thread conversion reads are conflict free).  A staging tile is busy
// from the copy until the finishing warps are done with the rows; three of them keep three such chains interleaved.
#ifndef MTVQ_STAGE_DIM          // experiments only: staging tiles sized for a smaller D leave room for a deeper ring
#define MTVQ_STAGE_DIM 64
#endif
#ifndef MTVQ_STAGE_TILES
#define MTVQ_STAGE_TILES 3
#endif
constexpr int kStageTileBytes = kTileRows * (4 * MTVQ_STAGE_DIM + 16);   // 34816
constexpr int kStageTiles     = MTVQ_STAGE_TILES;                        // <= 4
template <>
struct SmemMap<1, true> {
    static constexpr int kStages   = 1;
    static constexpr int kRec      = 512;
    static constexpr int kOffA     = 0;                                  // 1 tile x 16 KB
    static constexpr int kOffOnes  = 16384;
    static constexpr int kOffBData = kOffOnes + kBiasBytes;
    static constexpr int kOffBBias = kOffBData + kStages * kKblkBytes;
    static constexpr int kOffSlots = kOffBBias + kBiasStages * kBiasBytes;   // 12 KB of epilogue scratch (refine_epilogue)
    static constexpr int kOffState = kOffSlots + 12288;
    static constexpr int kOffAcc   = kOffState;                          // double [128][d]
    static constexpr int kOffStage = kOffAcc + kTileCodes * 64 * 8;
    static constexpr int kOffBars  = kOffStage + kStageTiles * kStageTileBytes;
    static constexpr int kNumBars  = 2 * kStages + 2 * kBiasStages + 4 + 4 + 4 + 4 + 8 + 4 + 4 + 4 + 4;
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
    uint64_t* dec_full;  // [4]  REFINE: a tile's decisions are in the ring (deciders -> finishers)
    uint64_t* dec_empty; // [4]
    uint64_t* p1_full;   // [4]  REFINE: the ROWS half has left a tile's zf in the staging tile (-> SUMS)
    uint64_t* sort_full; // [4]  REFINE: a tile's decided rows are listed in code order (-> SUMS)
    uint64_t* landed;    // [3]  REFINE: a staging tile's cp.async copies have landed (copy warps -> convert warps)
    uint64_t* stg_empty; // [3]  REFINE: the finishers are done with a staging tile (-> loader)
};

template <class SM>
__device__ __forceinline__ Bars carve_bars(unsigned char* smem) {
    uint64_t* b = reinterpret_cast<uint64_t*>(smem + SM::kOffBars);
    Bars r;
    r.bd_full = b;                 b += SM::kStages;
    r.bd_empty = b;                b += SM::kStages;
    r.bb_full = b;                 b += kBiasStages;
    r.bb_empty = b;                b += kBiasStages;
    r.a_full = b;                  b += 4;
    r.a_empty = b;                 b += 4;
    r.acc_full = b;                b += 4;
    r.acc_empty = b;               b += 4;
    r.dec_full = b;                b += 4;
    r.dec_empty = b;               b += 4;
    r.p1_full = b;                 b += 4;
    r.sort_full = b;               b += 4;
    r.landed = b;                  b += 4;
    r.stg_empty = b;
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

// GROUP (two-pass hard route): branch-free.  The thread only tracks the three largest 64-column maxima of its column
// half (one per 128-code chunk) and the chunks of the two largest; which COLUMN wins is left to the REFINE pass.
template <int NTOP>
__device__ __forceinline__ void consume_group(const float (&va)[32], const float (&vb)[32], int c, RowState<NTOP>& st) {
    const float x = fmaxf(max32(va), max32(vb));
    const bool p1 = x > st.m1, p2 = x > st.m2;
    st.m3 = fmaxf(st.m3, fminf(st.m2, x));
    st.m2 = fmaxf(st.m2, fminf(st.m1, x));
    st.chunk2 = p1 ? st.best_chunk : (p2 ? c : st.chunk2);
    st.m1 = fmaxf(st.m1, x);
    st.best_chunk = p1 ? c : st.best_chunk;
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

// bar.sync is warp-aligned: after divergent code (a lane-0 arrive, per-lane list writes) the warp is re-converged
// explicitly before it meets a named barrier (synccheck flags the barrier otherwise)
__device__ __forceinline__ void conv_bar_sync(int id, int nthreads) {
    __syncwarp();
    named_bar_sync(id, nthreads);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------------------------------------
// REFINE loader: the rows of a tile are scattered over z (the tile is 128 entries of the chunk-sorted order), so a
// row-per-thread loader would expose a DRAM round trip per tile.  Two of the four loader warps COPY: the 16 lanes of a
// group move whole 4 d-byte rows global -> shared with cp.async (no registers held; up to three tiles in flight, a
// staging tile being refilled as soon as the epilogue is done with its rows) and signal the tile's mbarrier when their
// copies have landed.  The other two CONVERT: thread c turns rows c and c + 64 into the fp16 operand planes -- the same
// values the GROUP pass fed to the tensor cores: x * rcp(norm), norm from the entry.  The copies never wait for the
// conversion and the conversion never waits for a staging tile to be released.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <class SM, bool TRACE>
__device__ __forceinline__ void refine_loader(const CoarseParams& p, unsigned char* smem, const Bars& bars, int r, int lane,
                                              int item0, int item_end) {
    // pipeline trace (mtvq_debug_set_trace, CTA 0): per tile [0] copies issued, [1] landed, [2] operand tile published,
    // [3] deciders saw the accumulator, [4] decisions published, [5] ROWS saw them, [6] ROWS done, [7] SUMS done
    const int d = p.d, vpr = d >> 2, nplanes = d >> 3;
    const int row_stride = 4 * d + 16;
    unsigned char* stage = smem + SM::kOffStage;
    const bool fuse = p.fuse_rows != 0;
    if (r < 64) {
        // ---------------------------------- COPY ----------------------------------
        const bool tracing = TRACE && p.trace != nullptr && blockIdx.x == 0 && r == 0;
        const int g = r >> 4, q = r & 15;                // group g copies rows g, g + 4, ... (32 rows per tile)
        uint32_t bi = 0;
        int buf = 0;                                     // bi % kStageTiles
        for (int item = item0; item < item_end; ++item, ++bi, buf = (buf == kStageTiles - 1 ? 0 : buf + 1)) {
            const int2 tt = __ldg(&p.tile_tab[item]);
            const int cnt = tt.x >> 16;
            // lane q knows the rows g + 4 q and g + 4 (q + 16) of the tile
            const int ra = g + 4 * q, rb = ra + 64;
            const int ea = ra < cnt ? __ldg(&p.entries[tt.y + ra]).x : -1;
            const int eb = rb < cnt ? __ldg(&p.entries[tt.y + rb]).x : -1;
            // the staging tile held tile bi - 3: with fused rows the epilogue reads the fp32 rows there (z makes one trip
            // from HBM), so it is refilled only once the SUMS step is done with it
            if (fuse && bi >= (uint32_t)kStageTiles) mbar_wait_backoff(&bars.stg_empty[buf], (((bi / kStageTiles) & 1u) ^ 1u), 100);
            // without fused rows: once the conversion of tile bi - 3 has been published (its MMAs need tile bi - 2's turn)
            if (!fuse && bi >= (uint32_t)kStageTiles) mbar_wait_backoff(&bars.a_empty[0], ((bi - kStageTiles) & 1u), 100);
            unsigned char* st = stage + buf * kStageTileBytes;
#pragma unroll 4
            for (int i = 0; i < 32; ++i) {
                const int ex = __shfl_sync(0xffffffffu, i < 16 ? ea : eb, (lane & 16) + (i & 15));
                const int rr = g + 4 * i;
                if (ex != -1 && q < vpr)
                    cp_async16(st + rr * row_stride + q * 16, p.z + (size_t)((unsigned)ex & kEntRowMask) * d + 4 * q);
            }
            cp_async_mbar_arrive_noinc(&bars.landed[buf]);
            if (TRACE && tracing && bi < 256u) p.trace[bi * 8 + 0] = clock64();
        }
        cp_async_wait<0>();
    } else {
        // ---------------------------------- CONVERT ----------------------------------
        const int c = r - 64;
        const bool tracing = TRACE && p.trace != nullptr && blockIdx.x == 0 && c == 0;
        auto fetch = [&](int item, int4& e0, int4& e1) {
            e0 = e1 = make_int4(-1, 0, 0x3f800000, 0);
            if (item < item_end) {
                const int2 tt = __ldg(&p.tile_tab[item]);
                const int cnt = tt.x >> 16;
                if (c < cnt) e0 = __ldg(&p.entries[tt.y + c]);
                if (c + 64 < cnt) e1 = __ldg(&p.entries[tt.y + c + 64]);
            }
        };
        int4 e0, e1;
        fetch(item0, e0, e1);
        uint32_t bi = 0;
        int buf = 0;
        for (int item = item0; item < item_end; ++item, ++bi, buf = (buf == kStageTiles - 1 ? 0 : buf + 1)) {
            const int4 c0 = e0, c1 = e1;
            fetch(item + 1, e0, e1);                      // the next tile's norms arrive while this one is converted
            mbar_wait(&bars.landed[buf], (bi / kStageTiles) & 1u);
            if (TRACE && tracing && bi < 256u) p.trace[bi * 8 + 1] = clock64();
            mbar_wait_backoff(&bars.a_empty[0], (bi & 1u) ^ 1u, 100);      // the previous tile's MMAs have read the operand tile
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const int4 e = h ? c1 : c0;
                const int row = c + 64 * h;
                const bool valid = e.x != -1;
                const float rden = __frcp_rn(__int_as_float(e.z));
                const float4* src = reinterpret_cast<const float4*>(stage + buf * kStageTileBytes + row * row_stride);
                unsigned char* tile = smem + SM::kOffA + row * 16;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (j < nplanes) {
                        uint4 pk = make_uint4(0, 0, 0, 0);
                        if (valid) {
                            const float4 x0 = src[2 * j], x1 = src[2 * j + 1];
                            __half2 h0 = __floats2half2_rn(x0.x * rden, x0.y * rden);
                            __half2 h1 = __floats2half2_rn(x0.z * rden, x0.w * rden);
                            __half2 h2 = __floats2half2_rn(x1.x * rden, x1.y * rden);
                            __half2 h3 = __floats2half2_rn(x1.z * rden, x1.w * rden);
                            pk.x = *reinterpret_cast<uint32_t*>(&h0);
                            pk.y = *reinterpret_cast<uint32_t*>(&h1);
                            pk.z = *reinterpret_cast<uint32_t*>(&h2);
                            pk.w = *reinterpret_cast<uint32_t*>(&h3);
                        }
                        *reinterpret_cast<uint4*>(tile + j * kPlaneBytes) = pk;
                    }
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars.a_full[0]);
            if (TRACE && tracing && bi < 256u) p.trace[bi * 8 + 2] = clock64();
        }
    }
}

// REFINE codebook producer / MMA issuers.  Consecutive tiles of the chunk-sorted order mostly share their chunk, so the
// single codebook stage (and the bias stage) is loaded once per RUN of tiles with the same chunk and handed back by all
// four issuers with tcgen05.commit -- each may still have MMAs of the run in flight -- after the run's last tile.
template <class SM>
__device__ __forceinline__ void refine_producer(const CoarseParams& p, unsigned char* smem, const Bars& bars, int item0,
                                                int item_end, bool use_bias) {
    const int nplanes = p.d >> 3;
    const int tile_b = (nplanes + 2) * kPlaneBytes;
    const unsigned char* img = static_cast<const unsigned char*>(p.b_image);
    const uint32_t bytes = (uint32_t)nplanes * kPlaneBytes;
    int prev = -1;
    uint32_t run = 0;
    for (int item = item0; item < item_end; ++item) {
        const int chunk = __ldg(&p.tile_tab[item]).x & 0xffff;
        if (chunk == prev) continue;
        prev = chunk;
        const unsigned char* tile = img + (size_t)chunk * tile_b;
        mbar_wait(&bars.bd_empty[0], (run & 1u) ^ 1u);
        mbar_arrive_expect_tx(&bars.bd_full[0], bytes);
        bulk_g2s(smem + SM::kOffBData, tile, bytes, &bars.bd_full[0]);
        if (use_bias) {
            const uint32_t sb = run % kBiasStages;
            mbar_wait(&bars.bb_empty[sb], ((run / kBiasStages) & 1u) ^ 1u);
            mbar_arrive_expect_tx(&bars.bb_full[sb], kBiasBytes);
            bulk_g2s(smem + SM::kOffBBias + sb * kBiasBytes, tile + (size_t)nplanes * kPlaneBytes, kBiasBytes, &bars.bb_full[sb]);
        }
        ++run;
    }
}

template <class SM>
__device__ __forceinline__ void refine_issuer(const CoarseParams& p, unsigned char* smem, const Bars& bars, uint32_t tmem_base,
                                              int j, int item0, int item_end, bool use_bias) {
    const int nsteps = p.d >> 4;                     // UMMA K-steps of 16 elements (D <= 64: one K-block)
    const uint32_t idesc = umma_idesc_f16(kTileRows, kTileCodes);
    const uint64_t dA = umma_smem_desc(smem_u32(smem + SM::kOffA), p.lbo_bytes, p.sbo_bytes);
    const uint64_t dOnes = umma_smem_desc(smem_u32(smem + SM::kOffOnes), p.lbo_bytes, p.sbo_bytes);
    const uint64_t dB = umma_smem_desc(smem_u32(smem + SM::kOffBData), p.lbo_bytes, p.sbo_bytes);
    const uint64_t dBB0 = umma_smem_desc(smem_u32(smem + SM::kOffBBias), p.lbo_bytes, p.sbo_bytes);
    const uint32_t tacc = tmem_base + (uint32_t)(j * kTileCodes);
    int prev = -1;
    uint32_t run = 0xffffffffu, uses = 0, bi = 0;
    for (int item = item0; item < item_end; ++item, ++bi) {
        const int chunk = __ldg(&p.tile_tab[item]).x & 0xffff;
        if (chunk != prev) {                          // every issuer sees every run (shared barriers: one phase per run)
            prev = chunk;
            ++run;
            mbar_wait(&bars.bd_full[0], run & 1u);
            if (use_bias) mbar_wait(&bars.bb_full[run % kBiasStages], (run / kBiasStages) & 1u);
        }
        const uint32_t sb = run % kBiasStages;
        // every issuer observes every phase of the shared operand-tile barrier (a parity wait is only meaningful within
        // one phase) and takes part in handing the tile back, so the loader cannot get a phase ahead of any of them
        mbar_wait(&bars.a_full[0], bi & 1u);
        if ((int)(bi & 3u) != j) {
            mbar_arrive(&bars.a_empty[0]);
        } else {                                      // tile bi of the CTA lives in accumulator slot bi & 3
            mbar_wait(&bars.acc_empty[j], (uses & 1u) ^ 1u);
            ++uses;
            tc_fence_after();
            uint64_t da = dA, db = dB;
            for (int ks = 0; ks < nsteps; ++ks) {
                umma_f16(tacc, da, db, idesc, ks != 0);
                da += (2 * kPlaneBytes) >> 4;
                db += (2 * kPlaneBytes) >> 4;
            }
            if (use_bias) umma_f16(tacc, dOnes, dBB0 + (uint64_t)(sb * (kBiasBytes >> 4)), idesc, 1u);
            umma_commit(&bars.acc_full[j]);
            umma_commit(&bars.a_empty[0]);
        }
        const bool last_of_run = (item + 1 == item_end) || ((__ldg(&p.tile_tab[item + 1]).x & 0xffff) != chunk);
        if (last_of_run) {
            umma_commit(&bars.bd_empty[0]);
            if (use_bias) umma_commit(&bars.bb_empty[sb]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// REFINE epilogue (two-pass hard route).  A tile is 128 entries of ONE chunk; tile bi of the CTA lives in accumulator
// slot bi & 3.  The 16 epilogue warps split into two groups of eight, each with 16 lanes per row:
//   warps 0-7  DECIDE tile t: both threads of a row (TMEM lane quarter w & 3, 64-column half w >> 2) locate the columns with
//     score >= M - margin (M = the row's best coarse score of the GROUP pass, which has shown that no other chunk -- for a
//     row listed twice: no third chunk -- reaches that bound); the thread of the left half decides: exactly one column
//     -> final index (+ histogram); several -> the row's candidate list for the exact fp32 re-scoring (a row listed under
//     two chunks appends to its list from both tiles).  Decisions go to a 4-tile ring in shared memory.  Then, fused rows
//     only, the same warps do SUMS of tile t - 1: the per-code sums of its decided rows into the fp64 accumulators
//     (128 codes x d in shared memory; the CTA's tiles are contiguous in the chunk-sorted order, so they are flushed only
//     when the chunk changes).  The 16-lane group g OWNS the codes with code % 16 == g, so the accumulators are updated
//     with plain loads and stores: shared-memory fp64 atomics are CAS loops at 2-4 cycles per lane and made the pass
//     three times slower.
//   warps 8-15 (fused rows only) ROWS of tile t, what sorted_rows_kernel does: zf = z / norm, z_q = zf + (w - zf), squared
//     error, on the fp32 rows still sitting in the loader's staging tile (z makes one trip from HBM); zf is left in the
//     staging tile for the sums.
// Both halves are bound by the latency of their own instruction streams (no pipe is more than a quarter busy), which is
// why the work is spread over all 16 warps and why the float -> double widening of the sums is done with integer
// instructions (F2F.F64.F32 runs on the quarter-rate conversion unit that the divisions already keep busy).
// ------------------------------------------------------------------------------------------------
constexpr int kRefHitOff  = 0;       // uint2  [half][128]      columns within the margin (bit masks)
constexpr int kRefSortOff = 2048;    // uint8  [4 tiles][168]   decided rows in code order [128] | first entry of each group [33]
constexpr int kRefSortSlot = 168;
constexpr int kRefCntOff  = 2720;    // int    [128]            counting-sort counters / offsets of the tile being decided
constexpr int kRefFlagOff = 3232;    // int    [1 + 210 + 1]    the CTA's near-tie rows (count first), flushed to flag_rows in batches
constexpr int kRefFlagCap = 210;
constexpr int kRefDecOff  = 4096;    // int4   [4 tiles][128]   {code or -1, row, norm bits, 0}

template <int OFF_SCRATCH, int OFF_ACC, int OFF_STAGE, bool TRACE>
__device__ __forceinline__ void refine_epilogue(const CoarseParams& p, unsigned char* smem, const Bars& bars, uint32_t tmem_base,
                                                int warp, int lane, int item0, int item_end) {
    const int d = p.d, vpr = d >> 2;
    const int row_stride = 4 * d + 16;
    const bool fuse = p.fuse_rows != 0;
    int4* dec_ring = reinterpret_cast<int4*>(smem + OFF_SCRATCH + kRefDecOff);
    const int tid_h = (warp & 7) * 32 + lane;         // within the half (0..255)
    const int g16 = tid_h >> 4, q = tid_h & 15;       // 16 groups of 16 lanes per half
    const bool lane_on = q < vpr;
    // SUMS (fused rows): the per-code sums of a tile's decided rows go into fp64 accumulators in shared memory (128 codes
    // x d; the CTA's tiles are contiguous in the chunk-sorted order, so they are flushed only when the chunk changes).
    // Each of the 32 sixteen-lane groups of the epilogue OWNS four consecutive codes of the chunk, and lane q of the group
    // the dimensions 4 q .. 4 q + 3 of them: every accumulator cell belongs to one thread -- no atomics (shared-memory fp64
    // atomics are CAS loops at 2-4 cycles per lane and made the pass three times slower), no barriers.  The tile's decided
    // rows arrive sorted by code, so a group adds up each code's rows in registers and touches a cell once per code and tile.
    const int G = (warp >> 3) * 16 + g16;
    double* acc = reinterpret_cast<double*>(smem + OFF_ACC);
    unsigned char* sort_s = smem + OFF_SCRATCH + kRefSortOff;         // per ring slot: order[128] | group offsets [33]
    const bool tr_d = TRACE && p.trace != nullptr && blockIdx.x == 0 && warp == 0 && lane == 0;
    int acc_chunk = -1;
    if (fuse && lane_on)
        for (int c4 = 0; c4 < 4; ++c4) {
            double2* a = reinterpret_cast<double2*>(acc + (4 * G + c4) * d + 4 * q);
            a[0] = make_double2(0.0, 0.0); a[1] = make_double2(0.0, 0.0);
        }
    auto flush_own = [&](int ch) {       // this thread's cells of chunk `ch` -> embed_sum (fp64 REDs), then zero
        if (!lane_on) return;
        for (int c4 = 0; c4 < 4; ++c4) {
            const int code = ch * kTileCodes + 4 * G + c4;
            double2* a = reinterpret_cast<double2*>(acc + (4 * G + c4) * d + 4 * q);
            const double2 x0 = a[0], x1 = a[1];
            if (code < p.k) {
                double* dst = p.stats + p.k + (size_t)code * d + 4 * q;
                if (x0.x != 0.0) atomicAdd(dst + 0, x0.x);
                if (x0.y != 0.0) atomicAdd(dst + 1, x0.y);
                if (x1.x != 0.0) atomicAdd(dst + 2, x1.x);
                if (x1.y != 0.0) atomicAdd(dst + 3, x1.y);
            }
            a[0] = make_double2(0.0, 0.0); a[1] = make_double2(0.0, 0.0);
        }
    };
    // SUMS of tile number `t` (chunk `ch`, staging tile `sbuf`); `release_dec`: also hand the decision ring slot back
    auto sums = [&](uint32_t t, int ch, int sbuf, bool release_dec) {
        if (acc_chunk >= 0 && ch != acc_chunk) flush_own(acc_chunk);
        acc_chunk = ch;
        const int slot = (int)(t & 3u);
        mbar_wait(&bars.p1_full[slot], (t >> 2) & 1u);            // the ROWS half has left zf in the staging tile
        mbar_wait(&bars.sort_full[slot], (t >> 2) & 1u);          // the tile's decided rows are listed in code order
        if (TRACE && tr_d && t + 1u < 256u) p.trace[(256 + t + 1u) * 8 + 6] = clock64();
        const int4* dec = dec_ring + slot * 128;
        const unsigned char* stg = smem + OFF_STAGE + sbuf * kStageTileBytes;
        const unsigned char* order = sort_s + slot * kRefSortSlot;
        const int lo = order[128 + G], hi = order[128 + G + 1];   // this group's codes: 4 G .. 4 G + 3
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int cur = -1;
        auto put = [&]() {
            if (cur >= 0 && lane_on) {
                double2* a = reinterpret_cast<double2*>(acc + (cur & (kTileCodes - 1)) * d + 4 * q);
                double2 x0 = a[0], x1 = a[1];
                x0.x += a0; x0.y += a1; x1.x += a2; x1.y += a3;
                a[0] = x0; a[1] = x1;
            }
        };
        // one row ahead: the next row's code and zf are fetched before the current row is added
        int r_n = lo < hi ? (int)order[lo] : 0;
        int code_n = lo < hi ? dec[r_n].x : -1;
        float4 z_n = (lo < hi && lane_on) ? *reinterpret_cast<const float4*>(stg + r_n * row_stride + q * 16)
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = lo; i < hi; ++i) {
            const int code = code_n;
            const float4 zf = z_n;
            if (i + 1 < hi) {
                r_n = (int)order[i + 1];
                code_n = dec[r_n].x;
                if (lane_on) z_n = *reinterpret_cast<const float4*>(stg + r_n * row_stride + q * 16);
            }
            if (code != cur) { put(); cur = code; a0 = a1 = a2 = a3 = 0.0; }
            a0 += (double)zf.x; a1 += (double)zf.y; a2 += (double)zf.z; a3 += (double)zf.w;
        }
        put();
        __syncwarp();
        if (TRACE && tr_d && t < 256u) p.trace[t * 8 + 7] = clock64();
        if (lane == 0) {
            mbar_arrive(&bars.stg_empty[sbuf]);                    // the staging tile may be refilled
            if (release_dec) mbar_arrive(&bars.dec_empty[slot]);
        }
    };
    if (warp < 8) {
        // ---------------------------------- DECIDE (+ SUMS of the previous tile) ----------------------------------
        const int quarter = warp & 3, half = warp >> 2;
        const int rq = quarter * 32 + lane;
        const float margin = p.meta[META_MARGIN];
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 64);
        uint2* hit_s = reinterpret_cast<uint2*>(smem + OFF_SCRATCH + kRefHitOff);
        int* cnt_s = reinterpret_cast<int*>(smem + OFF_SCRATCH + kRefCntOff);
        int* fl_s = reinterpret_cast<int*>(smem + OFF_SCRATCH + kRefFlagOff);    // [0] count, [1...] the CTA's near-tie rows
        if (warp == 0 && lane == 0) fl_s[0] = 0;
        // the CTA's flag list -> the global one (`nthr` threads with ids `tid` that meet at named barrier `bar`)
        auto flush_flags = [&](int tid, int nthr, int bar) {
            int* base_s = fl_s + 1 + kRefFlagCap;
            const int nfl = fl_s[0];
            if (tid == 0) *base_s = atomicAdd(&p.counters[CNT_FLAGGED], nfl);
            conv_bar_sync(bar, nthr);
            const int base = *base_s;
            for (int i = tid; i < nfl; i += nthr) p.flag_rows[base + i] = fl_s[1 + i];
            conv_bar_sync(bar, nthr);
            if (tid == 0) fl_s[0] = 0;
            conv_bar_sync(bar, nthr);
        };
        uint32_t bi = 0;
        int sbuf = 0, sbuf_prev = 0, chunk_prev = -1;
        for (int item = item0; item < item_end; ++item, ++bi) {
            const int2 tt = __ldg(&p.tile_tab[item]);
            const int chunk = tt.x & 0xffff, cnt = tt.x >> 16, start = tt.y;
            const int slot = (int)(bi & 3u);
            const uint32_t use_par = (bi >> 2) & 1u;
            const bool valid = rq < cnt;
            int4 e = make_int4(0, 0, 0, 0);                 // {row | flags, best coarse score, norm, -}
            if (valid) e = __ldg(&p.entries[start + rq]);
            const float thrF = __int_as_float(e.y) - margin;
            mbar_wait(&bars.acc_full[slot], use_par);
            if (TRACE && tr_d && bi < 256u) { p.trace[bi * 8 + 3] = clock64(); p.trace[(256 + bi) * 8 + 0] = clock64(); }
            tc_fence_after();
            uint32_t ha = 0, hb = 0;
            {
                float va[32], vb[32];
                tmem_ld32x2_wait(t_lane + (uint32_t)(slot * kTileCodes), va, vb);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_relaxed_addr(smem_u32(&bars.acc_empty[slot]));
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    ha |= (va[i] >= thrF) ? (1u << i) : 0u;
                    hb |= (vb[i] >= thrF) ? (1u << i) : 0u;
                }
            }
            if (!valid) ha = hb = 0u;
            hit_s[half * 128 + rq] = make_uint2(ha, hb);
            conv_bar_sync(1, 256);
            if (TRACE && tr_d && bi < 256u) p.trace[(256 + bi) * 8 + 1] = clock64();
            // left half: merge the two column halves of every row and decide
            uint32_t msk[4] = {0u, 0u, 0u, 0u};                      // columns [0,32) [32,64) [64,96) [96,128) of the chunk
            int nhit = 0, row = 0, code = -1;
            bool cross = false, second = false, decided = false;
            const int col0 = chunk * kTileCodes;
            if (half == 0) {
                const uint2 o = hit_s[128 + rq];
                msk[0] = ha; msk[1] = hb; msk[2] = o.x; msk[3] = o.y;
                nhit = __popc(ha) + __popc(hb) + __popc(o.x) + __popc(o.y);
                const int first = ha ? __ffs(ha) - 1 : (hb ? 31 + __ffs(hb) : (o.x ? 63 + __ffs(o.x) : (o.y ? 95 + __ffs(o.y) : -1)));
                row = (int)((unsigned)e.x & kEntRowMask);
                cross = ((unsigned)e.x & kEntCross) != 0u;
                second = ((unsigned)e.x & kEntSecond) != 0u;
                decided = valid && !cross && nhit == 1 && col0 + first < p.k;
                code = decided ? col0 + first : -1;
                if (fuse) {       // the other warps only need the decided rows: publish them before the near-tie bookkeeping
                    mbar_wait(&bars.dec_empty[slot], use_par ^ 1u);
                    dec_ring[slot * 128 + rq] = make_int4(code, row, e.z, 0);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars.dec_full[slot]);
                }
            }
            if (fuse) conv_bar_sync(3, 256);                          // the decisions are in the ring: the right half sorts them
            if (TRACE && tr_d && bi < 256u) { p.trace[bi * 8 + 4] = clock64(); p.trace[(256 + bi) * 8 + 3] = clock64(); }
            if (half == 0) {
                // Near-tie bookkeeping.  Nothing here waits for a global atomic: the flag list is collected per CTA in
                // shared memory (flushed with one reservation per ~100 rows), a row listed under two chunks gets its first
                // tile's columns from the front of its candidate list and the second tile's from the back, counted in
                // separate bit fields of cand_cnt by fire-and-forget reductions.
                bool flagged = false, over = false;
                if (valid && !decided) {
                    int* dst = p.cand + (size_t)row * p.cand_cap;
                    if (!cross) {
                        if (nhit >= 2 && nhit <= p.cand_cap) {
                            int at = 0;
#pragma unroll
                            for (int qq = 0; qq < 4; ++qq) {
                                uint32_t m = msk[qq];
                                while (m) { const int i = __ffs(m) - 1; m &= m - 1; dst[at++] = col0 + 32 * qq + i; }
                            }
                            p.cand_cnt[row] = nhit;
                            flagged = true;
                        } else {                                      // too many near-ties: the COLLECT pass / exact scan decides
                            flagged = true;                           // (best_val[row] still holds M from the GROUP pass)
                            over = true;
                        }
                    } else {
                        int at = 0;
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) {
                            uint32_t m = msk[qq];
                            while (m) {
                                const int i = __ffs(m) - 1;
                                m &= m - 1;
                                if (at < p.cand_cap) dst[second ? p.cand_cap - 1 - at : at] = col0 + 32 * qq + i;
                                ++at;
                            }
                        }
                        atomicAdd(&p.cand_cnt[row], second ? (min(nhit, 255) << kCandBackShift) : nhit);   // result unused: RED
                        flagged = !second;
                    }
                }
                const unsigned dall = __ballot_sync(0xffffffffu, decided);
                if (decided) {
                    p.idx_out[row] = (int64_t)code;
                    const unsigned same = __match_any_sync(dall, code);
                    if (lane == __ffs(same) - 1) atomicAdd(&p.bins[code], __popc(same));
                }
                const unsigned ball = __ballot_sync(0xffffffffu, flagged);
                if (ball) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&fl_s[0], __popc(ball));            // shared-memory counter
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (flagged) fl_s[1 + base + __popc(ball & ((1u << lane) - 1))] = row;
                }
                const unsigned oall = __ballot_sync(0xffffffffu, over);
                if (oall) {                                           // rare
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&p.counters[CNT_OVER], __popc(oall));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (over) p.over_rows[base + __popc(oall & ((1u << lane) - 1))] = row;
                }
                // flush the CTA's flag list before it can overflow (128 more rows fit after this point)
                conv_bar_sync(2, 128);
                if (fl_s[0] > kRefFlagCap - 128) flush_flags(warp * 32 + lane, 128, 2);
            } else if (fuse) {
                // right half: counting sort of the tile's decided rows by code (128 threads = 128 rows, 128 possible codes):
                // the SUMS step then adds up each code's rows in registers and touches an accumulator once per code and tile
                unsigned char* order = sort_s + slot * kRefSortSlot;
                code = dec_ring[slot * 128 + rq].x;
                decided = code >= 0;
                cnt_s[rq] = 0;
                conv_bar_sync(4, 128);
                const int lc = code & (kTileCodes - 1);
                const int rank = decided ? atomicAdd(&cnt_s[lc], 1) : 0;
                conv_bar_sync(4, 128);
                if (warp == 4) {
                    const int4 c4 = *reinterpret_cast<const int4*>(&cnt_s[4 * lane]);
                    const int tot4 = c4.x + c4.y + c4.z + c4.w;
                    int incl = tot4;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += v;
                    }
                    const int ex0 = incl - tot4;
                    *reinterpret_cast<int4*>(&cnt_s[4 * lane]) = make_int4(ex0, ex0 + c4.x, ex0 + c4.x + c4.y, ex0 + c4.x + c4.y + c4.z);
                    order[128 + lane] = (unsigned char)ex0;               // first entry of group `lane` (codes 4 lane ...)
                    if (lane == 31) order[128 + 32] = (unsigned char)incl;
                }
                conv_bar_sync(4, 128);
                if (decided) order[cnt_s[lc] + rank] = (unsigned char)rq;
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars.sort_full[slot]);
            }
            if (TRACE && tr_d && bi < 256u) p.trace[(256 + bi) * 8 + 4] = clock64();
            conv_bar_sync(1, 256);                                   // hit_s is rewritten by the next tile
            if (TRACE && tr_d && bi < 256u) p.trace[(256 + bi) * 8 + 5] = clock64();
            if (fuse && bi >= 1u) sums(bi - 1u, chunk_prev, sbuf_prev, false);
            if (TRACE && tr_d && bi < 256u) p.trace[(256 + bi) * 8 + 7] = clock64();
            chunk_prev = chunk;
            sbuf_prev = sbuf;
            sbuf = sbuf == kStageTiles - 1 ? 0 : sbuf + 1;
        }
        conv_bar_sync(1, 256);
        flush_flags(tid_h, 256, 1);
        if (fuse) {
            if (bi >= 1u) sums(bi - 1u, chunk_prev, sbuf_prev, false);
            if (acc_chunk >= 0) flush_own(acc_chunk);
        }
    } else if (fuse) {
        // ---------------------------------- ROWS ----------------------------------
        double sse = 0.0;
        uint32_t bi = 0;
        int sbuf = 0;                                     // bi % kStageTiles
        for (int item = item0; item < item_end; ++item, ++bi, sbuf = (sbuf == kStageTiles - 1 ? 0 : sbuf + 1)) {
            const int chunk = __ldg(&p.tile_tab[item]).x & 0xffff;
            const int slot = (int)(bi & 3u);
            mbar_wait(&bars.dec_full[slot], (bi >> 2) & 1u);
            const bool tr_f = TRACE && p.trace != nullptr && blockIdx.x == 0 && warp == 8 && lane == 0 && bi < 256u;
            if (TRACE && tr_f) p.trace[bi * 8 + 5] = clock64();
            const int4* dec = dec_ring + slot * 128;
            unsigned char* stg_w = smem + OFF_STAGE + sbuf * kStageTileBytes;
            // group g takes rows g, g + 16, ...; four codebook rows in flight at a time
#pragma unroll 1
            for (int t0 = 0; t0 < 8; t0 += 4) {
                int4 de[4];
                float4 c[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    de[u] = dec[g16 + 16 * (t0 + u)];
                    c[u] = (lane_on && de[u].x >= 0) ? __ldg(reinterpret_cast<const float4*>(p.w + (size_t)de[u].x * d) + q)
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (lane_on && de[u].x >= 0) {
                        float4* xp = reinterpret_cast<float4*>(stg_w + (g16 + 16 * (t0 + u)) * row_stride + q * 16);
                        const float4 x = *xp;
                        const float dn = __int_as_float(de[u].z);
                        float4 zf;
                        zf.x = __fdiv_rn(x.x, dn); zf.y = __fdiv_rn(x.y, dn); zf.z = __fdiv_rn(x.z, dn); zf.w = __fdiv_rn(x.w, dn);
                        *xp = zf;
                        const float d0 = __fsub_rn(c[u].x, zf.x), d1 = __fsub_rn(c[u].y, zf.y), d2 = __fsub_rn(c[u].z, zf.z),
                                    d3 = __fsub_rn(c[u].w, zf.w);
                        // the four squares of a lane are added in fp32 first (one widening per row and lane instead of four)
                        sse += (double)__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)),
                                                 __fadd_rn(__fmul_rn(d2, d2), __fmul_rn(d3, d3)));
                        float4 o;
                        o.x = __fadd_rn(zf.x, d0); o.y = __fadd_rn(zf.y, d1); o.z = __fadd_rn(zf.z, d2); o.w = __fadd_rn(zf.w, d3);
                        __stcs(reinterpret_cast<float4*>(p.z_q + (size_t)de[u].y * d) + q, o);
                    }
                }
            }
            __syncwarp();
            if (TRACE && tr_f) p.trace[bi * 8 + 6] = clock64();
            if (lane == 0) mbar_arrive(&bars.p1_full[slot]);
            // this half's share of the tile's sums right away (the deciders do theirs one tile later, after the next
            // decision): the staging tile is handed back a whole tile period earlier, and with three staging tiles that
            // period is what bounds the pass (copy -> convert -> MMA -> decide -> rows -> sums must fit into two periods)
            sums(bi, chunk, sbuf, true);
        }
        if (acc_chunk >= 0) flush_own(acc_chunk);
        sse = warp_sum(sse);
        if (lane == 0 && sse != 0.0) atomicAdd(&p.stats[(size_t)p.k + (size_t)p.k * d], sse);
    }
}

}  // namespace

template <int RT, int MODE, int NTOP, bool DBG>
__global__ void __launch_bounds__(kThreads, 1) coarse_kernel(const CoarseParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    using SM = SmemMap<RT, MODE == MODE_REFINE>;
    constexpr int kDataStages = SM::kStages;
    constexpr int kRing = SM::kStages;
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
    // REFINE walks the 128-entry tiles of the chunk-sorted entries.  A CTA takes a CONTIGUOUS range of tiles: it meets few
    // different chunks, so its shared-memory per-code accumulators are flushed rarely.
    int item0 = blockIdx.x, item_end = nitems, item_step = gridDim.x;
    if (MODE == MODE_REFINE) {
        const int nt = *p.n_tiles_dev;
        const int per = (nt + (int)gridDim.x - 1) / (int)gridDim.x;
        item0 = (int)blockIdx.x * per;
        item_end = min(nt, item0 + per);
        item_step = 1;
    }

    // bias K-step (-||w||^2/2): needed unless every code has the same squared norm (then it is a constant
    // shift of all scores); padding codes rely on it to lose, so it stays on when K is not a tile multiple
    const bool use_bias = !(DBG && (p.debug_flags & 2)) &&
                          (MODE == MODE_DUMP || p.k != p.kpad || !(p.meta[META_SPREAD] <= kBiasSkipTol));

    Bars bars = carve_bars<SM>(smem);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kOffTmem);

    // ---- one-time setup ------------------------------------------------------------------------
    if (tid_x == 0) {
        // every codebook stage is released by all four issuers (owners via tcgen05.commit, the others arrive)
        for (int i = 0; i < kDataStages; ++i) { mbar_init(&bars.bd_full[i], 1); mbar_init(&bars.bd_empty[i], 4); }
        for (int i = 0; i < kBiasStages; ++i) { mbar_init(&bars.bb_full[i], 1); mbar_init(&bars.bb_empty[i], 4); }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&bars.a_full[i], (MODE == MODE_REFINE) ? 2 : kLoadWarps);  // one elected lane per loader (REFINE: converting) warp
            mbar_init(&bars.a_empty[i], 4 / RT);     // one commit from each issuer that reads the tile
            mbar_init(&bars.acc_full[i], 1);
            mbar_init(&bars.acc_empty[i], 8);        // one elected lane per epilogue warp serving the slot
            mbar_init(&bars.dec_full[i], 4);         // REFINE: the four deciding warps of the left column half
            mbar_init(&bars.dec_empty[i], 8);        //         the eight finishing warps
            mbar_init(&bars.p1_full[i], 8);
            mbar_init(&bars.sort_full[i], 4);
            if (i < kStageTiles) { mbar_init(&bars.stg_empty[i], 16); mbar_init(&bars.landed[i], 64); }
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
      if constexpr (MODE == MODE_REFINE) {
        refine_epilogue<SM::kOffSlots, SM::kOffAcc, SM::kOffStage, DBG>(p, smem, bars, tmem_base, warp, lane, item0, item_end);
      } else {
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
        int* hist_s = reinterpret_cast<int*>(smem + kOffSlots);   // GROUP: this CTA's sorted entries per chunk
        if (MODE == MODE_GROUP) {
            for (int i = (int)tid_x; i < nchunks; i += kEpiWarps * 32) hist_s[i] = 0;
            named_bar_sync(5, kEpiWarps * 32);
        }
        for (int item = item0; item < item_end; item += item_step) {
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
                    } else if (MODE == MODE_GROUP) {
                        consume_group<NLIST>(va, vb, c, st[ss]);
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
            } else if (MODE == MODE_GROUP) {
#pragma unroll
                for (int ss = 0; ss < NS; ++ss)
                    state[ss * 512 + rec0] = make_float4(st[ss].m1, st[ss].m2,
                                                         __int_as_float(st[ss].best_chunk | (st[ss].chunk2 << 16)), st[ss].m3);
                named_bar_sync(bar_id, bar_n);
                if (leader) {
#pragma unroll
                    for (int ss = 0; ss < NS; ++ss) {
                        // the row's best chunk, and every OTHER chunk holding a 64-column group within the margin of the best
                        float M = -INFINITY;
                        int cb = 0;
#pragma unroll
                        for (int pp = 0; pp < NP; ++pp)
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                const int w2 = quarter + 4 * hh + 8 * ((RT == 1) ? pp : pair);
                                const float4 sv = state[ss * 512 + w2 * 32 + lane];
                                if (sv.x > M) { M = sv.x; cb = __float_as_int(sv.z) & 0xffff; }
                            }
                        const float thrF = M - margin;
                        int other = -1;
                        bool unknown = !(M > -INFINITY);          // no finite score (NaN row): the exact scan decides
#pragma unroll
                        for (int pp = 0; pp < NP; ++pp)
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                const int w2 = quarter + 4 * hh + 8 * ((RT == 1) ? pp : pair);
                                const float4 sv = state[ss * 512 + w2 * 32 + lane];
                                const int c1 = __float_as_int(sv.z) & 0xffff, c2 = (int)((unsigned)__float_as_int(sv.z) >> 16);
                                if (sv.w >= thrF) unknown = true;            // a group known only by its maximum
                                if (sv.x >= thrF && c1 != cb) { if (other < 0 || other == c1) other = c1; else unknown = true; }
                                if (sv.y >= thrF && c2 != cb) { if (other < 0 || other == c2) other = c2; else unknown = true; }
                            }
                        const bool valid = st[ss].valid;
                        if (valid) {
                            p.best_val[st[ss].row] = M;
                            p.grp[st[ss].row] = cb | ((other < 0 ? 0 : other) << kGrpOtherShift) |
                                                ((other >= 0 && !unknown) ? kGrpCrossBit : 0) | (unknown ? kGrpOverBit : 0);
                            if (!unknown) {
                                // rank of the entry within this CTA's share of its chunk: the group sort then places every
                                // entry without a single global atomic (base of (CTA, chunk) + rank)
                                p.grank[st[ss].row] = atomicAdd(&hist_s[cb], 1);
                                if (other >= 0) p.grank2[st[ss].row] = atomicAdd(&hist_s[other], 1);
                            }
                        }
                        // three or more chunks within the margin (or nothing finite): flagged + COLLECT, like ARGMAX's overflow
                        const bool over = valid && unknown;
                        const unsigned oall = __ballot_sync(0xffffffffu, over);
                        if (oall) {
                            int base = 0, base2 = 0;
                            if (lane == 0) {
                                base = atomicAdd(&p.counters[CNT_FLAGGED], __popc(oall));
                                base2 = atomicAdd(&p.counters[CNT_OVER], __popc(oall));
                            }
                            base = __shfl_sync(0xffffffffu, base, 0);
                            base2 = __shfl_sync(0xffffffffu, base2, 0);
                            const int at = __popc(oall & ((1u << lane) - 1));
                            if (over) { p.flag_rows[base + at] = (int)st[ss].row; p.over_rows[base2 + at] = (int)st[ss].row; }
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
        if (MODE == MODE_GROUP) {      // the CTA's histogram (every row block has been counted) goes to its row of the table
            named_bar_sync(5, kEpiWarps * 32);
            for (int i = (int)tid_x; i < nchunks; i += kEpiWarps * 32) {
                // reserve this CTA's share of the chunk: offset within the chunk's range, table [CTA][chunk]
                const int h = hist_s[i];
                if (h) p.ghist[(size_t)blockIdx.x * nchunks + i] = atomicAdd(&p.gtot[i], h);
            }
        }
      }
    } else if (MODE == MODE_REFINE && warp < kLoadWarp0 + kLoadWarps) {
        if constexpr (MODE == MODE_REFINE) refine_loader<SM, DBG>(p, smem, bars, (int)tid_x - kLoadWarp0 * 32, lane, item0, item_end);
    } else if (warp < kLoadWarp0 + kLoadWarps) {
        // ---------------------------------- A LOADER --------------------------------------------
        // one thread per row.  The exact reference-order norm goes to `denom` (later kernels divide by
        // it); the fp16 MMA operand itself only needs to be inside the coarse error bound, so it is
        // produced with one reciprocal and a multiply per element instead of an IEEE division.
        const int r = (int)tid_x - kLoadWarp0 * 32;
        uint32_t bi = 0;
        for (int item = item0; item < item_end; item += item_step, ++bi) {
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
                        // four planes (eight 16-byte loads) in flight at a time: one thread reads a whole row, and with one
                        // plane per iteration every iteration exposed a DRAM round trip (~10 us per tile)
                        for (int j0 = 0; j0 < nplanes; j0 += 4) {
                            float4 x[8];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const bool in = j0 + u < nplanes;
                                x[2 * u] = in ? __ldg(zr4 + 2 * (j0 + u)) : make_float4(0.f, 0.f, 0.f, 0.f);
                                x[2 * u + 1] = in ? __ldg(zr4 + 2 * (j0 + u) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                if (j0 + u < nplanes) {      // (in order of the planes: the canonical accumulation order)
                                    const float4 x0 = x[2 * u], x1 = x[2 * u + 1];
                                    a0 = __fadd_rn(a0, __fmul_rn(x0.x, x0.x)); a1 = __fadd_rn(a1, __fmul_rn(x0.y, x0.y));
                                    a2 = __fadd_rn(a2, __fmul_rn(x0.z, x0.z)); a3 = __fadd_rn(a3, __fmul_rn(x0.w, x0.w));
                                    a4 = __fadd_rn(a4, __fmul_rn(x1.x, x1.x)); a5 = __fadd_rn(a5, __fmul_rn(x1.y, x1.y));
                                    a6 = __fadd_rn(a6, __fmul_rn(x1.z, x1.z)); a7 = __fadd_rn(a7, __fmul_rn(x1.w, x1.w));
                                }
                            }
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
                    if (MODE == MODE_GROUP) p.cand_cnt[row] = 0;      // candidate lists start empty (REFINE / COLLECT append)
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
    } else if (MODE == MODE_REFINE && warp == kProdWarp) {
        if constexpr (MODE == MODE_REFINE) { if (elect_one()) refine_producer<SM>(p, smem, bars, item0, item_end, use_bias); }
    } else if (MODE == MODE_REFINE && warp >= kMmaWarp) {
        if constexpr (MODE == MODE_REFINE) { if (elect_one()) refine_issuer<SM>(p, smem, bars, tmem_base, warp - kMmaWarp, item0, item_end, use_bias); }
    } else if (warp == kProdWarp) {
        // ---------------------------------- B PRODUCER ------------------------------------------
        if (elect_one()) {
            const unsigned char* img = static_cast<const unsigned char*>(p.b_image);
            uint32_t itB = 0;
            uint32_t st = 0, ph = 0;      // ring position and phase, advanced per K-block (the ring need not be a power of 2)
            for (int item = item0; item < item_end; item += item_step) {
                int cbeg = (MODE == MODE_COLLECT) ? (item % ksplit) * cps : 0;
                int cend = min(nchunks, cbeg + cps);
                for (int c = cbeg; c < cend; ++c) {
                    const unsigned char* tile = img + (size_t)c * tile_b;
                    for (int kb = 0; kb < nkblk; ++kb) {
                        const uint32_t bytes = (uint32_t)min(8, nplanes - 8 * kb) * kPlaneBytes;
                        mbar_wait(&bars.bd_empty[st], ph ^ 1);
                        mbar_arrive_expect_tx(&bars.bd_full[st], bytes);
                        bulk_g2s(smem + kOffBData + st * kKblkBytes, tile + (size_t)kb * kKblkBytes, bytes,
                                 &bars.bd_full[st]);
                        if (++st == (uint32_t)kRing) { st = 0; ph ^= 1; }
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
            for (int item = item0; item < item_end; item += item_step, ++bi, chunk0 += (uint32_t)(cend - cbeg)) {
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
                    auto next_stage = [&](uint32_t& st_, uint32_t& ph_) { if (++st_ == (uint32_t)kRing) { st_ = 0; ph_ ^= 1; } };
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
    using SM = SmemMap<RT, MODE == MODE_REFINE>;
    static PerDeviceOnce once;         // one per instantiation and device
    bool& configured = once.flag();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kSmemBytes);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    kern<<<grid, kThreads, SM::kSmemBytes, stream>>>(p);
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
    if (mode == MODE_GROUP) {
        if (row_tiles != 4) return cudaErrorInvalidValue;
        // slot-loop trace of the GROUP pass (mtvq_debug_set_trace + MTVQ_TRACE_GROUP=1; the REFINE pass leaves the buffer alone then)
        // (MTVQ_DBG timing experiments 1 / 2 / 4 apply to this pass too; results are wrong when set)
        const bool dbg_g = (p.trace != nullptr && (p.debug_flags & 0x100)) || (p.debug_flags & 7);
        return dbg_g ? launch_one<4, MODE_GROUP, 5, true>(p, num_sms, stream) : launch_one<4, MODE_GROUP, 5, false>(p, num_sms, stream);
    }
    if (mode == MODE_REFINE)
        return p.trace != nullptr ? launch_one<1, MODE_REFINE, 5, true>(p, num_sms, stream)
                                  : launch_one<1, MODE_REFINE, 5, false>(p, num_sms, stream);
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

int coarse_smem_bytes() {
    int m = SmemMap<4>::kSmemBytes > SmemMap<1>::kSmemBytes ? SmemMap<4>::kSmemBytes : SmemMap<1>::kSmemBytes;
    return m > SmemMap<1, true>::kSmemBytes ? m : SmemMap<1, true>::kSmemBytes;
}

}  // namespace mtvq
