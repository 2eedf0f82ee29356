// common.cuh -- shared device helpers for libmedtok_vq (sm_100a only).
//
//  * "canonical arithmetic": the fp32 operation orders of the reference's CPU path
//    (DESIGN.md section 3).  Every function here mirrors oracle/vq_oracle.c
//    operation for operation; __fmul_rn/__fadd_rn/__fmaf_rn/__fdiv_rn/__fsqrt_rn are used so the
//    compiler can neither contract nor reassociate.
//  * thin PTX wrappers: mbarrier, cp.async.bulk, tcgen05 (alloc / mma / commit / ld / fences).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "workspace.h"

namespace mtvq {

constexpr float kNormEps     = 1e-12f;   // F.normalize eps
constexpr int   kCandMax     = 16;       // near-tie candidates kept per row before the full-scan fallback
constexpr int   kTileCodes   = 128;      // codes per MMA tile (UMMA N)
constexpr int   kTileRows    = 128;      // rows per MMA tile (UMMA M)
constexpr int   kBiasCols    = 16;       // one extra UMMA K-step carrying -||w||^2/2 as 3 fp16 terms
constexpr int   kChunk       = 32;       // epilogue granule: tcgen05.ld 32x32b.x32

// ------------------------------------------------------------------------------------------------
// canonical arithmetic
// ------------------------------------------------------------------------------------------------

// torch.linalg.vector_norm over a unit-stride row: one 8-lane accumulator, acc += fl(x*x),
// lanes summed 0..7, scalar tail (oracle: norm2_row).  `get(i)` returns element i.
template <class Get>
__device__ __forceinline__ float canon_norm8(Get get, int d) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f, a6 = 0.f, a7 = 0.f;
    const int main_ = d & ~7;
    for (int c = 0; c < main_; c += 8) {
        float x;
        x = get(c + 0); a0 = __fadd_rn(a0, __fmul_rn(x, x));
        x = get(c + 1); a1 = __fadd_rn(a1, __fmul_rn(x, x));
        x = get(c + 2); a2 = __fadd_rn(a2, __fmul_rn(x, x));
        x = get(c + 3); a3 = __fadd_rn(a3, __fmul_rn(x, x));
        x = get(c + 4); a4 = __fadd_rn(a4, __fmul_rn(x, x));
        x = get(c + 5); a5 = __fadd_rn(a5, __fmul_rn(x, x));
        x = get(c + 6); a6 = __fadd_rn(a6, __fmul_rn(x, x));
        x = get(c + 7); a7 = __fadd_rn(a7, __fmul_rn(x, x));
    }
    float s = a0;
    s = __fadd_rn(s, a1); s = __fadd_rn(s, a2); s = __fadd_rn(s, a3);
    s = __fadd_rn(s, a4); s = __fadd_rn(s, a5); s = __fadd_rn(s, a6); s = __fadd_rn(s, a7);
    for (int i = main_; i < d; ++i) { float x = get(i); s = __fadd_rn(s, __fmul_rn(x, x)); }
    return __fsqrt_rn(s);
}

// vector_norm over a strided view: plain sequential sum (oracle: norm2_row_seq)
template <class Get>
__device__ __forceinline__ float canon_norm_seq(Get get, int d) {
    float s = 0.f;
    for (int i = 0; i < d; ++i) { float x = get(i); s = __fadd_rn(s, __fmul_rn(x, x)); }
    return __fsqrt_rn(s);
}

// norm_mode is a small bit field (include/medtok_vq.h): bits 0-1 = MTVQ_NORM_* (0 contiguous, 1 strided, 2 none: the
// rows are used as they are, denominator exactly 1), bit 2 = MTVQ_SCORE_DOT (rank codes by the inner product alone)
__device__ __forceinline__ int  norm_kind(int norm_mode) { return norm_mode & 3; }
__device__ __forceinline__ bool score_dot(int norm_mode) { return (norm_mode & 4) != 0; }

template <class Get>
__device__ __forceinline__ float canon_denom(Get get, int d, int norm_mode) {
    if (norm_kind(norm_mode) == 2) return 1.0f;
    float nrm = norm_kind(norm_mode) ? canon_norm_seq(get, d) : canon_norm8(get, d);
    return nrm < kNormEps ? kNormEps : nrm;      // clamp_min(eps); NaN stays NaN like torch
}

// x.pow(2).sum(dim=1): ATen cascade_sum inner kernel, 4 accumulators x 8 lanes (oracle: sumsq_row)
template <class Get>
__device__ __forceinline__ float canon_sumsq(Get get, int d) {
    float acc[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 8; ++v) acc[u][v] = 0.f;
    const int nvec = d >> 3, n4 = nvec >> 2;
    for (int c = 0; c < n4; ++c) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                float e = get((c * 4 + u) * 8 + v);
                acc[u][v] = __fadd_rn(acc[u][v], __fmul_rn(e, e));
            }
    }
    for (int i = n4 * 4; i < nvec; ++i) {
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            float e = get(i * 8 + v);
            acc[0][v] = __fadd_rn(acc[0][v], __fmul_rn(e, e));
        }
    }
#pragma unroll
    for (int u = 1; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 8; ++v) acc[0][v] = __fadd_rn(acc[0][v], acc[u][v]);
    float fin = 0.f;
    for (int i = nvec * 8; i < d; ++i) { float e = get(i); fin = __fadd_rn(fin, __fmul_rn(e, e)); }
#pragma unroll
    for (int v = 0; v < 8; ++v) fin = __fadd_rn(fin, acc[0][v]);
    return fin;
}

// canon_sumsq of a row held in shared memory, with the 32 independent accumulator chains (4 accumulators x 8 lanes)
// spread over the lanes of a warp: the additions inside each chain and the order in which the chains are folded are
// exactly those of canon_sumsq, only the chains run side by side.  d % 8 == 0; every lane returns the sum.
__device__ __forceinline__ float warp_canon_sumsq(const float* xs, int d, int lane) {
    const int u = lane >> 3, v = lane & 7;
    const int nvec = d >> 3, n4 = nvec >> 2;
    float a = 0.f;
    for (int c = 0; c < n4; ++c) { const float e = xs[(c * 4 + u) * 8 + v]; a = __fadd_rn(a, __fmul_rn(e, e)); }
    if (u == 0)
        for (int i = n4 * 4; i < nvec; ++i) { const float e = xs[i * 8 + v]; a = __fadd_rn(a, __fmul_rn(e, e)); }
    float f = __shfl_sync(0xffffffffu, a, v);                                 // acc[0][v]
    f = __fadd_rn(f, __shfl_sync(0xffffffffu, a, 8 + v));
    f = __fadd_rn(f, __shfl_sync(0xffffffffu, a, 16 + v));
    f = __fadd_rn(f, __shfl_sync(0xffffffffu, a, 24 + v));
    float fin = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) fin = __fadd_rn(fin, __shfl_sync(0xffffffffu, f, t));
    return fin;
}

// d = fl( fl(zz + cc) - fl(2 * dot) ), dot = sequential fmaf chain (MKL sgemm, oracle: dist_row)
__device__ __forceinline__ float canon_dist(float zz, float cc, float dot) {
    return __fsub_rn(__fadd_rn(zz, cc), __fmul_rn(2.0f, dot));
}

// ------------------------------------------------------------------------------------------------
// warp helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------
// PTX: shared-memory addresses, mbarrier, bulk copy, proxy fences
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// try_wait may suspend the thread in hardware until the phase completes or a time limit expires; the
// optional hint (ns) raises that limit, so waiters sleep instead of re-polling every few dozen cycles
// (polling loops were ~30% of all issued instructions of the coarse kernel, profiles/r01).
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t hint_ns = 20000u) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity), "r"(hint_ns)
            : "memory");
    } while (!ok);
}

// long waits off the critical path (the A loader waits ~a whole row block for its tile to be released): poll with
// an explicit sleep in between, so the polling loop stops competing with the epilogue warps for issue slots
// (it was 13% of all warp instructions of the coarse kernel, profiles/r02)
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, uint32_t sleep_ns) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) break;
        __nanosleep(sleep_ns);
    }
}

// same, on a precomputed shared-space address (keeps address arithmetic out of hot loops)
__device__ __forceinline__ void mbar_wait_addr(uint32_t addr, uint32_t parity, uint32_t hint_ns = 20000u) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity), "r"(hint_ns)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_arrive_addr(uint32_t addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
// Arrive without release semantics: for barriers that only hand a tensor-memory slot back (the tcgen05 fences order
// the TMEM accesses).  A releasing arrive first waits for the thread's outstanding global stores to be performed --
// an L2 round trip on the slot loop's critical path whenever the epilogue has just written records to global memory.
__device__ __forceinline__ void mbar_arrive_relaxed_addr(uint32_t addr) {
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}

// 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// make generic-proxy smem writes (st.shared) visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Published bounds (assign_tc.cu): one thread keeps storing a value that only grows, another reads it without any
// synchronisation and only needs SOME earlier value.  Relaxed CTA-scope accesses make that well defined in the PTX
// memory model (a plain C++ load/store pair would be a data race); each 32-bit component is read or written whole.
__device__ __forceinline__ float ld_relaxed_smem(const float* p) {
    float v;
    asm volatile("ld.relaxed.cta.shared::cta.f32 %0, [%1];" : "=f"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_smem(float* p, float v) {
    asm volatile("st.relaxed.cta.shared::cta.f32 [%0], %1;" ::"r"(smem_u32(p)), "f"(v) : "memory");
}
__device__ __forceinline__ float4 ld_relaxed_smem4(const float4* p) {
    float4 v;
    asm volatile("ld.relaxed.cta.shared::cta.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_smem4(float4* p, float4 v) {
    asm volatile("st.relaxed.cta.shared::cta.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// non-blocking arrival of the whole (converged) warp on a named barrier that other warps bar.sync on
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------------------
// PTX: tcgen05 (5th-gen tensor cores, accumulators in tensor memory)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {        // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, fp16 operands, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// K-major, no-swizzle ("interleave") shared-memory matrix descriptor: 8x16-byte core matrices,
// LBO = byte step between core matrices along K, SBO = byte step along M/N.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version = 1 on sm_100
    return d;                 // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}
// instruction descriptor for kind::f16: A,B = fp16 K-major, D = fp32, shape M x N
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// 32 lanes x 32 consecutive fp32 columns of tensor memory -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// two adjacent 32-column granules in ONE asm statement (both loads, then the wait): the compiler can neither
// schedule work between the two loads nor borrow the second granule's registers as temporaries before it is issued
// (it did, under register pressure, and serialised the loads -- profiles/r02)
__device__ __forceinline__ void tmem_ld32x2_wait(uint32_t taddr, float (&a)[32], float (&b)[32]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(a);
    uint32_t* q = reinterpret_cast<uint32_t*>(b);
    asm volatile(
        "{\n\t.reg .b32 t1;\n\t"
        "add.u32 t1, %64, 32;\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%64];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [t1];\n\t"
        "tcgen05.wait::ld.sync.aligned;\n\t}"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
          "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]),
          "=r"(q[8]), "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]),
          "=r"(q[16]), "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]),
          "=r"(q[24]), "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace mtvq
