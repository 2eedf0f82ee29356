// microbench.cu -- two tiny measurements that size the epilogue of assign_tc.cu:
//   which = 0 : tcgen05.ld 32x32b.x32 throughput (bytes of tensor memory -> registers per second)
//   which = 1 : FMNMX3 tree throughput (fp32 scores reduced per second)
// Not part of the product path; reachable only through mtvq_microbench().
#include "common.cuh"
#include "microbench.h"

namespace mtvq {

__global__ void __launch_bounds__(1024, 1) tmem_ld_bench_kernel(int iters, float* sink) {
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int c = 0; c < 512; c += 32) {
            float v[32];
            tmem_ld32(base + ((c + warp * 32) & 511), v);
            tmem_ld_wait();
            acc += v[0] + v[17] + v[31];   // static indices: keeps v[] in registers
        }
    }
    if (acc == 123.456f) sink[threadIdx.x] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_ptr, 512); }
}

__global__ void __launch_bounds__(1024, 1) fmnmx_bench_kernel(int iters, const float* in, float* sink) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = in[(threadIdx.x + i) & 1023];
    float best = -INFINITY;
    for (int it = 0; it < iters; ++it) {
        float a = fmaxf(fmaxf(v[0], v[1]), v[2]), b = fmaxf(fmaxf(v[3], v[4]), v[5]);
        float c = fmaxf(fmaxf(v[6], v[7]), v[8]), d = fmaxf(fmaxf(v[9], v[10]), v[11]);
        float e = fmaxf(fmaxf(v[12], v[13]), v[14]), f = fmaxf(fmaxf(v[15], v[16]), v[17]);
        float g = fmaxf(fmaxf(v[18], v[19]), v[20]), h = fmaxf(fmaxf(v[21], v[22]), v[23]);
        float i2 = fmaxf(fmaxf(v[24], v[25]), v[26]), j = fmaxf(fmaxf(v[27], v[28]), v[29]);
        float k = fmaxf(v[30], v[31]);
        a = fmaxf(fmaxf(a, b), c); d = fmaxf(fmaxf(d, e), f); g = fmaxf(fmaxf(g, h), i2); j = fmaxf(j, k);
        const float m = fmaxf(fmaxf(fmaxf(a, d), g), j);
        best = fmaxf(best, m);
        v[0] = __int_as_float(__float_as_int(m) ^ it);          // perturb so iterations are not hoisted
        v[31] = __int_as_float(__float_as_int(v[31]) + it);
    }
    if (best == 123.456f) sink[threadIdx.x] = best;
}

// which = 2: back-to-back tcgen05.mma (M=128, N in {64,128,256}, K=16, fp16, operands in smem).
// which = 3: the same while 16 more warps stream tcgen05.ld.32x32b.x32 from the OTHER half of tensor memory
//            (contention between accumulator read-modify-write and epilogue loads).
// which = 4: the same while 16 more warps stream 16-byte shared-memory stores (operand-fetch contention).
__global__ void __launch_bounds__(640, 1) umma_bench_kernel(int iters, int n_cols, int mode, float* sink) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint32_t tmem_ptr;
    __shared__ uint64_t bar;
    __shared__ volatile int done;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;  // 1.0h
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); done = 0; }
    fence_proxy_async_smem();
    if (warp == 0) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    float acc = 0.f;
    if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_f16(128, n_cols);
            const uint64_t da = umma_smem_desc(smem_u32(sm), 2048, 128);
            const uint64_t db = umma_smem_desc(smem_u32(sm + 16384), 4096, 128);
            for (int it = 0; it < iters; ++it) {
                // 4 K-steps into one accumulator, like one tile of the real kernel (accumulators in columns 0..255)
                const uint32_t tacc = tmem_ptr + (uint32_t)((it & 1) * (256 - n_cols > 0 ? n_cols : 0));
                umma_f16(tacc, da, db, idesc, 0);
                umma_f16(tacc, da + 256, db + 512, idesc, 1);
                umma_f16(tacc, da + 512, db + 1024, idesc, 1);
                umma_f16(tacc, da + 768, db + 1536, idesc, 1);
            }
            umma_commit(&bar);
            mbar_wait(&bar, 0);
            done = 1;
        }
    } else if (warp >= 4 && mode == 3) {
        const uint32_t base = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16) + 256;   // columns 256..511
        while (!done) {
#pragma unroll 1
            for (int c = 0; c < 256; c += 32) {
                float v[32];
                tmem_ld32(base + c, v);
                tmem_ld_wait();
                acc += v[0] + v[31];
            }
        }
    } else if (warp >= 4 && mode == 4) {
        uint4* dst = reinterpret_cast<uint4*>(sm + 49152) + (threadIdx.x - 128);
        uint4 val = make_uint4(threadIdx.x, 1, 2, 3);
        while (!done) {
#pragma unroll
            for (int c = 0; c < 8; ++c) { dst[0] = val; val.x += 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_ptr, 512); }
    if (acc == 123.456f) sink[threadIdx.x] = acc;
}

// which = 5: `nissue` (1, 2 or 4) issuer warps, each feeding its own 128-column accumulator with 4-MMA tiles and
// waiting for its own tcgen05.commit before the next tile (the slot loop of the real kernel without an epilogue).
// result: aggregate TFLOP/s and, via result_host[1], flops.  Tells whether MMAs from different warps pipeline.
__global__ void __launch_bounds__(256, 1) umma_multi_issuer_kernel(int iters, int nissue, int wait_each, float* sink) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint32_t tmem_ptr;
    __shared__ uint64_t bar[4];
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
    fence_proxy_async_smem();
    if (warp == 0) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp >= 4 && warp < 4 + nissue && elect_one()) {
        const int j = warp - 4;
        const uint32_t idesc = umma_idesc_f16(128, 128);
        const uint64_t da = umma_smem_desc(smem_u32(sm), 2048, 128);
        const uint64_t db = umma_smem_desc(smem_u32(sm + 16384), 2048, 128);
        const uint32_t tacc = tmem_ptr + (uint32_t)(j * 128);
        const int my_iters = iters / nissue;
        for (int it = 0; it < my_iters; ++it) {
            umma_f16(tacc, da, db, idesc, 0);
            umma_f16(tacc, da + 256, db + 256, idesc, 1);
            umma_f16(tacc, da + 512, db + 512, idesc, 1);
            umma_f16(tacc, da + 768, db + 768, idesc, 1);
            if (wait_each) { umma_commit(&bar[j]); mbar_wait(&bar[j], it & 1, 0u); }
        }
        if (!wait_each) { umma_commit(&bar[j]); mbar_wait(&bar[j], 0, 0u); }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_ptr, 512); }
    if (sink == nullptr) sm[0] = 1;
}

// which = 7: functional probe of the two-CTA MMA (tcgen05.mma.cta_group::2, M = 256 over a CTA pair, each CTA holding
// its 128 rows of A and its 64-code half of B): one K-step on known data, every accumulator element checked.
// result_host[0] = number of wrong elements, result_host[1] = elements checked.
__global__ void __launch_bounds__(128, 1) umma_2cta_probe_kernel(float* out /* [256][128] */) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint32_t tmem_ptr;
    __shared__ uint64_t bar;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // A: 128 rows x 16 fp16 (2 planes of 128 x 16 B at sm + 0); row r = 128 * rank + tid: A[r][0] = (r % 7 + 1) / 8
    // B: 64 codes x 16 fp16 (2 planes of 64 x 16 B at sm + 8192); code n = 64 * rank + tid: B[n][0] = (n % 5 + 1) / 4
    for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0u;
    __syncthreads();
    {
        const int r = 128 * (int)rank + (int)threadIdx.x;
        reinterpret_cast<__half*>(sm + threadIdx.x * 16)[0] = __float2half((float)(r % 7 + 1) / 8.0f);
        if (threadIdx.x < 64) {
            const int n = 64 * (int)rank + (int)threadIdx.x;
            reinterpret_cast<__half*>(sm + 8192 + threadIdx.x * 16)[0] = __float2half((float)(n % 5 + 1) / 4.0f);
        }
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    fence_proxy_async_smem();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    tc_fence_after();
    if (rank == 0 && warp == 1 && elect_one()) {
        const uint32_t idesc = umma_idesc_f16(256, 128);
        const uint64_t da = umma_smem_desc(smem_u32(sm), 2048, 128);
        const uint64_t db = umma_smem_desc(smem_u32(sm + 8192), 1024, 128);
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_ptr), "l"(da), "l"(db), "r"(idesc), "r"(0u)
            : "memory");
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                         smem_u32(&bar)),
                     "h"((uint16_t)3)
                     : "memory");
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    {
        // warp w reads TMEM lanes [32 w, 32 w + 32): row = 128 * rank + 32 * w + lane, all 128 columns
        const uint32_t base = tmem_ptr + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < 128; c += 32) {
            float v[32];
            tmem_ld32(base + c, v);
            tmem_ld_wait();
            const int row = 128 * (int)rank + 32 * warp + lane;
#pragma unroll
            for (int i = 0; i < 32; ++i) out[row * 128 + c + i] = v[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_ptr), "r"(128u) : "memory");
    }
}

static cudaError_t run_2cta_probe(double* result_host, cudaStream_t stream) {
    float* out = nullptr;
    cudaError_t e = cudaMalloc(&out, 256 * 128 * sizeof(float));
    if (e != cudaSuccess) return e;
    cudaMemsetAsync(out, 0xff, 256 * 128 * sizeof(float), stream);
    cudaFuncSetAttribute(umma_2cta_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2, 1, 1);
    cfg.blockDim = dim3(128, 1, 1);
    cfg.dynamicSmemBytes = 32768;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, umma_2cta_probe_kernel, out);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) {
        static float host[256 * 128];
        e = cudaMemcpy(host, out, sizeof(host), cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int r = 0; r < 256 && e == cudaSuccess; ++r)
            for (int n = 0; n < 128; ++n) {
                const float want = ((float)(r % 7 + 1) / 8.0f) * ((float)(n % 5 + 1) / 4.0f);
                if (!(host[r * 128 + n] == want)) ++bad;
            }
        result_host[0] = (double)bad;
        result_host[1] = 256.0 * 128.0;
    }
    cudaFree(out);
    return e;
}

cudaError_t run_microbench(int which, int warps, int iters, int num_sms, double* result_host, cudaStream_t stream) {
    if (which == 7) return run_2cta_probe(result_host, stream);
    const int warps_raw = warps;
    if (warps < 4) warps = 4;
    if (warps > 32) warps = 32;
    warps &= ~3;
    float* buf = nullptr;
    cudaError_t e = cudaMalloc(&buf, 4096 * sizeof(float));   // diagnostics only: not on the product path
    if (e != cudaSuccess) return e;
    cudaMemsetAsync(buf, 0, 4096 * sizeof(float), stream);
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(t0, stream);
        if (which == 5 || which == 6) {
            cudaFuncSetAttribute(umma_multi_issuer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
            umma_multi_issuer_kernel<<<num_sms, 256, 65536, stream>>>(iters, warps_raw, which == 6, buf + 1024);
        } else if (which >= 2) {
            cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
            umma_bench_kernel<<<num_sms, which == 2 ? 128 : 640, 65536, stream>>>(iters, warps * 8, which, buf + 1024);
        } else if (which == 0) tmem_ld_bench_kernel<<<num_sms, warps * 32, 0, stream>>>(iters, buf + 1024);
        else fmnmx_bench_kernel<<<num_sms, warps * 32, 0, stream>>>(iters, buf, buf + 1024);
        cudaEventRecord(t1, stream);
    }
    e = cudaEventSynchronize(t1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t0, t1);
    result_host[0] = ms;
    if (which == 5 || which == 6) result_host[1] = (double)num_sms * (iters / warps_raw) * warps_raw * 4.0 * 128.0 * 128.0 * 16.0 * 2.0;
    else if (which >= 2) result_host[1] = (double)num_sms * iters * 4.0 * 128.0 * (warps * 8) * 16.0 * 2.0;   // flops
    else if (which == 0) result_host[1] = (double)num_sms * warps * iters * 16.0 * 32.0 * 32.0 * 4.0;   // bytes
    else result_host[1] = (double)num_sms * warps * 32.0 * iters * 32.0;                               // scores
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    cudaFree(buf);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace mtvq
