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
            acc += v[it & 31];      // one dependent use so the load cannot be dropped
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
        v[it & 31] = __int_as_float(__float_as_int(m) ^ it);   // perturb so iterations are not hoisted
    }
    if (best == 123.456f) sink[threadIdx.x] = best;
}

cudaError_t run_microbench(int which, int warps, int iters, int num_sms, double* result_host, cudaStream_t stream) {
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
        if (which == 0) tmem_ld_bench_kernel<<<num_sms, warps * 32, 0, stream>>>(iters, buf + 1024);
        else fmnmx_bench_kernel<<<num_sms, warps * 32, 0, stream>>>(iters, buf, buf + 1024);
        cudaEventRecord(t1, stream);
    }
    e = cudaEventSynchronize(t1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t0, t1);
    result_host[0] = ms;
    if (which == 0) result_host[1] = (double)num_sms * warps * iters * 16.0 * 32.0 * 32.0 * 4.0;   // bytes
    else result_host[1] = (double)num_sms * warps * 32.0 * iters * 32.0;                               // scores
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    cudaFree(buf);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace mtvq
