#pragma once
#include <cuda_runtime.h>
namespace mtvq {
cudaError_t run_microbench(int which, int warps, int iters, int num_sms, double* result_host, cudaStream_t stream);
}

// Diagnostic entry point, exported only by the `probe` build (python -m medtok_b200.build --variant probe
// -DMTVQ_WITH_MICROBENCH; tools_gpu_probe.py loads that library).  Not part of the product ABI (include/medtok_vq.h).
//   which = 0: tcgen05.ld throughput, 1: FMNMX3 tree throughput, 2-7: raw tcgen05.mma issue experiments;
//   result_host[0] = device ms, result_host[1] = bytes or ops processed
extern "C" __attribute__((visibility("default"))) int mtvq_microbench(int which, int warps, int iters, double* result_host,
                                                                       void* stream);
