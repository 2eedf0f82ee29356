#pragma once
#include <cuda_runtime.h>
namespace mtvq {
cudaError_t run_microbench(int which, int warps, int iters, int num_sms, double* result_host, cudaStream_t stream);
}
