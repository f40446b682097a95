// hawk_kernels.h -- internal declarations shared by the .cu files of libhawkscan
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hawkscan.h"


int hawk_fail(int code, const char* fmt, ...);
int hawk_check_cuda(cudaError_t err, const char* what);

// every kernel launch of the library is counted (bench.py reports it as gpu_launches)
void hawk_note_launch(int n);

// optional CUDA-event brackets around single kernels of the device layer (active while the
// host layer runs with hawk_ctx_set_profiling on; no-ops otherwise)
void hawk_prof_begin(cudaStream_t st, int kind);
void hawk_prof_end(cudaStream_t st);
