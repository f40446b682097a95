// hawk_kernels.h -- internal declarations shared by the .cu files of libhawkscan
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hawkscan.h"

// chunks per scan span (the unit one warp processes at a time): 1,024 chunks = 32,768 base
// slots = one 32-bit slice of the nz summary per lane
#define HAWK_SPAN_CHUNKS 1024

int hawk_fail(int code, const char* fmt, ...);
int hawk_check_cuda(cudaError_t err, const char* what);

// every kernel launch of the library is counted (bench.py reports it as gpu_launches)
void hawk_note_launch(int n);
