// Error reporting and small runtime helpers shared by the .cu translation units.
#pragma once
#include <cuda_runtime.h>

#include "../../include/thzdoe.h"

// Records `msg` as this thread's last error and returns `code`.
int thz_set_error(int code, const char* msg);
// Records "<what>: <cudaGetErrorString(e)>" and returns THZ_E_CUDA.
int thz_set_cuda_error(const char* what, cudaError_t e);
// Number of SMs of the current device (148 on B200); 148 if the query fails.
int thz_sm_count(void);
