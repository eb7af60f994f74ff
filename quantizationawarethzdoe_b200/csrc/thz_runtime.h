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

// Kernel classes for the launch counter / optional per-kernel event timing (thz_profile_*).
enum { THZ_KC_ROW_FWD = 0, THZ_KC_COL = 1, THZ_KC_ROW_INV = 2, THZ_KC_FFT2_COL = 3, THZ_KC_DOE = 4, THZ_KC_QUANT = 5,
       THZ_KC_CZT = 6, THZ_KC_TRAIN = 7, THZ_KC_COUNT = 8 };
// Called around every kernel launch: counts it and, when profiling is enabled, brackets it with CUDA events
// recorded on the launching stream.
void thz_launch_begin(cudaStream_t stream, int kernel_class);
void thz_launch_end(cudaStream_t stream, int kernel_class);
