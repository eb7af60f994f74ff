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
// THZ_KC_CZT = the CUDA-core Toeplitz GEMM (+ the prologue multiply), THZ_KC_CZT_TC = the tcgen05 kernel: separate classes so
// that a caller can tell which implementation produced a result (thz_launch_count_class).
// THZ_KC_COL_TMA is a side counter (thz_launch_note): launches of the kernel variants that move their output / input with the
// TMA (thz_p2_k1t and thz_p2_k2ft: tensor stores; thz_p2_k3t: bulk-copy staging of the permuted intermediate); those
// launches are counted and timed under their own class (0 / 1 / 2) like every other launch.
enum { THZ_KC_ROW_FWD = 0, THZ_KC_COL = 1, THZ_KC_ROW_INV = 2, THZ_KC_FFT2_COL = 3, THZ_KC_DOE = 4, THZ_KC_QUANT = 5,
       THZ_KC_CZT = 6, THZ_KC_TRAIN = 7, THZ_KC_CZT_TC = 8, THZ_KC_COL_TMA = 9, THZ_KC_COUNT = 10 };
// Called around every kernel launch: counts it and, when profiling is enabled, brackets it with CUDA events
// recorded on the launching stream.
void thz_launch_begin(cudaStream_t stream, int kernel_class);
void thz_launch_end(cudaStream_t stream, int kernel_class);
// Bumps the per-class counter only (no total, no timing): side counters such as THZ_KC_COL_TMA.
void thz_launch_note(int kernel_class);

// Makes the device that owns `device_ptr` current for the lifetime of the guard (restored afterwards): every entry point
// that launches work constructs one from its first device pointer, so a caller whose current device differs from the
// tensors' device (ASM_prop(device='cuda:1') under current device 0) gets correct launches, attributes and SM counts.
struct ThzDeviceGuard {
    int prev;
    bool switched;
    explicit ThzDeviceGuard(const void* device_ptr);
    ~ThzDeviceGuard();
    ThzDeviceGuard(const ThzDeviceGuard&) = delete;
    ThzDeviceGuard& operator=(const ThzDeviceGuard&) = delete;
};
