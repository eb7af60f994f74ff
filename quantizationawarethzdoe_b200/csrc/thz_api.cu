// Version, error reporting and host-side FFT planning helpers of the C ABI.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "thz_fft.cuh"
#include "thz_runtime.h"

static thread_local char g_last_error[512] = "";

int thz_set_error(int code, const char* msg) {
    snprintf(g_last_error, sizeof(g_last_error), "%s", msg ? msg : "");
    return code;
}

int thz_set_cuda_error(const char* what, cudaError_t e) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what ? what : "cuda", cudaGetErrorString(e));
    return THZ_E_CUDA;
}

int thz_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    return n;
}

extern "C" int thz_version(void) { return 100; /* 0.1.0 */ }

extern "C" const char* thz_last_error(void) { return g_last_error; }

extern "C" int thz_fft_plan_info(int32_t n, int32_t* radices, int32_t* nstages) {
    if (!radices || !nstages) return thz_set_error(THZ_E_NULL, "thz_fft_plan_info: null pointer");
    FftPlan P;
    if (thz_make_plan(n, &P) != 0) return thz_set_error(THZ_E_UNSUPPORTED, "thz_fft_plan_info: length has a prime factor > 7");
    for (int s = 0; s < THZ_MAX_STAGES; ++s) radices[s] = s < P.ns ? P.radix[s] : 0;
    *nstages = P.ns;
    return THZ_OK;
}

extern "C" int thz_fft_slot_to_bin(int32_t n, int32_t* slot_to_bin) {
    if (!slot_to_bin) return thz_set_error(THZ_E_NULL, "thz_fft_slot_to_bin: null pointer");
    FftPlan P;
    if (thz_make_plan(n, &P) != 0) return thz_set_error(THZ_E_UNSUPPORTED, "thz_fft_slot_to_bin: length has a prime factor > 7");
    for (int p = 0; p < n; ++p) slot_to_bin[p] = thz_pos_to_bin(P, p);
    return THZ_OK;
}

extern "C" int thz_fft_twiddles(int32_t n, float* tw) {
    if (!tw) return thz_set_error(THZ_E_NULL, "thz_fft_twiddles: null pointer");
    if (n < 1) return thz_set_error(THZ_E_SHAPE, "thz_fft_twiddles: n < 1");
    const double w = -2.0 * M_PI / (double)n;
    for (int m = 0; m < n; ++m) {
        tw[2 * m] = (float)cos(w * m);
        tw[2 * m + 1] = (float)sin(w * m);
    }
    return THZ_OK;
}
