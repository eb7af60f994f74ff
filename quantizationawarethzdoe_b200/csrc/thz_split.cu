// Outer decimation step for lines longer than 16384 points + its C ABI (bodies in thz_split.cuh).
#include "thz_split.cuh"
#include "thz_runtime.h"

#include <cstdio>

#define THZ_CHECK_LAUNCH(name)                                         \
    do {                                                               \
        cudaError_t e_ = cudaGetLastError();                           \
        if (e_ != cudaSuccess) return thz_set_cuda_error(name, e_);    \
    } while (0)

static int split_error(int code, const char* who, const char* what) {
    char msg[160];
    snprintf(msg, sizeof msg, "%s: %s", who, what);
    return thz_set_error(code, msg);
}

template <int PR, int PC>
__global__ void __launch_bounds__(256) thz_k_split_pre(SplitArgs A, const cpx* __restrict__ x, cpx* __restrict__ u) {
    const int Mr = A.Hp / PR, Mc = A.Wp / PC;
    const cpx* xf = x + (size_t)blockIdx.z * A.H * A.W;
    cpx* uf = u + (size_t)blockIdx.z * A.Hp * A.Wp;
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= Mc) return;
    for (int n = blockIdx.y; n < Mr; n += gridDim.y) thz_split_pre_point<PR, PC>(A, xf, uf, n, m);
}

template <int PR, int PC>
__global__ void __launch_bounds__(256) thz_k_split_post(SplitArgs A, const cpx* __restrict__ v, cpx* __restrict__ y) {
    const cpx* vf = v + (size_t)blockIdx.z * A.Hp * A.Wp;
    cpx* yf = y + (size_t)blockIdx.z * A.H * A.W;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= A.W) return;
    for (int i = blockIdx.y; i < A.H; i += gridDim.y) yf[(size_t)i * A.W + j] = thz_split_post_point<PR, PC>(A, vf, i, j);
}

static int thz_split_check(const char* who, const void* a, const void* b, int F, int H, int W, int r0, int c0, int Hp, int Wp, int Pr,
                           int Pc, const void* twr, const void* twc) {
    if (F < 0 || H < 0 || W < 0 || Hp < 1 || Wp < 1 || r0 < 0 || c0 < 0 || r0 + H > Hp || c0 + W > Wp)
        return split_error(THZ_E_SHAPE, who, "bad sizes");
    if ((Pr != 1 && Pr != 2 && Pr != 4) || (Pc != 1 && Pc != 2 && Pc != 4) || Hp % Pr || Wp % Pc)
        return split_error(THZ_E_UNSUPPORTED, who, "split factors must be 1, 2 or 4 and divide the canvas");
    if (F > 65535) return split_error(THZ_E_SHAPE, who, "more than 65535 fields per call");
    if (F == 0) return THZ_OK;
    if (!a || !b || !twr || !twc) return split_error(THZ_E_NULL, who, "null pointer");
    return THZ_OK;
}

#define THZ_SPLIT_DISPATCH(KERNEL, ...)                                                          \
    switch (Pr * 8 + Pc) {                                                                       \
        case 1 * 8 + 1: KERNEL<1, 1><<<grid, 256, 0, stream>>>(__VA_ARGS__); break;              \
        case 1 * 8 + 2: KERNEL<1, 2><<<grid, 256, 0, stream>>>(__VA_ARGS__); break;              \
        case 1 * 8 + 4: KERNEL<1, 4><<<grid, 256, 0, stream>>>(__VA_ARGS__); break;              \
        case 2 * 8 + 1: KERNEL<2, 1><<<grid, 256, 0, stream>>>(__VA_ARGS__); break;              \
        case 2 * 8 + 2: KERNEL<2, 2><<<grid, 256, 0, stream>>>(__VA_ARGS__); break;              \
        case 2 * 8 + 4: KERNEL<2, 4><<<grid, 256, 0, stream>>>(__VA_ARGS__); break;              \
        case 4 * 8 + 1: KERNEL<4, 1><<<grid, 256, 0, stream>>>(__VA_ARGS__); break;              \
        case 4 * 8 + 2: KERNEL<4, 2><<<grid, 256, 0, stream>>>(__VA_ARGS__); break;              \
        default: KERNEL<4, 4><<<grid, 256, 0, stream>>>(__VA_ARGS__); break;                     \
    }

extern "C" int thz_split_pre(const void* x, void* u, int32_t F, int32_t H, int32_t W, int32_t r0, int32_t c0, int32_t Hp, int32_t Wp,
                             int32_t Pr, int32_t Pc, const void* tw_r, const void* tw_c, int32_t conj_tw, float scale, void* stream_) {
    ThzDeviceGuard dev_guard(u);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int rc = thz_split_check("thz_split_pre", x, u, F, H, W, r0, c0, Hp, Wp, Pr, Pc, tw_r, tw_c);
    if (rc != THZ_OK || F == 0) return rc;
    const SplitArgs A = {Hp, Wp, H, W, r0, c0, (const cpx*)tw_r, (const cpx*)tw_c, conj_tw ? 1 : 0, scale};
    const int Mr = Hp / Pr, Mc = Wp / Pc;
    dim3 grid((Mc + 255) / 256, Mr < 32768 ? Mr : 32768, F);
    thz_launch_begin(stream, THZ_KC_DOE);
    THZ_SPLIT_DISPATCH(thz_k_split_pre, A, (const cpx*)x, (cpx*)u)
    thz_launch_end(stream, THZ_KC_DOE);
    THZ_CHECK_LAUNCH("thz_k_split_pre");
    return THZ_OK;
}

extern "C" int thz_split_post(const void* v, void* y, int32_t F, int32_t H, int32_t W, int32_t r0, int32_t c0, int32_t Hp, int32_t Wp,
                              int32_t Pr, int32_t Pc, const void* tw_r, const void* tw_c, int32_t conj_tw, float scale, void* stream_) {
    ThzDeviceGuard dev_guard(y);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int rc = thz_split_check("thz_split_post", v, y, F, H, W, r0, c0, Hp, Wp, Pr, Pc, tw_r, tw_c);
    if (rc != THZ_OK || F == 0 || H == 0 || W == 0) return rc;
    const SplitArgs A = {Hp, Wp, H, W, r0, c0, (const cpx*)tw_r, (const cpx*)tw_c, conj_tw ? 1 : 0, scale};
    dim3 grid((W + 255) / 256, H < 32768 ? H : 32768, F);
    thz_launch_begin(stream, THZ_KC_DOE);
    THZ_SPLIT_DISPATCH(thz_k_split_post, A, (const cpx*)v, (cpx*)y)
    thz_launch_end(stream, THZ_KC_DOE);
    THZ_CHECK_LAUNCH("thz_k_split_post");
    return THZ_OK;
}
