// DOE phase modulation and quantizer kernels + their C ABI (bodies in thz_doe.cuh).
#include "thz_doe.cuh"
#include "thz_runtime.h"

// ------------------------------------------------------------------------------- modulation
// y[b,c,p] = x[b,c,p] * p_c(h[p])            (Components/QuantizedDOE.py:113-118)
__global__ void __launch_bounds__(256) thz_k_doe_fwd(const cpx* __restrict__ x, cpx* __restrict__ y,
                                                     const float* __restrict__ hmap, const float4* __restrict__ coef,
                                                     float base, int BC, int C, size_t HW) {
    const size_t total = (size_t)BC * HW;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t f = i / HW, p = i - f * HW;
        const float4 cf = __ldg(coef + (f % C));
        y[i] = cmul(x[i], thz_doe_phase(__ldg(hmap + p), cf, base));
    }
}

// gx = g conj(p);  gh[p] = sum_{b,c} Re(conj(g) x p gamma_c)   -- one thread per pixel, fields walked in order
__global__ void __launch_bounds__(256) thz_k_doe_bwd(const cpx* __restrict__ g, const cpx* __restrict__ x,
                                                     const float* __restrict__ hmap, const float4* __restrict__ coef,
                                                     float base, cpx* __restrict__ gx, float* __restrict__ gh, int BC,
                                                     int C, size_t HW) {
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (size_t)gridDim.x * blockDim.x) {
        const float h = __ldg(hmap + p);
        float acc = 0.f;
        for (int f = 0; f < BC; ++f) {
            const float4 cf = __ldg(coef + (f % C));
            const cpx pp = thz_doe_phase(h, cf, base);
            const cpx gamma = cmake(-cf.x * (0.5f * cf.y * cf.z), -cf.x * cf.w);
            const size_t i = (size_t)f * HW + p;
            const cpx gv = g[i];
            if (gx) gx[i] = cmulc(gv, pp);
            if (gh) {
                const cpx xp = cmul(cmul(x[i], pp), gamma);
                acc += gv.x * xp.x + gv.y * xp.y;
            }
        }
        if (gh) gh[p] = acc;
    }
}

// ------------------------------------------------------------------------------- quantizers
__global__ void __launch_bounds__(256) thz_k_ste_fwd(const float* __restrict__ in, int from_weights, float hmax, float clampv,
                                                     const float* __restrict__ lut_g, int L, float* __restrict__ q,
                                                     int32_t* __restrict__ idx, float* __restrict__ h_pre, size_t n) {
    __shared__ float lut[THZ_MAX_LEVELS];
    for (int j = threadIdx.x; j < L; j += blockDim.x) lut[j] = lut_g[j];
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = in[i];
        const float h = from_weights ? thz_height_from_weight(v, hmax, clampv) : v;
        const int k = thz_ste_index(h, lut, L);
        q[i] = lut[k];
        if (idx) idx[i] = k;
        if (h_pre) h_pre[i] = h;
    }
}

__global__ void __launch_bounds__(256) thz_k_height_fwd(const float* __restrict__ w, float hmax, float clampv,
                                                        float* __restrict__ h, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        h[i] = thz_height_from_weight(w[i], hmax, clampv);
}

// gw = g * d h / d w  (sigmoid-of-clamp chain behind the straight-through quantizer)
__global__ void __launch_bounds__(256) thz_k_height_bwd(const float* __restrict__ g, const float* __restrict__ w, float hmax,
                                                        float clampv, float* __restrict__ gw, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        gw[i] = g[i] * thz_height_grad(w[i], hmax, clampv);
}

__global__ void __launch_bounds__(256) thz_k_nn_fwd(const float* __restrict__ x, const float* __restrict__ lut_g, int nlut,
                                                    const float* __restrict__ mid_g, int nmid, float* __restrict__ q,
                                                    int32_t* __restrict__ idx, size_t n) {
    __shared__ float lut[THZ_MAX_LEVELS + 1], mid[THZ_MAX_LEVELS];
    for (int j = threadIdx.x; j < nlut; j += blockDim.x) lut[j] = lut_g[j];
    for (int j = threadIdx.x; j < nmid; j += blockDim.x) mid[j] = mid_g[j];
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int k = thz_nn_index(x[i], mid, nmid);
        q[i] = lut[k];
        idx[i] = k;
    }
}

__global__ void __launch_bounds__(256) thz_k_nn_bwd(const float* __restrict__ g, const float* __restrict__ x,
                                                    const int32_t* __restrict__ idx, const float* __restrict__ lut_g, int nlut,
                                                    float s, int kind, float* __restrict__ gx, size_t n) {
    __shared__ float lut[THZ_MAX_LEVELS + 1];
    for (int j = threadIdx.x; j < nlut; j += blockDim.x) lut[j] = lut_g[j];
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int k = idx[i];
        gx[i] = g[i] * thz_nn_grad_factor(x[i], lut[k], k, lut, nlut, s, kind);
    }
}

__global__ void __launch_bounds__(256) thz_k_psq_fwd(const float* __restrict__ w, float hmax, int L, float tau,
                                                     float* __restrict__ out, float* __restrict__ dout_dw, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float o, d;
        thz_psq(w[i], hmax, L, tau, &o, &d);
        out[i] = o;
        if (dout_dw) dout_dw[i] = d;
    }
}

__global__ void __launch_bounds__(256) thz_k_gumbel_v3_fwd(const float* __restrict__ w, const float* __restrict__ lut_g,
                                                           const float* __restrict__ noise, GumbelV3Params P,
                                                           float* __restrict__ h_out, int32_t* __restrict__ idx,
                                                           float* __restrict__ dh_dw, size_t n) {
    __shared__ float lut[THZ_MAX_LEVELS];
    for (int j = threadIdx.x; j < P.L; j += blockDim.x) lut[j] = lut_g[j];
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float ho, dd;
        const int k = thz_gumbel_v3_pixel(w[i], lut, noise + i, n, P, &ho, dh_dw ? &dd : nullptr);
        h_out[i] = ho;
        if (idx) idx[i] = k;
        if (dh_dw) dh_dw[i] = dd;
    }
}

// logits, noise: [n, L] (level index fastest, QuantizedDOE.py:1010);  dq: [n, L] = d q / d logit
__global__ void __launch_bounds__(256) thz_k_gumbel_naive_fwd(const float* __restrict__ logits, const float* __restrict__ noise,
                                                              const float* __restrict__ lut_g, int L, float tau,
                                                              float* __restrict__ q, int32_t* __restrict__ idx,
                                                              float* __restrict__ dq, size_t n) {
    __shared__ float lut[THZ_MAX_LEVELS];
    for (int j = threadIdx.x; j < L; j += blockDim.x) lut[j] = lut_g[j];
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float lg[THZ_MAX_LEVELS], y[THZ_MAX_LEVELS];
        for (int j = 0; j < L; ++j) lg[j] = (logits[i * L + j] + noise[i * L + j]) / tau;
        float qq, sm;
        const int k = thz_gumbel_hard(lg, lut, L, y, &qq, &sm);
        q[i] = qq;
        if (idx) idx[i] = k;
        if (dq)
            for (int j = 0; j < L; ++j) dq[i * L + j] = y[j] * (lut[j] - sm) / tau;
    }
}

// ------------------------------------------------------------------------------- C ABI
static inline int grid_for(size_t n) {
    size_t b = (n + 255) / 256;
    const size_t cap = (size_t)thz_sm_count() * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
#define THZ_CHECK_LAUNCH(name)                                         \
    do {                                                               \
        cudaError_t e_ = cudaGetLastError();                           \
        if (e_ != cudaSuccess) return thz_set_cuda_error(name, e_);    \
    } while (0)

extern "C" int thz_doe_modulate_fwd(const void* x, void* y, const void* hmap, const void* coef, float base, int32_t B,
                                    int32_t C, int32_t H, int32_t W, void* stream) {
    ThzDeviceGuard dev_guard(y);
    if (!x || !y || !hmap || !coef) return thz_set_error(THZ_E_NULL, "thz_doe_modulate_fwd: null pointer");
    if (B < 1 || C < 1 || H < 1 || W < 1) return thz_set_error(THZ_E_SHAPE, "thz_doe_modulate_fwd: bad shape");
    const size_t HW = (size_t)H * W;
    thz_launch_begin((cudaStream_t)stream, THZ_KC_DOE);
    thz_k_doe_fwd<<<grid_for((size_t)B * C * HW), 256, 0, (cudaStream_t)stream>>>(
        (const cpx*)x, (cpx*)y, (const float*)hmap, (const float4*)coef, base, B * C, C, HW);
    thz_launch_end((cudaStream_t)stream, THZ_KC_DOE);
    THZ_CHECK_LAUNCH("thz_k_doe_fwd");
    return THZ_OK;
}

extern "C" int thz_doe_modulate_bwd(const void* g, const void* x, const void* hmap, const void* coef, float base, void* gx,
                                    void* gh, int32_t B, int32_t C, int32_t H, int32_t W, void* stream) {
    ThzDeviceGuard dev_guard(g);
    if (!g || !hmap || !coef) return thz_set_error(THZ_E_NULL, "thz_doe_modulate_bwd: null pointer");
    if (gh && !x) return thz_set_error(THZ_E_NULL, "thz_doe_modulate_bwd: grad_height needs the saved input field");
    if (B < 1 || C < 1 || H < 1 || W < 1) return thz_set_error(THZ_E_SHAPE, "thz_doe_modulate_bwd: bad shape");
    const size_t HW = (size_t)H * W;
    thz_launch_begin((cudaStream_t)stream, THZ_KC_DOE);
    thz_k_doe_bwd<<<grid_for(HW), 256, 0, (cudaStream_t)stream>>>((const cpx*)g, (const cpx*)x, (const float*)hmap,
                                                                  (const float4*)coef, base, (cpx*)gx, (float*)gh, B * C, C, HW);
    thz_launch_end((cudaStream_t)stream, THZ_KC_DOE);
    THZ_CHECK_LAUNCH("thz_k_doe_bwd");
    return THZ_OK;
}

extern "C" int thz_quant_ste_fwd(const void* in, int32_t from_weights, float hmax, float clampv, const void* lut, int32_t L,
                                 void* q, void* idx, void* h_pre, uint64_t n, void* stream) {
    ThzDeviceGuard dev_guard(in);
    if (n == 0) return THZ_OK;   // empty maps are legal (and have a NULL data pointer)
    if (!in || !lut || !q) return thz_set_error(THZ_E_NULL, "thz_quant_ste_fwd: null pointer");
    if (L < 1 || L > THZ_MAX_LEVELS) return thz_set_error(THZ_E_SHAPE, "thz_quant_ste_fwd: 1 <= levels <= 64");
    thz_launch_begin((cudaStream_t)stream, THZ_KC_QUANT);
    thz_k_ste_fwd<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>((const float*)in, from_weights, hmax, clampv, (const float*)lut,
                                                                 L, (float*)q, (int32_t*)idx, (float*)h_pre, n);
    thz_launch_end((cudaStream_t)stream, THZ_KC_QUANT);
    THZ_CHECK_LAUNCH("thz_k_ste_fwd");
    return THZ_OK;
}

extern "C" int thz_height_fwd(const void* w, float hmax, float clampv, void* h, uint64_t n, void* stream) {
    ThzDeviceGuard dev_guard(w);
    if (n == 0) return THZ_OK;   // empty maps are legal (and have a NULL data pointer)
    if (!w || !h) return thz_set_error(THZ_E_NULL, "thz_height_fwd: null pointer");
    thz_launch_begin((cudaStream_t)stream, THZ_KC_QUANT);
    thz_k_height_fwd<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>((const float*)w, hmax, clampv, (float*)h, n);
    thz_launch_end((cudaStream_t)stream, THZ_KC_QUANT);
    THZ_CHECK_LAUNCH("thz_k_height_fwd");
    return THZ_OK;
}

extern "C" int thz_height_bwd(const void* g, const void* w, float hmax, float clampv, void* gw, uint64_t n, void* stream) {
    ThzDeviceGuard dev_guard(g);
    if (n == 0) return THZ_OK;   // empty maps are legal (and have a NULL data pointer)
    if (!g || !w || !gw) return thz_set_error(THZ_E_NULL, "thz_height_bwd: null pointer");
    thz_launch_begin((cudaStream_t)stream, THZ_KC_QUANT);
    thz_k_height_bwd<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>((const float*)g, (const float*)w, hmax, clampv, (float*)gw, n);
    thz_launch_end((cudaStream_t)stream, THZ_KC_QUANT);
    THZ_CHECK_LAUNCH("thz_k_height_bwd");
    return THZ_OK;
}

extern "C" int thz_quant_nn_fwd(const void* x, const void* lut, int32_t nlut, const void* mid, int32_t nmid, void* q, void* idx,
                                uint64_t n, void* stream) {
    ThzDeviceGuard dev_guard(x);
    if (n == 0) return THZ_OK;   // empty maps are legal (and have a NULL data pointer)
    if (!x || !lut || !mid || !q || !idx) return thz_set_error(THZ_E_NULL, "thz_quant_nn_fwd: null pointer");
    if (nlut < 1 || nlut > THZ_MAX_LEVELS + 1 || nmid < 1 || nmid > THZ_MAX_LEVELS || nmid >= nlut + 1)
        return thz_set_error(THZ_E_SHAPE, "thz_quant_nn_fwd: bad lut sizes");
    thz_launch_begin((cudaStream_t)stream, THZ_KC_QUANT);
    thz_k_nn_fwd<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>((const float*)x, (const float*)lut, nlut, (const float*)mid, nmid,
                                                                (float*)q, (int32_t*)idx, n);
    thz_launch_end((cudaStream_t)stream, THZ_KC_QUANT);
    THZ_CHECK_LAUNCH("thz_k_nn_fwd");
    return THZ_OK;
}

extern "C" int thz_quant_nn_bwd(const void* g, const void* x, const void* idx, const void* lut, int32_t nlut, float s,
                                int32_t kind, void* gx, uint64_t n, void* stream) {
    ThzDeviceGuard dev_guard(g);
    if (n == 0) return THZ_OK;   // empty maps are legal (and have a NULL data pointer)
    if (!g || !x || !idx || !lut || !gx) return thz_set_error(THZ_E_NULL, "thz_quant_nn_bwd: null pointer");
    if (nlut < 1 || nlut > THZ_MAX_LEVELS + 1 || kind < 0 || kind > 2) return thz_set_error(THZ_E_SHAPE, "thz_quant_nn_bwd: bad arguments");
    thz_launch_begin((cudaStream_t)stream, THZ_KC_QUANT);
    thz_k_nn_bwd<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>((const float*)g, (const float*)x, (const int32_t*)idx,
                                                                (const float*)lut, nlut, s, kind, (float*)gx, n);
    thz_launch_end((cudaStream_t)stream, THZ_KC_QUANT);
    THZ_CHECK_LAUNCH("thz_k_nn_bwd");
    return THZ_OK;
}

extern "C" int thz_quant_psq_fwd(const void* w, float hmax, int32_t L, float tau, void* out, void* dout_dw, uint64_t n,
                                 void* stream) {
    ThzDeviceGuard dev_guard(w);
    if (n == 0) return THZ_OK;   // empty maps are legal (and have a NULL data pointer)
    if (!w || !out) return thz_set_error(THZ_E_NULL, "thz_quant_psq_fwd: null pointer");
    if (L < 2) return thz_set_error(THZ_E_SHAPE, "thz_quant_psq_fwd: levels >= 2");
    thz_launch_begin((cudaStream_t)stream, THZ_KC_QUANT);
    thz_k_psq_fwd<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>((const float*)w, hmax, L, tau, (float*)out, (float*)dout_dw, n);
    thz_launch_end((cudaStream_t)stream, THZ_KC_QUANT);
    THZ_CHECK_LAUNCH("thz_k_psq_fwd");
    return THZ_OK;
}

extern "C" int thz_quant_gumbel_v3_fwd(const void* w, const void* lut, int32_t L, const void* noise, float hmax, float kfac,
                                       float c_s, float tau, float tau_max, float s, float beta, float one_minus_beta,
                                       int32_t phase_input, void* h_out, void* idx, void* dh_dw, uint64_t n, void* stream) {
    ThzDeviceGuard dev_guard(w);
    if (n == 0) return THZ_OK;   // empty maps are legal (and have a NULL data pointer)
    if (!w || !lut || !noise || !h_out) return thz_set_error(THZ_E_NULL, "thz_quant_gumbel_v3_fwd: null pointer");
    if (L < 1 || L > THZ_MAX_LEVELS) return thz_set_error(THZ_E_SHAPE, "thz_quant_gumbel_v3_fwd: 1 <= levels <= 64");
    GumbelV3Params P;
    P.hmax = hmax;
    P.kfac = kfac;
    P.c_s = c_s;
    P.tau = tau;
    P.tau_max = tau_max;
    P.s = s;
    P.beta = beta;
    P.omb = one_minus_beta;
    P.L = L;
    P.phase_input = phase_input;
    thz_launch_begin((cudaStream_t)stream, THZ_KC_QUANT);
    thz_k_gumbel_v3_fwd<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>((const float*)w, (const float*)lut, (const float*)noise, P,
                                                                       (float*)h_out, (int32_t*)idx, (float*)dh_dw, n);
    thz_launch_end((cudaStream_t)stream, THZ_KC_QUANT);
    THZ_CHECK_LAUNCH("thz_k_gumbel_v3_fwd");
    return THZ_OK;
}

extern "C" int thz_quant_gumbel_naive_fwd(const void* logits, const void* noise, const void* lut, int32_t L, float tau, void* q,
                                          void* idx, void* dq, uint64_t n, void* stream) {
    ThzDeviceGuard dev_guard(logits);
    if (n == 0) return THZ_OK;   // empty maps are legal (and have a NULL data pointer)
    if (!logits || !noise || !lut || !q) return thz_set_error(THZ_E_NULL, "thz_quant_gumbel_naive_fwd: null pointer");
    if (L < 1 || L > THZ_MAX_LEVELS) return thz_set_error(THZ_E_SHAPE, "thz_quant_gumbel_naive_fwd: 1 <= levels <= 64");
    thz_launch_begin((cudaStream_t)stream, THZ_KC_QUANT);
    thz_k_gumbel_naive_fwd<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>((const float*)logits, (const float*)noise,
                                                                          (const float*)lut, L, tau, (float*)q, (int32_t*)idx,
                                                                          (float*)dq, n);
    thz_launch_end((cudaStream_t)stream, THZ_KC_QUANT);
    THZ_CHECK_LAUNCH("thz_k_gumbel_naive_fwd");
    return THZ_OK;
}

// ------------------------------------------------------------------------------- thickness-space softmax quantization
// stats (device float[4], caller-owned): [0] m = max |t - lut_j| over the map and the levels (bit pattern, via atomicMax on
// the non-negative float's integer image), [1] number of (pixel, level) pairs attaining it, [2] sum_pixels g dq/dm (backward).
__global__ void __launch_bounds__(256) thz_k_softmaxq_absmax(const float* __restrict__ t, const float* __restrict__ lut_g, int L,
                                                             float* __restrict__ stats, size_t n) {
    __shared__ float lut[THZ_MAX_LEVELS];
    __shared__ float red[8];
    for (int j = threadIdx.x; j < L; j += blockDim.x) lut[j] = lut_g[j];
    __syncthreads();
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = t[i];
        for (int j = 0; j < L; ++j) m = fmaxf(m, fabsf(thz_sub_rn(v, lut[j])));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) m = fmaxf(m, red[k]);
        atomicMax(reinterpret_cast<int*>(stats), __float_as_int(m));
    }
}

__global__ void __launch_bounds__(256) thz_k_softmaxq_fwd(const float* __restrict__ t, const float* __restrict__ lut_g,
                                                          const float* __restrict__ noise, SoftmaxQParams P,
                                                          float* __restrict__ stats, float* __restrict__ q,
                                                          int32_t* __restrict__ idx, float* __restrict__ A, float* __restrict__ Bm,
                                                          float* __restrict__ E, size_t n) {
    __shared__ float lut[THZ_MAX_LEVELS];
    for (int j = threadIdx.x; j < P.L; j += blockDim.x) lut[j] = lut_g[j];
    __syncthreads();
    P.m = stats[0];
    int ties = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float qq, a, b, e;
        int nt;
        const int k = thz_softmaxq_pixel(t[i], lut, noise ? noise + i : nullptr, n, P, &qq, A ? &a : nullptr, &b, &e, &nt);
        q[i] = qq;
        if (idx) idx[i] = k;
        if (A) {
            A[i] = a;
            Bm[i] = b;
            E[i] = e;
        }
        ties += nt;
    }
    if (ties) atomicAdd(stats + 1, (float)ties);
}

__global__ void __launch_bounds__(256) thz_k_softmaxq_bwd_reduce(const float* __restrict__ g, const float* __restrict__ Bm,
                                                                 float* __restrict__ stats, size_t n) {
    __shared__ float red[8];
    float acc = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) acc += g[i] * Bm[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) acc += red[k];
        atomicAdd(stats + 2, acc);
    }
}

__global__ void __launch_bounds__(256) thz_k_softmaxq_bwd(const float* __restrict__ g, const float* __restrict__ A,
                                                          const float* __restrict__ E, const float* __restrict__ stats,
                                                          float* __restrict__ gt, size_t n) {
    const float share = stats[1] > 0.f ? stats[2] / stats[1] : 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        gt[i] = g[i] * A[i] + share * E[i];
}

__global__ void __launch_bounds__(256) thz_k_score_thickness(const float* __restrict__ t, const float* __restrict__ lut_g, int L,
                                                             float s, int func, const float* __restrict__ stats,
                                                             float* __restrict__ scores, size_t n_per_b, size_t total) {
    __shared__ float lut[THZ_MAX_LEVELS];
    for (int j = threadIdx.x; j < L; j += blockDim.x) lut[j] = lut_g[j];
    __syncthreads();
    const float m = stats[0];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / n_per_b, p = i - b * n_per_b;
        const float v = t[i];
        for (int j = 0; j < L; ++j) scores[(b * L + j) * n_per_b + p] = thz_score_value(thz_sub_rn(v, lut[j]) / m, s, func);
    }
}

static int softmaxq_absmax(const void* t, const void* lut, int L, void* stats, uint64_t n, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(stats, 0, 4 * sizeof(float), stream);
    if (e != cudaSuccess) return thz_set_cuda_error("cudaMemsetAsync(stats)", e);
    thz_launch_begin(stream, THZ_KC_QUANT);
    thz_k_softmaxq_absmax<<<grid_for(n), 256, 0, stream>>>((const float*)t, (const float*)lut, L, (float*)stats, n);
    thz_launch_end(stream, THZ_KC_QUANT);
    THZ_CHECK_LAUNCH("thz_k_softmaxq_absmax");
    return THZ_OK;
}

extern "C" int thz_quant_softmax_fwd(const void* thickness, const void* lut, int32_t L, const void* noise, float c, float tau,
                                     float s, int32_t hard, void* q, void* idx, void* dq_dt, void* dq_dm, void* tie_sign,
                                     void* stats, uint64_t n, void* stream_) {
    ThzDeviceGuard dev_guard(thickness);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) return THZ_OK;
    if (!thickness || !lut || !q || !stats) return thz_set_error(THZ_E_NULL, "thz_quant_softmax_fwd: null pointer");
    if (dq_dt && (!dq_dm || !tie_sign)) return thz_set_error(THZ_E_NULL, "thz_quant_softmax_fwd: dq_dt needs dq_dm and tie_sign");
    if (L < 1 || L > THZ_MAX_LEVELS) return thz_set_error(THZ_E_SHAPE, "thz_quant_softmax_fwd: 1 <= levels <= 64");
    int rc = softmaxq_absmax(thickness, lut, L, stats, n, stream);
    if (rc != THZ_OK) return rc;
    SoftmaxQParams P;
    P.m = 0.f;
    P.s = s;
    P.c = c;
    P.tau = tau;
    P.L = L;
    P.hard = hard ? 1 : 0;
    P.gumbel = noise ? 1 : 0;
    thz_launch_begin(stream, THZ_KC_QUANT);
    thz_k_softmaxq_fwd<<<grid_for(n), 256, 0, stream>>>((const float*)thickness, (const float*)lut, (const float*)noise, P,
                                                        (float*)stats, (float*)q, (int32_t*)idx, (float*)dq_dt, (float*)dq_dm,
                                                        (float*)tie_sign, n);
    thz_launch_end(stream, THZ_KC_QUANT);
    THZ_CHECK_LAUNCH("thz_k_softmaxq_fwd");
    return THZ_OK;
}

extern "C" int thz_quant_softmax_bwd(const void* g, const void* dq_dt, const void* dq_dm, const void* tie_sign, void* stats,
                                     void* gt, uint64_t n, void* stream_) {
    ThzDeviceGuard dev_guard(g);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) return THZ_OK;
    if (!g || !dq_dt || !dq_dm || !tie_sign || !stats || !gt) return thz_set_error(THZ_E_NULL, "thz_quant_softmax_bwd: null pointer");
    cudaError_t e = cudaMemsetAsync((float*)stats + 2, 0, sizeof(float), stream);
    if (e != cudaSuccess) return thz_set_cuda_error("cudaMemsetAsync(stats)", e);
    thz_launch_begin(stream, THZ_KC_QUANT);
    thz_k_softmaxq_bwd_reduce<<<grid_for(n), 256, 0, stream>>>((const float*)g, (const float*)dq_dm, (float*)stats, n);
    thz_launch_end(stream, THZ_KC_QUANT);
    THZ_CHECK_LAUNCH("thz_k_softmaxq_bwd_reduce");
    thz_launch_begin(stream, THZ_KC_QUANT);
    thz_k_softmaxq_bwd<<<grid_for(n), 256, 0, stream>>>((const float*)g, (const float*)dq_dt, (const float*)tie_sign,
                                                        (const float*)stats, (float*)gt, n);
    thz_launch_end(stream, THZ_KC_QUANT);
    THZ_CHECK_LAUNCH("thz_k_softmaxq_bwd");
    return THZ_OK;
}

extern "C" int thz_score_thickness(const void* thickness, const void* lut, int32_t L, float s, int32_t func, void* scores,
                                   void* stats, int32_t batch, uint64_t n_per_b, void* stream_) {
    ThzDeviceGuard dev_guard(thickness);
    cudaStream_t stream = (cudaStream_t)stream_;
    const uint64_t total = (uint64_t)batch * n_per_b;
    if (batch < 0) return thz_set_error(THZ_E_SHAPE, "thz_score_thickness: bad batch");
    if (total == 0) return THZ_OK;
    if (!thickness || !lut || !scores || !stats) return thz_set_error(THZ_E_NULL, "thz_score_thickness: null pointer");
    if (L < 1 || L > THZ_MAX_LEVELS || func < 0 || func > 4) return thz_set_error(THZ_E_SHAPE, "thz_score_thickness: bad arguments");
    int rc = softmaxq_absmax(thickness, lut, L, stats, total, stream);
    if (rc != THZ_OK) return rc;
    thz_launch_begin(stream, THZ_KC_QUANT);
    thz_k_score_thickness<<<grid_for(total), 256, 0, stream>>>((const float*)thickness, (const float*)lut, L, s, func,
                                                               (const float*)stats, (float*)scores, n_per_b, total);
    thz_launch_end(stream, THZ_KC_QUANT);
    THZ_CHECK_LAUNCH("thz_k_score_thickness");
    return THZ_OK;
}

// ------------------------------------------------------------------------------- cached transfer-function table from angles
// kernel_mode 'cached' streams H' from a table that must carry the REFERENCE's own phase angles z sqrt(klam^2 - Kx^2 - Ky^2)
// (torch's CPU sqrt is not correctly rounded, see asm_host.inregister_deviation_estimate).  Kx^2 and Ky^2 are even in the
// frequency index, so the host evaluates the angles with the reference's library on the unique quarter only
// ([C][Hp/2+1][Wp/2+1], 4x less host work and upload) and this kernel expands it into the column-major slot-order table the
// column pass reads: table[c][sc][sr] = keep ? exp(i ang[c][|bin(sr)|][|bin(sc)|]) : 0, keep <=> Ky^2[sc] <= tau[sr] (the same
// bit-exact mask as the in-register mode).
__global__ void __launch_bounds__(256) thz_k_tf_table(const float* __restrict__ angq, int Hu, int Wu, const float2* __restrict__ rowtau,
                                                      const float* __restrict__ colk2, const int* __restrict__ rabs,
                                                      const int* __restrict__ cabs, int Hp, int Wp, cpx* __restrict__ table) {
    const int c = blockIdx.z, sc = blockIdx.y;
    const float ky2 = colk2[(size_t)c * Wp + sc];
    const float* aq = angq + (size_t)c * Hu * Wu + cabs[sc];
    cpx* out = table + ((size_t)c * Wp + sc) * Hp;
    for (int sr = blockIdx.x * blockDim.x + threadIdx.x; sr < Hp; sr += gridDim.x * blockDim.x) {
        const bool keep = ky2 <= rowtau[(size_t)c * Hp + sr].y;
        float sn = 0.f, cs = 0.f;
        if (keep) sincosf(aq[(size_t)rabs[sr] * Wu], &sn, &cs);
        out[sr] = cmake(cs, sn);
    }
}

extern "C" int thz_tf_table_from_angles(const void* angq, int32_t C, int32_t Hu, int32_t Wu, const void* rowtau, const void* colk2,
                                        const void* rabs, const void* cabs, int32_t Hp, int32_t Wp, void* table, void* stream_) {
    ThzDeviceGuard dev_guard(table);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (C < 1 || Hp < 1 || Wp < 1 || Hu < 1 || Wu < 1) return thz_set_error(THZ_E_SHAPE, "thz_tf_table_from_angles: bad sizes");
    if (!angq || !rowtau || !colk2 || !rabs || !cabs || !table) return thz_set_error(THZ_E_NULL, "thz_tf_table_from_angles: null pointer");
    if (Wp > 65535 || C > 65535) return thz_set_error(THZ_E_SHAPE, "thz_tf_table_from_angles: grid too large");
    dim3 grid((Hp + 255) / 256 > 8 ? 8 : (Hp + 255) / 256, Wp, C);
    thz_launch_begin(stream, THZ_KC_DOE);
    thz_k_tf_table<<<grid, 256, 0, stream>>>((const float*)angq, Hu, Wu, (const float2*)rowtau, (const float*)colk2, (const int*)rabs,
                                             (const int*)cabs, Hp, Wp, (cpx*)table);
    thz_launch_end(stream, THZ_KC_DOE);
    THZ_CHECK_LAUNCH("thz_k_tf_table");
    return THZ_OK;
}

// ------------------------------------------------------------------------------- pointwise optical elements
// y[f, p] = x[f, p] * m[(f % C) * per_channel, p]  (m complex, optionally conjugated) or  * mask[p]  (real):
// thin lens (Components/Thin_Lens.py:66-72) and aperture (Components/Aperture.py:126), forward and adjoint.
__global__ void __launch_bounds__(256) thz_k_field_mul(const cpx* __restrict__ x, const void* __restrict__ m, cpx* __restrict__ y,
                                                       int BC, int C, size_t HW, int per_channel, int m_real, int conj_m) {
    const size_t total = (size_t)BC * HW;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t f = i / HW, p = i - f * HW;
        const size_t mi = (per_channel ? (f % C) * HW : 0) + p;
        const cpx v = x[i];
        if (m_real) {
            y[i] = cscale(v, __ldg((const float*)m + mi));
        } else {
            const cpx w = __ldg((const cpx*)m + mi);
            y[i] = conj_m ? cmulc(v, w) : cmul(v, w);
        }
    }
}

extern "C" int thz_field_mul(const void* x, const void* m, void* y, int32_t BC, int32_t C, uint64_t HW, int32_t per_channel,
                             int32_t m_real, int32_t conj_m, void* stream_) {
    ThzDeviceGuard dev_guard(x);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (BC < 0 || C < 1) return thz_set_error(THZ_E_SHAPE, "thz_field_mul: bad sizes");
    if (BC == 0 || HW == 0) return THZ_OK;
    if (!x || !m || !y) return thz_set_error(THZ_E_NULL, "thz_field_mul: null pointer");
    const size_t total = (size_t)BC * HW, want = (total + 255) / 256, cap = (size_t)thz_sm_count() * 16;
    thz_launch_begin(stream, THZ_KC_DOE);
    thz_k_field_mul<<<(unsigned)(want < cap ? want : cap), 256, 0, stream>>>((const cpx*)x, m, (cpx*)y, BC, C, (size_t)HW, per_channel,
                                                                           m_real, conj_m);
    thz_launch_end(stream, THZ_KC_DOE);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return thz_set_cuda_error("thz_field_mul", e);
    return THZ_OK;
}
