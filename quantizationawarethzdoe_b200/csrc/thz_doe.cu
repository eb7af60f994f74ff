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
    if (n == 0) return THZ_OK;   // empty maps are legal (and have a NULL data pointer)
    if (!w || !h) return thz_set_error(THZ_E_NULL, "thz_height_fwd: null pointer");
    thz_launch_begin((cudaStream_t)stream, THZ_KC_QUANT);
    thz_k_height_fwd<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>((const float*)w, hmax, clampv, (float*)h, n);
    thz_launch_end((cudaStream_t)stream, THZ_KC_QUANT);
    THZ_CHECK_LAUNCH("thz_k_height_fwd");
    return THZ_OK;
}

extern "C" int thz_height_bwd(const void* g, const void* w, float hmax, float clampv, void* gw, uint64_t n, void* stream) {
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
