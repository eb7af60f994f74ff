// Loss and optimizer kernels of the optimisation loop around the propagation:
//   out_amp = normalize(abs(y)**2)            utils/Helper_Functions.py:185-193 (divide by the per-batch maximum)
//   loss    = nn.MSELoss()(out_amp, target)   experiment_four_focal_spots.ipynb cell 8
//   optimizer.step()                          torch.optim.Adam / AdamW, lr 0.02 (same cell)
// The loss is evaluated together with its gradient wrt the complex field (what autograd would hand to the ASM
// adjoint), and the Adam step keeps its step counter on the device, so a whole iteration can be captured in a CUDA
// graph and replayed.  All of it is elementwise / reduction work bound by HBM: y is read twice (maximum, loss +
// gradient), the gradient written once.
#include "thz_common.cuh"
#include "thz_runtime.h"

// ------------------------------------------------------------------------------- per-batch maximum of |y|^2
// key = (float bits of I) << 32 | ~index: I >= 0, so unsigned order of the bits is numeric order, and among equal maxima
// the SMALLEST index wins -- torch.max(dim) returns the first maximal element, and its backward sends the gradient there.
__global__ void __launch_bounds__(256) thz_k_intensity_max(const cpx* __restrict__ y, unsigned long long* __restrict__ key,
                                                           size_t n_per_b) {
    const int b = blockIdx.y;
    const cpx* yb = y + (size_t)b * n_per_b;
    unsigned long long best = 0ull;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_per_b; i += (size_t)gridDim.x * blockDim.x) {
        const cpx v = yb[i];
        const float a = hypotf(v.x, v.y);                      // torch.abs(y) ** 2: the magnitude is rounded first
        const float I = a * a;
        const unsigned long long k = ((unsigned long long)__float_as_uint(I) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
        best = k > best ? k : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    __shared__ unsigned long long sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) best = sm[w] > best ? sm[w] : best;
        atomicMax(key + b, best);
    }
}

// ------------------------------------------------------------------------------- loss + gradient
// a = I / m_b, diff = a - t, loss += diff^2 / N;  dL/dI_k = 2 diff_k / (N m) - [k == argmax_b] S_b / m^2 with
// S_b = sum_j (2 diff_j / N) I_j;  dL/dy = 2 dL/dI y  (torch's convention for a real loss of a complex tensor).
// The argmax correction needs the finished S_b, so it is applied by thz_k_normmse_fixup afterwards.
__global__ void __launch_bounds__(256) thz_k_normmse(const cpx* __restrict__ y, const float* __restrict__ target,
                                                     const unsigned long long* __restrict__ key, float* __restrict__ loss,
                                                     float* __restrict__ S, cpx* __restrict__ gy, size_t n_per_b, float inv_n) {
    const int b = blockIdx.y;
    const float m = __uint_as_float((unsigned)(key[b] >> 32));
    const cpx* yb = y + (size_t)b * n_per_b;
    const float* tb = target + (size_t)b * n_per_b;
    float l = 0.f, s = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_per_b; i += (size_t)gridDim.x * blockDim.x) {
        const cpx v = yb[i];
        const float a = hypotf(v.x, v.y);
        const float I = a * a;
        const float diff = __fdiv_rn(I, m) - tb[i];
        l = fmaf(diff, diff, l);
        const float ga = 2.f * diff * inv_n;
        s = fmaf(ga, I, s);
        if (gy) {
            const float c = 2.f * __fdiv_rn(ga, m);
            gy[(size_t)b * n_per_b + i] = cmake(c * v.x, c * v.y);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        l += __shfl_xor_sync(0xffffffffu, l, o);
        s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    __shared__ float sl[8], ss[8];
    if ((threadIdx.x & 31) == 0) {
        sl[threadIdx.x >> 5] = l;
        ss[threadIdx.x >> 5] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            l += sl[w];
            s += ss[w];
        }
        atomicAdd(loss, l * inv_n);
        atomicAdd(S + b, s);
    }
}

__global__ void thz_k_normmse_fixup(const cpx* __restrict__ y, const unsigned long long* __restrict__ key,
                                    const float* __restrict__ S, cpx* __restrict__ gy, size_t n_per_b, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float m = __uint_as_float((unsigned)(key[b] >> 32));
    const size_t i = (size_t)b * n_per_b + (size_t)(0xFFFFFFFFu - (unsigned)(key[b] & 0xFFFFFFFFull));
    const float c = -2.f * __fdiv_rn(S[b], m * m);
    const cpx v = y[i];
    gy[i] = cmake(gy[i].x + c * v.x, gy[i].y + c * v.y);
}

extern "C" int thz_normmse_loss(const void* y, const void* target, int32_t B, uint64_t n_per_b, void* scratch, void* loss,
                                void* gy, void* stream_) {
    ThzDeviceGuard dev_guard(y);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B < 0) return thz_set_error(THZ_E_SHAPE, "thz_normmse_loss: negative batch");
    if (B == 0 || n_per_b == 0) return THZ_OK;
    if (!y || !target || !scratch || !loss) return thz_set_error(THZ_E_NULL, "thz_normmse_loss: null pointer");
    if (n_per_b > 0xFFFFFFFFull) return thz_set_error(THZ_E_UNSUPPORTED, "thz_normmse_loss: more than 2^32 samples per batch entry");
    unsigned long long* key = (unsigned long long*)scratch;     // [B] keys, then [B] float sums
    float* S = (float*)(key + B);
    cudaError_t e = cudaMemsetAsync(scratch, 0, (size_t)B * 12, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(loss, 0, sizeof(float), stream);
    if (e != cudaSuccess) return thz_set_cuda_error("thz_normmse_loss: memset", e);
    const size_t want = (n_per_b + 255) / 256;
    const size_t cap = (size_t)thz_sm_count() * 8 / (size_t)(B < 8 ? B : 8) + 1;
    dim3 grid((unsigned)(want < cap ? want : cap), (unsigned)B);
    const float inv_n = (float)(1.0 / ((double)B * (double)n_per_b));
    thz_launch_begin(stream, THZ_KC_TRAIN);
    thz_k_intensity_max<<<grid, 256, 0, stream>>>((const cpx*)y, key, (size_t)n_per_b);
    thz_launch_end(stream, THZ_KC_TRAIN);
    thz_launch_begin(stream, THZ_KC_TRAIN);
    thz_k_normmse<<<grid, 256, 0, stream>>>((const cpx*)y, (const float*)target, key, (float*)loss, S, (cpx*)gy, (size_t)n_per_b, inv_n);
    thz_launch_end(stream, THZ_KC_TRAIN);
    if (gy) {
        thz_launch_begin(stream, THZ_KC_TRAIN);
        thz_k_normmse_fixup<<<(B + 63) / 64, 64, 0, stream>>>((const cpx*)y, key, S, (cpx*)gy, (size_t)n_per_b, B);
        thz_launch_end(stream, THZ_KC_TRAIN);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return thz_set_cuda_error("thz_normmse_loss", e);
    return THZ_OK;
}

// ------------------------------------------------------------------------------- per-entry losses (loss-landscape sweeps)
// losses[b] = mean_i (I_b[i] / max_i I_b[i] - target[i])^2: what VisTools/calc_loss.py:35-39 evaluates once per grid point
// (output / torch.max(output), then nn.MSELoss), for B candidate outputs at once; the target may be shared by all entries.
__global__ void __launch_bounds__(256) thz_k_normmse_each(const cpx* __restrict__ y, const float* __restrict__ target,
                                                          size_t t_bstride, const unsigned long long* __restrict__ key,
                                                          float* __restrict__ losses, size_t n_per_b, float inv_n) {
    const int b = blockIdx.y;
    const float m = __uint_as_float((unsigned)(key[b] >> 32));
    const cpx* yb = y + (size_t)b * n_per_b;
    const float* tb = target + (size_t)b * t_bstride;
    float l = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_per_b; i += (size_t)gridDim.x * blockDim.x) {
        const cpx v = yb[i];
        const float a = hypotf(v.x, v.y);
        const float diff = __fdiv_rn(a * a, m) - tb[i];
        l = fmaf(diff, diff, l);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    __shared__ float sl[8];
    if ((threadIdx.x & 31) == 0) sl[threadIdx.x >> 5] = l;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) l += sl[w];
        atomicAdd(losses + b, l * inv_n);
    }
}

extern "C" int thz_normmse_loss_each(const void* y, const void* target, int32_t target_shared, int32_t B, uint64_t n_per_b,
                                     void* scratch, void* losses, void* stream_) {
    ThzDeviceGuard dev_guard(y);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B < 0) return thz_set_error(THZ_E_SHAPE, "thz_normmse_loss_each: negative batch");
    if (B == 0 || n_per_b == 0) return THZ_OK;
    if (!y || !target || !scratch || !losses) return thz_set_error(THZ_E_NULL, "thz_normmse_loss_each: null pointer");
    if (n_per_b > 0xFFFFFFFFull) return thz_set_error(THZ_E_UNSUPPORTED, "thz_normmse_loss_each: more than 2^32 samples per entry");
    unsigned long long* key = (unsigned long long*)scratch;     // [B] keys
    cudaError_t e = cudaMemsetAsync(scratch, 0, (size_t)B * 8, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(losses, 0, (size_t)B * sizeof(float), stream);
    if (e != cudaSuccess) return thz_set_cuda_error("thz_normmse_loss_each: memset", e);
    const size_t want = (n_per_b + 255) / 256;
    const size_t cap = (size_t)thz_sm_count() * 8 / (size_t)(B < 8 ? B : 8) + 1;
    dim3 grid((unsigned)(want < cap ? want : cap), (unsigned)B);
    thz_launch_begin(stream, THZ_KC_TRAIN);
    thz_k_intensity_max<<<grid, 256, 0, stream>>>((const cpx*)y, key, (size_t)n_per_b);
    thz_launch_end(stream, THZ_KC_TRAIN);
    thz_launch_begin(stream, THZ_KC_TRAIN);
    thz_k_normmse_each<<<grid, 256, 0, stream>>>((const cpx*)y, (const float*)target, target_shared ? 0 : (size_t)n_per_b, key,
                                                 (float*)losses, (size_t)n_per_b, (float)(1.0 / (double)n_per_b));
    thz_launch_end(stream, THZ_KC_TRAIN);
    e = cudaGetLastError();
    if (e != cudaSuccess) return thz_set_cuda_error("thz_normmse_loss_each", e);
    return THZ_OK;
}

// ------------------------------------------------------------------------------- Adam / AdamW
// torch.optim.Adam (single-tensor path): m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g, g, 1-b2);
// p.addcdiv_(m, sqrt(v)/sqrt(1-b2^t) + eps, value = -lr/(1-b1^t)); weight decay: g += wd p (Adam) or p *= 1 - lr wd
// (AdamW).  t = *step + 1 is read from the device and incremented by a second one-thread kernel, so the launch
// sequence does not depend on the iteration number (CUDA-graph capturable).
__global__ void __launch_bounds__(256) thz_k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                  float* __restrict__ v, const int* __restrict__ step, size_t n, float lr,
                                                  float b1, float b2, float eps, float wd, int decoupled) {
    const float t = (float)(*step + 1);
    const float bc1 = 1.f - powf(b1, t), bc2s = sqrtf(1.f - powf(b2, t));
    const float step_size = lr / bc1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float pi = p[i], gi = g[i];
        if (wd != 0.f) {
            if (decoupled) pi *= 1.f - lr * wd;
            else gi = fmaf(wd, pi, gi);
        }
        const float mi = m[i] + (1.f - b1) * (gi - m[i]);
        const float vi = fmaf(gi * gi, 1.f - b2, v[i] * b2);
        m[i] = mi;
        v[i] = vi;
        const float denom = __fdiv_rn(sqrtf(vi), bc2s) + eps;
        p[i] = pi - step_size * __fdiv_rn(mi, denom);
    }
}
__global__ void thz_k_step_inc(int* step) { *step += 1; }

extern "C" int thz_adam_step(void* p, const void* g, void* m, void* v, void* step, uint64_t n, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int32_t decoupled, int32_t advance, void* stream_) {
    ThzDeviceGuard dev_guard(p);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) return THZ_OK;
    if (!p || !g || !m || !v || !step) return thz_set_error(THZ_E_NULL, "thz_adam_step: null pointer");
    const size_t want = (n + 255) / 256, cap = (size_t)thz_sm_count() * 8;
    thz_launch_begin(stream, THZ_KC_TRAIN);
    thz_k_adam<<<(unsigned)(want < cap ? want : cap), 256, 0, stream>>>((float*)p, (const float*)g, (float*)m, (float*)v,
                                                                      (const int*)step, (size_t)n, lr, beta1, beta2, eps,
                                                                      weight_decay, decoupled);
    thz_launch_end(stream, THZ_KC_TRAIN);
    if (advance) {
        thz_launch_begin(stream, THZ_KC_TRAIN);
        thz_k_step_inc<<<1, 1, 0, stream>>>((int*)step);
        thz_launch_end(stream, THZ_KC_TRAIN);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return thz_set_cuda_error("thz_adam_step", e);
    return THZ_OK;
}
