// __global__ wrappers, launch code and C ABI for the fused ASM pipeline (see thz_asm.cuh).
#include <stdio.h>
#include <string.h>

#include "thz_asm_host.h"
#include "thz_asm_p2_launch.h"
#include "thz_runtime.h"

// ------------------------------------------------------------------------------- kernels
template <bool MIXED>
__global__ void __launch_bounds__(256) thz_k1_row_fwd(const __grid_constant__ RowFwdArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* s = reinterpret_cast<cpx*>(smem_raw);
    const int tid = threadIdx.x, nt = blockDim.x, bx = blockIdx.x;
    k1_load(a, s, bx, tid, nt);
    __syncthreads();
    const int pitch = thz_padded_len(a.Wp);
    for (int st = 0; st < a.plan.ns; ++st) {
        fft_stage_all<MIXED, false>(a.plan, st, s, a.lines, pitch, false, tid, nt, a.tw);
        __syncthreads();
    }
    k1_store(a, s, bx, tid, nt);
}

template <bool MIXED>
__global__ void __launch_bounds__(512) thz_k2_col(const __grid_constant__ ColArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* s = reinterpret_cast<cpx*>(smem_raw);
    const int tid = threadIdx.x, nt = blockDim.x, bx = blockIdx.x, by = blockIdx.y;
    k2_load(a, s, bx, by, tid, nt);
    __syncthreads();
    const int last = a.plan.ns - 1;
    for (int st = 0; st < last; ++st) {
        fft_stage_all<MIXED, false>(a.plan, st, s, a.cols, 0, true, tid, nt, a.tw);
        __syncthreads();
    }
    k2_middle<MIXED>(a, s, bx, by, tid, nt);
    __syncthreads();
    for (int st = last - 1; st >= 0; --st) {
        fft_stage_all<MIXED, true>(a.plan, st, s, a.cols, 0, true, tid, nt, a.tw);
        __syncthreads();
    }
    k2_store(a, s, bx, by, tid, nt);
}

template <bool MIXED>
__global__ void __launch_bounds__(512) thz_k3_row_inv(const __grid_constant__ RowInvArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* s = reinterpret_cast<cpx*>(smem_raw);
    const int tid = threadIdx.x, nt = blockDim.x, bx = blockIdx.x;
    const int pitch = thz_padded_len(a.Wp);
    float acc[THZ_K3_OWN];
#pragma unroll
    for (int k = 0; k < THZ_K3_OWN; ++k) acc[k] = 0.f;
    const int f_lo = blockIdx.y * a.bc_per_cta;
    const int f_hi = min(a.nbc, f_lo + a.bc_per_cta);
    for (int f = f_lo; f < f_hi; ++f) {
        k3_load(a, s, bx, f, tid, nt);
        __syncthreads();
        for (int st = a.plan.ns - 1; st >= 0; --st) {
            fft_stage_all<MIXED, true>(a.plan, st, s, a.lines, pitch, false, tid, nt, a.tw);
            __syncthreads();
        }
        k3_epilogue(a, s, bx, f, tid, nt, acc);
        __syncthreads();
    }
    k3_flush(a, bx, tid, nt, acc);
}

template <bool MIXED>
__global__ void __launch_bounds__(512) thz_k2f_col_fft(const __grid_constant__ ColFftArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* s = reinterpret_cast<cpx*>(smem_raw);
    const int tid = threadIdx.x, nt = blockDim.x, bx = blockIdx.x, by = blockIdx.y;
    k2f_load(a, s, bx, by, tid, nt);
    __syncthreads();
    for (int st = 0; st < a.plan.ns; ++st) {
        fft_stage_all<MIXED, false>(a.plan, st, s, a.cols, 0, true, tid, nt, a.tw);
        __syncthreads();
    }
    k2f_store(a, s, bx, by, tid, nt);
}

// ------------------------------------------------------------------------------- launch helpers
template <typename K>
static int set_smem(K kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return THZ_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return thz_set_cuda_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize)", e);
    return THZ_OK;
}

#define THZ_LAUNCH(kern, cls, mixed, grid, block, smem, stream, args)                           \
    do {                                                                                        \
        int rc_;                                                                                \
        if (mixed) {                                                                            \
            if ((rc_ = set_smem(kern<true>, smem)) != THZ_OK) return rc_;                       \
            thz_launch_begin(stream, cls);                                                      \
            kern<true><<<grid, block, smem, stream>>>(args);                                    \
        } else {                                                                                \
            if ((rc_ = set_smem(kern<false>, smem)) != THZ_OK) return rc_;                      \
            thz_launch_begin(stream, cls);                                                      \
            kern<false><<<grid, block, smem, stream>>>(args);                                   \
        }                                                                                       \
        thz_launch_end(stream, cls);                                                            \
        cudaError_t e_ = cudaGetLastError();                                                    \
        if (e_ != cudaSuccess) return thz_set_cuda_error(#kern, e_);                            \
    } while (0)

extern "C" uint64_t thz_asm_workspace_bytes(const thz_asm_desc* d) {
    if (!d) return 0;
    return thz_asm_ws_bytes(d);
}

extern "C" int thz_asm_propagate(const thz_asm_desc* d, void* stream_) {
    ThzDeviceGuard dev_guard(d ? d->x : nullptr);
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = thz_asm_validate(d);
    if (rc != THZ_OK) return thz_set_error(rc, "thz_asm_propagate: invalid descriptor");
    if (d->slab_parts <= 1 && d->ws_bytes < thz_asm_ws_bytes(d))
        return thz_set_error(THZ_E_WORKSPACE, "thz_asm_propagate: workspace too small");
    const int nbc_all = d->B * d->C;
    const int chunk = (int)thz_asm_chunk_fields(d);
    const int sm_count = thz_sm_count();
    const int nchunks = (nbc_all + chunk - 1) / chunk;
    const int stages = d->stages ? d->stages : 7;
    bool zeroed = false;
    for (int f0 = 0; f0 < nbc_all; f0 += chunk) {
        const int nbc = nbc_all - f0 < chunk ? nbc_all - f0 : chunk;
        AsmLaunch L;
        rc = thz_asm_plan_chunk(d, f0, nbc, sm_count, &L);
        if (rc == THZ_E_UNSUPPORTED)
            return thz_set_error(rc, "thz_asm_propagate: transform length has a prime factor > 7 (or output row too wide)");
        if (rc != THZ_OK) return thz_set_error(rc, "thz_asm_propagate: line does not fit in shared memory");
        // K2 -> K3 intermediate with permuted columns + TMA-staged row-iFFT kernel (thz_p2_k3t): whole pipeline on the static
        // kernels, fast column kernel with 2-column tiles, row-major second buffer (THZ_NO_K3TMA=1: off)
        if (!thz_env_is_1("THZ_NO_K3TMA")) {
            L.k2.t2_perm = thz_asm_t2_perm_radix(d, &L, stages);
            L.k3.t2_perm = L.k2.t2_perm;
        }
        L.k3.gh_atomic = (d->doe_mode == 2 && d->doe_gh_mode == 1) ? 2 : (d->doe_mode == 2 && (nchunks > 1 || L.k3_gridy > 1)) ? 1 : 0;
        if ((stages & 4) && L.k3.gh_atomic == 1 && !zeroed) {
            cudaError_t e = cudaMemsetAsync(d->doe_gh, 0, (size_t)d->outH * d->outW * sizeof(float), stream);
            if (e != cudaSuccess) return thz_set_cuda_error("cudaMemsetAsync(gh)", e);
            zeroed = true;
        }
        if (!(stages & 1)) {
        } else if (L.p2_w) {
            if ((rc = thz_p2_launch_k1(L.k1, L.k1_grid, L.k1_threads, L.k1_smem, stream)) != THZ_OK) return rc;
        } else {
            THZ_LAUNCH(thz_k1_row_fwd, THZ_KC_ROW_FWD, L.mixed_w, L.k1_grid, L.k1_threads, L.k1_smem, stream, L.k1);
        }
        if (!(stages & 2)) {
        } else if (L.p2_h) {
            if ((rc = thz_p2_launch_k2(L.k2, L.k2_gridx, nbc, L.k2_threads, L.k2_smem, stream)) != THZ_OK) return rc;
        } else {
            THZ_LAUNCH(thz_k2_col, THZ_KC_COL, L.mixed_h, dim3(L.k2_gridx, nbc), L.k2_threads, L.k2_smem, stream, L.k2);
        }
        if (!(stages & 4)) {
        } else if (L.p2_w) {
            if ((rc = thz_p2_launch_k3(L.k3, L.k3_gridx, L.k3_gridy, L.k3_threads, L.k3_smem, stream)) != THZ_OK) return rc;
        } else {
            THZ_LAUNCH(thz_k3_row_inv, THZ_KC_ROW_INV, L.mixed_w, dim3(L.k3_gridx, L.k3_gridy), L.k3_threads, L.k3_smem, stream, L.k3);
        }
    }
    return THZ_OK;
}

// ------------------------------------------------------------------------------- stand-alone fft2
extern "C" int thz_fft2_c2c(const void* x, void* y, int32_t batch, int32_t H, int32_t W, int32_t inverse, int32_t ortho,
                            const void* tw_h, const void* tw_w, void* ws, uint64_t ws_bytes, void* stream_) {
    ThzDeviceGuard dev_guard(x);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!x || !y || !tw_h || !tw_w || !ws) return thz_set_error(THZ_E_NULL, "thz_fft2_c2c: null pointer");
    if (batch < 1 || H < 1 || W < 1) return thz_set_error(THZ_E_SHAPE, "thz_fft2_c2c: bad shape");
    if (ws_bytes < (uint64_t)batch * H * W * sizeof(cpx)) return thz_set_error(THZ_E_WORKSPACE, "thz_fft2_c2c: workspace too small");
    FftPlan pw, ph;
    if (thz_make_plan(W, &pw) != 0 || thz_make_plan(H, &ph) != 0)
        return thz_set_error(THZ_E_UNSUPPORTED, "thz_fft2_c2c: transform length has a prime factor > 7");
    const size_t lw = (size_t)thz_padded_len(W) * sizeof(cpx), lh = (size_t)thz_padded_len(H) * sizeof(cpx);
    if (lw > THZ_SMEM_BUDGET || lh > THZ_SMEM_BUDGET) return thz_set_error(THZ_E_SMEM, "thz_fft2_c2c: line does not fit in shared memory");

    RowFwdArgs a1;
    memset(&a1, 0, sizeof(a1));
    a1.x = (const cpx*)x;
    a1.T = (cpx*)ws;
    a1.nbc = batch;
    a1.rowsT = H;
    a1.c0 = 0;
    a1.C = 1;
    a1.inH = H;
    a1.inW = W;
    a1.Wp = W;
    a1.in_c0 = 0;
    a1.plan = pw;
    a1.tw = (const cpx*)tw_w;
    a1.doe.hmap = nullptr;
    a1.elem.mask = nullptr;
    a1.elem.mul = nullptr;
    a1.doe.hstride = 0;
    a1.doe.b0 = 0;
    a1.conj_in = inverse ? 1 : 0;
    int lines = thz_imax(1, 4096 / W);
    lines = thz_imin(lines, 16);
    while (lines > 1 && lines * lw > 64 * 1024) --lines;
    a1.lines = lines;
    const int g1 = (batch * H + lines - 1) / lines;
    THZ_LAUNCH(thz_k1_row_fwd, THZ_KC_ROW_FWD, pw.mixed, g1, 256, lines * lw, stream, a1);

    ColFftArgs a2;
    memset(&a2, 0, sizeof(a2));
    a2.T = (const cpx*)ws;
    a2.y = (cpx*)y;
    a2.H = H;
    a2.W = W;
    int cols = 16;
    while (cols > 1 && cols * lh > 72 * 1024) cols >>= 1;
    a2.cols = cols;
    const double nrm = ortho ? 1.0 / sqrt((double)H * (double)W) : (inverse ? 1.0 / ((double)H * (double)W) : 1.0);
    a2.scale = (float)nrm;
    a2.conj_out = inverse ? 1 : 0;
    a2.plan = ph;
    a2.planW = pw;
    a2.tw = (const cpx*)tw_h;
    const int work = (H / 16 + 1) * cols;
    const int threads = work >= 1024 ? 512 : (work >= 384 ? 256 : 128);
    THZ_LAUNCH(thz_k2f_col_fft, THZ_KC_FFT2_COL, ph.mixed, dim3((W + cols - 1) / cols, batch), threads, cols * lh, stream, a2);
    return THZ_OK;
}
