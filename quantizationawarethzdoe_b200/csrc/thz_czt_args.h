// Kernel arguments of the Toeplitz complex GEMM, shared by the CUDA-core (thz_czt.cu) and tcgen05 (thz_czt_tc.cu) paths.
#pragma once
#include <cuda_runtime.h>

#include "thz_common.cuh"

struct ToeplitzGemmArgs {
    int batch, M, N, K;
    const cpx* g;             // [batch][L]
    int L, off, sm, sk, conj_g;
    const cpx* B;             // element (b, k, n) at B[b*sb_b + k*sb_k + n*sb_n]
    long long sb_b, sb_k, sb_n;
    const cpx* pro;           // optional prologue factor, same indexing as B
    int conj_pro;
    cpx* C;                   // element (b, m, n) at C[b*sc_b + m*sc_m + n*sc_n]
    long long sc_b, sc_m, sc_n;
    const cpx* epi;           // optional epilogue factor, same indexing as C
    int conj_epi;
    int debug_mode;           // THZ_CZT_DEBUG: 0 normal; 1 producers skip the smem stores; 2 MMA warp skips the MMAs;
                              // 3 / 4 clock64 timeline of CTA 0 (tools/tc_timeline.py) -- bottleneck experiments only
};

// tcgen05 / TMEM implementation (thz_czt_tc.cu)
// thz_toeplitz_gemm_tc_ineligible: NULL if the kernel can serve the call, else the reason; the launch refuses ineligible calls.
const char* thz_toeplitz_gemm_tc_ineligible(const ToeplitzGemmArgs& a, const void* scratch);
int thz_toeplitz_gemm_tc_launch(const ToeplitzGemmArgs& a, void* scratch, cudaStream_t stream);
