// Launch entry points of the power-of-two fast path (defined in thz_asm_p2_k1.cu / _k2.cu / _k3.cu via thz_asm_p2_kernels.inc, used by thz_asm.cu).
#pragma once
#include <cuda_runtime.h>

#include "thz_asm.cuh"

int thz_p2_launch_k1(const RowFwdArgs& a, int grid, int threads, size_t smem, cudaStream_t stream);
int thz_p2_launch_k2(const ColArgs& a, int gridx, int gridy, int threads, size_t smem, cudaStream_t stream);
int thz_p2_launch_k3(const RowInvArgs& a, int gridx, int gridy, int threads, size_t smem, cudaStream_t stream);
