// Micro-benchmark: throughput of scalar FADD/FFMA vs the packed sm_100 forms (FADD2/FFMA2) per SM and clock.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o f32x2_bench f32x2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    const float2 b = make_float2(s, s * 0.5f), c = make_float2(0.25f, 0.125f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {          // scalar FADD x2
                a[i].x += b.x;
                a[i].y += b.y;
            } else if (MODE == 1) {   // FADD2
                a[i] = __fadd2_rn(a[i], b);
            } else if (MODE == 2) {   // scalar FFMA x2
                a[i].x = fmaf(a[i].x, b.x, c.x);
                a[i].y = fmaf(a[i].y, b.y, c.y);
            } else if (MODE == 3) {   // FFMA2
                a[i] = __ffma2_rn(a[i], b, c);
            } else if (MODE == 4) {   // complex multiply, scalar: 2 FMUL + 2 FFMA
                const float2 x = a[i];
                a[i].x = fmaf(-x.y, b.y, x.x * b.x);
                a[i].y = fmaf(x.y, b.x, x.x * b.y);
            } else if (MODE == 5) {   // complex multiply, packed: FMUL2 + FFMA2 on a swapped copy
                const float2 x = a[i];
                const float2 t = __fmul2_rn(x, make_float2(b.x, b.x));
                a[i] = __ffma2_rn(make_float2(x.y, x.x), make_float2(-b.y, b.y), t);
            }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, float elems_per_iter) {
    float* out;
    int sms = 148, iters = 20000;
    cudaMalloc(&out, sms * 8 * 256 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<sms * 8, 256>>>(out, 100, 1.0001f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<sms * 8, 256>>>(out, iters, 1.0001f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)sms * 8 * 256 * iters * elems_per_iter;
    printf("%-28s %8.3f ms  %8.2f G elem-ops/s  (%6.1f per SM per clk at 1.9 GHz)\n", name, ms, ops / ms * 1e-6, ops / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(out);
}

int main() {
    run<0>("scalar FADD (2 per cpx)", 16);
    run<1>("FADD2", 16);
    run<2>("scalar FFMA (2 per cpx)", 16);
    run<3>("FFMA2", 16);
    run<4>("cmul scalar (4 instr)", 8);
    run<5>("cmul packed (FMUL2+FFMA2)", 8);
    return 0;
}
