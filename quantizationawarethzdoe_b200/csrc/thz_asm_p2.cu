// __global__ wrappers and size dispatch of the power-of-two fast path (bodies: thz_asm_p2.cuh).
#include "thz_asm_host.h"
#include "thz_asm_p2_launch.h"
#include "thz_runtime.h"

// ------------------------------------------------------------------------------- stage recursion with barriers
template <int N, int S, int S1, int LINES>
__device__ __forceinline__ void fwd_rows(cpx* s, int tid, int nt, const cpx* tw) {
    if constexpr (S < S1) {
        p2_stage_rows<N, S, false, LINES>(s, tid, nt, tw);
        __syncthreads();
        fwd_rows<N, S + 1, S1, LINES>(s, tid, nt, tw);
    }
}
template <int N, int S, int S0, int LINES>
__device__ __forceinline__ void inv_rows(cpx* s, int tid, int nt, const cpx* tw) {   // stages S, S-1, ..., S0
    if constexpr (S >= S0) {
        p2_stage_rows<N, S, true, LINES>(s, tid, nt, tw);
        __syncthreads();
        inv_rows<N, S - 1, S0, LINES>(s, tid, nt, tw);
    }
}
template <int N, int S, int S1, int COLS>
__device__ __forceinline__ void fwd_cols(cpx* s, int tid, int nt, const cpx* tw) {
    if constexpr (S < S1) {
        p2_stage_cols<N, S, false, COLS, p2_col_threads(N)>(s, tid, nt, tw);
        __syncthreads();
        fwd_cols<N, S + 1, S1, COLS>(s, tid, nt, tw);
    }
}
template <int N, int S, int S0, int COLS>
__device__ __forceinline__ void inv_cols(cpx* s, int tid, int nt, const cpx* tw) {
    if constexpr (S >= S0) {
        p2_stage_cols<N, S, true, COLS, p2_col_threads(N)>(s, tid, nt, tw);
        __syncthreads();
        inv_cols<N, S - 1, S0, COLS>(s, tid, nt, tw);
    }
}

// ------------------------------------------------------------------------------- kernels
template <int N, bool ELEM>
__global__ void __launch_bounds__(p2_row_threads(N), p2_min_blocks(N)) thz_p2_k1(const __grid_constant__ RowFwdArgs a) {
    constexpr int LINES = p2_row_lines(N);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* s = reinterpret_cast<cpx*>(smem_raw);
    cpx* tws = s + LINES * p2_pitch(N);                                  // shared-memory copy of the stage twiddles
    cpx* xs = tws + p2_tw_count(N);                                      // staged raw rows
    float* hs = reinterpret_cast<float*>(xs + (size_t)LINES * a.inW);      // staged height-map rows
    // the launch always uses p2_row_threads(N) threads (thz_asm_apply_p2): a compile-time block size lets every work loop
    // of the phase functions resolve its trip count
    // (measured: helps the radix-16 lengths, hurts the radix-25 ones -- more unrolling, more register pressure -- so those
    // keep the run-time value)
    const int nt = p2_radix(N, 0) == 16 ? p2_row_threads(N) : (int)blockDim.x;
    const int tid = threadIdx.x;
    const int ngroups = (a.nbc * a.inH + LINES - 1) / LINES;
    int grp = blockIdx.x;
    p2_tw_fill<N>(tws, a.tw, tid, nt);                                  // visible after the first barrier below
    if constexpr (!p2_row_pipelined(N)) {     // no room for staging: plain load -> transform -> store per group
        for (; grp < ngroups; grp += gridDim.x) {
            p2k1_first<N, ELEM>(a, s, grp, tid, nt);
            __syncthreads();
            fwd_rows<N, 1, p2_stages(N), LINES>(s, tid, nt, tws);
            p2k1_store<N>(a, s, grp, tid, nt);
            __syncthreads();
        }
        return;
    }
    // Staging of the next group's raw rows.  When the rows are 16-byte granular (even width, width % 4 == 0 with a height map,
    // aligned bases) ONE thread hands them to the TMA copy engine -- a single cp.async.bulk for the x rows of the group (they
    // are contiguous) and one per height-map row -- and everybody waits on an mbarrier phase; otherwise every thread issues
    // 16 / 8 / 4-byte cp.async copies as before.
    __shared__ unsigned long long bar;
    const bool bulk = (a.inW % 2 == 0) && (((size_t)a.x & 15) == 0) &&
                      (!a.doe.hmap || (a.inW % 4 == 0 && ((size_t)a.doe.hmap & 15) == 0 && (a.doe.hstride % 4) == 0));
    const int total_lines = a.nbc * a.inH;
    auto stage_bulk = [&](int g) {                     // called by thread 0 only
        const int gl0 = g * LINES;
        const int nvalid = min(LINES, total_lines - gl0);
        const unsigned xbytes = (unsigned)nvalid * a.inW * sizeof(cpx);
        const unsigned hbytes = a.doe.hmap ? (unsigned)nvalid * a.inW * sizeof(float) : 0u;
        thz_mbar_expect_tx(&bar, xbytes + hbytes);
        thz_bulk_g2s(xs, a.x + (size_t)gl0 * a.inW, xbytes, &bar);
        if (a.doe.hmap)
            for (int line = 0; line < nvalid; ++line) {
                const int gl = gl0 + line;
                thz_bulk_g2s(hs + (size_t)line * a.inW, thz_doe_map(a.doe, a.c0, a.C, gl / a.inH) + (size_t)(gl % a.inH) * a.inW,
                             a.inW * (unsigned)sizeof(float), &bar);
            }
    };
    unsigned parity = 0;
    if (bulk) {
        if (tid == 0) thz_mbar_init(&bar, 1);
        __syncthreads();
        if (tid == 0 && grp < ngroups) stage_bulk(grp);
    } else {
        if (grp < ngroups) p2k1_prefetch<N>(a, xs, hs, grp, tid, nt);
        thz_cp_async_commit();
    }
    for (; grp < ngroups; grp += gridDim.x) {
        if (bulk) {
            thz_mbar_wait(&bar, parity);
            parity ^= 1u;
        } else {
            thz_cp_async_wait_all();
        }
        __syncthreads();                      // staging complete; previous group's store has drained the line buffer
        p2k1_first_staged<N, ELEM>(a, s, xs, hs, grp, tid, nt);
        __syncthreads();                      // staging consumed
        const int next = grp + (int)gridDim.x;
        if (bulk) {
            if (tid == 0 && next < ngroups) stage_bulk(next);
        } else {
            if (next < ngroups) p2k1_prefetch<N>(a, xs, hs, next, tid, nt);
            thz_cp_async_commit();
        }
        fwd_rows<N, 1, p2_stages(N), LINES>(s, tid, nt, tws);
        p2k1_store<N>(a, s, grp, tid, nt);
    }
}

template <int N>
__global__ void __launch_bounds__(p2_col_threads(N), p2_min_blocks(N)) thz_p2_k2(const __grid_constant__ ColArgs a) {
    constexpr int COLS = p2_col_cols(N), NS = p2_stages(N);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* s = reinterpret_cast<cpx*>(smem_raw);
    cpx* tws = s + COLS * p2_pitch(N);              // shared-memory copy of the twiddles the in-smem stages use
    const int tid = threadIdx.x, nt = blockDim.x, bx = blockIdx.x, by = blockIdx.y;
    p2_tw_fill<N>(tws, a.tw, tid, nt);
    p2k2_first<N, COLS>(a, s, bx, by, tid, nt);
    __syncthreads();
    fwd_cols<N, 1, NS - 1, COLS>(s, tid, nt, tws);
    p2k2_middle<N, COLS>(a, s, bx, by, tid, nt);
    __syncthreads();
    inv_cols<N, NS - 2, 1, COLS>(s, tid, nt, tws);
    p2k2_last<N, COLS>(a, s, tws, bx, by, tid, nt);
}

// fast path of the column kernel (ColArgs.fast, see thz_asm_p2.cuh): same phases, everything static
template <int N, int TFM>
__global__ void __launch_bounds__(p2_col_threads(N), p2_min_blocks(N)) thz_p2_k2f(const __grid_constant__ ColArgs a) {
    constexpr int COLS = p2_col_cols(N), NS = p2_stages(N), NT = p2_col_threads(N);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* s = reinterpret_cast<cpx*>(smem_raw);
    cpx* tws = s + COLS * p2_pitch(N);
    const int tid = threadIdx.x, bx = blockIdx.x, by = blockIdx.y;
    p2_tw_fill<N>(tws, a.tw, tid, NT);
    p2k2f_first<N, COLS, NT>(a, s, bx, by, tid);
    __syncthreads();
    fwd_cols<N, 1, NS - 1, COLS>(s, tid, NT, tws);
    p2k2f_middle<N, COLS, NT, TFM>(a, s, bx, by, tid);
    __syncthreads();
    inv_cols<N, NS - 2, 1, COLS>(s, tid, NT, tws);
    p2k2f_last<N, COLS, NT>(a, s, tws, bx, by, tid);
}

template <int N, bool ELEM>
__global__ void __launch_bounds__(p2_row_threads(N), p2_min_blocks(N)) thz_p2_k3(const __grid_constant__ RowInvArgs a) {
    constexpr int NACC = p2k3_acc<N>(), LINES = p2_row_lines(N), BUF = LINES * p2_pitch(N);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* s = reinterpret_cast<cpx*>(smem_raw);      // two line buffers: [0, BUF) and [BUF, 2 BUF)
    cpx* tws = s + (p2_row_pipelined(N) ? 2 : 1) * BUF;                  // shared-memory copy of the stage twiddles
    constexpr int nt = p2_row_threads(N);
    const int tid = threadIdx.x, bx = blockIdx.x;
    p2_tw_fill<N>(tws, a.tw, tid, nt);                                  // visible after the first barrier below
    float acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.f;
    const int f_lo = blockIdx.y * a.bc_per_cta;
    const int f_hi = min(a.nbc, f_lo + a.bc_per_cta);
    int cur = 0;
    if constexpr (!p2_row_pipelined(N)) {     // single line buffer
        for (int f = f_lo; f < f_hi; ++f) {
            p2k3_load<N>(a, s, bx, f, tid, nt);
            __syncthreads();
            inv_rows<N, p2_stages(N) - 1, 1, LINES>(s, tid, nt, tws);
            p2k3_last<N, NACC, ELEM>(a, s, tws, bx, f, tid, nt, acc);
            __syncthreads();
        }
        p2k3_flush<N, NACC>(a, bx, tid, nt, acc);
        return;
    }
    if (f_lo < f_hi) p2k3_prefetch<N>(a, s, bx, f_lo, tid, nt);
    thz_cp_async_commit();
    for (int f = f_lo; f < f_hi; ++f, cur ^= 1) {
        cpx* sc = s + cur * BUF;
        thz_cp_async_wait_all();
        __syncthreads();                      // rows of field f have landed; the other buffer is free (its epilogue ran)
        if (f + 1 < f_hi) p2k3_prefetch<N>(a, s + (cur ^ 1) * BUF, bx, f + 1, tid, nt, 0, 2);
#ifndef THZ_NO_EPI_PREFETCH
        p2k3_prefetch_epilogue<N>(a, bx, f, tid, nt);
#endif
        inv_rows<N, p2_stages(N) - 1, p2_stages(N) - 1, LINES>(sc, tid, nt, tws);
        if (f + 1 < f_hi) p2k3_prefetch<N>(a, s + (cur ^ 1) * BUF, bx, f + 1, tid, nt, 1, 2);
        thz_cp_async_commit();
        inv_rows<N, p2_stages(N) - 2, 1, LINES>(sc, tid, nt, tws);
        p2k3_last<N, NACC, ELEM>(a, sc, tws, bx, f, tid, nt, acc);
    }
    p2k3_flush<N, NACC>(a, bx, tid, nt, acc);
}

// ------------------------------------------------------------------------------- dispatch
template <typename K>
static int set_smem_p2(K kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return THZ_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return thz_set_cuda_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize)", e);
    return THZ_OK;
}

template <typename K, typename A>
static int launch_p2(K kernel, const char* name, int cls, dim3 grid, int block, size_t smem, cudaStream_t stream, const A& args) {
    int rc = set_smem_p2(kernel, smem);
    if (rc != THZ_OK) return rc;
    thz_launch_begin(stream, cls);
    kernel<<<grid, block, smem, stream>>>(args);
    thz_launch_end(stream, cls);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return thz_set_cuda_error(name, e);
    return THZ_OK;
}
// one case per entry of THZ_SP_SIZES (thz_asm_p2.cuh), so that thz_sp_instantiated() and the dispatch cannot disagree
#define THZ_P2_SWITCH(KERN, n, cls, grid, block, smem, stream, args)                                        \
    switch (n) {                                                                                            \
        THZ_SP_SIZES(THZ_P2_X)                                                                              \
    default:                                                                                                \
        return thz_set_error(THZ_E_UNSUPPORTED, "static fast path: size not instantiated");                 \
    }

int thz_p2_launch_k1(const RowFwdArgs& a, int grid, int threads, size_t smem, cudaStream_t stream) {
    // persistent CTAs: a few per SM, each walking line groups bx, bx + grid, ... (software pipeline inside)
    const int resident = thz_sm_count() * thz_p2_min_blocks_rt(a.Wp);
    if (grid > resident) grid = resident;
    if (a.elem.mask || a.elem.mul) {       // pointwise elements in front: the instantiation that multiplies on load
#define THZ_P2_X(NN) case NN: return launch_p2(thz_p2_k1<NN, true>, "thz_p2_k1", THZ_KC_ROW_FWD, dim3(grid), threads, smem, stream, a);
        THZ_P2_SWITCH(thz_p2_k1, a.Wp, THZ_KC_ROW_FWD, grid, threads, smem, stream, a)
#undef THZ_P2_X
    }
#define THZ_P2_X(NN) case NN: return launch_p2(thz_p2_k1<NN, false>, "thz_p2_k1", THZ_KC_ROW_FWD, dim3(grid), threads, smem, stream, a);
    THZ_P2_SWITCH(thz_p2_k1, a.Wp, THZ_KC_ROW_FWD, grid, threads, smem, stream, a)
#undef THZ_P2_X
}
template <int N>
static int launch_k2f(const ColArgs& a, int gridx, int gridy, int threads, size_t smem, cudaStream_t stream) {
    if constexpr (p2_k2_fast_ok(N)) {
        if (a.tf.mode == 0) return launch_p2(thz_p2_k2f<N, 0>, "thz_p2_k2f", THZ_KC_COL, dim3(gridx, gridy), threads, smem, stream, a);
        if (a.tf.mode == 1) return launch_p2(thz_p2_k2f<N, 1>, "thz_p2_k2f", THZ_KC_COL, dim3(gridx, gridy), threads, smem, stream, a);
        return launch_p2(thz_p2_k2f<N, 2>, "thz_p2_k2f", THZ_KC_COL, dim3(gridx, gridy), threads, smem, stream, a);
    } else {
        return thz_set_error(THZ_E_UNSUPPORTED, "fast column path: size not served");
    }
}

int thz_p2_launch_k2(const ColArgs& a, int gridx, int gridy, int threads, size_t smem, cudaStream_t stream) {
    if (a.fast) {
#define THZ_P2_X(NN) case NN: return launch_k2f<NN>(a, gridx, gridy, threads, smem, stream);
        THZ_P2_SWITCH(thz_p2_k2f, a.Hp, THZ_KC_COL, dim3(gridx, gridy), threads, smem, stream, a)
#undef THZ_P2_X
    }
#define THZ_P2_X(NN) case NN: return launch_p2(thz_p2_k2<NN>, "thz_p2_k2", THZ_KC_COL, dim3(gridx, gridy), threads, smem, stream, a);
    THZ_P2_SWITCH(thz_p2_k2, a.Hp, THZ_KC_COL, dim3(gridx, gridy), threads, smem, stream, a)
#undef THZ_P2_X
}
int thz_p2_launch_k3(const RowInvArgs& a, int gridx, int gridy, int threads, size_t smem, cudaStream_t stream) {
    if (a.elem.mask || a.elem.mul) {       // adjoint of pointwise elements: conjugate multiply in the epilogue
#define THZ_P2_X(NN) case NN: return launch_p2(thz_p2_k3<NN, true>, "thz_p2_k3", THZ_KC_ROW_INV, dim3(gridx, gridy), threads, smem, stream, a);
        THZ_P2_SWITCH(thz_p2_k3, a.Wp, THZ_KC_ROW_INV, dim3(gridx, gridy), threads, smem, stream, a)
#undef THZ_P2_X
    }
#define THZ_P2_X(NN) case NN: return launch_p2(thz_p2_k3<NN, false>, "thz_p2_k3", THZ_KC_ROW_INV, dim3(gridx, gridy), threads, smem, stream, a);
    THZ_P2_SWITCH(thz_p2_k3, a.Wp, THZ_KC_ROW_INV, dim3(gridx, gridy), threads, smem, stream, a)
#undef THZ_P2_X
}
