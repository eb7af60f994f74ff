// CPU replay harness for the CUDA kernel bodies (TEST INFRASTRUCTURE, never loaded by the product).
//
// The kernels in quantizationawarethzdoe_b200/csrc are written as __host__ __device__ phase
// functions separated by block barriers.  This file replays them block by block, thread by thread,
// on HOST pointers, with the very same launch planning (thz_asm_host.h), so that index arithmetic,
// digit-reversal bookkeeping, pruning, padding and the epilogues can be checked against the oracle
// in the GPU-less build container (`pytest -m "not gpu"`).  Built by tests/emul/build.py with g++.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../quantizationawarethzdoe_b200/csrc/thz_asm_host.h"

template <typename F>
static void for_threads(int nthreads, F f) {
    for (int t = 0; t < nthreads; ++t) f(t);
}

template <bool MIXED>
static void run_k1(const RowFwdArgs& a, int grid, int nt, size_t smem) {
    std::vector<unsigned char> buf(smem);
    cpx* s = (cpx*)buf.data();
    const int pitch = thz_padded_len(a.Wp);
    for (int bx = 0; bx < grid; ++bx) {
        for_threads(nt, [&](int t) { k1_load(a, s, bx, t, nt); });
        for (int st = 0; st < a.plan.ns; ++st)
            for_threads(nt, [&](int t) { fft_stage_all<MIXED, false>(a.plan, st, s, a.lines, pitch, false, t, nt, a.tw); });
        for_threads(nt, [&](int t) { k1_store(a, s, bx, t, nt); });
    }
}

template <bool MIXED>
static void run_k2(const ColArgs& a, int gx, int gy, int nt, size_t smem) {
    std::vector<unsigned char> buf(smem);
    cpx* s = (cpx*)buf.data();
    const int last = a.plan.ns - 1;
    for (int by = 0; by < gy; ++by)
        for (int bx = 0; bx < gx; ++bx) {
            for_threads(nt, [&](int t) { k2_load(a, s, bx, by, t, nt); });
            for (int st = 0; st < last; ++st)
                for_threads(nt, [&](int t) { fft_stage_all<MIXED, false>(a.plan, st, s, a.cols, 0, true, t, nt, a.tw); });
            for_threads(nt, [&](int t) { k2_middle<MIXED>(a, s, bx, by, t, nt); });
            for (int st = last - 1; st >= 0; --st)
                for_threads(nt, [&](int t) { fft_stage_all<MIXED, true>(a.plan, st, s, a.cols, 0, true, t, nt, a.tw); });
            for_threads(nt, [&](int t) { k2_store(a, s, bx, by, t, nt); });
        }
}

template <bool MIXED>
static void run_k3(const RowInvArgs& a, int gx, int gy, int nt, size_t smem) {
    std::vector<unsigned char> buf(smem);
    cpx* s = (cpx*)buf.data();
    const int pitch = thz_padded_len(a.Wp);
    typedef float acc_t[THZ_K3_OWN];
    std::vector<float> accs((size_t)nt * THZ_K3_OWN);
    for (int by = 0; by < gy; ++by)
        for (int bx = 0; bx < gx; ++bx) {
            std::fill(accs.begin(), accs.end(), 0.f);
            const int f_lo = by * a.bc_per_cta;
            const int f_hi = thz_imin(a.nbc, f_lo + a.bc_per_cta);
            for (int f = f_lo; f < f_hi; ++f) {
                for_threads(nt, [&](int t) { k3_load(a, s, bx, f, t, nt); });
                for (int st = a.plan.ns - 1; st >= 0; --st)
                    for_threads(nt, [&](int t) { fft_stage_all<MIXED, true>(a.plan, st, s, a.lines, pitch, false, t, nt, a.tw); });
                for_threads(nt, [&](int t) { k3_epilogue(a, s, bx, f, t, nt, *(acc_t*)&accs[(size_t)t * THZ_K3_OWN]); });
            }
            for_threads(nt, [&](int t) { k3_flush(a, bx, t, nt, *(acc_t*)&accs[(size_t)t * THZ_K3_OWN]); });
        }
}

template <bool MIXED>
static void run_k2f(const ColFftArgs& a, int gx, int gy, int nt, size_t smem) {
    std::vector<unsigned char> buf(smem);
    cpx* s = (cpx*)buf.data();
    for (int by = 0; by < gy; ++by)
        for (int bx = 0; bx < gx; ++bx) {
            for_threads(nt, [&](int t) { k2f_load(a, s, bx, by, t, nt); });
            for (int st = 0; st < a.plan.ns; ++st)
                for_threads(nt, [&](int t) { fft_stage_all<MIXED, false>(a.plan, st, s, a.cols, 0, true, t, nt, a.tw); });
            for_threads(nt, [&](int t) { k2f_store(a, s, bx, by, t, nt); });
        }
}


// ------------------------------------------------------------------------------- power-of-two fast path replay
template <int N, int S, int S1, int LINES>
static void e_fwd_rows(cpx* s, int nt, const cpx* tw) {
    if constexpr (S < S1) {
        for_threads(nt, [&](int t) { p2_stage_rows<N, S, false, LINES>(s, t, nt, tw); });
        e_fwd_rows<N, S + 1, S1, LINES>(s, nt, tw);
    }
}
template <int N, int S, int S0, int LINES>
static void e_inv_rows(cpx* s, int nt, const cpx* tw) {
    if constexpr (S >= S0) {
        for_threads(nt, [&](int t) { p2_stage_rows<N, S, true, LINES>(s, t, nt, tw); });
        e_inv_rows<N, S - 1, S0, LINES>(s, nt, tw);
    }
}
template <int N, int S, int S1, int COLS>
static void e_fwd_cols(cpx* s, int nt, const cpx* tw) {
    if constexpr (S < S1) {
        for_threads(nt, [&](int t) { p2_stage_cols<N, S, false, COLS, p2_col_threads(N)>(s, t, nt, tw); });
        e_fwd_cols<N, S + 1, S1, COLS>(s, nt, tw);
    }
}
template <int N, int S, int S0, int COLS>
static void e_inv_cols(cpx* s, int nt, const cpx* tw) {
    if constexpr (S >= S0) {
        for_threads(nt, [&](int t) { p2_stage_cols<N, S, true, COLS, p2_col_threads(N)>(s, t, nt, tw); });
        e_inv_cols<N, S - 1, S0, COLS>(s, nt, tw);
    }
}

template <int N>
static void run_p2_k1(const RowFwdArgs& a, int grid, int nt, size_t smem) {
    std::vector<cpx> tws(p2_tw_count(N));            // padded twiddle copy, as the kernels keep it in shared memory
    p2_tw_fill<N>(tws.data(), a.tw, 0, 1);
    constexpr int LINES = p2_row_lines(N);
    std::vector<unsigned char> buf(smem);
    cpx* s = (cpx*)buf.data();
    cpx* xs = s + LINES * p2_pitch(N) + p2_tw_count(N);
    float* hs = (float*)(xs + (size_t)LINES * a.inW);
    const int ngroups = (a.nbc * a.inH + LINES - 1) / LINES;
    const int g = grid < 5 ? grid : 5;   // a few persistent "CTAs", each walking groups bx, bx+g, ...
    for (int bx = 0; bx < g; ++bx) {
        int grp = bx;
        if constexpr (!p2_row_pipelined(N)) {
            for (; grp < ngroups; grp += g) {
                if (a.doe.lphase) for_threads(nt, [&](int t) { p2k1_first<N, 2>(a, s, grp, t, nt); });
                else for_threads(nt, [&](int t) { p2k1_first<N, 1>(a, s, grp, t, nt); });
                e_fwd_rows<N, 1, p2_stages(N), LINES>(s, nt, tws.data());
                for_threads(nt, [&](int t) { p2k1_store<N>(a, s, grp, t, nt); });
            }
            continue;
        }
        if (grp < ngroups) for_threads(nt, [&](int t) { p2k1_prefetch<N>(a, xs, hs, grp, t, nt); });
        for (; grp < ngroups; grp += g) {
            // the replay copies synchronously, so the "next" prefetch must not clobber the staging buffer before it is consumed
            if (a.doe.lphase) for_threads(nt, [&](int t) { p2k1_first_staged<N, 2>(a, s, xs, hs, grp, t, nt); });
            else for_threads(nt, [&](int t) { p2k1_first_staged<N, 1>(a, s, xs, hs, grp, t, nt); });
            if (grp + g < ngroups) for_threads(nt, [&](int t) { p2k1_prefetch<N>(a, xs, hs, grp + g, t, nt); });
            e_fwd_rows<N, 1, p2_stages(N), LINES>(s, nt, tws.data());
            for_threads(nt, [&](int t) { p2k1_store<N>(a, s, grp, t, nt); });
        }
    }
}
template <int N>
static void run_p2_k2(const ColArgs& a, int gx, int gy, int nt, size_t smem) {
    std::vector<cpx> tws(p2_tw_count(N));            // padded twiddle copy, as the kernels keep it in shared memory
    p2_tw_fill<N>(tws.data(), a.tw, 0, 1);
    constexpr int COLS = p2_col_cols(N), NS = p2_stages(N);
    std::vector<unsigned char> buf(smem);
    cpx* s = (cpx*)buf.data();
    for (int by = 0; by < gy; ++by)
        for (int bx = 0; bx < gx; ++bx) {
            if constexpr (p2_k2_fast_ok(N)) {
                if (a.fast) {          // the specialised column kernel thz_p2_k2f
                    constexpr int NT = p2_col_threads(N);
                    for_threads(nt, [&](int t) { p2k2f_first<N, COLS, NT>(a, s, bx, by, t); });
                    e_fwd_cols<N, 1, NS - 1, COLS>(s, nt, tws.data());
                    if (a.tf.mode == 0) for_threads(nt, [&](int t) { p2k2f_middle<N, COLS, NT, 0>(a, s, bx, by, t); });
                    else if (a.tf.mode == 1) for_threads(nt, [&](int t) { p2k2f_middle<N, COLS, NT, 1>(a, s, bx, by, t); });
                    else for_threads(nt, [&](int t) { p2k2f_middle<N, COLS, NT, 2>(a, s, bx, by, t); });
                    e_inv_cols<N, NS - 2, 1, COLS>(s, nt, tws.data());
                    for_threads(nt, [&](int t) { p2k2f_last<N, COLS, NT>(a, s, tws.data(), bx, by, t); });
                    continue;
                }
            }
            for_threads(nt, [&](int t) { p2k2_first<N, COLS>(a, s, bx, by, t, nt); });
            e_fwd_cols<N, 1, NS - 1, COLS>(s, nt, tws.data());
            for_threads(nt, [&](int t) { p2k2_middle<N, COLS>(a, s, bx, by, t, nt); });
            e_inv_cols<N, NS - 2, 1, COLS>(s, nt, tws.data());
            for_threads(nt, [&](int t) { p2k2_last<N, COLS>(a, s, tws.data(), bx, by, t, nt); });
        }
}
template <int N>
static void run_p2_k3(const RowInvArgs& a, int gx, int gy, int nt, size_t smem) {
    std::vector<cpx> tws(p2_tw_count(N));            // padded twiddle copy, as the kernels keep it in shared memory
    p2_tw_fill<N>(tws.data(), a.tw, 0, 1);
    constexpr int NACC = p2k3_acc<N>(), LINES = p2_row_lines(N), BUF = LINES * p2_pitch(N);
    typedef float acc_t[NACC];
    std::vector<unsigned char> buf(smem);
    cpx* s = (cpx*)buf.data();
    std::vector<float> accs((size_t)nt * NACC);
    for (int by = 0; by < gy; ++by)
        for (int bx = 0; bx < gx; ++bx) {
            std::fill(accs.begin(), accs.end(), 0.f);
            const int f_lo = by * a.bc_per_cta, f_hi = thz_imin(a.nbc, f_lo + a.bc_per_cta);
            int cur = 0;
            if constexpr (!p2_row_pipelined(N)) {
                for (int f = f_lo; f < f_hi; ++f) {
                    for_threads(nt, [&](int t) { p2k3_load<N>(a, s, bx, f, t, nt); });
                    e_inv_rows<N, p2_stages(N) - 1, 1, LINES>(s, nt, tws.data());
                    if (a.doe.lphase) for_threads(nt, [&](int t) { p2k3_last<N, NACC, 2>(a, s, tws.data(), bx, f, t, nt, *(acc_t*)&accs[(size_t)t * NACC]); });
                    else for_threads(nt, [&](int t) { p2k3_last<N, NACC, 1>(a, s, tws.data(), bx, f, t, nt, *(acc_t*)&accs[(size_t)t * NACC]); });
                }
                for_threads(nt, [&](int t) { p2k3_flush<N, NACC>(a, bx, t, nt, *(acc_t*)&accs[(size_t)t * NACC]); });
                continue;
            }
            if constexpr (p2_k3_tma_ok(N)) {
                if (a.t2_perm) {       // thz_p2_k3t: dense staged rows (the bulk copy is a memcpy here) -> first butterfly -> padded line buffer
                    cpx* g = s + BUF;
                    const int r0 = bx * LINES;
                    const int nvalid = (a.outH - r0) < LINES ? (a.outH - r0) : LINES;
                    for (int f = f_lo; f < f_hi; ++f) {
                        memcpy(g, a.T + ((size_t)f * a.rowsT + r0) * N, (size_t)nvalid * N * sizeof(cpx));
                        for_threads(nt, [&](int t) { p2k3_first_from_dense<N>(s, g, t); });
                        e_inv_rows<N, p2_stages(N) - 2, 1, LINES>(s, nt, tws.data());
                        if (a.doe.lphase) for_threads(nt, [&](int t) { p2k3_last<N, NACC, 2>(a, s, tws.data(), bx, f, t, nt, *(acc_t*)&accs[(size_t)t * NACC]); });
                        else for_threads(nt, [&](int t) { p2k3_last<N, NACC, 1>(a, s, tws.data(), bx, f, t, nt, *(acc_t*)&accs[(size_t)t * NACC]); });
                    }
                    for_threads(nt, [&](int t) { p2k3_flush<N, NACC>(a, bx, t, nt, *(acc_t*)&accs[(size_t)t * NACC]); });
                    continue;
                }
            }
            if (f_lo < f_hi) for_threads(nt, [&](int t) { p2k3_prefetch<N>(a, s, bx, f_lo, t, nt); });
            for (int f = f_lo; f < f_hi; ++f, cur ^= 1) {
                cpx* sc = s + cur * BUF;
                // as in thz_p2_k3: the next line is fetched in two portions, around the first in-smem stage
                if (f + 1 < f_hi) for_threads(nt, [&](int t) { p2k3_prefetch<N>(a, s + (cur ^ 1) * BUF, bx, f + 1, t, nt, 0, 2); });
                e_inv_rows<N, p2_stages(N) - 1, p2_stages(N) - 1, LINES>(sc, nt, tws.data());
                if (f + 1 < f_hi) for_threads(nt, [&](int t) { p2k3_prefetch<N>(a, s + (cur ^ 1) * BUF, bx, f + 1, t, nt, 1, 2); });
                e_inv_rows<N, p2_stages(N) - 2, 1, LINES>(sc, nt, tws.data());
                if (a.doe.lphase) for_threads(nt, [&](int t) { p2k3_last<N, NACC, 2>(a, sc, tws.data(), bx, f, t, nt, *(acc_t*)&accs[(size_t)t * NACC]); });
                else for_threads(nt, [&](int t) { p2k3_last<N, NACC, 1>(a, sc, tws.data(), bx, f, t, nt, *(acc_t*)&accs[(size_t)t * NACC]); });
            }
            for_threads(nt, [&](int t) { p2k3_flush<N, NACC>(a, bx, t, nt, *(acc_t*)&accs[(size_t)t * NACC]); });
        }
}

// one case per entry of THZ_SP_SIZES, like the library's own dispatch
#define P2_CASE_X(NN) case NN: P2_FN<NN> P2_ARGS; break;
#define P2_DISPATCH(n)                               \
    switch (n) {                                     \
        THZ_SP_SIZES(P2_CASE_X)                      \
    default: return THZ_E_UNSUPPORTED;               \
    }

// column of slot p in the column-permuted K2 -> K3 intermediate (thz_asm.cuh) and the radix the library would permute with
extern "C" int thz_emul_t2_perm_col(int p, int R, int W) { return thz_t2_perm_col(p, R, W); }
extern "C" int thz_emul_k3_tma_radix(int W) {
    if (!thz_sp_instantiated(W)) return 0;
    int r = 0;
#define THZ_EMUL_R_X(NN) if (W == NN) r = p2_k3_tma_ok(NN) ? p2_radix(NN, p2_stages(NN) - 1) : 0;
    THZ_SP_SIZES(THZ_EMUL_R_X)
#undef THZ_EMUL_R_X
    return r;
}
static int g_last_t2_perm = 0;
// radix of the column permutation the last thz_emul_asm_propagate chunk ran with (0: natural column order)
extern "C" int thz_emul_last_t2_perm(void) { return g_last_t2_perm; }

extern "C" int thz_emul_asm_propagate(const thz_asm_desc* d, int sm_count) {
    int rc = thz_asm_validate(d);
    if (rc != THZ_OK) return rc;
    if (d->slab_parts <= 1 && d->ws_bytes < thz_asm_ws_bytes(d)) return THZ_E_WORKSPACE;
    const int nbc_all = d->B * d->C;
    const int chunk = (int)thz_asm_chunk_fields(d);
    const int nchunks = (nbc_all + chunk - 1) / chunk;
    const int stages = d->stages ? d->stages : 7;
    bool zeroed = false;
    for (int f0 = 0; f0 < nbc_all; f0 += chunk) {
        const int nbc = nbc_all - f0 < chunk ? nbc_all - f0 : chunk;
        AsmLaunch L;
        rc = thz_asm_plan_chunk(d, f0, nbc, sm_count, &L);
        if (rc != THZ_OK) return rc;
        if (thz_env_is_1("THZ_EMUL_T2_PERM")) {      // column-permuted K2 -> K3 intermediate, as thz_asm_propagate sets it up
            L.k2.t2_perm = thz_asm_t2_perm_radix(d, &L, stages);
            L.k3.t2_perm = L.k2.t2_perm;
        }
        g_last_t2_perm = L.k3.t2_perm;
        L.k3.gh_atomic = (d->doe_mode == 2 && (nchunks > 1 || L.k3_gridy > 1)) ? 1 : 0;
        if ((stages & 4) && L.k3.gh_atomic && !zeroed) {
            memset(d->doe_gh, 0, (size_t)d->outH * d->outW * sizeof(float));
            zeroed = true;
        }
        if (!(stages & 1)) {
        } else if (L.p2_w) {
#define P2_FN run_p2_k1
#define P2_ARGS (L.k1, L.k1_grid, L.k1_threads, L.k1_smem)
            P2_DISPATCH(d->Wp)
#undef P2_FN
#undef P2_ARGS
        } else if (L.mixed_w) run_k1<true>(L.k1, L.k1_grid, L.k1_threads, L.k1_smem);
        else run_k1<false>(L.k1, L.k1_grid, L.k1_threads, L.k1_smem);
        if (!(stages & 2)) {
        } else if (L.p2_h) {
#define P2_FN run_p2_k2
#define P2_ARGS (L.k2, L.k2_gridx, nbc, L.k2_threads, L.k2_smem)
            P2_DISPATCH(d->Hp)
#undef P2_FN
#undef P2_ARGS
        } else if (L.mixed_h) run_k2<true>(L.k2, L.k2_gridx, nbc, L.k2_threads, L.k2_smem);
        else run_k2<false>(L.k2, L.k2_gridx, nbc, L.k2_threads, L.k2_smem);
        if (!(stages & 4)) {
        } else if (L.p2_w) {
#define P2_FN run_p2_k3
#define P2_ARGS (L.k3, L.k3_gridx, L.k3_gridy, L.k3_threads, L.k3_smem)
            P2_DISPATCH(d->Wp)
#undef P2_FN
#undef P2_ARGS
        } else if (L.mixed_w) run_k3<true>(L.k3, L.k3_gridx, L.k3_gridy, L.k3_threads, L.k3_smem);
        else run_k3<false>(L.k3, L.k3_gridx, L.k3_gridy, L.k3_threads, L.k3_smem);
    }
    return THZ_OK;
}

extern "C" int thz_emul_fft2_c2c(const void* x, void* y, int32_t batch, int32_t H, int32_t W, int32_t inverse,
                                 int32_t ortho, const void* tw_h, const void* tw_w, void* ws) {
    FftPlan pw, ph;
    if (thz_make_plan(W, &pw) != 0 || thz_make_plan(H, &ph) != 0) return THZ_E_UNSUPPORTED;
    const size_t lw = (size_t)thz_padded_len(W) * sizeof(cpx), lh = (size_t)thz_padded_len(H) * sizeof(cpx);
    RowFwdArgs a1;
    memset(&a1, 0, sizeof(a1));
    a1.x = (const cpx*)x;
    a1.T = (cpx*)ws;
    a1.nbc = batch;
    a1.rowsT = H;
    a1.C = 1;
    a1.inH = H;
    a1.inW = W;
    a1.Wp = W;
    a1.plan = pw;
    a1.tw = (const cpx*)tw_w;
    a1.conj_in = inverse ? 1 : 0;
    int lines = thz_imax(1, 4096 / W);
    lines = thz_imin(lines, 16);
    a1.lines = lines;
    if (pw.mixed) run_k1<true>(a1, (batch * H + lines - 1) / lines, 256, lines * lw);
    else run_k1<false>(a1, (batch * H + lines - 1) / lines, 256, lines * lw);
    ColFftArgs a2;
    memset(&a2, 0, sizeof(a2));
    a2.T = (const cpx*)ws;
    a2.y = (cpx*)y;
    a2.H = H;
    a2.W = W;
    int cols = 16;
    while (cols > 1 && cols * lh > 72 * 1024) cols >>= 1;
    a2.cols = cols;
    a2.scale = (float)(ortho ? 1.0 / sqrt((double)H * (double)W) : (inverse ? 1.0 / ((double)H * (double)W) : 1.0));
    a2.conj_out = inverse ? 1 : 0;
    a2.plan = ph;
    a2.planW = pw;
    a2.tw = (const cpx*)tw_h;
    if (ph.mixed) run_k2f<true>(a2, (W + cols - 1) / cols, batch, 128, cols * lh);
    else run_k2f<false>(a2, (W + cols - 1) / cols, batch, 128, cols * lh);
    return THZ_OK;
}

// host-side planning helpers, same as the product exports them (thz_api.cu) but without CUDA
extern "C" int thz_emul_slot_to_bin(int32_t n, int32_t* out) {
    FftPlan P;
    if (thz_make_plan(n, &P) != 0) return THZ_E_UNSUPPORTED;
    for (int p = 0; p < n; ++p) out[p] = thz_pos_to_bin(P, p);
    return THZ_OK;
}
extern "C" int thz_emul_plan_info(int32_t n, int32_t* radices, int32_t* ns) {
    FftPlan P;
    if (thz_make_plan(n, &P) != 0) return THZ_E_UNSUPPORTED;
    for (int s = 0; s < THZ_MAX_STAGES; ++s) radices[s] = s < P.ns ? P.radix[s] : 0;
    *ns = P.ns;
    return THZ_OK;
}

// thickness-space softmax quantizer (thz_doe.cuh thz_softmaxq_pixel), forward + backward, on host pointers
#include "../../quantizationawarethzdoe_b200/csrc/thz_doe.cuh"
extern "C" int thz_emul_softmaxq(const float* t, const float* lut, int32_t L, const float* noise, float c, float tau, float s,
                                 int32_t hard, float* q, int32_t* idx, float* A, float* Bm, float* E, float* stats, uint64_t n) {
    float m = 0.f;
    for (uint64_t i = 0; i < n; ++i)
        for (int j = 0; j < L; ++j) m = fmaxf(m, fabsf(thz_sub_rn(t[i], lut[j])));
    SoftmaxQParams P;
    P.m = m;
    P.s = s;
    P.c = c;
    P.tau = tau;
    P.L = L;
    P.hard = hard;
    P.gumbel = noise ? 1 : 0;
    int ties = 0;
    for (uint64_t i = 0; i < n; ++i) {
        int nt;
        idx[i] = thz_softmaxq_pixel(t[i], lut, noise ? noise + i : nullptr, n, P, q + i, A + i, Bm + i, E + i, &nt);
        ties += nt;
    }
    stats[0] = m;
    stats[1] = (float)ties;
    return THZ_OK;
}
extern "C" int thz_emul_softmaxq_bwd(const float* g, const float* A, const float* Bm, const float* E, float* stats, float* gt,
                                     uint64_t n) {
    double S = 0.0;
    for (uint64_t i = 0; i < n; ++i) S += (double)g[i] * Bm[i];
    stats[2] = (float)S;
    const float share = stats[1] > 0.f ? stats[2] / stats[1] : 0.f;
    for (uint64_t i = 0; i < n; ++i) gt[i] = g[i] * A[i] + share * E[i];
    return THZ_OK;
}
extern "C" int thz_emul_score_thickness(const float* t, const float* lut, int32_t L, float s, int32_t func, float* scores,
                                        int32_t batch, uint64_t n_per_b) {
    const uint64_t total = (uint64_t)batch * n_per_b;
    float m = 0.f;
    for (uint64_t i = 0; i < total; ++i)
        for (int j = 0; j < L; ++j) m = fmaxf(m, fabsf(thz_sub_rn(t[i], lut[j])));
    for (uint64_t i = 0; i < total; ++i) {
        const uint64_t b = i / n_per_b, p = i - b * n_per_b;
        for (int j = 0; j < L; ++j) scores[(b * L + j) * n_per_b + p] = thz_score_value(thz_sub_rn(t[i], lut[j]) / m, s, func);
    }
    return THZ_OK;
}

// outer decimation step for lines above 16384 points (thz_split.cuh bodies), on host pointers
#include "../../quantizationawarethzdoe_b200/csrc/thz_split.cuh"
template <int PR, int PC>
static void emul_split_pre(const SplitArgs& A, const cpx* x, cpx* u, int F) {
    for (int f = 0; f < F; ++f)
        for (int n = 0; n < A.Hp / PR; ++n)
            for (int m = 0; m < A.Wp / PC; ++m)
                thz_split_pre_point<PR, PC>(A, x + (size_t)f * A.H * A.W, u + (size_t)f * A.Hp * A.Wp, n, m);
}
template <int PR, int PC>
static void emul_split_post(const SplitArgs& A, const cpx* v, cpx* y, int F) {
    for (int f = 0; f < F; ++f)
        for (int i = 0; i < A.H; ++i)
            for (int j = 0; j < A.W; ++j)
                y[((size_t)f * A.H + i) * A.W + j] = thz_split_post_point<PR, PC>(A, v + (size_t)f * A.Hp * A.Wp, i, j);
}
#define EMUL_SPLIT_DISPATCH(FN, ...)                    \
    switch (Pr * 8 + Pc) {                              \
        case 1 * 8 + 1: FN<1, 1>(__VA_ARGS__); break;   \
        case 1 * 8 + 2: FN<1, 2>(__VA_ARGS__); break;   \
        case 1 * 8 + 4: FN<1, 4>(__VA_ARGS__); break;   \
        case 2 * 8 + 1: FN<2, 1>(__VA_ARGS__); break;   \
        case 2 * 8 + 2: FN<2, 2>(__VA_ARGS__); break;   \
        case 2 * 8 + 4: FN<2, 4>(__VA_ARGS__); break;   \
        case 4 * 8 + 1: FN<4, 1>(__VA_ARGS__); break;   \
        case 4 * 8 + 2: FN<4, 2>(__VA_ARGS__); break;   \
        case 4 * 8 + 4: FN<4, 4>(__VA_ARGS__); break;   \
        default: return THZ_E_UNSUPPORTED;              \
    }
extern "C" int thz_emul_split_pre(const void* x, void* u, int32_t F, int32_t H, int32_t W, int32_t r0, int32_t c0, int32_t Hp,
                                  int32_t Wp, int32_t Pr, int32_t Pc, const void* twr, const void* twc, int32_t conj_tw, float scale) {
    const SplitArgs A = {Hp, Wp, H, W, r0, c0, (const cpx*)twr, (const cpx*)twc, conj_tw ? 1 : 0, scale};
    EMUL_SPLIT_DISPATCH(emul_split_pre, A, (const cpx*)x, (cpx*)u, F)
    return THZ_OK;
}
extern "C" int thz_emul_split_post(const void* v, void* y, int32_t F, int32_t H, int32_t W, int32_t r0, int32_t c0, int32_t Hp,
                                   int32_t Wp, int32_t Pr, int32_t Pc, const void* twr, const void* twc, int32_t conj_tw, float scale) {
    const SplitArgs A = {Hp, Wp, H, W, r0, c0, (const cpx*)twr, (const cpx*)twc, conj_tw ? 1 : 0, scale};
    EMUL_SPLIT_DISPATCH(emul_split_post, A, (const cpx*)v, (cpx*)y, F)
    return THZ_OK;
}
