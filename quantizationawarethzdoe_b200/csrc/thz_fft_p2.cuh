// Compile-time specialised power-of-two FFT stages (fast path for N = 256 ... 16384).
//
// Same algorithm and the same shared-memory layout as thz_fft.cuh (in-place DIF forward / DIT inverse,
// digit-reversed spectra, one padding slot per 16), but the line length N is a template parameter, so
//   * radices, strides and twiddle steps are constants: a butterfly addresses its R slots as
//     base + immediate offsets (no per-element index arithmetic, no integer divisions);
//   * the first forward stage can take its inputs straight from global memory and the last inverse
//     stage can hand its outputs straight to global memory (functor arguments), which removes one
//     shared-memory round trip and one barrier at each end of a kernel.
// The ncu profile of the generic engine showed ~38 % of executed instructions to be integer/index work
// (profiles/r01_*): this file exists to get rid of them.  thz_make_plan() (host) and p2_pick() (here)
// must agree on the factorisation; tests/emul checks that.
#pragma once
#include "thz_fft.cuh"

// ---------------------------------------------------------------- compile-time plan (mirrors thz_make_plan)
// (the name p2_* is historical: the same machinery serves every length whose plan keeps the padded-slot offsets
//  of a butterfly compile-time constants, see sp_static_ok below -- powers of two and e.g. 400 = 25*16, 2000 = 25*20*4)
THZ_HD constexpr int p2_pick(int rem) {
    if ((rem & (rem - 1)) == 0)
        return (rem >= 256 || rem == 16) ? 16 : (rem == 128 ? 16 : (rem == 64 ? 8 : (rem == 32 ? 8 : rem)));
    constexpr int cand[14] = {25, 20, 15, 14, 12, 10, 9, 8, 7, 6, 5, 4, 3, 2};
    for (int i = 0; i < 14; ++i)
        if (rem % cand[i] == 0) return cand[i];
    return rem;
}
THZ_HD constexpr int p2_L(int N, int s) {   // block length entering stage s
    int rem = N;
    for (int i = 0; i < s; ++i) rem /= p2_pick(rem);
    return rem;
}
THZ_HD constexpr int p2_radix(int N, int s) { return p2_L(N, s) > 1 ? p2_pick(p2_L(N, s)) : 1; }
THZ_HD constexpr int p2_stages(int N) {
    int rem = N, k = 0;
    while (rem > 1) {
        rem /= p2_pick(rem);
        ++k;
    }
    return k;
}
// slot offset of element t of a butterfly with sub-block length M (relative to the slot of element 0)
THZ_HD constexpr int p2_coff(int M, int t) { return t * M + ((t * M) >> 4); }

// A length can use the static path iff, in every stage, the padded slot of element t of a butterfly is
// slot(element 0) + a constant: that needs  (p0 & 15) + ((t M) & 15) < 16  for every butterfly, which holds when the
// block length L and the sub-block length M are each a multiple of 16 or a divisor of 16.
THZ_HD constexpr bool sp_len_ok(int v) { return (v % 16 == 0) || (v < 16 && 16 % v == 0); }
THZ_HD constexpr bool sp_static_ok(int N) {
    for (int s = 0; s < p2_stages(N); ++s) {
        const int L = p2_L(N, s), R = p2_radix(N, s);
        if (R > 25 || !sp_len_ok(L) || !sp_len_ok(L / R)) return false;
    }
    return p2_stages(N) >= 2;
}

// Twiddles of every stage but the first are read from a shared-memory copy of tw[0 .. p2_tw_count(N)): stage s uses
// tw[j * N / L_s] with j < L_s / R_s, i.e. indices below N / R_s.  (An inverse stage multiplies BEFORE its butterfly,
// so nothing hides the load: from L2 it was ~7 % of the column kernel's stall samples, profiles/README.md.)
//   The copy is padded like the data (entry i at i + (i >> 4)): a later stage reads entries j * WT with WT a multiple of
//   16, i.e. 128-byte strides that would all fall on one bank.
THZ_HD constexpr int p2_twi(int i) { return i + (i >> 4); }
THZ_HD constexpr int p2_tw_entries(int N) {      // table entries copied
    int n = 1;
    for (int s = 0; s < p2_stages(N); ++s) {
        const int L = p2_L(N, s), R = p2_radix(N, s);
        if (L / R > 1 && N / R > n) n = N / R;
    }
    return n;
}
THZ_HD constexpr int p2_tw_count(int N) { return (p2_twi(p2_tw_entries(N) - 1) + 2) & ~1; }   // shared-memory slots (even: what follows stays 16-byte aligned)
// fill (all threads of the CTA / the host replay): tws[p2_twi(i)] = tw[i]
template <int N>
THZ_HD void p2_tw_fill(cpx* tws, const cpx* tw, int tid, int nthreads) {
    for (int i = tid; i < p2_tw_entries(N); i += nthreads) tws[p2_twi(i)] = tw[i];
}

template <int N, int S>
struct P2Stage {
    static_assert(sp_static_ok(N), "length cannot use the static-offset FFT path");
    static constexpr int R = p2_radix(N, S);
    static constexpr int L = p2_L(N, S);
    static constexpr int M = L / R;
    static constexpr int WT = N / L;      // twiddle table step and bin weight of this digit
    static constexpr int NB = N / R;      // butterflies per line
};

template <int R>
THZ_HD void p2_apply_twiddles(cpx (&v)[R], cpx w1) {
    cpx w[R];
    twiddle_powers<R>(w1, w);
#pragma unroll
    for (int q = 1; q < R; ++q) v[q] = cmul(v[q], w[q]);
}

// ---------------------------------------------------------------- in-shared-memory stage S, one butterfly
//   base: pointer to slot 0 of this line (row-major, STRIDE = 1) or to column l of slot 0 (column tile,
//   STRIDE = number of columns in the tile)
template <int N, int S, bool INV, int STRIDE>
THZ_HD void p2_butterfly(cpx* base, int u, const cpx* tw) {
    typedef P2Stage<N, S> St;
    constexpr int R = St::R, M = St::M, L = St::L;
    const int b = u / M, j = u % M;       // M is a power of two: shift / mask
    const int p0 = b * L + j;
    cpx* p = base + (p0 + (p0 >> 4)) * STRIDE;
    cpx v[R];
#pragma unroll
    for (int t = 0; t < R; ++t) v[t] = p[p2_coff(M, t) * STRIDE];
    if (!INV) {
        Dft<R, false>::run(v);
        if (M > 1) p2_apply_twiddles<R>(v, tw[p2_twi(j * St::WT)]);  // tw: the padded shared-memory copy
    } else {
        if (M > 1) p2_apply_twiddles<R>(v, cconj(tw[p2_twi(j * St::WT)]));
        Dft<R, true>::run(v);
    }
#pragma unroll
    for (int t = 0; t < R; ++t) p[p2_coff(M, t) * STRIDE] = v[t];
}

// Same butterfly with the twiddle powers w[q] = w1^q (already conjugated for INV) supplied by the caller: when the
// butterflies a thread runs in one stage all share their sub-block offset j, the power tree is built once per stage.
template <int N, int S, bool INV, int STRIDE>
THZ_HD void p2_butterfly_w(cpx* base, int u, const cpx (&w)[P2Stage<N, S>::R]) {
    typedef P2Stage<N, S> St;
    constexpr int R = St::R, M = St::M, L = St::L;
    const int b = u / M, j = u % M;
    const int p0 = b * L + j;
    cpx* p = base + (p0 + (p0 >> 4)) * STRIDE;
    cpx v[R];
#pragma unroll
    for (int t = 0; t < R; ++t) v[t] = p[p2_coff(M, t) * STRIDE];
    if (!INV) {
        Dft<R, false>::run(v);
#pragma unroll
        for (int q = 1; q < R; ++q) v[q] = cmul(v[q], w[q]);
    } else {
#pragma unroll
        for (int q = 1; q < R; ++q) v[q] = cmul(v[q], w[q]);
        Dft<R, true>::run(v);
    }
#pragma unroll
    for (int t = 0; t < R; ++t) p[p2_coff(M, t) * STRIDE] = v[t];
}

// ---------------------------------------------------------------- first forward stage, inputs from a functor
//   load(pos) returns the input at logical position pos in [0, N).  Stage 0 has a single block (L = N), so
//   butterfly j touches positions j + t*M.
template <int N, int STRIDE, bool HALF = false, typename Load>
THZ_HD void p2_first_stage_from(cpx* base, int j, const cpx* __restrict__ tw, Load load) {
    typedef P2Stage<N, 0> St;
    constexpr int R = St::R, M = St::M;
    cpx v[R];
    if constexpr (HALF && R == 16) {        // centred 2x padding: only t = 4..11 are live, no bounds checks (load.live)
        cpx in[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) in[t] = load.live(j + (4 + t) * M);
        dft16_half_in<false>(in, v);
    } else if constexpr (HALF && R == 25 && M % 4 == 0) {
        // centred 2x padding with a radix-25 first stage (N = 25 M: 400, 800, 1600, 2000, ...): element t is live iff
        // 6.25 M <= j + t M < 18.75 M -- t = 7..17 always, t = 6 iff j >= M / 4, t = 18 iff j < 3 M / 4, the other 12 never:
        // no bounds checks, no loads of padding (the butterfly itself is not pruned: a 3-of-5 sparse DFT5 costs what the
        // full one does)
#pragma unroll
        for (int t = 0; t < R; ++t) {
            if (t >= 7 && t <= 17) v[t] = load.live(j + t * M);
            else if (t == 6) v[t] = (j >= M / 4) ? load.live(j + t * M) : cmake(0.f, 0.f);
            else if (t == 18) v[t] = (j < 3 * M / 4) ? load.live(j + t * M) : cmake(0.f, 0.f);
            else v[t] = cmake(0.f, 0.f);
        }
        Dft<R, false>::run(v);
    } else {
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = load(j + t * M);
        Dft<R, false>::run(v);
    }
    p2_apply_twiddles<R>(v, thz_ldg(tw + j));
    cpx* p = base + (j + (j >> 4)) * STRIDE;
#pragma unroll
    for (int t = 0; t < R; ++t) p[p2_coff(M, t) * STRIDE] = v[t];
}

// ---------------------------------------------------------------- last inverse stage, outputs to a functor
//   A store functor with an epilogue that READS global memory (the DOE adjoint) may define PF > 0 and prefetch(pos, t):
//   the loads of output t + PF are then issued while output t is being finished (software pipeline over the R outputs;
//   after unrolling the functor's PF-deep ring lives in registers).
template <typename Store>
struct P2StorePF {
    template <typename S>
    static constexpr auto get(int) -> decltype(S::PF) { return S::PF; }
    template <typename S>
    static constexpr int get(...) { return 0; }
    static constexpr int value = get<Store>(0);
};
template <int N, int STRIDE, bool HALF = false, typename Store>
THZ_HD void p2_last_inverse_stage_to(const cpx* base, int j, const cpx* tw, Store store) {
    typedef P2Stage<N, 0> St;
    constexpr int R = St::R, M = St::M, PF = P2StorePF<Store>::value;
    const cpx* p = base + (j + (j >> 4)) * STRIDE;
    cpx v[R];
#pragma unroll
    for (int t = 0; t < R; ++t) v[t] = p[p2_coff(M, t) * STRIDE];
    if constexpr (HALF && R == 16) {        // centred crop to half the line: only outputs 4..11 exist (store.live)
        if constexpr (PF > 0) {
#pragma unroll
            for (int t = 0; t < PF && t < 8; ++t) store.prefetch_live(j + (4 + t) * M, t);
        }
        p2_apply_twiddles<R>(v, cconj(tw[p2_twi(j)]));
        cpx out[8];
        dft16_half_out<true>(v, out);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            store.live(j + (4 + t) * M, t, 4 + t, out[t]);
            if constexpr (PF > 0) {
                if (t + PF < 8) store.prefetch_live(j + (4 + t + PF) * M, t + PF);
            }
        }
    } else if constexpr (HALF && R == 25 && M % 4 == 0 && PF <= 1) {
        // centred crop with a radix-25 last stage: outputs t = 7..17 always survive, t = 6 iff j >= M / 4, t = 18 iff
        // j < 3 M / 4, the rest never (see p2_first_stage_from): no bounds checks, no divergent skips
        const bool lo = j >= M / 4, hi = j < 3 * M / 4;
        if constexpr (PF > 0) {
            if (lo) store.prefetch_live(j + 6 * M, 0);
            else store.prefetch_live(j + 7 * M, 0);
        }
        p2_apply_twiddles<R>(v, cconj(tw[p2_twi(j)]));
        Dft<R, true>::run(v);
#pragma unroll
        for (int t = 6; t <= 18; ++t) {
            if (t == 6 && !lo) continue;
            if (t == 18 && !hi) continue;
            store.live(j + t * M, 0, t, v[t]);
            if constexpr (PF > 0) {
                if (t < 17 || (t == 17 && hi)) store.prefetch_live(j + (t + 1) * M, 0);
            }
        }
    } else {
        if constexpr (PF > 0) {
#pragma unroll
            for (int t = 0; t < PF && t < R; ++t) store.prefetch(j + t * M, t);
        }
        p2_apply_twiddles<R>(v, cconj(tw[p2_twi(j)]));
        Dft<R, true>::run(v);
#pragma unroll
        for (int t = 0; t < R; ++t) {
            store(j + t * M, t, v[t]);
            if constexpr (PF > 0) {
                if (t + PF < R) store.prefetch(j + (t + PF) * M, t + PF);
            }
        }
    }
}

// ---------------------------------------------------------------- all butterflies of in-smem stage S
//   ROWS: LINES lines of pitch PITCH slots (row kernels), work item w -> (line = w / NB, u = w % NB)
template <int N, int S, bool INV, int LINES>
THZ_HD void p2_stage_rows(cpx* s, int tid, int nthreads, const cpx* tw) {
    constexpr int NB = P2Stage<N, S>::NB;
    constexpr int PITCH = N + (N >> 4);
    for (int w = tid; w < LINES * NB; w += nthreads) {
        const int line = w / NB, u = w % NB;
        p2_butterfly<N, S, INV, 1>(s + line * PITCH, u, tw);
    }
}
//   COLUMN TILE: COLS lines interleaved (slot * COLS + l), work item w -> (u = w / COLS, l = w % COLS)
//   NT = the kernel's compile-time block size (0: unknown).  A thread's work items are w = tid + k NT, i.e. u = u0 +
//   k NT / COLS: if NT / COLS is a multiple of the sub-block length M, all of them have the same j = u % M and hence
//   the same twiddle powers -- computed once here instead of once per butterfly (14 complex multiplies for radix 16).
template <int N, int S, bool INV, int COLS, int NT = 0>
THZ_HD void p2_stage_cols(cpx* s, int tid, int nthreads, const cpx* tw) {
    typedef P2Stage<N, S> St;
    constexpr int NB = St::NB, M = St::M, R = St::R;
    if constexpr (NT > 0 && M > 1 && NT % COLS == 0 && (NT / COLS) % M == 0 && (COLS * NB) > NT) {
        const cpx w1 = tw[p2_twi(((tid / COLS) % M) * St::WT)];
        cpx w[R];
        twiddle_powers<R>(INV ? cconj(w1) : w1, w);
        for (int w_ = tid; w_ < COLS * NB; w_ += NT) p2_butterfly_w<N, S, INV, COLS>(s + w_ % COLS, w_ / COLS, w);
    } else {
        for (int w = tid; w < COLS * NB; w += nthreads) {
            const int u = w / COLS, l = w % COLS;
            p2_butterfly<N, S, INV, COLS>(s + l, u, tw);
        }
    }
}
