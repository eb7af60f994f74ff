// Outer decimation step for lines longer than the in-shared-memory plans hold (> 16384 points): per-point bodies, shared
// by the kernels (thz_split.cu) and the CPU replay (tests/emul/emul.cpp).
//
// A length-N line, N = P M, splits into P interleaved length-M transforms (decimation in frequency):
//     X[P k + a] = sum_{n<M} u_a[n] w_M^{k n},       u_a[n] = sum_{s<P} x[n + s M] w_N^{a (n + s M)},     w_N = exp(-2 pi i / N)
// and back (the adjoint of the same map, divided by P):
//     x[i] = 1/P sum_{a<P} conj(w_N^{a i}) v_a[i mod M],     v_a = inverse length-M transform of X[P . + a].
// In two dimensions u_{ab} and v_{ab} carry one factor per axis.  The canvas is never materialised: the forward map reads the
// live region of the zero-padded field, the backward map writes the cropped region only.
#pragma once
#include "thz_common.cuh"

struct SplitArgs {
    int Hp, Wp;           // canvas size (Pr Mr x Pc Mc)
    int H, W, r0, c0;     // live (pre) / wanted (post) region of the canvas
    const cpx* twr;       // w_Hp^j, j < Hp
    const cpx* twc;       // w_Wp^j, j < Wp
    int conj_tw;          // 1: the conjugate twiddles (inverse transforms)
    float scale;
};

THZ_HD cpx thz_split_tw(const cpx* tw, int idx, int conj) {
    const cpx w = tw[idx];
    return conj ? cconj(w) : w;
}

// one (n, m) of one field: reads the <= PR x PC canvas samples congruent to (n, m) and writes u_ab[n, m] for every (a, b);
// x = the field's live region [H][W], u = the field's stack [PR PC][Mr][Mc]
template <int PR, int PC>
THZ_HD void thz_split_pre_point(const SplitArgs& A, const cpx* x, cpx* u, int n, int m) {
    const int Mr = A.Hp / PR, Mc = A.Wp / PC;
    cpx tmp[PR][PC];
    for (int s = 0; s < PR; ++s) {
        for (int b = 0; b < PC; ++b) tmp[s][b] = cmake(0.f, 0.f);
        const int li = n + s * Mr - A.r0;
        if ((unsigned)li >= (unsigned)A.H) continue;
        for (int t = 0; t < PC; ++t) {
            const int j = m + t * Mc, lj = j - A.c0;
            if ((unsigned)lj >= (unsigned)A.W) continue;
            const cpx v = x[(size_t)li * A.W + lj];
            tmp[s][0] = cadd(tmp[s][0], v);
            for (int b = 1; b < PC; ++b) tmp[s][b] = cadd(tmp[s][b], cmul(v, thz_split_tw(A.twc, (int)(((long long)b * j) % A.Wp), A.conj_tw)));
        }
    }
    for (int a = 0; a < PR; ++a)
        for (int b = 0; b < PC; ++b) {
            cpx acc = cmake(0.f, 0.f);
            for (int s = 0; s < PR; ++s) {
                const int i = n + s * Mr;
                acc = cadd(acc, a == 0 ? tmp[s][b] : cmul(tmp[s][b], thz_split_tw(A.twr, (int)(((long long)a * i) % A.Hp), A.conj_tw)));
            }
            u[((size_t)(a * PC + b) * Mr + n) * Mc + m] = cscale(acc, A.scale);
        }
}

// one wanted sample (region row i, column j) of one field from the stack v [PR PC][Mr][Mc]
template <int PR, int PC>
THZ_HD cpx thz_split_post_point(const SplitArgs& A, const cpx* v, int i, int j) {
    const int Mr = A.Hp / PR, Mc = A.Wp / PC;
    const int I = i + A.r0, J = j + A.c0, n = I % Mr, m = J % Mc;
    cpx acc = cmake(0.f, 0.f);
    for (int a = 0; a < PR; ++a) {
        cpx row = cmake(0.f, 0.f);
        for (int b = 0; b < PC; ++b) {
            const cpx val = v[((size_t)(a * PC + b) * Mr + n) * Mc + m];
            row = cadd(row, b == 0 ? val : cmul(val, thz_split_tw(A.twc, (int)(((long long)b * J) % A.Wp), !A.conj_tw)));
        }
        acc = cadd(acc, a == 0 ? row : cmul(row, thz_split_tw(A.twr, (int)(((long long)a * I) % A.Hp), !A.conj_tw)));
    }
    return cscale(acc, A.scale);
}
