// In-shared-memory mixed-radix FFT engine (sm_100a, also replayable on the host by tests/emul).
//
// Formulation.  A line of n complex points lives in shared memory and is transformed IN PLACE by a
// decimation-in-frequency (Gentleman-Sande) pass per radix R_s: butterfly (b, j) of stage s reads the
// R_s points  p = b*L_s + j + t*M_s  (L_s = n / (R_0..R_{s-1}), M_s = L_s / R_s), takes an R_s-point
// DFT in registers, multiplies output q by w_{L_s}^{j q} and writes the same R_s slots back.  Reads
// and writes of a butterfly hit the same slots, so a stage has no cross-thread hazard and any number
// of butterflies per thread is legal; one __syncthreads separates stages.  The spectrum comes out in
// mixed-radix digit-reversed order:  slot  p = sum_s q_s M_s  holds bin  k = sum_s q_s (R_0..R_{s-1}).
// The inverse runs the exact mirror image (decimation in time, stages in reverse, conjugate twiddles
// before the butterfly), so a scrambled spectrum goes back to a natural-order signal and the
// pipeline FFT -> pointwise multiply -> iFFT never needs a reordering pass.  All four fftshifts and
// both 'ortho' scalings of the reference (utils/Helper_Functions.py:150) cancel algebraically.
//
// Twiddles: one table  tw[m] = exp(-2 pi i m / n)  per length (built on the host in float64); a
// butterfly loads w^1 and forms w^2..w^{R-1} by a depth-log2(R) product tree in registers.
#pragma once
#include "thz_common.cuh"
#include "thz_fft_consts.cuh"

#define THZ_MAX_STAGES 16

struct FftPlan {
    int n;                        // line length
    int ns;                       // number of stages
    int radix[THZ_MAX_STAGES];    // R_s
    int L[THZ_MAX_STAGES];        // block length entering stage s
    int M[THZ_MAX_STAGES];        // L_s / R_s
    int wt[THZ_MAX_STAGES];       // R_0 * ... * R_{s-1}  (bin weight of digit s; also n / L_s)
    int mshift[THZ_MAX_STAGES];   // log2(M_s) if M_s is a power of two, else -1
    int mixed;                    // 1 if any radix outside {2,4,8,16}
};

// Shared-memory slot of logical position p: one padding slot per 16 keeps the stride-R accesses of
// the late stages (M_s small) off the same banks.
THZ_HD int thz_pad(int p) { return p + (p >> 4); }
THZ_HD int thz_padded_len(int n) { return thz_pad(n - 1) + 1; }

// position -> DFT bin (digit reversal of the plan)
THZ_HD int thz_pos_to_bin(const FftPlan& P, int pos) {
    int bin = 0;
    for (int s = 0; s < P.ns; ++s) {
        int q = pos / P.M[s];
        pos -= q * P.M[s];
        bin += q * P.wt[s];
    }
    return bin;
}

// ---------------------------------------------------------------- multiply by a constant root of unity
// a * exp(-/+ 2 pi i e / R)   (- forward, + inverse); e is a compile-time constant after unrolling.
template <int R, bool INV>
THZ_HD cpx mul_root(cpx a, int e) {
    e %= R;
    if (e == 0) return a;
    if (2 * e == R) return cmake(-a.x, -a.y);
    if (4 * e == R) return INV ? cmul_pi(a) : cmul_mi(a);
    if (4 * e == 3 * R) return INV ? cmul_mi(a) : cmul_pi(a);
    const float c = cw_cos<R>(e), s = cw_sin<R>(e);
    return INV ? cmul(a, cmake(c, s)) : cmulc(a, cmake(c, s));
}

// ---------------------------------------------------------------- register butterflies
template <int R, bool INV>
struct Dft;

template <bool INV>
struct Dft<2, INV> {
    static THZ_HD void run(cpx (&v)[2]) {
        cpx a = cadd(v[0], v[1]), b = csub(v[0], v[1]);
        v[0] = a;
        v[1] = b;
    }
};

template <bool INV>
struct Dft<3, INV> {
    static THZ_HD void run(cpx (&v)[3]) {
        const float S = 0.8660254037844386f;
        cpx t1 = cadd(v[1], v[2]);
        cpx t2 = caxpy(-0.5f, t1, v[0]);
        cpx t3 = cscale(csub(v[1], v[2]), S);
        cpx it3 = INV ? cmul_pi(t3) : cmul_mi(t3);   // -/+ i t3
        v[0] = cadd(v[0], t1);
        v[1] = cadd(t2, it3);
        v[2] = csub(t2, it3);
    }
};

template <bool INV>
struct Dft<4, INV> {
    static THZ_HD void run(cpx (&v)[4]) {
        cpx a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
        cpx c = cadd(v[1], v[3]), d = csub(v[1], v[3]);
        cpx id = INV ? cmul_pi(d) : cmul_mi(d);
        v[0] = cadd(a, c);
        v[2] = csub(a, c);
        v[1] = cadd(b, id);
        v[3] = csub(b, id);
    }
};

template <bool INV>
struct Dft<5, INV> {
    static THZ_HD void run(cpx (&v)[5]) {
        const float C1 = 0.30901699437494745f, C2 = -0.8090169943749475f;
        const float S1 = 0.9510565162951535f, S2 = 0.5877852522924731f;
        cpx t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]);
        cpx t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
        cpx a1 = caxpy(C2, t2, caxpy(C1, t1, v[0]));
        cpx a2 = caxpy(C1, t2, caxpy(C2, t1, v[0]));
        cpx b1 = caxpy(S2, t4, cscale(t3, S1));
        cpx b2 = caxpy(-S1, t4, cscale(t3, S2));
        cpx ib1 = INV ? cmul_pi(b1) : cmul_mi(b1);
        cpx ib2 = INV ? cmul_pi(b2) : cmul_mi(b2);
        v[0] = cadd(v[0], cadd(t1, t2));
        v[1] = cadd(a1, ib1);
        v[4] = csub(a1, ib1);
        v[2] = cadd(a2, ib2);
        v[3] = csub(a2, ib2);
    }
};

// direct O(R^2) butterfly for the rare prime 7
template <bool INV>
struct Dft<7, INV> {
    static THZ_HD void run(cpx (&v)[7]) {
        cpx o[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            cpx acc = v[0];
#pragma unroll
            for (int t = 1; t < 7; ++t) acc = cadd(acc, mul_root<7, INV>(v[t], t * k));
            o[k] = acc;
        }
#pragma unroll
        for (int k = 0; k < 7; ++k) v[k] = o[k];
    }
};

// Cooley-Tukey composition in registers: R = R1*R2, input index R2*n1 + n2, output index k1 + R1*k2.
template <int R1, int R2, bool INV>
THZ_HD void dft_ct(cpx (&v)[R1 * R2]) {
    constexpr int R = R1 * R2;
    cpx y[R];
#pragma unroll
    for (int n2 = 0; n2 < R2; ++n2) {
        cpx t[R1];
#pragma unroll
        for (int n1 = 0; n1 < R1; ++n1) t[n1] = v[R2 * n1 + n2];
        Dft<R1, INV>::run(t);
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) y[k1 * R2 + n2] = mul_root<R, INV>(t[k1], n2 * k1);
    }
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) {
        cpx t[R2];
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) t[n2] = y[k1 * R2 + n2];
        Dft<R2, INV>::run(t);
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) v[k1 + R1 * k2] = t[k2];
    }
}

// ---------------------------------------------------------------- pruned radix-16 butterflies (2x zero padding)
// A line that is zero-padded to twice its length, centred, is live in [N/4, 3N/4): element t of a first-stage butterfly
// (stride N/16) is live iff 4 <= t < 12, for every butterfly -- and likewise only outputs 4..11 of a last inverse
// stage butterfly survive the crop.  In the 4 x 4 Cooley-Tukey form:
//   inputs 0..3, 12..15 zero : every first-layer radix-4 sees (0, a, b, 0): 4 complex adds instead of 8;
//   outputs 4..11 only       : every second-layer radix-4 needs outputs 1 and 2 only: 6 instead of 8.
template <bool INV>
THZ_HD void dft16_half_in(const cpx (&in)[8], cpx (&v)[16]) {       // in[i] = v[4 + i]
    cpx y[16];
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
        const cpx a = in[n2], b = in[4 + n2];                       // n1 = 1 and n1 = 2
        const cpx ia = INV ? cmul_pi(a) : cmul_mi(a);               // w4 a
        cpx t[4];
        t[0] = cadd(a, b);
        t[1] = csub(ia, b);                                         // w4 a + w4^2 b
        t[2] = csub(b, a);                                          // w4^2 a + w4^4 b
        t[3] = csub(cmake(-ia.x, -ia.y), b);                        // w4^3 a + w4^6 b
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) y[k1 * 4 + n2] = mul_root<16, INV>(t[k1], n2 * k1);
    }
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        cpx t[4];
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) t[n2] = y[k1 * 4 + n2];
        Dft<4, INV>::run(t);
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) v[k1 + 4 * k2] = t[k2];
    }
}
template <bool INV>
THZ_HD void dft16_half_out(const cpx (&v)[16], cpx (&out)[8]) {     // out[i] = V[4 + i]
    cpx y[16];
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
        cpx t[4];
#pragma unroll
        for (int n1 = 0; n1 < 4; ++n1) t[n1] = v[4 * n1 + n2];
        Dft<4, INV>::run(t);
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) y[k1 * 4 + n2] = mul_root<16, INV>(t[k1], n2 * k1);
    }
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {                                // outputs k1 + 4 k2 for k2 = 1, 2
        const cpx a = cadd(y[k1 * 4 + 0], y[k1 * 4 + 2]), b = csub(y[k1 * 4 + 0], y[k1 * 4 + 2]);
        const cpx c = cadd(y[k1 * 4 + 1], y[k1 * 4 + 3]), d = csub(y[k1 * 4 + 1], y[k1 * 4 + 3]);
        const cpx id = INV ? cmul_pi(d) : cmul_mi(d);
        out[k1] = cadd(b, id);                                      // k2 = 1 -> V[4 + k1]
        out[4 + k1] = csub(a, c);                                   // k2 = 2 -> V[8 + k1]
    }
}

#define THZ_DFT_CT(R, R1, R2)                                             \
    template <bool INV>                                                   \
    struct Dft<R, INV> {                                                  \
        static THZ_HD void run(cpx (&v)[R]) { dft_ct<R1, R2, INV>(v); }   \
    };
THZ_DFT_CT(6, 2, 3)
THZ_DFT_CT(8, 2, 4)
THZ_DFT_CT(9, 3, 3)
THZ_DFT_CT(10, 2, 5)
THZ_DFT_CT(12, 4, 3)
THZ_DFT_CT(14, 2, 7)
THZ_DFT_CT(15, 3, 5)
THZ_DFT_CT(16, 4, 4)
THZ_DFT_CT(20, 4, 5)
THZ_DFT_CT(25, 5, 5)
#undef THZ_DFT_CT

// ---------------------------------------------------------------- twiddle powers w^0..w^{R-1}
template <int R>
THZ_HD void twiddle_powers(cpx w1, cpx (&w)[R]) {
    w[0] = cmake(1.f, 0.f);
    if (R > 1) w[1] = w1;
#pragma unroll
    for (int q = 2; q < R; ++q) w[q] = cmul(w[(q + 1) >> 1], w[q >> 1]);
}

// Split a butterfly index into (block b, offset j).
THZ_HD void thz_split(int u, int M, int mshift, int& b, int& j) {
    if (mshift >= 0) {
        b = u >> mshift;
        j = u & (M - 1);
    } else {
        b = u / M;
        j = u - b * M;
    }
}

// ---------------------------------------------------------------- one butterfly of one stage, in place
//   s       shared-memory line storage
//   base    slot offset of this line;  pstride  slot stride between consecutive positions
//   u       butterfly index in [0, n/R)
template <int R, bool INV>
THZ_HD void fft_butterfly(cpx* s, int base, int pstride, int u, int M, int mshift, int L, int wt, const cpx* tw) {
    int b, j;
    thz_split(u, M, mshift, b, j);
    const int p0 = b * L + j;
    cpx v[R];
#pragma unroll
    for (int t = 0; t < R; ++t) v[t] = s[base + thz_pad(p0 + t * M) * pstride];
    if (!INV) {
        Dft<R, false>::run(v);
        if (M > 1) {
            cpx w[R];
            twiddle_powers<R>(thz_ldg(tw + j * wt), w);
#pragma unroll
            for (int q = 1; q < R; ++q) v[q] = cmul(v[q], w[q]);
        }
    } else {
        if (M > 1) {
            cpx w[R];
            twiddle_powers<R>(cconj(thz_ldg(tw + j * wt)), w);
#pragma unroll
            for (int q = 1; q < R; ++q) v[q] = cmul(v[q], w[q]);
        }
        Dft<R, true>::run(v);
    }
#pragma unroll
    for (int t = 0; t < R; ++t) s[base + thz_pad(p0 + t * M) * pstride] = v[t];
}

// Runtime radix dispatch (block-uniform branch).  MIXED=false instantiates only the power-of-two
// butterflies so that the pow-2 kernels keep a small register footprint.
template <bool MIXED, bool INV>
THZ_HD void fft_stage_butterfly(const FftPlan& P, int st, cpx* s, int base, int pstride, int u, const cpx* tw) {
    const int M = P.M[st], ms = P.mshift[st], L = P.L[st], wt = P.wt[st];
    switch (P.radix[st]) {
    case 16: fft_butterfly<16, INV>(s, base, pstride, u, M, ms, L, wt, tw); break;
    case 8: fft_butterfly<8, INV>(s, base, pstride, u, M, ms, L, wt, tw); break;
    case 4: fft_butterfly<4, INV>(s, base, pstride, u, M, ms, L, wt, tw); break;
    case 2: fft_butterfly<2, INV>(s, base, pstride, u, M, ms, L, wt, tw); break;
    default:
        if (MIXED) {
            switch (P.radix[st]) {
            case 25: fft_butterfly<25, INV>(s, base, pstride, u, M, ms, L, wt, tw); break;
            case 20: fft_butterfly<20, INV>(s, base, pstride, u, M, ms, L, wt, tw); break;
            case 15: fft_butterfly<15, INV>(s, base, pstride, u, M, ms, L, wt, tw); break;
            case 14: fft_butterfly<14, INV>(s, base, pstride, u, M, ms, L, wt, tw); break;
            case 12: fft_butterfly<12, INV>(s, base, pstride, u, M, ms, L, wt, tw); break;
            case 10: fft_butterfly<10, INV>(s, base, pstride, u, M, ms, L, wt, tw); break;
            case 9: fft_butterfly<9, INV>(s, base, pstride, u, M, ms, L, wt, tw); break;
            case 7: fft_butterfly<7, INV>(s, base, pstride, u, M, ms, L, wt, tw); break;
            case 6: fft_butterfly<6, INV>(s, base, pstride, u, M, ms, L, wt, tw); break;
            case 5: fft_butterfly<5, INV>(s, base, pstride, u, M, ms, L, wt, tw); break;
            case 3: fft_butterfly<3, INV>(s, base, pstride, u, M, ms, L, wt, tw); break;
            default: break;
            }
        }
        break;
    }
}

// All butterflies of stage `st` for `lines` lines, distributed over `nthreads` threads.
//   col_major=false: line l occupies slots [l*line_pitch, ...), unit position stride (row kernels)
//   col_major=true : position p of line l is slot pad(p)*lines + l (column tiles; l fastest)
template <bool MIXED, bool INV>
THZ_HD void fft_stage_all(const FftPlan& P, int st, cpx* s, int lines, int line_pitch, bool col_major, int tid,
                          int nthreads, const cpx* tw) {
    const int nb = P.n / P.radix[st];
    const int total = nb * lines;
    for (int w = tid; w < total; w += nthreads) {
        int l, u;
        if (col_major) {
            u = w / lines;
            l = w - u * lines;
            fft_stage_butterfly<MIXED, INV>(P, st, s, l, lines, u, tw);
        } else {
            l = w / nb;
            u = w - l * nb;
            fft_stage_butterfly<MIXED, INV>(P, st, s, l * line_pitch, 1, u, tw);
        }
    }
}

// ---------------------------------------------------------------- host-side planning
// Greedy factorisation into the supported radices, largest first.  Returns 0 on success, -1 if n has
// a prime factor > 7 (unsupported by the register butterflies).
static inline int thz_make_plan(int n, FftPlan* P) {
    static const int cand[] = {16, 25, 20, 15, 14, 12, 10, 9, 8, 7, 6, 5, 4, 3, 2};
    if (n < 1) return -1;
    P->n = n;
    P->ns = 0;
    P->mixed = 0;
    int rem = n;
    // pull out powers of two as 16s first when the rest is a pure power of two, otherwise prefer
    // the big mixed radices so that 2000 -> 25*20*4, 3000 -> 25*20*6, 400 -> 25*16.
    int radices[THZ_MAX_STAGES];
    int ns = 0;
    while (rem > 1) {
        int pick = 0;
        bool pow2 = (rem & (rem - 1)) == 0;
        if (pow2) {
            // balance the tail: avoid a trailing radix-2 (e.g. 32 -> 8*4 not 16*2)
            if (rem >= 256 || rem == 16) pick = 16;
            else if (rem == 128) pick = 16;   // 16*8
            else if (rem == 64) pick = 8;     // 8*8
            else if (rem == 32) pick = 8;     // 8*4
            else pick = rem;                  // 8, 4, 2
        } else {
            for (unsigned i = 0; i < sizeof(cand) / sizeof(cand[0]); ++i) {
                if (cand[i] == 16) continue;
                if (rem % cand[i] == 0) {
                    pick = cand[i];
                    break;
                }
            }
            if (!pick) return -1;
        }
        if (ns >= THZ_MAX_STAGES) return -1;
        radices[ns++] = pick;
        rem /= pick;
    }
    if (ns == 0) {   // n == 1: a single trivial stage keeps the kernels uniform
        return -1;
    }
    int L = n, wt = 1;
    for (int s = 0; s < ns; ++s) {
        int R = radices[s];
        P->radix[s] = R;
        P->L[s] = L;
        P->M[s] = L / R;
        P->wt[s] = wt;
        int M = L / R;
        P->mshift[s] = -1;
        if ((M & (M - 1)) == 0) {
            int sh = 0;
            while ((1 << sh) < M) ++sh;
            P->mshift[s] = sh;
        }
        if (!(R == 2 || R == 4 || R == 8 || R == 16)) P->mixed = 1;
        L = M;
        wt *= R;
    }
    P->ns = ns;
    return 0;
}
