// Static ("p2": historically power-of-two) fast path of the fused ASM pipeline: the same three kernels as thz_asm.cuh
// (same argument structs, same results), with the transform length a template parameter and the first / last FFT
// stage of each kernel fused with its global-memory traffic.  On top of that, where it applies:
//   * the K1 -> K2 intermediate is stored in 4-column blocks and the K2 -> K3 one row-major in a second buffer
//     (thz_t_tiled_index in thz_asm.cuh: the column kernel's loads then touch 4 instead of 16 cache lines per warp);
//   * for centred 2x padding the first forward / last inverse radix-16 stage is pruned (half_in / half_out);
//   * stage twiddles come from a padded shared-memory copy, and a thread's butterflies share one power tree when they
//     can (p2_stage_cols); the row vectors of the transfer function are read in a warp-contiguous "chunked" layout;
//   * the row kernels are software-pipelined with cp.async (K1 persistent; K3 double-buffered, copies issued in portions);
//   * multi-GPU slab FFT: K1 can scatter its rows into, and K3 gather them from, column slabs in peer memory (SlabArgs).
//
//   K1<N>       x row --[DOE phase]--> registers -> stage 0 -> smem -> stages 1.. -> smem -> T row
//   K2<N,COLS>  T column tile -> registers -> stage 0 -> smem -> ... -> (last stage . H . last stage^-1)
//               -> ... -> inverse stage 0 -> registers -> cropped rows of T
//   K3<N>       T row -> smem -> inverse stages ..1 -> inverse stage 0 -> registers -> crop/scale/epilogue -> y
//
// Every phase is a per-thread function; the __global__ wrappers (thz_asm_p2_kernels.inc) and the CPU replay
// (tests/emul) put the barriers between them.
#pragma once
#include "thz_asm.cuh"
#include "thz_fft_p2.cuh"

// lengths with a compiled static path: powers of two 256 .. 16384, the multiples of 16 of the form 25*16*2^a / 25*20*4*2^a
// that the BASELINE configs and the reference notebooks produce (200 -> 400, 1000 -> 2000, ...), and 3 * 2^a (padding_scale 2)
#define THZ_SP_SIZES(X) X(256) X(512) X(1024) X(2048) X(4096) X(8192) X(16384) X(400) X(800) X(1600) X(2000) X(3200) X(4000) X(768) X(1536) X(3072) X(6144)
THZ_HD constexpr bool thz_sp_instantiated(int n) {
#define THZ_SP_CMP(NN) if (n == NN) return true;
    THZ_SP_SIZES(THZ_SP_CMP)
#undef THZ_SP_CMP
    return false;
}

// CTAs per SM the static kernels are compiled for (register cap = 65536 / (threads x this)): 3 (80 registers) for the
// radix-16 lengths; the lengths with a radix-25 stage hold 25 complex values per butterfly and spill at 80 registers
#ifndef THZ_R25_BLOCKS
#define THZ_R25_BLOCKS 3
#endif
THZ_HD constexpr bool p2_has_radix25(int N) {
    for (int s = 0; s < p2_stages(N); ++s)
        if (p2_radix(N, s) == 25) return true;
    return false;
}
THZ_HD constexpr int p2_min_blocks(int N) { return N >= 8192 ? 1 : (p2_has_radix25(N) ? THZ_R25_BLOCKS : 3); }
THZ_HD constexpr int p2_row_lines(int N) { return N >= 4096 ? 1 : 4096 / N; }        // lines per CTA, row kernels
THZ_HD constexpr int p2_row_threads(int N) { return N >= 8192 ? 512 : 256; }
THZ_HD constexpr int p2_col_cols(int N) { return N > 8192 ? 1 : (N > 2048 ? 2 : (N > 1024 ? 4 : (N > 512 ? 8 : 16))); }
THZ_HD constexpr int p2_col_threads(int N) { return N >= 8192 ? 512 : 256; }
THZ_HD constexpr int p2_pitch(int N) { return N + (N >> 4); }
// the software pipelines of the row kernels need extra shared memory; 16384-point lines do without
THZ_HD constexpr bool p2_row_pipelined(int N) { return N <= 8192; }

// =============================================================================== K1<N>
// MODE of the row kernels (separate instantiations: the common case pays nothing for the others):
//   0 plain / DOE transmission evaluated per pixel;  1 pointwise elements in front (aperture mask, lens kernel);
//   2 quantised DOE: the staged "height" row holds int32 level indices and the transmission comes from a [C][levels] table
//   3 (row-iFFT kernels only) plain output -- no DOE adjoint, no elements: the gradient accumulators, the epilogue's prefetch ring
//     and its code are compiled out (forward passes and plain adjoints; the radix-25 lengths spill 300 bytes less)
template <int MODE>
struct K1Loader {
    const cpx* xr;      // row of x (NULL: line beyond the end of the batch)
    const float* hr;    // row of the height map (NULL: no DOE)
    const float* mr;    // row of the aperture mask (NULL: none)
    const cpx* kr;      // row of this wavelength's lens kernel (NULL: none)
    const cpx* lut;     // MODE 2: this wavelength's level transmissions
    float4 cf;
    float base;
    int in_c0, inW, conj_in;
    THZ_HD cpx elem(cpx v, int c) const {
        if constexpr (MODE == 1) {
            if (mr) v = cscale(v, thz_ldg(mr + c));
            if (kr) v = cmul(v, thz_ldg(kr + c));
        }
        return v;
    }
    THZ_HD cpx doe(cpx v, int c) const {
        if (!hr) return v;
        if constexpr (MODE == 2) return cmul(v, thz_ldg(lut + thz_float_bits(hr[c])));
        else return cmul(v, thz_doe_phase(hr[c], cf, base));
    }
    THZ_HD cpx operator()(int pos) const {
        const int c = pos - in_c0;
        if (xr == nullptr || (unsigned)c >= (unsigned)inW) return cmake(0.f, 0.f);   // one compare: c < 0 wraps
        cpx v = xr[c];
        if (conj_in) v.y = -v.y;
        return doe(elem(v, c), c);
    }
    THZ_HD cpx live(int pos) const {             // pos known to be inside [in_c0, in_c0 + inW)
        if (xr == nullptr) return cmake(0.f, 0.f);
        const int c = pos - in_c0;
        cpx v = xr[c];
        if (conj_in) v.y = -v.y;
        return doe(elem(v, c), c);
    }
};
// rows of the pointwise elements for line gl of a row kernel's chunk (field f = gl / inH, row r)
template <int MODE>
THZ_HD void p2k1_elem_rows(const RowFwdArgs& a, K1Loader<MODE>& ld, int f, int r) {
    if constexpr (MODE == 2) ld.lut = a.doe.lphase + (size_t)((a.c0 + f) % a.C) * a.doe.nlev;
    if constexpr (MODE == 1) {
        ld.mr = a.elem.mask ? a.elem.mask + (size_t)r * a.inW : nullptr;
        ld.kr = a.elem.mul ? a.elem.mul + ((size_t)((a.c0 + f) % a.C) * a.inH + r) * a.inW : nullptr;
    }
}

template <int N, int MODE = 0>
THZ_HD void p2k1_first(const RowFwdArgs& a, cpx* s, int bx, int tid, int nt) {
    constexpr int LINES = p2_row_lines(N), NB = P2Stage<N, 0>::NB, PITCH = p2_pitch(N);
    const int total_lines = a.nbc * a.inH;
    for (int w = tid; w < LINES * NB; w += nt) {
        const int line = w / NB, j = w % NB;
        const int gl = bx * LINES + line;
        K1Loader<MODE> ld;
        ld.lut = nullptr;
        ld.xr = nullptr;
        ld.hr = nullptr;
        ld.mr = nullptr;
        ld.kr = nullptr;
        ld.cf = cmake4(0.f);
        ld.base = a.doe.base;
        ld.in_c0 = a.in_c0;
        ld.inW = a.inW;
        ld.conj_in = a.conj_in;
        if (gl < total_lines) {
            const int f = gl / a.inH, r = gl - f * a.inH;
            p2k1_elem_rows(a, ld, f, r);
            ld.xr = a.x + (size_t)gl * a.inW;
            if (a.doe.hmap) {
                ld.cf = thz_ldg(a.doe.coef + (a.c0 + f) % a.C);
                ld.hr = thz_doe_map(a.doe, a.c0, a.C, f) + (size_t)r * a.inW;
            }
        }
        if (a.half_in) p2_first_stage_from<N, 1, true>(s + line * PITCH, j, a.tw, ld);
        else p2_first_stage_from<N, 1>(s + line * PITCH, j, a.tw, ld);
    }
}

// ---- software-pipelined variant: the CTA is persistent over line groups; the raw input rows (and the
// height-map rows in DOE mode) of the NEXT group are staged in shared memory by cp.async while the
// current group is transformed.  Staging layout: xs[LINES][inW] complex, then hs[LINES][inW] float.
template <int N>
THZ_HD void p2k1_prefetch(const RowFwdArgs& a, cpx* xs, float* hs, int grp, int tid, int nt) {
    constexpr int LINES = p2_row_lines(N);
    const int total_lines = a.nbc * a.inH;
    for (int line = 0; line < LINES; ++line) {
        const int gl = grp * LINES + line;
        if (gl >= total_lines) break;
        const cpx* xr = a.x + (size_t)gl * a.inW;
        cpx* xd = xs + (size_t)line * a.inW;
        if ((a.inW & 1) == 0) {
            for (int c = tid * 2; c < a.inW; c += nt * 2) thz_cp_async16(xd + c, xr + c);
        } else {
            for (int c = tid; c < a.inW; c += nt) thz_cp_async8(xd + c, xr + c);
        }
        if (a.doe.hmap) {
            const float* hr = thz_doe_map(a.doe, a.c0, a.C, gl / a.inH) + (size_t)(gl % a.inH) * a.inW;
            float* hd = hs + (size_t)line * a.inW;
            if ((a.inW & 3) == 0) {
                for (int c = tid * 4; c < a.inW; c += nt * 4) thz_cp_async16(hd + c, hr + c);
            } else {
                for (int c = tid; c < a.inW; c += nt) thz_cp_async4(hd + c, hr + c);
            }
        }
    }
}

template <int N, int MODE = 0>
THZ_HD void p2k1_first_staged(const RowFwdArgs& a, cpx* s, const cpx* xs, const float* hs, int grp, int tid, int nt) {
    constexpr int LINES = p2_row_lines(N), NB = P2Stage<N, 0>::NB, PITCH = p2_pitch(N);
    const int total_lines = a.nbc * a.inH;
    for (int w = tid; w < LINES * NB; w += nt) {
        const int line = w / NB, j = w % NB;
        const int gl = grp * LINES + line;
        K1Loader<MODE> ld;
        ld.lut = nullptr;
        ld.xr = nullptr;
        ld.hr = nullptr;
        ld.mr = nullptr;
        ld.kr = nullptr;
        ld.cf = cmake4(0.f);
        ld.base = a.doe.base;
        ld.in_c0 = a.in_c0;
        ld.inW = a.inW;
        ld.conj_in = a.conj_in;
        if (gl < total_lines) {
            p2k1_elem_rows(a, ld, gl / a.inH, gl % a.inH);
            ld.xr = xs + (size_t)line * a.inW;
            if (a.doe.hmap) {
                ld.cf = thz_ldg(a.doe.coef + (a.c0 + gl / a.inH) % a.C);
                ld.hr = hs + (size_t)line * a.inW;
            }
        }
        if (a.half_in) p2_first_stage_from<N, 1, true>(s + line * PITCH, j, a.tw, ld);
        else p2_first_stage_from<N, 1>(s + line * PITCH, j, a.tw, ld);
    }
}

template <int N>
THZ_HD void p2k1_store(const RowFwdArgs& a, const cpx* s, int bx, int tid, int nt) {
    constexpr int LINES = p2_row_lines(N), PITCH = p2_pitch(N);
    const int total_lines = a.nbc * a.inH;
    if constexpr (LINES > 1) {
        // short lines, several per group: ONE flattened loop over the group's elements instead of a loop per line -- the
        // per-line set-up (an integer division, a 64-bit block index) and the half-empty second trip of each line's loop
        // were 36 instructions per stored element on the 400-point lines (profiles/r02_donn_*: 24 % of the kernel)
        if (a.slab.parts <= 1 && a.t_tiled == 2) {
            const int gl0 = bx * LINES;
            const int f0 = gl0 / a.inH, r0 = gl0 - f0 * a.inH;
            const int nvalid = (total_lines - gl0) < LINES ? (total_lines - gl0) : LINES;
            const size_t fstride = (size_t)(N >> 2) * a.rowsT * 4;          // elements of one field in the blocked intermediate
            const unsigned rs4 = (unsigned)a.rowsT * 4;
            if (nvalid == LINES && r0 + LINES <= a.inH) {
                // all lines of the group are consecutive rows of ONE field: in the blocked layout [c/4][r][c%4] the LINES rows
                // of a 4-column block are LINES * 32 contiguous bytes, so walking (block, row, column) makes a warp's stores
                // contiguous runs of 32 * LINES bytes instead of 32-byte pieces LINES rows apart
                cpx* tb = a.T + (size_t)f0 * fstride + (size_t)r0 * 4;
                constexpr int RUN = LINES * 4;
                for (int e = tid; e < (N >> 2) * RUN; e += nt) {
                    const int b = e / RUN, j = e - b * RUN;
                    const int l = j >> 2, p = 4 * b + (j & 3);
                    tb[(size_t)((unsigned)b * rs4 + (unsigned)j)] = s[l * PITCH + p + (p >> 4)];
                }
                return;
            }
            for (int e = tid; e < nvalid * N; e += nt) {
                const int l = e / N, p = e - l * N;
                int r = r0 + l, f = f0;
                while (r >= a.inH) {
                    r -= a.inH;
                    ++f;
                }
                a.T[(size_t)f * fstride + (size_t)((unsigned)(p >> 2) * rs4 + (unsigned)((r << 2) + (p & 3)))] = s[l * PITCH + p + (p >> 4)];
            }
            return;
        }
    }
#pragma unroll
    for (int l = 0; l < LINES; ++l) {
        const int gl = bx * LINES + l;
        if (gl >= total_lines) break;
        const int f = gl / a.inH, r = gl - f * a.inH;
        const cpx* sl = s + l * PITCH;
        if (a.slab.parts > 1) {           // slab FFT: the transpose happens here, segment d of the row goes to GPU d
            for (int p = tid; p < N; p += nt) *thz_slab_addr(a.slab, f, r, p) = sl[p + (p >> 4)];
            continue;
        }
        if (a.t_tiled) {
            // blocked store: consecutive p of a thread are nt apart, i.e. nt >> k blocks further -- one pointer step
            const int k = a.t_tiled;
            cpx* q = a.T + thz_t_tiled_index(f, r, tid, a.rowsT, N, k);
            const size_t step = ((size_t)(nt >> k) * a.rowsT) << k;          // nt is a multiple of the block width
            for (int p = tid; p < N; p += nt, q += step) *q = sl[p + (p >> 4)];
            continue;
        }
        cpx* tr = a.T + ((size_t)f * a.rowsT + r) * N;
        for (int p = tid; p < N; p += nt) tr[p] = sl[p + (p >> 4)];
    }
}

// =============================================================================== K2<N, COLS>
struct K2Loader {
    const cpx* col;     // T + field offset + column (NULL if the column is outside the grid)
    int in_r0, inH, Wp;   // Wp: distance between consecutive rows of this column in T (elements)
    THZ_HD cpx operator()(int pos) const {
        const int r = pos - in_r0;
        if (col == nullptr || (unsigned)r >= (unsigned)inH) return cmake(0.f, 0.f);  // one compare: r < 0 wraps
        return col[(size_t)r * Wp];
    }
    THZ_HD cpx live(int pos) const { return col ? col[(size_t)(pos - in_r0) * Wp] : cmake(0.f, 0.f); }
};
struct K2Storer {
    cpx* col;
    int out_r0, outH, Wp;
    THZ_HD void operator()(int pos, int, cpx v) const {
        const int r = pos - out_r0;
        if (col != nullptr && (unsigned)r < (unsigned)outH) col[(size_t)r * Wp] = v;
    }
    THZ_HD void live(int pos, int, int, cpx v) const {
        if (col != nullptr) col[(size_t)(pos - out_r0) * Wp] = v;
    }
};

template <int N, int COLS>
THZ_HD void p2k2_first(const ColArgs& a, cpx* s, int bx, int by, int tid, int nt) {
    constexpr int NB = P2Stage<N, 0>::NB;
    for (int w = tid; w < COLS * NB; w += nt) {
        const int j = w / COLS, l = w % COLS;
        const int c = bx * COLS + l;
        K2Loader ld;
        ld.col = c >= a.Wp ? nullptr : (a.t_tiled ? a.T + thz_t_tiled_index(by, 0, c, a.rowsT, a.Wp, a.t_tiled) : a.T + (size_t)by * a.rowsT * a.Wp + c);
        ld.in_r0 = a.in_r0;
        ld.inH = a.inH;
        ld.Wp = a.t_tiled ? (1 << a.t_tiled) : a.Wp;
        if (a.half_in) p2_first_stage_from<N, COLS, true>(s + l, j, a.tw, ld);
        else p2_first_stage_from<N, COLS>(s + l, j, a.tw, ld);
    }
}

// last forward stage . H . first inverse stage, in registers (see thz_asm.cuh k2_middle_butterfly)
template <int N, int COLS>
THZ_HD void p2k2_middle(const ColArgs& a, cpx* s, int bx, int by, int tid, int nt) {
    constexpr int S = p2_stages(N) - 1;
    typedef P2Stage<N, S> St;
    constexpr int R = St::R, NB = St::NB;
    static_assert(St::M == 1, "last stage has unit sub-blocks");
    // the column (hence its bin and its column-vector entry) is fixed per thread when nt % COLS == 0
    const int c_chan = (a.c0 + by) % a.C;
    for (int w = tid; w < COLS * NB; w += nt) {
        const int u = w / COLS, l = w % COLS;
        const int col = bx * COLS + l;
        const int p0 = u * R;
        cpx* p = s + (p0 + (p0 >> 4)) * COLS + l;
        cpx v[R];
        // issue the transfer-function loads first: the forward butterfly below hides their latency
        const bool gen = a.tf.mode == 0 && col < a.Wp;
        float4 rv4[R / 2];
        float cv = 0.f;
        float2 sc = cmake(0.f, 0.f);
        if (gen) {
            if (a.tf.row_chunked) {      // [c][q][u] float4: the lanes of a warp (consecutive u) read contiguous memory
                const float4* rvp = reinterpret_cast<const float4*>(a.tf.rowvec + (size_t)c_chan * N) + u;
#pragma unroll
                for (int q = 0; q < R / 2; ++q) rv4[q] = thz_ldg(rvp + (size_t)q * NB);
            } else {
                const float4* rvp = reinterpret_cast<const float4*>(a.tf.rowvec + (size_t)c_chan * N + p0);
#pragma unroll
                for (int q = 0; q < R / 2; ++q) rv4[q] = thz_ldg(rvp + q);
            }
            cv = thz_ldg(a.tf.colvec + (size_t)c_chan * a.Wp + col);
            sc = thz_ldg(a.tf.scal + c_chan);
        }
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = p[t * COLS];
        Dft<R, false>::run(v);
        if (a.tf.mode != 2 && col < a.Wp) {
            if (a.tf.mode == 0) {
#pragma unroll
                for (int q = 0; q < R / 2; ++q) {
                    v[2 * q] = cmul(v[2 * q], thz_tf_value(cmake(rv4[q].x, rv4[q].y), cv, sc, a.tf.conj));
                    v[2 * q + 1] = cmul(v[2 * q + 1], thz_tf_value(cmake(rv4[q].z, rv4[q].w), cv, sc, a.tf.conj));
                }
            } else {
                // table[c][slot_c][slot_r]: this thread's R rows are R * 8 contiguous bytes (R even: 16-byte loads)
                const float4* tp = reinterpret_cast<const float4*>(a.tf.table + ((size_t)c_chan * a.Wp + col) * N + p0);
                static_assert(R % 2 == 0 || R == 1, "static column kernels end in an even radix");
#pragma unroll
                for (int q = 0; q < R / 2; ++q) {
                    const float4 h2 = thz_ldg(tp + q);
                    const cpx h0 = cmake(h2.x, h2.y), h1 = cmake(h2.z, h2.w);
                    v[2 * q] = a.tf.conj ? cmulc(v[2 * q], h0) : cmul(v[2 * q], h0);
                    v[2 * q + 1] = a.tf.conj ? cmulc(v[2 * q + 1], h1) : cmul(v[2 * q + 1], h1);
                }
            }
        }
        Dft<R, true>::run(v);
#pragma unroll
        for (int t = 0; t < R; ++t) p[t * COLS] = v[t];
    }
}

template <int N, int COLS>
THZ_HD void p2k2_last(const ColArgs& a, const cpx* s, const cpx* tw, int bx, int by, int tid, int nt) {
    constexpr int NB = P2Stage<N, 0>::NB;
    for (int w = tid; w < COLS * NB; w += nt) {
        const int j = w / COLS, l = w % COLS;
        const int c = bx * COLS + l;
        K2Storer st;
        st.col = c >= a.Wp ? nullptr : (a.tout_tiled ? a.Tout + thz_t_tiled_index(by, 0, c, a.rowsT, a.Wp, a.tout_tiled)
                                                    : (a.Tout ? a.Tout : a.T) + (size_t)by * a.rowsT * a.Wp + c);   // Tout: not in place
        st.out_r0 = a.out_r0;
        st.outH = a.outH;
        st.Wp = a.tout_tiled ? (1 << a.tout_tiled) : a.Wp;
        if (a.half_out) p2_last_inverse_stage_to<N, COLS, true>(s + l, j, tw, st);
        else p2_last_inverse_stage_to<N, COLS>(s + l, j, tw, st);
    }
}

// =============================================================================== K2 fast path <N, COLS, NT[, TFM]>
// The column kernel of the benchmarked configuration (ColArgs.fast): centred 2x padding on the column axis, input in the
// 4-column blocked intermediate, row-major output, whole tiles.  Same arithmetic, same order, same results as the general
// phases above (the parity tests run both), but every decision the general code takes at run time is a constant here:
// the work loops have compile-time trip counts (NT = block size), the 8 live first-stage inputs of a butterfly sit at
// immediate offsets from one pointer (rows j + t N/16 of a 32-byte-pitch column), the 8 live outputs advance one 64-bit
// pointer by a precomputed step, and no bounds check survives.  profiles/README.md (round 2): the general kernel spends
// 35 % / 27 % of its first / last phase on integer and control instructions; this path removes most of them.
#ifndef THZ_K2F_UNROLL
#define THZ_K2F_UNROLL 1
#endif
constexpr int K2F_UNROLL = THZ_K2F_UNROLL;
THZ_HD constexpr bool p2_k2_fast_ok(int N) {
    // radix-16 first stage (pruned half butterflies) or radix-25 first stage with M % 4 == 0 (static liveness, see
    // p2_first_stage_from); an even last radix (16-byte transfer-function loads)
    return sp_static_ok(N) && N >= 256 && (p2_radix(N, 0) == 16 || (p2_radix(N, 0) == 25 && (N / 25) % 4 == 0)) &&
           p2_radix(N, p2_stages(N) - 1) % 2 == 0;
}

// loads / stores of the fast column path: row 0 of the tile is canvas row N / 4, rows are 4 complex apart on the way in
// (blocked intermediate) and `rs` complex apart on the way out (row-major)
template <int N>
struct K2FastLoader {
    const cpx* p0;
    THZ_HD cpx live(int pos) const { return thz_ld_stream(p0 + (pos - N / 4) * 4); }
    THZ_HD cpx operator()(int pos) const { return (unsigned)(pos - N / 4) < (unsigned)(N / 2) ? live(pos) : cmake(0.f, 0.f); }
};
template <int N>
struct K2FastStorer {
    cpx* q0;
    size_t rs;
    THZ_HD void live(int pos, int, int, cpx v) const { thz_st_stream(q0 + (size_t)(pos - N / 4) * rs, v); }
    THZ_HD void operator()(int pos, int, cpx v) const {
        if ((unsigned)(pos - N / 4) < (unsigned)(N / 2)) live(pos, 0, 0, v);
    }
};

template <int N>
THZ_HD constexpr bool sp_k2_fast_ok_or_false() { return p2_k2_fast_ok(N); }

template <int N, int COLS, int NT>
THZ_HD void p2k2f_first(const ColArgs& a, cpx* s, int bx, int by, int tid) {
    typedef P2Stage<N, 0> St;
    constexpr int NB = St::NB, M = St::M, ITEMS = (COLS * NB + NT - 1) / NT;
    const cpx* tile = a.T + ((((size_t)by * (a.Wp >> 2) + ((bx * COLS) >> 2)) * a.rowsT) << 2) + ((bx * COLS) & 3);
    // one butterfly after the other: with both butterflies' 16 loads in flight per thread the L1 request queue backs up and the
    // kernel is SLOWER (measured twice, profiles/README.md), so the item loop is deliberately not unrolled
#pragma unroll K2F_UNROLL
    for (int k = 0; k < ITEMS; ++k) {
        const int w = tid + k * NT;
        if ((COLS * NB) % NT != 0 && w >= COLS * NB) break;
        const int j = w / COLS, l = w % COLS;
        // element 4 + t of butterfly j is canvas row j + (4 + t) M = live row j + t M; rows are 4 complex (32 bytes) apart
        const cpx* p0 = tile + (COLS > 4 ? (((size_t)(l >> 2) * a.rowsT) << 2) + (l & 3) : (size_t)l);
        if constexpr (St::R == 16) {
            static_assert(4 * M == N / 4, "radix-16 first stage: live elements 4..11");
            p0 += (size_t)j * 4;
            cpx in[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) in[t] = thz_ld_stream(p0 + t * M * 4);
            cpx v[16];
            dft16_half_in<false>(in, v);
            p2_apply_twiddles<16>(v, thz_ldg(a.tw + j));
            cpx* p = s + (j + (j >> 4)) * COLS + l;
#pragma unroll
            for (int t = 0; t < 16; ++t) p[p2_coff(M, t) * COLS] = v[t];
        } else {        // radix-25 first stage: the general stage function with the static-address loader
            K2FastLoader<N> ld;
            ld.p0 = p0;
            p2_first_stage_from<N, COLS, true>(s + l, j, a.tw, ld);
        }
    }
}

template <int N, int COLS, int NT, int TFM>
THZ_HD void p2k2f_middle(const ColArgs& a, cpx* s, int bx, int by, int tid) {
    constexpr int S = p2_stages(N) - 1;
    typedef P2Stage<N, S> St;
    constexpr int R = St::R, NB = St::NB, ITEMS = (COLS * NB + NT - 1) / NT;
    static_assert(St::M == 1 && R % 2 == 0, "fast column path: unit last stage, even radix");
    const int c_chan = (a.c0 + by) % a.C;
#pragma unroll 1
    for (int k = 0; k < ITEMS; ++k) {
        const int w = tid + k * NT;
        if ((COLS * NB) % NT != 0 && w >= COLS * NB) break;
        const int u = w / COLS, l = w % COLS;
        const int col = bx * COLS + l;
        const int p0 = u * R;
        cpx* p = s + (p0 + (p0 >> 4)) * COLS + l;
        cpx v[R];
        float4 h4[R / 2];           // TFM 0: {Kx^2, tau} of two rows per entry; TFM 1: two complex table entries per entry
        float cv = 0.f;
        float2 sc = cmake(0.f, 0.f);
        if constexpr (TFM == 0) {
            const float4* rvp = reinterpret_cast<const float4*>(a.tf.rowvec + (size_t)c_chan * N) + u;      // chunked layout
#pragma unroll
            for (int q = 0; q < R / 2; ++q) h4[q] = thz_ldg_keep(rvp + (size_t)q * NB);
            cv = thz_ldg(a.tf.colvec + (size_t)c_chan * a.Wp + col);
            sc = thz_ldg(a.tf.scal + c_chan);
        } else if constexpr (TFM == 1) {
            const float4* tp = reinterpret_cast<const float4*>(a.tf.table + ((size_t)c_chan * a.Wp + col) * N + p0);
#pragma unroll
            for (int q = 0; q < R / 2; ++q) h4[q] = thz_ldg(tp + q);
        }
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = p[t * COLS];
        Dft<R, false>::run(v);
        if constexpr (TFM == 0) {
#pragma unroll
            for (int q = 0; q < R / 2; ++q) {
                v[2 * q] = cmul(v[2 * q], thz_tf_value(cmake(h4[q].x, h4[q].y), cv, sc, a.tf.conj));
                v[2 * q + 1] = cmul(v[2 * q + 1], thz_tf_value(cmake(h4[q].z, h4[q].w), cv, sc, a.tf.conj));
            }
        } else if constexpr (TFM == 1) {
            if (a.tf.conj) {
#pragma unroll
                for (int q = 0; q < R / 2; ++q) {
                    v[2 * q] = cmulc(v[2 * q], cmake(h4[q].x, h4[q].y));
                    v[2 * q + 1] = cmulc(v[2 * q + 1], cmake(h4[q].z, h4[q].w));
                }
            } else {
#pragma unroll
                for (int q = 0; q < R / 2; ++q) {
                    v[2 * q] = cmul(v[2 * q], cmake(h4[q].x, h4[q].y));
                    v[2 * q + 1] = cmul(v[2 * q + 1], cmake(h4[q].z, h4[q].w));
                }
            }
        }
        Dft<R, true>::run(v);
#pragma unroll
        for (int t = 0; t < R; ++t) p[t * COLS] = v[t];
    }
}

template <int N, int COLS, int NT>
THZ_HD void p2k2f_last(const ColArgs& a, const cpx* s, const cpx* tws, int bx, int by, int tid) {
    typedef P2Stage<N, 0> St;
    constexpr int NB = St::NB, M = St::M, ITEMS = (COLS * NB + NT - 1) / NT;
    const int c_out = a.t2_perm ? thz_t2_perm_col(bx * COLS, a.t2_perm, a.Wp) : bx * COLS;   // (permuted: COLS == 2, an even / odd pair)
    cpx* tile = a.Tout + (size_t)by * a.rowsT * a.Wp + (size_t)c_out;           // row-major, row 0 = canvas row N / 4
    const size_t step = (size_t)M * a.Wp;
#pragma unroll K2F_UNROLL
    for (int k = 0; k < ITEMS; ++k) {
        const int w = tid + k * NT;
        if ((COLS * NB) % NT != 0 && w >= COLS * NB) break;
        const int j = w / COLS, l = w % COLS;
        if constexpr (St::R == 16) {
            const cpx* p = s + (j + (j >> 4)) * COLS + l;
            cpx v[16];
#pragma unroll
            for (int t = 0; t < 16; ++t) v[t] = p[p2_coff(M, t) * COLS];
            p2_apply_twiddles<16>(v, cconj(tws[p2_twi(j)]));
            cpx o[8];
            dft16_half_out<true>(v, o);
            cpx* q = tile + (size_t)j * a.Wp + l;             // output 4 + t is canvas row j + (4 + t) M = live row j + t M
#pragma unroll
            for (int t = 0; t < 8; ++t, q += step) thz_st_stream(q, o[t]);
        } else {        // radix-25: the general stage function with the static-address storer
            K2FastStorer<N> st;
            st.q0 = tile + l;
            st.rs = (size_t)a.Wp;
            p2_last_inverse_stage_to<N, COLS, true>(s + l, j, tws, st);
        }
    }
}

// =============================================================================== K3<N>
template <int N>
THZ_HD void p2k3_load(const RowInvArgs& a, cpx* s, int bx, int f, int tid, int nt) {
    constexpr int LINES = p2_row_lines(N), PITCH = p2_pitch(N);
#pragma unroll
    for (int l = 0; l < LINES; ++l) {
        const int r = bx * LINES + l;
        if (r >= a.outH) break;
        cpx* sl = s + l * PITCH;
        if (a.slab.parts > 1) {           // slab FFT: gather the row from the column slabs of all GPUs
            for (int p = tid; p < N; p += nt) sl[p + (p >> 4)] = *thz_slab_addr(a.slab, f, r, p);
            continue;
        }
        if (a.t_tiled) {
            for (int p = tid; p < N; p += nt) sl[p + (p >> 4)] = a.T[thz_t_tiled_index(f, r, p, a.rowsT, N, a.t_tiled)];
            continue;
        }
        const cpx* tr = a.T + ((size_t)f * a.rowsT + r) * N;
        for (int p = tid; p < N; p += nt) sl[p + (p >> 4)] = tr[p];
    }
}

// asynchronous variant of the load: T rows of field f -> padded line buffer `s` (8-byte cp.async: the padded
// slots of odd 16-groups are only 8-byte aligned)
template <int N>
// (part, nparts): the copies are issued in nparts interleaved portions, between the stages of the current line, so that
// they do not arrive at the LSU queue in one burst
THZ_HD void p2k3_prefetch(const RowInvArgs& a, cpx* s, int bx, int f, int tid, int nt, int part = 0, int nparts = 1) {
    constexpr int LINES = p2_row_lines(N), PITCH = p2_pitch(N);
    const int p_lo = tid + part * nt, p_step = nparts * nt;
    if constexpr (LINES > 1) {
        // short lines: the group's rows are ONE contiguous run of the row-major intermediate -- a single flattened loop
        if (a.slab.parts <= 1 && !a.t_tiled) {
            const int r0 = bx * LINES;
            const int nvalid = (a.outH - r0) < LINES ? (a.outH - r0) : LINES;
            const cpx* tb = a.T + ((size_t)f * a.rowsT + r0) * N;
            for (int e = p_lo; e < nvalid * N; e += p_step) {
                const int l = e / N, p = e - l * N;
                thz_cp_async8(s + l * PITCH + p + (p >> 4), tb + e);
            }
            return;
        }
    }
#pragma unroll
    for (int l = 0; l < LINES; ++l) {
        const int r = bx * LINES + l;
        if (r >= a.outH) break;
        cpx* sl = s + l * PITCH;
        if (a.slab.parts > 1) {
            for (int p = p_lo; p < N; p += p_step) thz_cp_async8(sl + p + (p >> 4), thz_slab_addr(a.slab, f, r, p));
            continue;
        }
        if (a.t_tiled) {
            for (int p = p_lo; p < N; p += p_step) thz_cp_async8(sl + p + (p >> 4), a.T + thz_t_tiled_index(f, r, p, a.rowsT, N, a.t_tiled));
            continue;
        }
        const cpx* tr = a.T + ((size_t)f * a.rowsT + r) * N;
        for (int p = p_lo; p < N; p += p_step) thz_cp_async8(sl + p + (p >> 4), tr + p);
    }
}

// ---- TMA-staged variant (thz_p2_k3t, thz_asm_p2_kernels.inc): which lengths it serves, and its first butterfly, which reads
// the dense staged copy of a column-permuted row (thz_t2_perm_col) and writes the padded line buffer
THZ_HD constexpr bool p2_k3_tma_ok(int N) {
    if (!sp_static_ok(N) || (N & (N - 1)) != 0 || N < 2048 || !p2_row_pipelined(N)) return false;
    const int R = p2_radix(N, p2_stages(N) - 1);
    return R % 2 == 0 && R <= 16 && (p2_row_lines(N) * (N / R)) % p2_row_threads(N) == 0;
}
template <int N>
THZ_HD void p2k3_first_from_dense(cpx* line, const cpx* dense, int tid) {
    constexpr int S = p2_stages(N) - 1, R = p2_radix(N, S), NBU = N / R, LINES = p2_row_lines(N), NT = p2_row_threads(N);
    constexpr int ITEMS = LINES * NBU / NT, PITCH = p2_pitch(N);
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int w = tid + k * NT;
        const int l = w / NBU, u = w % NBU, p0 = u * R;
        const float4* g = reinterpret_cast<const float4*>(dense + l * N) + u;
        cpx v[R];
#pragma unroll
        for (int h = 0; h < R / 2; ++h) {
            const float4 q = g[h * NBU];
            v[2 * h] = cmake(q.x, q.y);
            v[2 * h + 1] = cmake(q.z, q.w);
        }
        Dft<R, true>::run(v);
        cpx* p = line + l * PITCH + p0 + (p0 >> 4);
#pragma unroll
        for (int t = 0; t < R; ++t) p[t] = v[t];
    }
}

// number of register accumulators a thread needs: one per output of each of its stage-0 butterflies
template <int N>
THZ_HD constexpr int p2k3_acc() {
    return ((p2_row_lines(N) * P2Stage<N, 0>::NB + p2_row_threads(N) - 1) / p2_row_threads(N)) * P2Stage<N, 0>::R;
}

template <int PF_, int MODE = 0>
struct K3Storer {
    static constexpr int PF = PF_;   // epilogue loads (saved field, height map) run this many outputs ahead
    cpx* yrow;            // output row (forward output or grad wrt field; may be NULL in DOE mode)
    const cpx* xrow;      // saved input row (DOE mode)
    const float* hrow;    // height-map row (DOE mode; NULL = plain forward)
    const float* mrow;    // aperture-mask row (adjoint of pointwise elements; NULL: none)
    const cpx* krow;      // lens-kernel row of this wavelength (NULL: none)
    const cpx* lut;       // MODE 2: this wavelength's level transmissions (hrow then holds int32 level indices)
    float4 cf;
    cpx gamma;
    float base, scale;
    int out_c0, outW;
    float* acc;           // this butterfly's R accumulators
    cpx xq[PF];           // prefetch ring (registers after unrolling)
    float hq[PF];
    THZ_HD void prefetch(int pos, int t) {
        if constexpr (MODE == 3) return;
        const int c = pos - out_c0;
        if (hrow == nullptr || (unsigned)c >= (unsigned)outW) return;
        hq[t % PF] = thz_ldg(hrow + c);
        xq[t % PF] = thz_ldg(xrow + c);
    }
    THZ_HD void prefetch_live(int pos, int slot) {      // pos known to be inside the crop
        if constexpr (MODE == 3) return;
        if (hrow == nullptr) return;
        const int c = pos - out_c0;
        hq[slot % PF] = thz_ldg(hrow + c);
        xq[slot % PF] = thz_ldg(xrow + c);
    }
    THZ_HD void live(int pos, int slot, int t, cpx v) {
        emit(pos - out_c0, slot, t, v);
    }
    THZ_HD void operator()(int pos, int t, cpx v) {
        const int c = pos - out_c0;
        if ((unsigned)c >= (unsigned)outW) return;
        emit(c, t, t, v);
    }
    THZ_HD cpx elem(cpx v, int c, bool conj) const {     // pointwise elements in front of the DOE (thz_elem_apply per row)
        if constexpr (MODE == 1) {
            if (mrow) v = cscale(v, thz_ldg(mrow + c));
            if (krow) {
                const cpx m = thz_ldg(krow + c);
                v = conj ? cmulc(v, m) : cmul(v, m);
            }
        }
        return v;
    }
    THZ_HD void emit(int c, int slot, int t, cpx v) {
        v = cscale(v, scale);
        if constexpr (MODE == 3) {                 // plain output known at compile time: no epilogue state is kept alive
            yrow[c] = v;
            return;
        }
        if (hrow == nullptr) {
            yrow[c] = elem(v, c, true);            // plain output, or the adjoint of pointwise elements alone
            return;
        }
        cpx p;
        if constexpr (MODE == 2) p = thz_ldg(lut + thz_float_bits(hq[slot % PF]));
        else p = thz_doe_phase(hq[slot % PF], cf, base);
        const cpx q = cmulc(v, p);                 // grad wrt the field that entered the DOE (= x m)
        if (yrow) yrow[c] = elem(q, c, true);      // grad wrt x: conj(m) on top
        // gh += Re(conj(v) (x m) p gamma) = Re(conj(q) (x m) gamma)
        const cpx xg = cmul(elem(xq[slot % PF], c, false), gamma);
        acc[t] += q.x * xg.x + q.y * xg.y;
    }
};

// DOE adjoint only: ask the L2 for the saved-field and height-map sectors this thread's epilogue will read, three FFT
// stages before it reads them (one request per 32-byte sector), so that those loads find L2 instead of HBM latency.
template <int N>
THZ_HD void p2k3_prefetch_epilogue(const RowInvArgs& a, int bx, int f, int tid, int nt) {
    if (!a.doe.hmap) return;
    constexpr int LINES = p2_row_lines(N), NB = P2Stage<N, 0>::NB, R = P2Stage<N, 0>::R, M = P2Stage<N, 0>::M;
    if constexpr (LINES > 1) {
        // short lines: the group's saved-field rows and height-map rows are contiguous runs -- one request per 32-byte sector
        // in a flat loop (any thread may ask for any sector: the L2 is shared) instead of a 25-way per-butterfly index walk
        const int r0 = bx * LINES;
        const int nvalid = (a.outH - r0) < LINES ? (a.outH - r0) : LINES;
        if (nvalid <= 0) return;
        const cpx* xb = a.xsaved + ((size_t)f * a.outH + r0) * a.outW;
        const float* hb = a.doe.hmap + (size_t)r0 * a.outW;
        const int nx = nvalid * a.outW;
        for (int e = tid * 4; e < nx; e += nt * 4) thz_prefetch_l2(xb + e);
        for (int e = tid * 8; e < nx; e += nt * 8) thz_prefetch_l2(hb + e);
        return;
    }
    for (int w = tid; w < LINES * NB; w += nt) {
        const int line = w / NB, j = w % NB;
        const int r = bx * LINES + line;
        if (r >= a.outH) break;
        const cpx* xrow = a.xsaved + ((size_t)f * a.outH + r) * a.outW;
        const float* hrow = a.doe.hmap + (size_t)r * a.outW;
#pragma unroll
        for (int t = 0; t < R; ++t) {
            const int c = j + t * M - a.out_c0;
            if ((unsigned)c >= (unsigned)a.outW) continue;
            if ((c & 3) == 0) thz_prefetch_l2(xrow + c);
            if ((c & 7) == 0) thz_prefetch_l2(hrow + c);
        }
    }
}

// inverse stage 0 + crop + scale + epilogue for field f; acc has p2k3_acc<N>() entries
template <int N, int NACC, int MODE = 0>
THZ_HD void p2k3_last(const RowInvArgs& a, const cpx* s, const cpx* tw, int bx, int f, int tid, int nt, float (&acc)[NACC]) {
    constexpr int LINES = p2_row_lines(N), NB = P2Stage<N, 0>::NB, PITCH = p2_pitch(N), R = P2Stage<N, 0>::R;
    K3Storer<(R <= 16 ? 4 : 1), MODE> st;     // radix-25 butterflies have no registers to spare for a deeper ring
    st.cf = cmake4(0.f);
    st.gamma = cmake(0.f, 0.f);
    st.lut = nullptr;
    if constexpr (MODE == 2) st.lut = a.doe.lphase + (size_t)((a.c0 + f) % a.C) * a.doe.nlev;
    if (a.doe.hmap) {
        st.cf = thz_ldg(a.doe.coef + (a.c0 + f) % a.C);
        st.gamma = cmake(-st.cf.x * (0.5f * st.cf.y * st.cf.z), -st.cf.x * st.cf.w);
    }
    st.base = a.doe.base;
    st.scale = a.scale;
    st.out_c0 = a.out_c0;
    st.outW = a.outW;
    int k = 0;
#pragma unroll
    for (int w0 = 0; w0 < LINES * NB; w0 += p2_row_threads(N), ++k) {
        const int w = w0 + tid;
        if (w >= LINES * NB) break;
        const int line = w / NB, j = w % NB;
        const int r = bx * LINES + line;
        if (r >= a.outH) break;
        const size_t o = ((size_t)f * a.outH + r) * a.outW;
        st.yrow = a.y ? a.y + o : nullptr;
        st.xrow = a.xsaved ? a.xsaved + o : nullptr;
        st.hrow = a.doe.hmap ? a.doe.hmap + (size_t)r * a.outW : nullptr;
        if constexpr (MODE == 1) {
            st.mrow = a.elem.mask ? a.elem.mask + (size_t)r * a.outW : nullptr;
            st.krow = a.elem.mul ? a.elem.mul + ((size_t)((a.c0 + f) % a.C) * a.outH + r) * a.outW : nullptr;
        }
        st.acc = &acc[k * R];
        if (a.half_out) p2_last_inverse_stage_to<N, 1, true>(s + line * PITCH, j, tw, st);
        else p2_last_inverse_stage_to<N, 1>(s + line * PITCH, j, tw, st);
    }
}

template <int N, int NACC>
THZ_HD void p2k3_flush(const RowInvArgs& a, int bx, int tid, int nt, const float (&acc)[NACC]) {
    if (!a.doe.hmap) return;
    constexpr int LINES = p2_row_lines(N), NB = P2Stage<N, 0>::NB, R = P2Stage<N, 0>::R, M = P2Stage<N, 0>::M;
    int k = 0;
#pragma unroll
    for (int w0 = 0; w0 < LINES * NB; w0 += p2_row_threads(N), ++k) {
        const int w = w0 + tid;
        if (w >= LINES * NB) break;
        const int line = w / NB, j = w % NB;
        const int r = bx * LINES + line;
        if (r >= a.outH) break;
#pragma unroll
        for (int t = 0; t < R; ++t) {
            const int c = j + t * M - a.out_c0;
            if ((unsigned)c >= (unsigned)a.outW) continue;
            float* g = a.gh + (size_t)r * a.outW + c;
            thz_gh_commit(g, acc[k * R + t], a.gh_atomic);
        }
    }
}
