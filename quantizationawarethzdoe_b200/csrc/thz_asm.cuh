// Fused band-limited angular-spectrum propagation (+ DOE phase modulation) -- kernel phase bodies.
//
// Replaces the reference op chain  pad -> fftshift -> fft2(ortho) -> fftshift -> create_kernel ->
// multiply -> ifftshift -> ifft2 -> ifftshift -> CenterCrop  (Props/ASM_Prop.py:340-361,
// utils/Helper_Functions.py:150) and, when a height map is given, DOELayer.modulate
// (Components/QuantizedDOE.py:92-126) with three kernels:
//
//   K1 row pass   : x rows (live rows only) --[* exp(-a h) e^{-i phi h}]--> zero-pad in smem -> row FFT -> T
//   K2 column pass: T column tile -> zero-pad in smem -> column FFT -> * H(kx,ky,lambda,z) generated
//                   in registers (or conj(H) for the adjoint) -> column iFFT -> cropped rows -> T (in place)
//   K3 row pass   : T rows -> row iFFT -> crop -> scale -> y          (forward)
//                                                 -> gx = g' conj(p), gh += Re(conj(g') x p gamma)   (adjoint)
//
// T holds spectra in the plans' digit-reversed order along both axes; nothing ever reorders it.
#pragma once
#include "thz_fft.cuh"

struct TfArgs {
    int mode;                 // 0: generate H in registers from separable vectors; 1: cached table; 2: H == 1
    int conj;                 // 1: use conj(H) (adjoint / backward)
    const float2* rowvec;     // [C][Hp] {Kx^2, tau} in SLOT order: bin (r, c) is kept iff Ky^2[c] <= tau[r]
    const float* colvec;      // [C][Wp] Ky^2 in SLOT order
    const float2* scal;       // [C]     {klam^2, z}
    const cpx* table;         // [C][Hp][Wp] in (slot_r, slot_c) scrambled layout (mode 1)
    int row_chunked;          // 1: rowvec is [C][R/2][Hp/R] float4 pairs (thz_asm_desc.tf_row_chunked), static column kernels only
};

struct DoeArgs {
    const float* hmap;        // [inH][inW] height map, NULL = no DOE
    const float4* coef;       // [C] {k_c, tand, sqrt(eps), sqrt(eps)-1}
    float base;               // BASE_PLANE_THICKNESS (Components/QuantizedDOE.py:23)
    int b0;                   // batch entry of the chunk's first field (f0 / C)
    long long hstride;        // 0: one map for every batch entry; else floats between the maps of consecutive batch entries
                              // (forward only: a sweep of DOE candidates over one input field)
    const cpx* lphase;        // quantised DOE, static row kernels: [C][nlev] transmission of every level; hmap then points at the
    int nlev;                 //   int32 LEVEL INDEX map (same 4-byte elements) and the kernels look p up instead of evaluating it
};
// Fixed pointwise optical elements in front of the (DOE +) propagation -- aperture masks and thin-lens kernels
// (Components/Aperture.py:105-135, Components/Thin_Lens.py:31-85) -- folded into the row-FFT prologue (forward:
// x m) and the row-iFFT epilogue of the adjoint (gx conj(m); the DOE's grad_height sees x m), instead of a pass of their own.
struct ElemArgs {
    const float* mask;        // real [H][W] or NULL
    const cpx* mul;           // complex [C][H][W], one kernel per wavelength, or NULL
};
THZ_HD cpx thz_elem_apply(const ElemArgs& e, cpx v, int chan, size_t hw, size_t rc, bool conj) {
    if (e.mask) v = cscale(v, thz_ldg(e.mask + rc));
    if (e.mul) {
        const cpx m = thz_ldg(e.mul + (size_t)chan * hw + rc);
        v = conj ? cmulc(v, m) : cmul(v, m);
    }
    return v;
}

// height map of field i of a chunk whose first field has wavelength index c0 (batch entry b0 + (c0 + i) / C)
THZ_HD const float* thz_doe_map(const DoeArgs& d, int c0, int C, int i) {
    return d.hstride ? d.hmap + (size_t)(d.b0 + (c0 + i) / C) * (size_t)d.hstride : d.hmap;
}

// DOE transmission p = exp(-0.5 k (h+b) tand sqrt(eps)) * exp(-i k (h+b) (sqrt(eps)-1)), evaluated in the
// reference's fp32 rounding order (Components/QuantizedDOE.py:73-77).
THZ_HD cpx thz_doe_phase(float h, float4 cf, float base) {
    const float hb = thz_add_rn(h, base);
    const float la = thz_mul_rn(thz_mul_rn(thz_mul_rn(thz_mul_rn(-0.5f, cf.x), hb), cf.y), cf.z);
    const float ph = thz_mul_rn(thz_mul_rn(-cf.x, hb), cf.w);
    float sn, cs;
    thz_sincos_fast(ph, &sn, &cs);      // |ph| is a few tens of radians: exact 2 pi reduction + SFU, abs error ~4e-7
    const float a = thz_exp_fast(la);
    return cmake(a * cs, a * sn);
}

// Transfer function value for row entry rv = {Kx^2, tau} and column entry ky2 (Props/ASM_Prop.py:249-301).
// The evanescent cut (:262) and both band-limit constraints (:297-306) are monotone in Ky^2 for a fixed row,
// so the host folds them, evaluated with the reference's exact fp32 expressions, into one threshold tau per
// row: the kept set is exactly {Ky^2 <= tau}.  The phase follows the reference's rounding order.
// Branch-free: dropped bins are computed too (their d may be negative -> NaN) and replaced by 0 with a select at the end,
// which costs less than a divergent early return around the sqrt / sincos sequence.
THZ_HD cpx thz_tf_value(float2 rv, float ky2, float2 sc, int conj) {
    const float K2 = thz_add_rn(rv.x, ky2);
    const float d = thz_sub_rn(sc.x, K2);
    const float ang = thz_mul_rn(sc.y, thz_sqrt_pos(d));
    float sn, cs;
    thz_sincos_fast(ang, &sn, &cs);
    const bool keep = ky2 <= rv.y;
    return cmake(keep ? cs : 0.f, keep ? (conj ? -sn : sn) : 0.f);
}

// =============================================================================== K1: row forward
// Column slabs of a multi-GPU slab FFT (see thz_asm_desc.slab_*): element p of global row (row0 + r) of field f lives in
// slab d = p / Wc, column cc = p % Wc, at ptr[d][(f * rows + row0 + r) * Wc + cc] (row-major slabs, what the row-iFFT
// kernel gathers from) or, for the slabs the row-FFT kernel scatters into, in 4-column blocks
// ptr[d][((f * Wc/4 + cc/4) * rows + row0 + r) * 4 + cc % 4]; ptr[d] may be peer memory.
struct SlabArgs {
    int parts, row0, rows, Wc;
    int blocked;
    cpx* ptr[8];
};

// Blocked layout of the intermediate between the row-FFT and the column kernel (static kernels, whole-pipeline runs): the
// columns are grouped in blocks of 4 (32 bytes = one sector) and all rows of a block are contiguous,
//     T[f][c / 4][r][c % 4]   instead of   T[f][r][c].
// The column pass, which walks down 2 (or 1, 4, ...) columns, then touches 4 cache lines per warp-wide load instead of
// 16 (rows 32 bytes apart instead of Wp * 8) -- the L1 data pipe is the busiest unit of that kernel
// (profiles/README.md) -- at the price of 8 instead of 2 lines per warp-wide STORE in the row-FFT kernel, which costs
// little because stores do not stall.  The column kernel's output goes row-major to a second buffer: blocked READS in
// the row-iFFT kernel cost more than the column kernel's stores gain (measured).
THZ_HD size_t thz_t_tiled_index(int f, int r, int c, int rowsT, int Wp, int k = 2) {     // blocks of 2^k columns
    return ((((size_t)f * (Wp >> k) + (c >> k)) * rowsT + r) << k) + (c & ((1 << k) - 1));
}

THZ_HD cpx* thz_slab_addr(const SlabArgs& sl, int f, int r, int p) {
    const int d = p / sl.Wc, cc = p - d * sl.Wc;
    if (sl.blocked) return sl.ptr[d] + thz_t_tiled_index(f, sl.row0 + r, cc, sl.rows, sl.Wc, 2);
    return sl.ptr[d] + ((size_t)f * sl.rows + (sl.row0 + r)) * sl.Wc + cc;
}

struct RowFwdArgs {
    const cpx* x;             // [nbc][inH][inW] (already offset to the chunk)
    cpx* T;                   // [nbc][rowsT][Wp]
    int nbc;                  // fields in this chunk
    int rowsT;                // row pitch of one field in T (max(inH, outH))
    int c0, C;                // wavelength index of field i is (c0 + i) % C
    int inH, inW, Wp, in_c0;
    int lines;                // rows per CTA
    FftPlan plan;             // length Wp
    const cpx* tw;
    DoeArgs doe;
    ElemArgs elem;            // pointwise elements applied on load (forward passes only)
    int conj_in;              // conjugate on load (inverse transforms via conj . FFT . conj)
    SlabArgs slab;            // parts > 1: scatter the output rows into column slabs instead of T
    int half_in;              // 1: the input columns are exactly [Wp/4, 3Wp/4) (centred 2x padding): pruned first stage
    int t_tiled;              // 1: T is stored in 4-column blocks, see thz_t_tiled_index
};

THZ_HD void k1_load(const RowFwdArgs& a, cpx* s, int bx, int tid, int nthreads) {
    const int pitch = thz_padded_len(a.Wp);
    const int total_lines = a.nbc * a.inH;
    for (int l = 0; l < a.lines; ++l) {
        const int gl = bx * a.lines + l;
        cpx* sl = s + l * pitch;
        if (gl >= total_lines) {
            for (int p = tid; p < a.Wp; p += nthreads) sl[thz_pad(p)] = cmake(0.f, 0.f);
            continue;
        }
        const int f = gl / a.inH, r = gl - f * a.inH;
        const cpx* xr = a.x + (size_t)gl * a.inW;
        float4 cf = cmake4(0.f);
        const float* hr = nullptr;
        if (a.doe.hmap) {
            cf = thz_ldg(a.doe.coef + (a.c0 + f) % a.C);
            hr = thz_doe_map(a.doe, a.c0, a.C, f) + (size_t)r * a.inW;
        }
        const bool has_elem = a.elem.mask || a.elem.mul;
        for (int p = tid; p < a.Wp; p += nthreads) {
            const int c = p - a.in_c0;
            cpx v = cmake(0.f, 0.f);
            if (c >= 0 && c < a.inW) {
                v = xr[c];
                if (a.conj_in) v.y = -v.y;
                if (has_elem) v = thz_elem_apply(a.elem, v, (a.c0 + f) % a.C, (size_t)a.inH * a.inW, (size_t)r * a.inW + c, false);
                if (hr) v = cmul(v, thz_doe_phase(thz_ldg(hr + c), cf, a.doe.base));
            }
            sl[thz_pad(p)] = v;
        }
    }
}

THZ_HD void k1_store(const RowFwdArgs& a, const cpx* s, int bx, int tid, int nthreads) {
    const int pitch = thz_padded_len(a.Wp);
    const int total_lines = a.nbc * a.inH;
    for (int l = 0; l < a.lines; ++l) {
        const int gl = bx * a.lines + l;
        if (gl >= total_lines) break;
        const cpx* sl = s + l * pitch;
        const int f = gl / a.inH, r = gl - f * a.inH;
        cpx* tr = a.T + ((size_t)f * a.rowsT + r) * a.Wp;
        for (int p = tid; p < a.Wp; p += nthreads) tr[p] = sl[thz_pad(p)];
    }
}

// =============================================================================== K2: column pass
struct ColArgs {
    cpx* T;                   // [nbc][rowsT][Wp] in place; input rows [0,inH), output rows [0,outH)
    int nbc, c0, C;
    int rowsT;                // max(inH, outH): row pitch of one field in T is rowsT*Wp
    int inH, outH, Hp, Wp;
    int in_r0, out_r0;
    int cols;                 // columns per CTA tile
    FftPlan plan;             // length Hp (column transform)
    const cpx* tw;            // length Hp
    TfArgs tf;
    int t_tiled;              // k > 0: the INPUT rows are read from T in 2^k-column blocks (thz_t_tiled_index) and the output
                              //    rows are written to Tout (a different buffer: not in place)
    cpx* Tout;                // separate output buffer (NULL: in place in T)
    int half_in, half_out;    // 1: input / output rows are exactly [Hp/4, 3Hp/4): pruned first / last stage
    int tout_tiled;           // k > 0: Tout is blocked as well (experiment), 0: row-major
    int pf_blocks;            // > 0 (static kernels, blocked input): every CTA asks the L2 for the 4-column blocks that lie this
                              //    many blocks ahead of its own -- the tile of the CTA that will take its place on the SM
    int t2_perm;              // R > 0 (fast kernel, 2 columns per CTA, set by thz_asm_propagate): the columns of Tout are stored in the
                              //    order the TMA-staged row-iFFT kernel reads them, thz_t2_perm_col below (0: natural order)
    int fast;                 // 1: the configuration the specialised column kernel serves (thz_p2_k2f): centred 2x padding on
                              //    the column axis (half_in && half_out), 4-column blocked input, row-major output in Tout,
                              //    whole column tiles, first radix 16 -- every address offset is then a compile-time constant
};

THZ_HD void k2_load(const ColArgs& a, cpx* s, int bx, int by, int tid, int nthreads) {
    const int col0 = bx * a.cols;
    const cpx* Tf = a.T + (size_t)by * a.rowsT * a.Wp;
    const int total = a.Hp * a.cols;
    for (int w = tid; w < total; w += nthreads) {
        const int p = w / a.cols, l = w - p * a.cols;
        const int r = p - a.in_r0, c = col0 + l;
        cpx v = cmake(0.f, 0.f);
        if (r >= 0 && r < a.inH && c < a.Wp) v = Tf[(size_t)r * a.Wp + c];
        s[thz_pad(p) * a.cols + l] = v;
    }
}

// Last forward stage + transfer-function multiply + first inverse stage, fused in registers: the DIF
// forward's final butterfly and the DIT inverse's first butterfly touch the same R slots.
template <int R>
THZ_HD void k2_middle_butterfly(const ColArgs& a, cpx* s, int l, int u, int col, int f) {
    const int p0 = u * R;   // M == 1 in the last stage
    cpx v[R];
#pragma unroll
    for (int t = 0; t < R; ++t) v[t] = s[thz_pad(p0 + t) * a.cols + l];
    Dft<R, false>::run(v);
    if (a.tf.mode != 2 && col < a.Wp) {
        const int c = (a.c0 + f) % a.C;
        if (a.tf.mode == 0) {
            const float cv = thz_ldg(a.tf.colvec + (size_t)c * a.Wp + col);
            const float2 sc = thz_ldg(a.tf.scal + c);
            const float2* rvp = a.tf.rowvec + (size_t)c * a.Hp + p0;
            float2 rv[R];
#pragma unroll
            for (int q = 0; q < R; ++q) rv[q] = thz_ldg(rvp + q);
#pragma unroll
            for (int q = 0; q < R; ++q) v[q] = cmul(v[q], thz_tf_value(rv[q], cv, sc, a.tf.conj));
        } else {
            const cpx* tp = a.tf.table + ((size_t)c * a.Wp + col) * a.Hp + p0;     // table[c][slot_c][slot_r]: the R rows are contiguous
#pragma unroll
            for (int q = 0; q < R; ++q) {
                cpx h = thz_ldg(tp + q);
                v[q] = a.tf.conj ? cmulc(v[q], h) : cmul(v[q], h);
            }
        }
    }
    Dft<R, true>::run(v);
#pragma unroll
    for (int t = 0; t < R; ++t) s[thz_pad(p0 + t) * a.cols + l] = v[t];
}

template <bool MIXED>
THZ_HD void k2_middle(const ColArgs& a, cpx* s, int bx, int by, int tid, int nthreads) {
    const int R = a.plan.radix[a.plan.ns - 1];
    const int nb = a.Hp / R;
    const int total = nb * a.cols;
    const int col0 = bx * a.cols;
    for (int w = tid; w < total; w += nthreads) {
        const int u = w / a.cols, l = w - u * a.cols;
        const int col = col0 + l;
        switch (R) {
        case 16: k2_middle_butterfly<16>(a, s, l, u, col, by); break;
        case 8: k2_middle_butterfly<8>(a, s, l, u, col, by); break;
        case 4: k2_middle_butterfly<4>(a, s, l, u, col, by); break;
        case 2: k2_middle_butterfly<2>(a, s, l, u, col, by); break;
        default:
            if (MIXED) {
                switch (R) {
                case 25: k2_middle_butterfly<25>(a, s, l, u, col, by); break;
                case 20: k2_middle_butterfly<20>(a, s, l, u, col, by); break;
                case 15: k2_middle_butterfly<15>(a, s, l, u, col, by); break;
                case 14: k2_middle_butterfly<14>(a, s, l, u, col, by); break;
                case 12: k2_middle_butterfly<12>(a, s, l, u, col, by); break;
                case 10: k2_middle_butterfly<10>(a, s, l, u, col, by); break;
                case 9: k2_middle_butterfly<9>(a, s, l, u, col, by); break;
                case 7: k2_middle_butterfly<7>(a, s, l, u, col, by); break;
                case 6: k2_middle_butterfly<6>(a, s, l, u, col, by); break;
                case 5: k2_middle_butterfly<5>(a, s, l, u, col, by); break;
                case 3: k2_middle_butterfly<3>(a, s, l, u, col, by); break;
                default: break;
                }
            }
            break;
        }
    }
}

THZ_HD void k2_store(const ColArgs& a, const cpx* s, int bx, int by, int tid, int nthreads) {
    const int col0 = bx * a.cols;
    cpx* Tf = a.T + (size_t)by * a.rowsT * a.Wp;
    const int total = a.outH * a.cols;
    for (int w = tid; w < total; w += nthreads) {
        const int r = w / a.cols, l = w - r * a.cols;
        const int c = col0 + l;
        if (c < a.Wp) Tf[(size_t)r * a.Wp + c] = s[thz_pad(r + a.out_r0) * a.cols + l];
    }
}

// =============================================================================== K3: row inverse + epilogue
#define THZ_K3_OWN 16  // max output elements a thread owns (register gradient accumulators)

struct RowInvArgs {
    const cpx* T;             // [nbc][rowsT][Wp]
    cpx* y;                   // [nbc][outH][outW]   forward output, or grad wrt field (may be NULL in DOE mode)
    int nbc, c0, C;
    int rowsT, outH, outW, Wp, out_c0;
    int lines;                // rows per CTA
    int bc_per_cta;           // fields a CTA walks through (gridDim.y = ceil(nbc / bc_per_cta))
    float scale;              // 1 / (Hp Wp)
    FftPlan plan;             // length Wp
    const cpx* tw;
    // DOE adjoint epilogue (mode = hmap != NULL)
    DoeArgs doe;
    const cpx* xsaved;        // [nbc][outH][outW] field that entered the DOE
    float* gh;                // [outH][outW]
    int gh_atomic;            // how partial sums reach gh (thz_gh_commit): 0 store, 1 atomicAdd, 2 multimem.red on a multicast address
    ElemArgs elem;            // adjoint passes: gx gets conj(m), the DOE's grad_height sees the saved field times m
    SlabArgs slab;            // parts > 1: gather the input rows from column slabs instead of T
    int half_out;             // 1: the output columns are exactly [Wp/4, 3Wp/4): pruned last stage
    int t_tiled;              // k > 0: T is stored in 2^k-column blocks (experiment)
    int pf_groups;            // > 0 (static kernel, row-major T): every CTA asks the L2 for the first rows of the CTA that will
                              //    take its place -- the row group this many groups further on (same field range)
    int t2_perm;              // R > 0: the columns of T are permuted (thz_t2_perm_col) and the kernel is thz_p2_k3t
};

// Column order of the K2 -> K3 intermediate when the row-iFFT kernel stages its rows with one TMA bulk copy (thz_p2_k3t).
// That kernel's first butterfly (last FFT stage, radix R, unit sub-block length) owns the R consecutive slots R u .. R u + R - 1
// of a line; reading them from a DENSE copy of the row would put all threads of a warp on one bank.  With slot p = R u + t
// stored at column  (t / 2) (2 W / R) + 2 u + (t % 2)  thread u reads 16 bytes at [t / 2][u]: consecutive threads, consecutive
// addresses.  Even / odd slot pairs stay adjacent, which is what the column kernel's 2-column tiles write.
THZ_HD int thz_t2_perm_col(int p, int R, int W) {
    const int t = p % R, u = p / R;
    return (t >> 1) * (2 * (W / R)) + 2 * u + (t & 1);
}

THZ_HD void k3_load(const RowInvArgs& a, cpx* s, int bx, int f, int tid, int nthreads) {
    const int pitch = thz_padded_len(a.Wp);
    for (int l = 0; l < a.lines; ++l) {
        const int r = bx * a.lines + l;
        cpx* sl = s + l * pitch;
        if (r >= a.outH) {
            for (int p = tid; p < a.Wp; p += nthreads) sl[thz_pad(p)] = cmake(0.f, 0.f);
            continue;
        }
        const cpx* tr = a.T + ((size_t)f * a.rowsT + r) * a.Wp;
        for (int p = tid; p < a.Wp; p += nthreads) sl[thz_pad(p)] = tr[p];
    }
}

// Epilogue for field f.  acc[] are the per-thread partial sums of grad_height for the elements the
// thread owns (element e = tid + k*nthreads of the CTA's lines*outW block).
THZ_HD void k3_epilogue(const RowInvArgs& a, const cpx* s, int bx, int f, int tid, int nthreads, float (&acc)[THZ_K3_OWN]) {
    const int pitch = thz_padded_len(a.Wp);
    const int block_elems = a.lines * a.outW;
    float4 cf = cmake4(0.f);
    cpx gamma = cmake(0.f, 0.f);
    if (a.doe.hmap) {
        cf = thz_ldg(a.doe.coef + (a.c0 + f) % a.C);
        // gamma_c = -k (0.5 tand sqrt(eps) + i (sqrt(eps) - 1))
        gamma = cmake(-cf.x * (0.5f * cf.y * cf.z), -cf.x * cf.w);
    }
    const bool has_elem = a.elem.mask || a.elem.mul;
    const int chan = (a.c0 + f) % a.C;
    const size_t hw = (size_t)a.outH * a.outW;
    if (!a.doe.hmap) {   // plain output (or the adjoint of pointwise elements only): no accumulators, any block size
        for (int e = tid; e < block_elems; e += nthreads) {
            const int l = e / a.outW, c = e - l * a.outW;
            const int r = bx * a.lines + l;
            if (r >= a.outH) break;
            cpx v = cscale(s[l * pitch + thz_pad(c + a.out_c0)], a.scale);
            if (has_elem) v = thz_elem_apply(a.elem, v, chan, hw, (size_t)r * a.outW + c, true);
            a.y[((size_t)f * a.outH + r) * a.outW + c] = v;
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < THZ_K3_OWN; ++k) {
        const int e = tid + k * nthreads;
        if (e >= block_elems) break;
        const int l = e / a.outW, c = e - l * a.outW;
        const int r = bx * a.lines + l;
        if (r >= a.outH) break;
        const cpx v = cscale(s[l * pitch + thz_pad(c + a.out_c0)], a.scale);
        const size_t o = ((size_t)f * a.outH + r) * a.outW + c;
        const cpx p = thz_doe_phase(thz_ldg(a.doe.hmap + (size_t)r * a.outW + c), cf, a.doe.base);
        cpx xs = a.xsaved[o];
        if (has_elem) xs = thz_elem_apply(a.elem, xs, chan, hw, (size_t)r * a.outW + c, false);     // the DOE saw x m
        if (a.y) {                                           // gx = g' conj(p) conj(m)
            cpx q = cmulc(v, p);
            if (has_elem) q = thz_elem_apply(a.elem, q, chan, hw, (size_t)r * a.outW + c, true);
            a.y[o] = q;
        }
        const cpx xp = cmul(cmul(xs, p), gamma);             // x m p gamma
        acc[k] += v.x * xp.x + v.y * xp.y;                   // Re(conj(g') x m p gamma)
    }
}

THZ_HD void k3_flush(const RowInvArgs& a, int bx, int tid, int nthreads, const float (&acc)[THZ_K3_OWN]) {
    if (!a.doe.hmap) return;
    const int block_elems = a.lines * a.outW;
#pragma unroll
    for (int k = 0; k < THZ_K3_OWN; ++k) {
        const int e = tid + k * nthreads;
        if (e >= block_elems) break;
        const int l = e / a.outW, c = e - l * a.outW;
        const int r = bx * a.lines + l;
        if (r >= a.outH) break;
        float* g = a.gh + (size_t)r * a.outW + c;
        thz_gh_commit(g, acc[k], a.gh_atomic);
    }
}

// =============================================================================== stand-alone fft2 column pass
// Forward column FFT of a tile, then un-scramble both axes on store:  y[bin_r][bin_c] = scale * s[slot_r][slot_c].
struct ColFftArgs {
    const cpx* T;             // [batch][H][W] row-transformed, slot order along W
    cpx* y;                   // [batch][H][W] natural order
    int H, W, cols;
    float scale;
    int conj_out;
    FftPlan plan;             // length H
    FftPlan planW;            // length W
    const cpx* tw;
};

THZ_HD void k2f_load(const ColFftArgs& a, cpx* s, int bx, int by, int tid, int nthreads) {
    const int col0 = bx * a.cols;
    const cpx* Tf = a.T + (size_t)by * a.H * a.W;
    const int total = a.H * a.cols;
    for (int w = tid; w < total; w += nthreads) {
        const int p = w / a.cols, l = w - p * a.cols;
        const int c = col0 + l;
        s[thz_pad(p) * a.cols + l] = c < a.W ? Tf[(size_t)p * a.W + c] : cmake(0.f, 0.f);
    }
}

THZ_HD void k2f_store(const ColFftArgs& a, const cpx* s, int bx, int by, int tid, int nthreads) {
    const int col0 = bx * a.cols;
    cpx* yf = a.y + (size_t)by * a.H * a.W;
    const int total = a.H * a.cols;
    for (int w = tid; w < total; w += nthreads) {
        const int p = w / a.cols, l = w - p * a.cols;
        const int c = col0 + l;
        if (c >= a.W) continue;
        cpx v = cscale(s[thz_pad(p) * a.cols + l], a.scale);
        if (a.conj_out) v.y = -v.y;
        yf[(size_t)thz_pos_to_bin(a.plan, p) * a.W + thz_pos_to_bin(a.planW, c)] = v;
    }
}
