/* thzdoe.h -- C ABI of the B200-native (sm_100a) field-propagation library.
 *
 * The reference (sihan-shao/QuantizationAwareTHzDOE) is pure Python and has no FFI layer; its
 * boundary for this path is the torch.nn.Module surface.  Each entry point
 * below names the reference code it replaces.  The Python modules under
 * quantizationawarethzdoe_b200/{Props,Components,DataType,utils}/ mirror the reference classes and
 * call these functions through ctypes with raw device pointers and the caller's CUDA stream.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to contiguous row-major data unless it says "host";
 *    complex64 = interleaved (re, im) float pairs; no torch types cross this boundary;
 *  - every function returns 0 (THZ_OK) or a negative THZ_E_* code; thz_last_error() gives a
 *    thread-local human-readable message; no exceptions cross the ABI;
 *  - functions are asynchronous on `stream` (a cudaStream_t passed as void*), never synchronise
 *    the device, never allocate device memory (workspaces and twiddle tables are caller-owned),
 *    and keep no mutable global state, so they are re-entrant and safe on autograd's backward thread.
 */
#ifndef THZDOE_H
#define THZDOE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define THZ_OK 0
#define THZ_E_NULL -1         /* a required pointer is NULL                                  */
#define THZ_E_SHAPE -2        /* inconsistent sizes / offsets / modes (reference: ValueError) */
#define THZ_E_UNSUPPORTED -3  /* e.g. a transform length with a prime factor > 7              */
#define THZ_E_WORKSPACE -4    /* workspace too small                                          */
#define THZ_E_CUDA -5         /* a CUDA runtime call failed (see thz_last_error)              */
#define THZ_E_SMEM -6         /* a line does not fit in shared memory                         */

/* Library version (major*10000 + minor*100 + patch). */
int thz_version(void);
/* Thread-local message of the last failing call on this thread ("" if none). */
const char* thz_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * FFT planning helpers (host side, no device work).
 * thz_fft_plan_info: radices of the in-shared-memory mixed-radix plan for length n.
 *   radices: host int32[16] out;  *nstages out.  Returns THZ_E_UNSUPPORTED if n has a prime factor > 7.
 * thz_fft_slot_to_bin: host int32[n] out; slot_to_bin[p] = DFT bin stored at slot p of a transformed
 *   line (the plans leave spectra in digit-reversed order; cached transfer-function tables must be
 *   laid out as table[c][slot_c][slot_r] = H'[c][bin(slot_r)][bin(slot_c)]).
 * thz_fft_twiddles: host float[2n] out; tw[m] = exp(-2 pi i m / n) rounded from float64.
 * ------------------------------------------------------------------------------------------- */
int thz_fft_plan_info(int32_t n, int32_t* radices, int32_t* nstages);
int thz_fft_slot_to_bin(int32_t n, int32_t* slot_to_bin);
int thz_fft_twiddles(int32_t n, float* tw);
/* 1 if length n is served by the compile-time specialised ("static") kernels, else 0 (runtime-planned engine). */
int thz_fft_is_static(int32_t n);

/* thz_tf_row_thresholds (host-only helper of tf_mode 0): folds the reference's per-bin keep conditions
 *   (rowvec.y[r] + colvec.y[c] <= 1) && (rowvec.z[r] + colvec.z[c] <= 1) && !(klam2 - (rowvec.x[r] + colvec.x[c]) < 0)
 * (Props/ASM_Prop.py:262, :297-301, fp32, one rounding per op) into one threshold per row, tau[c][r] = the
 * largest Ky^2 = colvec.x that is kept (-1 if none), so that keep(r, c) == (colvec.x[c] <= tau[r]) exactly.
 * rowvec float[C][Hp][4], colvec float[C][Wp][4], scal float[C][2] = {klam2, z}, tau float[C][Hp]; all host
 * pointers, natural FFT-bin order.  O((Hp + Wp) log Wp) per wavelength: a depth sweep calls it for every z.
 * Returns THZ_E_UNSUPPORTED if the band-limit quotients are not monotone in Ky^2 (use tf_mode 1 then). */
int thz_tf_row_thresholds(int32_t C, int32_t Hp, int32_t Wp, const float* rowvec, const float* colvec,
                          const float* scal, float* tau);

/* ---------------------------------------------------------------------------------------------
 * thz_asm_propagate -- fused band-limited angular-spectrum propagation, optionally fused with the
 * DOE phase modulation in front of it (forward) or its adjoint behind it (backward).
 *
 * Replaces  ASM_prop.forward                      (Props/ASM_Prop.py:314-378)
 *           ASM_prop.create_kernel                (Props/ASM_Prop.py:212-311)   [tf_mode 0: in registers]
 *           ft2 / ift2 / perform_ft               (utils/Helper_Functions.py:99-160)
 *           DOELayer.modulate / phase_shift_...   (Components/QuantizedDOE.py:46-79, 92-126) [doe_mode 1]
 *           and the autograd replay of all of the above                          [tf_conj 1, doe_mode 2]
 *
 *   y = crop_out( ifft2( H' . fft2( place_in( x [. p(h)] ) ) ) )            doe_mode 0 / 1
 *   g' = crop_out( ifft2( conj(H') . fft2( place_in( g ) ) ) ),
 *        y = g' . conj(p)  (if y != NULL),   gh[r,c] = sum_{b,ch} Re( conj(g') x p gamma_ch )    doe_mode 2
 *
 * with plain (unshifted, backward-normalised) transforms: the reference's four fftshifts and two
 * 'ortho' factors cancel exactly.  H' = ifftshift of the reference's centred kernel.
 * ------------------------------------------------------------------------------------------- */
typedef struct thz_asm_desc {
    /* geometry */
    int32_t B, C;              /* batch, wavelengths; field f = b*C + c                                   */
    int32_t inH, inW;          /* input region (ASM_Prop.py:332)                                          */
    int32_t Hp, Wp;            /* padded transform size (compute_padding, ASM_Prop.py:119-136)            */
    int32_t in_r0, in_c0;      /* offset of the input inside the canvas (= pad, ASM_Prop.py:342)          */
    int32_t outH, outW;        /* output region (CenterCrop, ASM_Prop.py:359-361; = Hp,Wp if no unpad)    */
    int32_t out_r0, out_c0;
    /* data */
    const void* x;             /* complex64 [B,C,inH,inW]                                                  */
    void* y;                   /* complex64 [B,C,outH,outW]; may be NULL iff doe_mode == 2                 */
    /* transfer function */
    int32_t tf_mode;           /* 0 separable vectors -> H in registers; 1 cached table; 2 identity        */
    int32_t tf_conj;           /* 1: multiply by conj(H) (adjoint)                                         */
    const void* tf_rowvec;     /* float32 [C,Hp,2] {Kx^2, tau} in slot order (thz_fft_slot_to_bin): bin (r,c) is kept
                                  iff Ky^2[c] <= tau[r]; tau folds the evanescent cut and both band-limit
                                  constraints (ASM_Prop.py:262,297-306), evaluated on the host in the reference's
                                  fp32 op order, so the mask is bit-identical to the reference's               */
    const void* tf_colvec;     /* float32 [C,Wp]   Ky^2 in slot order                                           */
    const void* tf_scal;       /* float32 [C,2]    {klam^2, z}                                             */
    const void* tf_table;      /* complex64 [C,Wp,Hp] = table[c][slot_c][slot_r] (slot order, see thz_fft_slot_to_bin;
                                  COLUMN-major: a thread of the column pass multiplies R consecutive rows of one column) */
    /* DOE */
    int32_t doe_mode;          /* 0 none; 1 multiply x by p(h) on load; 2 adjoint epilogue                 */
    float doe_base;            /* BASE_PLANE_THICKNESS (Components/QuantizedDOE.py:23)                     */
    const void* doe_hmap;      /* float32 [H,W] height map (mode 1: inH x inW; mode 2: outH x outW)        */
    const void* doe_coef;      /* float32 [C,4] {k_c = 2 pi / lambda_c, tand, sqrt(eps), sqrt(eps) - 1}     */
    const void* doe_xsaved;    /* complex64 [B,C,outH,outW] field that entered the DOE (mode 2)            */
    void* doe_gh;              /* float32 [outH,outW] gradient wrt the height map (mode 2, overwritten)     */
    /* plumbing */
    const void* tw_h;          /* complex64 [Hp] twiddles (thz_fft_twiddles)                               */
    const void* tw_w;          /* complex64 [Wp]                                                           */
    void* ws;                  /* workspace, >= thz_asm_workspace_bytes()                                  */
    uint64_t ws_bytes;
    int32_t bc_chunk;          /* fields processed per kernel group (0 = all); keeps T L2-resident          */
    int32_t tune_k2_cols;      /* 0 = auto; column-tile width override                                     */
    int32_t tune_lines;        /* 0 = auto; rows per CTA override for the row kernels                      */
    int32_t stages;            /* 0 = whole pipeline; else bit mask 1 = row FFT (x -> ws), 2 = column pass (ws in
                                  place), 4 = row iFFT + epilogue (ws -> y): the slab-decomposed multi-GPU FFT runs
                                  the three stages separately around its all-to-all transposes                 */
    /* Slab-decomposed FFT over peer memory (NVLink P2P), slab_parts > 1, one stage per call (stages == 1, 2 or 4).
       Every GPU d owns TWO column slabs of Wc = Wp / slab_parts columns, complex64 [B*C][slab_rows][Wc]:
         S1[d]  row spectra of ITS columns; row-major, or, if slab_blocked, in 4-column blocks [B*C][Wc/4][slab_rows][4]
                (what the column kernel reads best, but 32-byte remote stores: slower over NVLink, measured);
         S2[d]  the column pass output, row-major.
       stages == 1: slab_ptrs[d] = S1[d] (peer-mapped).  The row-FFT kernel writes segment d of each of its rows (local
                    row r = global row slab_row0 + r) straight into S1[d]: the transpose happens in the kernel's stores.
       stages == 2: (descriptor of the local column pass: Wp = Wc) ws = S1[own], slab_ptrs[0] = S2[own].
       stages == 4: slab_ptrs[d] = S2[d] (peer-mapped).  The row-iFFT kernel gathers segment d of its rows from S2[d].
       The caller orders the stages across GPUs: a barrier after stage 1 and one after stage 2 (none between calls:
       a stage only writes the buffer that no GPU can still be reading).  Static-path lengths only
       (THZ_E_UNSUPPORTED otherwise); Wc must be a multiple of 4.                                                        */
    int32_t slab_parts;        /* 0 or 1: off                                                               */
    int32_t slab_row0;         /* global row index of this GPU's first row                                  */
    int32_t slab_rows;         /* rows per field in every column slab                                       */
    int32_t slab_blocked;      /* layout of S1 (stages 1 and 2 must agree)                                    */
    void* slab_ptrs[8];
    /* tf_mode 0 only.  0: tf_rowvec is [C][Hp][2] in slot order.  1: "chunked" for the static column kernels, whose
       threads each own the R consecutive slots 16u .. of one last-stage butterfly: [C][R/2][Hp/R] float4 entries, entry
       (q, u) = {Kx^2, tau}[R u + 2 q], {Kx^2, tau}[R u + 2 q + 1] with R = the last radix of thz_fft_plan_info(Hp), so that
       the 32 lanes of a warp read 512 contiguous bytes instead of 32 different cache lines.  Needs thz_fft_is_static(Hp). */
    int32_t tf_row_chunked;
    int32_t reserved2;
    /* doe_mode 1 only.  0: one height map for every batch entry.  Else: floats between the maps of consecutive batch
       entries, doe_hmap = float32 [B][inH][inW] -- B candidate DOEs evaluated in one pass (loss-landscape sweeps,
       VisTools/calc_loss.py:8-55).  The adjoint (doe_mode 2) takes a shared map only. */
    int64_t doe_hmap_bstride;
    /* Fixed pointwise optical elements in FRONT of the (DOE +) propagation, fused like the DOE instead of a pass of their
       own: an aperture mask (Components/Aperture.py:105-135) and / or a thin-lens kernel per wavelength
       (Components/Thin_Lens.py:31-85).  elem_mode 0: none; 1 (forward calls): x is multiplied by mask * mul on load, before
       the DOE phase; 2 (adjoint calls): the output is multiplied by mask * conj(mul) and, with doe_mode 2, grad_height is
       formed against xsaved * mask * mul.  Sizes follow the region they apply to (forward: inH x inW, adjoint: outH x outW). */
    int32_t elem_mode;
    /* doe_mode 2 only.  0: doe_gh is ordinary device memory, overwritten with this call's grad_height.  1: doe_gh is an NVLS
       MULTICAST address (cuMulticast* / torch symmetric memory `multicast_ptr`) of a float32 [outH,outW] buffer replicated on
       every GPU of a data-parallel group: the epilogue ADDS its partial sums with multimem.red, i.e. the sum over the ranks
       forms inside the NVSwitch while the kernel runs and replaces the all-reduce of the weight gradient.  The caller zeroes
       the replicas beforehand and synchronises the ranks before reading them (parallel.FusedGradReduce does both). */
    int32_t doe_gh_mode;
    const void* elem_mask;     /* float32 [H,W] or NULL      */
    const void* elem_mul;      /* complex64 [C,H,W] or NULL  */
    /* Quantised DOEs (level selection fused into the propagation): when the height map takes only `doe_levels` distinct values,
       h[i,j] = lut[idx[i,j]] (STE / nearest-level / hard Gumbel layers without tolerance noise, Components/QuantizedDOE.py:
       1239-1253), pass the int32 level map the quantiser kernels emit and the transmission of every level for every wavelength,
       doe_level_phase[c][l] = p_c(lut[l]) (thz_doe_modulate_fwd on a unit field over the LUT evaluates exactly that).  The
       static row kernels then LOOK the transmission UP instead of evaluating exp / sincos per pixel and wavelength, forward
       (doe_mode 1) and adjoint (doe_mode 2); results are bit-identical to the height-map path.  doe_hmap must still be
       given: every other code path (run-time planned lengths, pointwise elements, per-entry maps, slabs) uses it. */
    const void* doe_level_idx;     /* int32 [H,W] or NULL */
    const void* doe_level_phase;   /* complex64 [C][doe_levels] or NULL */
    int32_t doe_levels;
    int32_t reserved4;
} thz_asm_desc;

/* Bytes of `ws` a call with this descriptor needs (uses B, C, inH, outH, Hp, Wp, bc_chunk, stages, slab_parts): one
 * intermediate of the live rows, [fields][max(inH, outH)][Wp] complex64 -- or two of them when both transform lengths are
 * served by the static kernels and the whole pipeline runs (the row spectra are then kept in 4-column blocks for the
 * column kernel, whose row-major output goes to the second half). */
/* thz_tf_table_from_angles: builds the tf_mode 1 table on the device from the reference's phase angles evaluated on the HOST
 * for the unique quarter of the (even) frequency grid: angq float32 [C][Hp/2+1][Wp/2+1] = z * sqrt(klam^2 - Kx^2[|i|] - Ky^2[|j|])
 * computed with the reference's own library (Props/ASM_Prop.py:257).  rowtau float32 [C][Hp][2] = {Kx^2, tau} and colk2
 * float32 [C][Wp] in slot order as for tf_mode 0 (the keep mask Ky^2 <= tau is bit-identical to the reference's); rabs int32
 * [Hp] / cabs int32 [Wp] = |centred frequency index| of every slot.  table complex64 [C][Wp][Hp] (out).  Hu x Wu is the extent of
 * angq and must cover every rabs / cabs entry; it equals (Hp/2+1) x (Wp/2+1) for a whole grid and is larger when the table is a
 * decimated sub-grid of a longer canvas (thz_split_pre below). */
int thz_tf_table_from_angles(const void* angq, int32_t C, int32_t Hu, int32_t Wu, const void* rowtau, const void* colk2,
                             const void* rabs, const void* cabs, int32_t Hp, int32_t Wp, void* table, void* stream);
uint64_t thz_asm_workspace_bytes(const thz_asm_desc* desc);
int thz_asm_propagate(const thz_asm_desc* desc, void* stream);

/* ---------------------------------------------------------------------------------------------
 * thz_split_pre / thz_split_post -- the outer decimation step for canvases with an edge above 16384 points (the longest line
 * the in-shared-memory plans hold).  The reference hands any size to torch.fft.fft2 / ifft2 (utils/Helper_Functions.py:141-150,
 * Props/ASM_Prop.py:329-341); here an edge N = P M (P = 1, 2 or 4 per axis) is split into P interleaved length-M problems
 *     X[P k + a] = FFT_M(u_a)[k],    u_a[n] = sum_s x[n + s M] w_N^{a (n + s M)},    w_N = exp(-2 pi i / N)
 * so that  ifft2(H . fft2(pad(x)))  becomes Pr Pc independent un-padded thz_asm_propagate problems of size Mr x Mc (one extra
 * "channel" per (a, b), transfer function = the decimated one), between
 *   thz_split_pre : x complex64 [F][H][W] = the live region at (r0, c0) of the Hp x Wp canvas (zero elsewhere, never stored)
 *                   -> u complex64 [F][Pr Pc][Hp/Pr][Wp/Pc],  times `scale`;
 *   thz_split_post: v complex64 [F][Pr Pc][Hp/Pr][Wp/Pc] -> y complex64 [F][H][W] = the region at (r0, c0) of
 *                   sum_ab conj(w^{a i}) conj(w^{b j}) v_ab[i mod Mr][j mod Mc],  times `scale` (pass 1 / (Pr Pc));
 * post is the adjoint of pre, so the adjoint propagation is the same chain with the two regions swapped.  tw_r / tw_c:
 * complex64 [Hp] / [Wp] = w^j; conj_tw = 1 uses the conjugate twiddles in both (inverse transforms).
 * ------------------------------------------------------------------------------------------- */
int thz_split_pre(const void* x, void* u, int32_t F, int32_t H, int32_t W, int32_t r0, int32_t c0, int32_t Hp, int32_t Wp,
                  int32_t Pr, int32_t Pc, const void* tw_r, const void* tw_c, int32_t conj_tw, float scale, void* stream);
int thz_split_post(const void* v, void* y, int32_t F, int32_t H, int32_t W, int32_t r0, int32_t c0, int32_t Hp, int32_t Wp,
                   int32_t Pr, int32_t Pc, const void* tw_r, const void* tw_c, int32_t conj_tw, float scale, void* stream);

/* ---------------------------------------------------------------------------------------------
 * thz_fft2_c2c -- stand-alone batched 2-D complex FFT in natural order (tests, ft2/ift2 helpers).
 * Replaces torch.fft.fft2 / ifft2 as used by utils/Helper_Functions.py:141-150 (norm: 0 backward,
 * 1 ortho).  x, y: complex64 [batch,H,W]; ws >= batch*H*W*8 bytes; tw_h/tw_w as above.
 * ------------------------------------------------------------------------------------------- */
int thz_fft2_c2c(const void* x, void* y, int32_t batch, int32_t H, int32_t W, int32_t inverse, int32_t ortho,
                 const void* tw_h, const void* tw_w, void* ws, uint64_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Stand-alone DOE phase modulation (a DOE that is not directly followed by ASM_prop).
 * Replaces DOELayer.phase_shift_according_to_height + .modulate (Components/QuantizedDOE.py:46-79,
 * 92-126) and their autograd:   y = x * p_c(h),   gx = g conj(p),   gh = sum_{b,c} Re(conj(g) x p gamma_c).
 *   x, y, g, gx: complex64 [B,C,H,W];  hmap, gh: float32 [H,W];  coef: float32 [C,4] as in thz_asm_desc.
 *   gx or gh may be NULL (not needed); x may be NULL iff gh is NULL.
 * ------------------------------------------------------------------------------------------- */
int thz_doe_modulate_fwd(const void* x, void* y, const void* hmap, const void* coef, float base, int32_t B, int32_t C,
                         int32_t H, int32_t W, void* stream);
int thz_doe_modulate_bwd(const void* g, const void* x, const void* hmap, const void* coef, float base, void* gx, void* gh,
                         int32_t B, int32_t C, int32_t H, int32_t W, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Height-map construction and quantized level selection (all float32 [n] unless noted; idx int32 [n]).
 *
 * thz_height_fwd / _bwd: h = hmax * sigmoid(clamp(w, -c, c)) and gw = g * dh/dw
 *     (FullPrecisionDOELayer.preprocessed_height_map, Components/QuantizedDOE.py:276-283; c = 8).
 * thz_quant_ste_fwd: STEQuantizationFunction.forward (Components/QuantizedDOE.py:1239-1246):
 *     idx = argmin_j |h - lut_j| (first minimum), q = lut[idx].  from_weights=1 fuses the sigmoid
 *     height construction of STEQuantizedDOELayer (:1379-1388) in front; h_pre (optional) gets h.
 *     Backward is the identity (:1248-1253) followed by thz_height_bwd.
 * thz_quant_nn_fwd / _bwd: NearestNeighborSearch / PolyGrad / SigmoidGrad
 *     (Components/quantization.py:59-122, utils/Helper_Functions.py:390-398):
 *     idx = bucketize(x, mid, right=True) % nmid, q = lut[idx]; kind 0 identity, 1 poly, 2 sigmoid.
 * thz_quant_psq_fwd: PSQuantizedDOELayer.preprocessed_height_map (Components/QuantizedDOE.py:1193-1207);
 *     dout_dw (optional) receives d out / d w so that backward is one multiply.
 * thz_quant_gumbel_v3_fwd: SoftGumbelQuantizedDOELayerv3.preprocessed_height_map
 *     (Components/QuantizedDOE.py:794-860) for iter_frac > 0.3, Gumbel noise supplied:
 *     noise float32 [L,n]; beta >= 1 means pure quantized output; dh_dw (optional) = d out / d w;
 *     phase_input=1: `w` is the phase parameter itself (SoftGumbelQuantizedDOELayer v1, :436-446).
 * thz_quant_gumbel_naive_fwd: NaiveGumbelQuantizedDOELayer.preprocessed_height_map (:1022-1031):
 *     logits, noise float32 [n,L]; dq (optional) float32 [n,L] = d q / d logit.
 * ------------------------------------------------------------------------------------------- */
int thz_height_fwd(const void* w, float hmax, float clampv, void* h, uint64_t n, void* stream);
int thz_height_bwd(const void* g, const void* w, float hmax, float clampv, void* gw, uint64_t n, void* stream);
int thz_quant_ste_fwd(const void* in, int32_t from_weights, float hmax, float clampv, const void* lut, int32_t L, void* q,
                      void* idx, void* h_pre, uint64_t n, void* stream);
int thz_quant_nn_fwd(const void* x, const void* lut, int32_t nlut, const void* mid, int32_t nmid, void* q, void* idx,
                     uint64_t n, void* stream);
int thz_quant_nn_bwd(const void* g, const void* x, const void* idx, const void* lut, int32_t nlut, float s, int32_t kind,
                     void* gx, uint64_t n, void* stream);
int thz_quant_psq_fwd(const void* w, float hmax, int32_t L, float tau, void* out, void* dout_dw, uint64_t n, void* stream);
int thz_quant_gumbel_v3_fwd(const void* w, const void* lut, int32_t L, const void* noise, float hmax, float kfac, float c_s,
                            float tau, float tau_max, float s, float beta, float one_minus_beta, int32_t phase_input,
                            void* h_out, void* idx, void* dh_dw, uint64_t n, void* stream);
int thz_quant_gumbel_naive_fwd(const void* logits, const void* noise, const void* lut, int32_t L, float tau, void* q,
                               void* idx, void* dq, uint64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Thickness-space softmax quantization: SoftmaxBasedQuantization.forward + score_thickness
 * (Components/quantization.py:128-161 and :36-55), reached through Quantization(method='*softmax*' | '*gumbel*') (:164-207).
 *
 * thz_quant_softmax_fwd: thickness float32 [n] (one map: the reference normalises by the maximum of |thickness - lut_j| over
 *   the WHOLE tensor, :41), lut float32 [L] (the caller passes lut[:-1] like :182), noise float32 [L, n] = the Gumbel noise
 *   F.gumbel_softmax would draw, or NULL for the plain-softmax branch (:146-152); c, tau, s = tau_max / tau; hard as :134.
 *   Writes q float32 [n], idx int32 [n] (argmax level; optional) and, when dq_dt != NULL, the three per-pixel factors of the
 *   backward (dq_dt, dq_dm, tie_sign: float32 [n] each).  stats: device float32 [4] scratch that links forward and backward
 *   ({max |diff|, number of ties attaining it, sum g dq/dm, unused}).
 * thz_quant_softmax_bwd: gt = g dq_dt + (sum g dq_dm) / ties * tie_sign -- autograd through the out-of-place form
 *   diff / max|diff| (torch.max splits its gradient evenly over ties).  The reference's own backward raises (its in-place
 *   `diff /= max`, :41, invalidates what abs() saved), so this gradient is pinned by the oracle restatement, the forward by
 *   the reference.
 * thz_score_thickness: the scoring function alone, scores float32 [batch, L, n_per_b] from thickness [batch, 1, n_per_b];
 *   func 0 sigmoid, 1 log, 2 poly, 3 sine, 4 chirp (:43-53).
 * ------------------------------------------------------------------------------------------- */
int thz_quant_softmax_fwd(const void* thickness, const void* lut, int32_t L, const void* noise, float c, float tau, float s,
                          int32_t hard, void* q, void* idx, void* dq_dt, void* dq_dm, void* tie_sign, void* stats, uint64_t n,
                          void* stream);
int thz_quant_softmax_bwd(const void* g, const void* dq_dt, const void* dq_dm, const void* tie_sign, void* stats, void* gt,
                          uint64_t n, void* stream);
int thz_score_thickness(const void* thickness, const void* lut, int32_t L, float s, int32_t func, void* scores, void* stats,
                        int32_t batch, uint64_t n_per_b, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Launch accounting (used by bench.py).  thz_launch_count: kernels launched by this library in this
 * process so far.  thz_profile_enable(1) makes every launch record a CUDA event pair on its stream
 * (small overhead, off by default); thz_profile_read sums elapsed milliseconds and launches per
 * kernel class (0 row-FFT forward, 1 column pass, 2 row-iFFT + epilogue, 3 fft2 column pass, 4 DOE
 * modulation, 5 quantizers, 6 CZT) after synchronising the recorded events; enable(0) clears them.
 * ------------------------------------------------------------------------------------------- */
uint64_t thz_launch_count(void);
/* launches of one kernel class so far: 0 row FFT, 1 column pass, 2 row iFFT, 3 fft2 column pass, 4 DOE modulation / field
 * multiply, 5 quantizers, 6 CUDA-core Toeplitz GEMM, 7 loss / optimizer, 8 tcgen05 Toeplitz GEMM (+ its prologue
 * multiply), 9 those launches of classes 0 / 1 / 2 (counted there as well) that ran the TMA variants of the kernels:
 * tensor stores of the row-FFT and column kernels, bulk-copy staging of the row-iFFT kernel. */
uint64_t thz_launch_count_class(int32_t kernel_class);
int thz_profile_enable(int32_t on);
int thz_profile_read(int32_t nclasses, float* ms_sum, int32_t* count);

/* ---------------------------------------------------------------------------------------------
 * thz_toeplitz_gemm -- one separable pass of the chirp-z (Bluestein) propagator as a batched complex
 * GEMM whose left operand is a Toeplitz matrix expanded on the fly from the chirp filter.
 *
 * Replaces CZT_prop.Bluestein_method + compute_fft (Props/CZT_Prop.py:132-225) for one axis, the
 * field * F product (:238) as `pro`, and the F0 * U * z dx dy lambda product (:248) as `epi`;
 * the adjoint passes use the same entry point with conj_* = 1 and the Toeplitz steps swapped.
 *
 *   C[b](m,n) = epi[b](m,n) * sum_k T[b](m,k) * ( pro[b](k,n) * B[b](k,n) ),
 *   T[b](m,k) = g[b][ (off + sm*m + sk*k) mod L ]   (conjugated if conj_g)
 *
 * All operands complex64.  B / pro are addressed as base + b*sb_b + k*sb_k + n*sb_n (elements),
 * C / epi as base + b*sc_b + m*sc_m + n*sc_n, so either orientation of either operand works without
 * a transposition pass.  pro and epi may be NULL.
 * ------------------------------------------------------------------------------------------- */
typedef struct thz_toeplitz_gemm_desc {
    int32_t batch, M, N, K;
    const void* g;             /* complex64 [batch, L] chirp filter (1/h zero-extended to np2, CZT_Prop.py:161) */
    int32_t L, off, sm, sk;    /* Toeplitz index = (off + sm*m + sk*k) mod L, sm, sk in {+1, -1}              */
    int32_t conj_g, conj_pro, conj_epi;
    int32_t impl;              /* 0 auto (tcgen05 kernel if eligible, else CUDA cores + one stderr warning); 1 tcgen05 or
                                  THZ_E_UNSUPPORTED; 2 CUDA-core kernel.  Eligible: sm == -sk, L >= 64, and, with `pro`, a dense
                                  B and a `scratch` buffer.  thz_launch_count_class(8 / 6) counts the two kinds of launch */
    const void* B;
    int64_t sb_b, sb_k, sb_n;
    const void* pro;
    void* C;
    int64_t sc_b, sc_m, sc_n;
    const void* epi;
    void* scratch;             /* optional complex64 [batch*K*N]: lets the tensor-core kernel fold `pro` into B first */
} thz_toeplitz_gemm_desc;

int thz_toeplitz_gemm(const thz_toeplitz_gemm_desc* desc, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The optimisation loop around the propagation (experiment_four_focal_spots.ipynb cell 8):
 *     out_amp = normalize(torch.abs(out_field.data) ** 2)     utils/Helper_Functions.py:185-193
 *     loss = nn.MSELoss()(out_amp, target);  loss.backward();  optimizer.step()   (torch.optim.Adam / AdamW)
 *
 * thz_normmse_loss: y complex64 [B, n_per_b], target float32 [B, n_per_b]; normalize divides each batch entry
 *   by its maximum intensity.  Writes loss float32[1] and, if gy != NULL, gy complex64 [B, n_per_b] = d loss / d y
 *   exactly as autograd forms it: through the division, through max() into the FIRST maximal element, and
 *   through abs()**2 (gy = 2 dL/dI y).  scratch: >= 12*B bytes of device memory.  n_per_b <= 2^32.
 * thz_adam_step: one Adam (decoupled = 0, weight decay added to the gradient) or AdamW (decoupled = 1) update of
 *   n float32 parameters; m, v are the running moments; step is a device int32 holding the number of updates
 *   done so far (bias correction uses *step + 1) and is incremented afterwards if `advance` != 0 -- pass
 *   advance = 1 for the last tensor of a parameter group.  No host-side state, so the launch sequence can be
 *   captured in a CUDA graph.
 * ------------------------------------------------------------------------------------------- */
/* thz_field_mul: y[f][p] = x[f][p] * m, the pointwise optical elements between propagations -- the thin lens phase
 * (Components/Thin_Lens.py:31-85: m complex64 [C][HW], per_channel = 1) and the aperture mask (Components/Aperture.py:
 * 105-135: m float32 [HW], m_real = 1).  conj_m = 1 gives the adjoint (gradient wrt x).  x, y complex64 [BC][HW]. */
int thz_field_mul(const void* x, const void* m, void* y, int32_t BC, int32_t C, uint64_t HW, int32_t per_channel,
                  int32_t m_real, int32_t conj_m, void* stream);
int thz_normmse_loss(const void* y, const void* target, int32_t B, uint64_t n_per_b, void* scratch, void* loss,
                     void* gy, void* stream);
/* thz_normmse_loss_each: losses[b] = mean((|y_b|^2 / max |y_b|^2 - target)^2) for B candidate outputs at once -- the loss
 * of VisTools/calc_loss.py:35-39 per grid point of a loss-landscape sweep.  y complex64 [B, n_per_b]; target float32
 * [n_per_b] (target_shared = 1) or [B, n_per_b]; scratch >= 8*B bytes; losses float32 [B] (overwritten). */
int thz_normmse_loss_each(const void* y, const void* target, int32_t target_shared, int32_t B, uint64_t n_per_b, void* scratch,
                          void* losses, void* stream);
int thz_adam_step(void* p, const void* g, void* m, void* v, void* step, uint64_t n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int32_t decoupled, int32_t advance, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* THZDOE_H */
