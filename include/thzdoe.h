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
 *   laid out as table[c][slot_r][slot_c] = H'[c][bin(slot_r)][bin(slot_c)]).
 * thz_fft_twiddles: host float[2n] out; tw[m] = exp(-2 pi i m / n) rounded from float64.
 * ------------------------------------------------------------------------------------------- */
int thz_fft_plan_info(int32_t n, int32_t* radices, int32_t* nstages);
int thz_fft_slot_to_bin(int32_t n, int32_t* slot_to_bin);
int thz_fft_twiddles(int32_t n, float* tw);

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
    const void* tf_rowvec;     /* float32 [C,Hp,4] {Kx^2, Kx^2/(2 pi u_lim)^2, Kx^2/klam^2, 0}, FFT-bin order */
    const void* tf_colvec;     /* float32 [C,Wp,4] {Ky^2, Ky^2/klam^2, Ky^2/(2 pi v_lim)^2, 0}, FFT-bin order */
    const void* tf_scal;       /* float32 [C,2]    {klam^2, z}                                             */
    const void* tf_table;      /* complex64 [C,Hp,Wp] in slot order (see thz_fft_slot_to_bin)              */
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
    int32_t reserved;
} thz_asm_desc;

uint64_t thz_asm_workspace_bytes(const thz_asm_desc* desc);
int thz_asm_propagate(const thz_asm_desc* desc, void* stream);

/* ---------------------------------------------------------------------------------------------
 * thz_fft2_c2c -- stand-alone batched 2-D complex FFT in natural order (tests, ft2/ift2 helpers).
 * Replaces torch.fft.fft2 / ifft2 as used by utils/Helper_Functions.py:141-150 (norm: 0 backward,
 * 1 ortho).  x, y: complex64 [batch,H,W]; ws >= batch*H*W*8 bytes; tw_h/tw_w as above.
 * ------------------------------------------------------------------------------------------- */
int thz_fft2_c2c(const void* x, void* y, int32_t batch, int32_t H, int32_t W, int32_t inverse, int32_t ortho,
                 const void* tw_h, const void* tw_w, void* ws, uint64_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* THZDOE_H */
