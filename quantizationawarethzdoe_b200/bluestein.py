"""Transforms of ANY length for the propagators: Bluestein's chirp-z algorithm on top of the fused pipeline.

The reference runs torch.fft on whatever grid it is given (utils/Helper_Functions.py:141-150); the in-shared-memory plans
of this package cover lengths whose prime factors are <= 7.  For every other length n (101 -> 202 = 2 x 101, 134 -> 268 =
4 x 67, ...) the DFT is rewritten as a circular convolution with a chirp,

    X_k = w_k  sum_j (x_j w_j) conj(w_{k-j}),      w_j = exp(-i pi j^2 / n),

and the convolution -- zero-pad to L >= 2n - 1 (a power of two), FFT, multiply by the chirp's spectrum, inverse FFT, crop --
is exactly what `thz_asm_propagate` does with a cached transfer-function table (tf_mode 1).  So one 2-D DFT of an
unsupported size is   pointwise chirp (thz_field_mul)  ->  thz_asm_propagate on an L1 x L2 canvas  ->  pointwise chirp,
all on the sm_100a kernels, no torch.fft, no CPU.  For the angular-spectrum pipeline the forward transform's trailing chirp
and the inverse transform's leading chirp cancel (|w_k| = 1), which leaves

    y = wbar_out/(Hp Wp) . conv_hbar( H' . conv_h( w_in . x ) )

-- two fused convolutions around one per-wavelength multiply by the reference's own transfer function H' (built on the
host with the reference's torch ops, as in kernel_mode='cached': bit-identical H).  Chirp phases are reduced exactly
(j^2 mod 2n in integers) and evaluated in float64 on the host, once per geometry.
"""
import numpy as np
import torch

from . import _native as N
from . import functional as Fn
from . import longline as LL

MAX_CONV = 16384          # longest line the in-shared-memory plans hold (longline.max_line(); longer canvases are split once)


def length_supported(n):
    """True if the in-shared-memory plans transform length n directly (prime factors <= 7, line fits shared memory)."""
    import ctypes
    rad, ns = (ctypes.c_int32 * 16)(), ctypes.c_int32(0)
    return int(n) <= LL.max_line() and N.lib().thz_fft_plan_info(int(n), rad, ctypes.byref(ns)) == 0


def conv_length(n):
    """Power-of-two canvas length of the chirp convolution for a length-n transform (>= 2n - 1, and >= 16)."""
    L = 16
    while L < 2 * n - 1:
        L *= 2
    top = 4 * LL.max_line()                  # one outer split (longline.py) on top of the in-shared-memory plans
    if L > top:
        raise NotImplementedError("transform length %d needs a %d-point chirp convolution; convolutions are served up to %d points "
                                  "(lengths up to %d)" % (n, L, top, (top + 1) // 2))
    return L


def chirp(n, inverse=False):
    """w_j = exp(-i pi j^2 / n), j = 0..n-1 (conjugated for the inverse transform), complex128.  j^2 is reduced mod 2n in
    integer arithmetic first, so the phase is exact to float64 rounding for any n."""
    j = np.arange(n, dtype=np.int64)
    ph = (j * j) % (2 * n)
    w = np.exp(-1j * np.pi * ph.astype(np.float64) / n)
    return np.conj(w) if inverse else w


def kernel_spectrum(H, W, L1, L2, inverse=False):
    """fft2 of the 2-D chirp kernel h[m1, m2] = conj(w1_m1) conj(w2_m2), |m| < n, embedded circularly in an L1 x L2 canvas
    (complex128 [L1, L2], natural bin order)."""
    def line(n, L):
        h = np.zeros(L, dtype=np.complex128)
        c = np.conj(chirp(n, inverse))
        h[:n] = c
        h[L - n + 1:] = c[1:][::-1]            # h[-m] = h[m]
        return h
    return np.fft.fft(line(H, L1))[:, None] * np.fft.fft(line(W, L2))[None, :]


def table_slot_order(Hn):
    """Natural-order spectrum [C, L1, L2] -> thz_asm_desc.tf_table layout table[c][slot_c][slot_r]."""
    pr = N.slot_to_bin(Hn.shape[-2])
    pc = N.slot_to_bin(Hn.shape[-1])
    return Hn[:, pr][:, :, pc].transpose(1, 2).contiguous()


class _Conv:
    """One chirp convolution on the fused pipeline: input region (inH x inW at r0, c0 of the canvas), output region
    (outH x outW at or0, oc0), kernel spectrum cached as a transfer-function table shared by all fields."""

    def __init__(self, H, W, region_in, region_out, inverse, device):
        L1, L2 = conv_length(H), conv_length(W)
        spec = torch.from_numpy(kernel_spectrum(H, W, L1, L2, inverse).astype(np.complex64))[None]
        inH, inW, r0, c0 = region_in
        self.plan = LL.table_plan(1, 1, inH, inW, r0, c0, L1, L2, True, device, spec.to(device))
        self.plan.outH, self.plan.outW, self.plan.out_r0, self.plan.out_c0 = region_out
        self.out_shape = region_out[:2]

    def run(self, x, conj):
        """x [F, 1, ., .] -> convolution (conj=0) or its adjoint (conj=1; regions swapped, conjugate spectrum)."""
        p = self.plan
        oh, ow = (p.outH, p.outW) if not conj else (p.H, p.W)
        y = torch.empty(x.shape[0], 1, oh, ow, dtype=torch.complex64, device=x.device)
        p.run(x, y, conj=conj)
        return y


def _outer(a, b, scale=1.0):
    return torch.from_numpy((np.outer(a, b) * scale).astype(np.complex64)).reshape(-1).contiguous()


class BluesteinFft2:
    """Natural-order 2-D DFT (or inverse) of fields [., H, W] of any size through two pointwise chirps and one fused
    convolution.  norm: 'backward' (torch.fft default) or 'ortho'."""

    def __init__(self, H, W, inverse, ortho, device):
        w1, w2 = chirp(H, inverse), chirp(W, inverse)
        scale = (1.0 / np.sqrt(H * W)) if ortho else ((1.0 / (H * W)) if inverse else 1.0)
        self.pre = _outer(w1, w2).to(device)
        self.post = _outer(w1, w2, scale).to(device)
        self.conv = _Conv(H, W, (H, W, 0, 0), (H, W, 0, 0), inverse, device)
        self.H, self.W = H, W

    def __call__(self, x):
        shape = x.shape
        f = x.reshape(-1, 1, self.H, self.W)
        a = Fn.FieldMulFn._run(f, self.pre, False, False, 0)
        b = self.conv.run(a, 0)
        return Fn.FieldMulFn._run(b, self.post, False, False, 0).reshape(shape)


class BluesteinAsmPlan:
    """Band-limited angular-spectrum propagation on a padded grid whose edge lengths the radix plans do not cover
    (see the module docstring).  Same geometry arguments as functional.AsmPlan; `Hn` = ifftshift of the reference's
    centred transfer function, complex64 [C, Hp, Wp] (host)."""

    def __init__(self, C, H, W, pad_h, pad_w, Hp, Wp, unpad, device, Hn):
        self.C, self.H, self.W, self.Hp, self.Wp = C, H, W, Hp, Wp
        self.outH, self.outW, self.out_r0, self.out_c0 = (H, W, pad_h, pad_w) if unpad else (Hp, Wp, 0, 0)
        w1, w2 = chirp(Hp), chirp(Wp)
        self.w_in = _outer(w1[pad_h:pad_h + H], w2[pad_w:pad_w + W]).to(device)
        self.w_out = _outer(np.conj(w1)[self.out_r0:self.out_r0 + self.outH], np.conj(w2)[self.out_c0:self.out_c0 + self.outW],
                            1.0 / (Hp * Wp)).to(device)
        self.Hn = Hn.reshape(C, Hp * Wp).contiguous().to(device)
        self.conv1 = _Conv(Hp, Wp, (H, W, pad_h, pad_w), (Hp, Wp, 0, 0), False, device)
        self.conv2 = _Conv(Hp, Wp, (Hp, Wp, 0, 0), (self.outH, self.outW, self.out_r0, self.out_c0), True, device)
        self.B = 1
        self.tf_mode = 1

    def forward(self, x):
        B, C = x.shape[0], self.C
        a = Fn.FieldMulFn._run(x.reshape(B * C, 1, self.H, self.W), self.w_in, False, False, 0)
        X = self.conv1.run(a, 0).reshape(B, C, self.Hp, self.Wp)
        Y = Fn.FieldMulFn._run(X, self.Hn, False, True, 0)
        y = self.conv2.run(Y.reshape(B * C, 1, self.Hp, self.Wp), 0)
        return Fn.FieldMulFn._run(y, self.w_out, False, False, 0).reshape(B, C, self.outH, self.outW)

    def adjoint(self, g):
        B, C = g.shape[0], self.C
        g1 = Fn.FieldMulFn._run(g.reshape(B * C, 1, self.outH, self.outW), self.w_out, False, False, 1)
        G = self.conv2.run(g1, 1).reshape(B, C, self.Hp, self.Wp)
        G2 = Fn.FieldMulFn._run(G, self.Hn, False, True, 1)
        ga = self.conv1.run(G2.reshape(B * C, 1, self.Hp, self.Wp), 1)
        return Fn.FieldMulFn._run(ga, self.w_in, False, False, 1).reshape(B, C, self.H, self.W)


class BluesteinAsmFn(torch.autograd.Function):
    """y = ASM(x) on a grid of unsupported edge lengths; backward = the explicit adjoint chain."""

    @staticmethod
    def forward(ctx, x, plan):
        x = Fn._c64(x, "field.data")
        ctx.plan = plan
        return plan.forward(x)

    @staticmethod
    def backward(ctx, g):
        return ctx.plan.adjoint(Fn._c64(g, "grad_output")), None
