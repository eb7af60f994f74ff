"""Host-side preparation for the chirp-z (Bluestein) propagator as separable Toeplitz GEMMs.

The reference (Props/CZT_Prop.py) runs two 1-D Bluestein passes with FFTs.  Algebraically
    out[c] = F0[c] * ( Wy[c] . (x[0,c] * F[c]) . Wx[c]^T ) * z dx' dy' lambda_c
with  W[k, i] = post[k] * g[(m + k - i) mod np2] * pre[i],  g = 1/h zero-extended to np2  (SURVEY 8 a-6).
Everything O(m + M) (chirp vectors) and the two O(H W) / O(M^2) Rayleigh-Sommerfeld factor tables are
built HERE, once per geometry, with the reference's own torch CPU expressions and dtype rules -- the
reference's fp32 chirps are only ~4e-3 accurate against exact math, so parity requires reproducing them,
not improving them.  The O(M H W) work (the GEMMs) runs on the GPU (thz_toeplitz_gemm).
"""
import numpy as np
import torch


def rs_kernel(z, meshx, meshy, wavelengths):
    """Props/CZT_Prop.py:44-57."""
    lam = wavelengths[None, :, None, None]
    k = 2 * torch.pi / lam
    r = torch.sqrt(meshx ** 2 + meshy ** 2 + z ** 2)
    factor = 1 / (2 * torch.pi) * z / r ** 2 * (1 / r - 1j * k)
    return torch.exp(1j * k * r) * factor


def build_grid(z, wavelengths, in_h, in_w, in_dx, in_dy, out_h, out_w, out_dx, out_dy):
    """Props/CZT_Prop.py:95-118."""
    x_in = torch.linspace(-in_h * in_dx / 2, in_h * in_dx / 2, in_h)
    y_in = torch.linspace(-in_w * in_dy / 2, in_w * in_dy / 2, in_w)
    in_mx, in_my = torch.meshgrid(x_in, y_in, indexing="ij")
    x_out = torch.linspace(-out_h * out_dx / 2, out_h * out_dx / 2, out_h)
    y_out = torch.linspace(-out_w * out_dy / 2, out_w * out_dy / 2, out_w)
    out_mx, out_my = torch.meshgrid(x_out, y_out, indexing="ij")
    Dm = wavelengths[None, :, None, None] * z / in_dx
    return (in_mx, in_my, out_mx, out_my, Dm, x_out[0] + Dm / 2, x_out[-1] + Dm / 2, y_out[0] + Dm / 2, y_out[-1] + Dm / 2)


def next_pow2(x):
    """Props/CZT_Prop.py:130."""
    return int(2 ** (np.ceil(np.log2(x))).astype(int))


def chirp_vectors(f1, f2, Dm, m, M_out):
    """One Bluestein pass (Props/CZT_Prop.py:149-155, 164, 200-206, 211-221) reduced to three vectors:
    pre [C,m], post [C,M_out] (window * frequency shift), g [C,np2] (the chirp filter as the FFT sees it)."""
    D1 = f1 + (M_out * Dm + f2 - f1) / (2 * M_out)
    D2 = f2 + (M_out * Dm + f2 - f1) / (2 * M_out)
    mp = m + M_out - 1
    np2 = next_pow2(mp)
    A = torch.exp(1j * 2 * torch.pi * D1 / Dm)
    W = torch.exp(-1j * 2 * torch.pi * (D1 - D2) / (M_out * Dm))
    e = torch.arange(-m + 1, max(M_out - 1, m - 1) + 1)
    h = W ** (e ** 2 / 2)                                    # [1,C,1,len]  (the reference's h[:mp+1] slices dim 0: no-op)
    pre = A ** (-(torch.arange(0, m))) * h[..., torch.arange(m - 1, 2 * m - 1)]
    l = torch.linspace(0, M_out - 1, M_out)[None, None, None, :]
    l = l / M_out * (D2 - D1) + D1
    mshift = torch.exp(-1j * 2 * torch.pi * l * (-m / 2 + 1 / 2) / Dm)
    post = h[..., m - 1:mp] * mshift
    hh = h[0, :, 0, :]
    g = torch.zeros(hh.shape[0], np2, dtype=hh.dtype)
    n = min(hh.shape[-1], np2)                               # fft(1/h, n=np2) truncates or zero-extends (:161)
    g[:, :n] = (1 / hh)[:, :n]
    return pre[0, :, 0, :].contiguous(), post[0, :, 0, :].contiguous(), g.contiguous(), np2


class CztPlan:
    """Static data of one CZT geometry (host tensors, complex64):
       P  [C,H,W]   = F * pre_y[h] * pre_x[w]            input-side table
       Q  [C,M1,M2] = F0 * post_y[k1] * post_x[k2] * s_c  output-side table (s_c = z dx' dy' lambda_c)
       gy [C,Ly], gx [C,Lx]                               chirp filters (Toeplitz generators)"""

    def __init__(self, wavelengths, spacing, z, in_h, in_w, out_h, out_w, out_dx, out_dy):
        wavelengths = torch.as_tensor(wavelengths).detach().cpu().reshape(-1)
        if not wavelengths.is_floating_point():
            wavelengths = wavelengths.float()
        spacing = torch.as_tensor(spacing).detach().cpu().reshape(-1)
        z = torch.as_tensor(z).detach().cpu()
        in_dx, in_dy = spacing[0], spacing[1]
        if out_h != out_w:
            # the reference multiplies a [.., outW, outH] result by a [.., outH, outW] kernel (CZT_Prop.py:248)
            raise RuntimeError("CZT_prop: outputHeight must equal outputWidth (the reference fails with a shape error otherwise)")
        (in_mx, in_my, out_mx, out_my, Dm, fx1, fx2, fy1, fy2) = build_grid(
            z, wavelengths, in_h, in_w, in_dx, in_dy, out_h, out_w, out_dx, out_dy)
        F0 = rs_kernel(z, out_mx, out_my, wavelengths)[0]
        Fi = rs_kernel(z, in_mx, in_my, wavelengths)[0]
        pre_y, post_y, gy, Ly = chirp_vectors(fy1, fy2, Dm, in_h, out_w)     # acts on the H axis (CZT_Prop.py:243)
        pre_x, post_x, gx, Lx = chirp_vectors(fx1, fx2, Dm, in_w, out_h)     # acts on the W axis (:246)
        scale = (z * out_dx * out_dy * wavelengths)[:, None, None]
        self.P = (Fi * pre_y[:, :, None] * pre_x[:, None, :]).to(torch.complex64).contiguous()
        self.Q = (F0 * post_y[:, :, None] * post_x[:, None, :] * scale).to(torch.complex64).contiguous()
        self.gy, self.gx = gy.to(torch.complex64), gx.to(torch.complex64)
        self.Ly, self.Lx = Ly, Lx
        self.C, self.H, self.W = wavelengths.numel(), in_h, in_w
        self.M1, self.M2 = out_w, out_h                       # rows of the output come from the H-axis pass
        self.out_spacing = [out_dx, out_dy]
