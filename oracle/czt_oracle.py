"""CPU oracle for chirp-z (Bluestein) zoomed Rayleigh-Sommerfeld propagation.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Restates (torch CPU, reference dtype rules and op order):
  Props/CZT_Prop.py:44-57    RS_kernel
  Props/CZT_Prop.py:59-118   build_CZT_grid
  Props/CZT_Prop.py:120-130  compute_np2
  Props/CZT_Prop.py:132-177  compute_fft
  Props/CZT_Prop.py:179-225  Bluestein_method
  Props/CZT_Prop.py:227-250  CZT
  Props/CZT_Prop.py:252-314  forward
`czt_forward` is the FFT-based algorithm exactly as the reference runs it; `czt_forward_dense`
is the algebraically equivalent separable dense form  F0 * (Wy . (x*F) . Wx^T) * s_c  that the
CUDA GEMM path implements, built from the same reference-order chirp vectors.
Parity is PINNED by tests/test_oracle_golden.py and tests/golden/czt_*.npz.
"""
import numpy as np
import torch


def rs_kernel(z, meshx, meshy, wavelengths):
    """CZT_Prop.py:44-57."""
    lam = wavelengths[None, :, None, None]
    k = 2 * torch.pi / lam
    r = torch.sqrt(meshx ** 2 + meshy ** 2 + z ** 2)
    factor = 1 / (2 * torch.pi) * z / r ** 2 * (1 / r - 1j * k)
    return torch.exp(1j * k * r) * factor


def build_grid(z, wavelengths, in_h, in_w, in_dx, in_dy, out_h, out_w, out_dx, out_dy):
    """CZT_Prop.py:95-118."""
    x_in = torch.linspace(-in_h * in_dx / 2, in_h * in_dx / 2, in_h)
    y_in = torch.linspace(-in_w * in_dy / 2, in_w * in_dy / 2, in_w)
    in_mx, in_my = torch.meshgrid(x_in, y_in, indexing="ij")
    x_out = torch.linspace(-out_h * out_dx / 2, out_h * out_dx / 2, out_h)
    y_out = torch.linspace(-out_w * out_dy / 2, out_w * out_dy / 2, out_w)
    out_mx, out_my = torch.meshgrid(x_out, y_out, indexing="ij")
    Dm = wavelengths[None, :, None, None] * z / in_dx
    fx1 = x_out[0] + Dm / 2
    fx2 = x_out[-1] + Dm / 2
    fy1 = y_out[0] + Dm / 2
    fy2 = y_out[-1] + Dm / 2
    return in_mx, in_my, out_mx, out_my, Dm, fx1, fx2, fy1, fy2


def next_pow2(x):
    """CZT_Prop.py:130."""
    return int(2 ** (np.ceil(np.log2(x))).astype(int))


def chirp_vectors(f1, f2, Dm, m, M_out):
    """The O(m+M) vectors of one Bluestein pass, with the reference's expressions (:149-155, :164,
    :200-202, :214-221).  Shapes [1,C,1,*].  Returns dict(h, pre, post, g, np2, mp)."""
    D1 = f1 + (M_out * Dm + f2 - f1) / (2 * M_out)
    D2 = f2 + (M_out * Dm + f2 - f1) / (2 * M_out)
    mp = m + M_out - 1
    np2 = next_pow2(mp)
    A = torch.exp(1j * 2 * torch.pi * D1 / Dm)
    W = torch.exp(-1j * 2 * torch.pi * (D1 - D2) / (M_out * Dm))
    e = torch.arange(-m + 1, max(M_out - 1, m - 1) + 1)
    h = W ** (e ** 2 / 2)                                   # [1,C,1,len]; h[:mp+1] slices dim 0 -> no-op (:157)
    pre = A ** (-(torch.arange(0, m))) * h[..., torch.arange(m - 1, 2 * m - 1)]
    l = torch.linspace(0, M_out - 1, M_out)[None, None, None, :]
    l = l / M_out * (D2 - D1) + D1
    mshift = torch.exp(-1j * 2 * torch.pi * l * (-m / 2 + 1 / 2) / Dm)
    post = h[..., m - 1:mp] * mshift
    return dict(h=h, pre=pre, post=post, mshift=mshift, np2=np2, mp=mp)


def bluestein(x, f1, f2, Dm, M_out):
    """CZT_Prop.py:179-225 incl. compute_fft (:132-177): transforms dim -2, returns transposed."""
    _, _, m, n = x.shape
    v = chirp_vectors(f1, f2, Dm, m, M_out)
    h, np2, mp = v["h"], v["np2"], v["mp"]
    ft = torch.fft.fft(1 / h, n=np2, dim=-1)
    tmp = torch.tile(v["pre"], (1, 1, n, 1)).transpose(-2, -1)
    b = torch.fft.fft(x * tmp, np2, dim=-2)
    b = torch.fft.ifft(b * torch.tile(ft, (1, 1, n, 1)).transpose(-2, -1), dim=-2)
    b = b[..., m:mp + 1, 0:n].transpose(-2, -1) * torch.tile(h[..., m - 1:mp], (1, 1, n, 1))
    return b * torch.tile(v["mshift"], (1, 1, n, 1))


def _resolve(x, spacing, out_h, out_w, out_dx, out_dy):
    in_h, in_w = x.shape[-2], x.shape[-1]
    spacing = torch.as_tensor(spacing, dtype=torch.float32).reshape(-1)
    in_dx, in_dy = spacing[0], spacing[1]
    out_h = in_h if out_h is None else out_h
    out_w = in_w if out_w is None else out_w
    out_dx = in_dx if out_dx is None else out_dx
    out_dy = in_dy if out_dy is None else out_dy
    return in_h, in_w, in_dx, in_dy, out_h, out_w, out_dx, out_dy


def czt_forward(x, wavelengths, spacing, z, out_h=None, out_w=None, out_dx=None, out_dy=None):
    """CZT_Prop.py:252-314 / :227-250 on a raw [1,C,H,W] complex tensor (FFT-based, as the reference)."""
    wavelengths = torch.as_tensor(wavelengths).reshape(-1)
    if not wavelengths.is_floating_point():
        wavelengths = wavelengths.float()
    z = torch.as_tensor(z)
    in_h, in_w, in_dx, in_dy, out_h, out_w, out_dx, out_dy = _resolve(x, spacing, out_h, out_w, out_dx, out_dy)
    in_mx, in_my, out_mx, out_my, Dm, fx1, fx2, fy1, fy2 = build_grid(
        z, wavelengths, in_h, in_w, in_dx, in_dy, out_h, out_w, out_dx, out_dy)
    F0 = rs_kernel(z, out_mx, out_my, wavelengths)
    Fi = rs_kernel(z, in_mx, in_my, wavelengths)
    u = x * Fi
    u = bluestein(u, fy1, fy2, Dm, out_w)
    u = bluestein(u, fx1, fx2, Dm, out_h)
    return F0 * u * z * out_dx * out_dy * wavelengths[None, :, None, None]


def toeplitz_matrix(v, m, M_out):
    """Dense [C, M_out, m] matrix of one Bluestein pass: post[k] * g[(m+k-i) mod np2] * pre[i],
    g = (1/h) truncated / zero-extended to np2 exactly as fft(1/h, n=np2) does (:161)."""
    h, np2 = v["h"][0, :, 0, :], v["np2"]
    C = h.shape[0]
    g = torch.zeros(C, np2, dtype=h.dtype)
    L = min(h.shape[-1], np2)
    g[:, :L] = (1 / h)[:, :L]
    k = torch.arange(M_out)[:, None]
    i = torch.arange(m)[None, :]
    T = g[:, (m + k - i) % np2]                              # [C, M_out, m]
    return v["post"][0, :, 0, :, None] * T * v["pre"][0, :, 0, None, :]


def czt_forward_dense(x, wavelengths, spacing, z, out_h=None, out_w=None, out_dx=None, out_dy=None):
    """Separable dense form: out[c] = F0[c] * (Wy[c] @ (x[0,c]*F[c]) @ Wx[c]^T) * z dx' dy' lambda_c."""
    wavelengths = torch.as_tensor(wavelengths).reshape(-1)
    if not wavelengths.is_floating_point():
        wavelengths = wavelengths.float()
    z = torch.as_tensor(z)
    in_h, in_w, in_dx, in_dy, out_h, out_w, out_dx, out_dy = _resolve(x, spacing, out_h, out_w, out_dx, out_dy)
    in_mx, in_my, out_mx, out_my, Dm, fx1, fx2, fy1, fy2 = build_grid(
        z, wavelengths, in_h, in_w, in_dx, in_dy, out_h, out_w, out_dx, out_dy)
    F0 = rs_kernel(z, out_mx, out_my, wavelengths)
    Fi = rs_kernel(z, in_mx, in_my, wavelengths)
    Wy = toeplitz_matrix(chirp_vectors(fy1, fy2, Dm, in_h, out_w), in_h, out_w)   # acts on the H axis (:243)
    Wx = toeplitz_matrix(chirp_vectors(fx1, fx2, Dm, in_w, out_h), in_w, out_h)   # acts on the W axis (:246)
    u = (x * Fi)[0]
    u = torch.matmul(torch.matmul(Wy, u), Wx.transpose(-2, -1))
    return F0 * u[None] * z * out_dx * out_dy * wavelengths[None, :, None, None]
