"""CPU restatement of the reference's Rayleigh-Sommerfeld convolution (TEST INFRASTRUCTURE ONLY).

Follows Props/RSC_Prop.py:129-215 (RSC_prop) and :218-321 (VRS_prop) with torch CPU ops, torch.fft included; pinned by
tests/golden/rsc_*.npz, which oracle/make_golden.py produces with the unmodified reference.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import this file.
"""
import torch


def rs_kernel(Hp, Wp, dx, wavelengths, z):
    """Props/RSC_Prop.py:79-87, 156-163 (both axes are spaced with dx in the reference)."""
    x = torch.linspace(-Hp * dx / 2, Hp * dx / 2, Hp)
    y = torch.linspace(-Wp * dx / 2, Wp * dx / 2, Wp)
    meshx, meshy = torch.meshgrid(x, y, indexing="ij")
    k = 2 * torch.pi / wavelengths[:, None, None]
    r = torch.sqrt(meshx ** 2 + meshy ** 2 + z ** 2)
    factor = 1 / (2 * torch.pi) * z / r ** 2 * (1 / r - 1j * k)
    return torch.exp(1j * k * r) * factor


def rsc_forward(x, wavelengths, spacing, z):
    """x [B,C,H,W] complex64 -> [B,C,H,W]; Props/RSC_Prop.py:196-207."""
    B, C, H, W = x.shape
    wl = torch.as_tensor(wavelengths, dtype=torch.float32)
    sp = torch.as_tensor(spacing, dtype=torch.float32).reshape(-1)
    sp = sp.repeat(2) if sp.numel() == 1 else sp
    Hp, Wp = H + 2 * (H // 2), W + 2 * (W // 2)
    K = rs_kernel(Hp, Wp, sp[0], wl, torch.as_tensor(z, dtype=torch.float32))[None]
    U = torch.zeros(B, C, Hp, Wp, dtype=x.dtype)
    U[..., 0:H, 0:W] = x
    spec = torch.fft.fft2(U) * torch.fft.fft2(K) * sp[0] * sp[1]
    return torch.fft.ifft2(spec)[..., H:, W:]


def vrs_forward(x, wavelengths, spacing, z):
    """x [>=2,C,H,W] (Ex, Ey, ...) -> [3,C,H,W]; Props/RSC_Prop.py:281-304."""
    _, C, H, W = x.shape
    sp = torch.as_tensor(spacing, dtype=torch.float32).reshape(-1)
    sp = sp.repeat(2) if sp.numel() == 1 else sp
    xs = torch.linspace(-H * sp[0] / 2, H * sp[0] / 2, H)
    ys = torch.linspace(-W * sp[0] / 2, W * sp[0] / 2, W)
    meshx, meshy = torch.meshgrid(xs, ys, indexing="ij")
    r = torch.sqrt(meshx ** 2 + meshy ** 2 + torch.as_tensor(z, dtype=torch.float32) ** 2)
    Ex, Ey = x[[0]], x[[1]]
    Ez = Ex * meshx / r + Ey * meshy / r
    return rsc_forward(torch.cat((Ex, Ey, Ez), dim=0), wavelengths, spacing, z)
