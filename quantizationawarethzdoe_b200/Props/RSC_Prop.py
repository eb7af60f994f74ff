"""Rayleigh-Sommerfeld convolution -- drop-in for the reference's Props/RSC_Prop.py.

Same constructor, attributes and forward(field) -> ElectricField surface as `RSC_prop` (Props/RSC_Prop.py:15-215) and
`VRS_prop` (:218-321).  The reference computes

    U[..., :H, :W] = x;   y = ifft2( fft2(U) * fft2(K_rs) * dx * dy )[..., H:, W:]          (:196-207, Shen & Wang Eq. 11-15)

with the spatial impulse response K_rs(x, y) = exp(i k r) z / (2 pi r^2) (1 / r - i k) on the padded grid (:161-163).  That is
the pad -> FFT -> multiply -> iFFT -> crop skeleton of the fused ASM pipeline with the input in the upper-left corner, the
crop in the lower-right one and a cached transfer function: the spectrum of K_rs is formed once per
(shape, spacing, wavelengths, z) -- the impulse response with the reference's own torch ops, its FFT with this package's
own kernels -- and streamed by the column kernel (tf_mode 1).  Backward is the explicit adjoint pipeline (conjugate table,
regions swapped).  No torch.fft / cuFFT on this path.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import _native as N
from .. import functional as Fn
from .. import longline as LL
from ..DataType.ElectricField import ElectricField

mm = 1e-3


class RSC_prop(nn.Module):

    def __init__(self, z_distance=0.0, device=None):
        super().__init__()
        self.do_padding = True
        self.DEFAULT_PADDING_SCALE = torch.tensor([1, 1])
        self.device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self._z = torch.tensor(z_distance, device=self.device)
        self.shape = None
        self.meshx = None
        self.meshy = None
        self.check_Zc = True
        self._plan_key = None
        self._plan = None

    @property
    def z(self):
        return self._z

    @z.setter
    def z(self, z):
        if not isinstance(z, torch.Tensor):
            z = torch.tensor(z, device=self.device)
        elif z.device != self.device:
            z = z.to(self.device)
        self._z = z

    def compute_padding(self, H, W, return_size_of_padding=False):
        if not self.do_padding:
            pad_h, pad_w, Hp, Wp = 0, 0, int(H), int(W)
        else:
            pad_h = int(np.floor(float(self.DEFAULT_PADDING_SCALE[0]) * int(H) / 2))     # Props/RSC_Prop.py:69-72
            pad_w = int(np.floor(float(self.DEFAULT_PADDING_SCALE[1]) * int(W) / 2))
            Hp, Wp = int(H) + 2 * pad_h, int(W) + 2 * pad_w
        return (pad_h, pad_w) if return_size_of_padding else (Hp, Wp)

    def create_spatial_grid(self, H, W, dx, dy):
        """Props/RSC_Prop.py:79-87 (the reference spaces BOTH axes with dx)."""
        x = torch.linspace(-H * dx / 2, H * dx / 2, H)
        y = torch.linspace(-W * dx / 2, W * dx / 2, W)
        meshx, meshy = torch.meshgrid(x, y, indexing="ij")
        return meshx.to(device=self.device), meshy.to(device=self.device)

    def check_RS_minimum_z(self, quality_factor=1, dx=None, dy=None, wavelength=None):
        """Props/RSC_Prop.py:89-127: prints the minimum trustworthy distance (energy conservation / sampling)."""
        range_x, range_y = self.shape[-2] * dx, self.shape[-1] * dy
        dr_real = torch.sqrt(dx ** 2 + dy ** 2)
        rmax = torch.sqrt(range_x ** 2 + range_y ** 2)
        factor = (((quality_factor * dr_real + rmax) ** 2 - wavelength ** 2 - rmax ** 2) / (2 * wavelength)) ** 2 - rmax ** 2
        z_min1 = torch.sqrt(factor) if factor > 0 else torch.zeros(())
        print("Minimum propagation distance to satisfy energy conservation: {:.3f} mm".format(float(z_min1) / mm))
        z_min2 = self.meshx.shape[0] * dx ** 2 / wavelength * torch.sqrt(1 - (wavelength / (2 * dx)) ** 2)
        print("Minimum propagation distance to satisfy sampling for FT: {:.3f} mm".format(float(z_min2) / mm))
        if float(self._z) > min(float(z_min1), float(z_min2)):
            print("The simulation will be accurate !")
        else:
            print("The propagation distance should be larger than minimum propagation distance to keep simulation accurate!")

    def _kernel_host(self, Hp, Wp, dx, dy, wavelengths, z):
        """Spatial impulse response [C,Hp,Wp] complex64 with the reference's own torch ops, on the host (:146-163)."""
        x = torch.linspace(-Hp * dx / 2, Hp * dx / 2, Hp)
        y = torch.linspace(-Wp * dx / 2, Wp * dx / 2, Wp)
        meshx, meshy = torch.meshgrid(x, y, indexing="ij")
        k = 2 * torch.pi / wavelengths[:, None, None]
        r = torch.sqrt(meshx ** 2 + meshy ** 2 + z ** 2)
        factor = 1 / (2 * torch.pi) * z / r ** 2 * (1 / r - 1j * k)
        return (torch.exp(1j * k * r) * factor).to(torch.complex64)

    def create_kernel(self, field):
        """[1,C,Hp,Wp] spatial kernel as the reference returns it (Props/RSC_Prop.py:129-168); for inspection."""
        Hp, Wp = self.compute_padding(field.shape[-2], field.shape[-1])
        dx, dy = field.spacing[0].detach().cpu(), field.spacing[1].detach().cpu()
        self.meshx, self.meshy = self.create_spatial_grid(Hp, Wp, float(dx), float(dy))
        K = self._kernel_host(Hp, Wp, dx, dy, field.wavelengths.detach().cpu().float(), self._z.detach().cpu().float())
        if self.check_Zc:
            self.check_RS_minimum_z(quality_factor=1, dx=dx, dy=dy, wavelength=torch.min(field.wavelengths.detach().cpu()))
            self.check_Zc = False
        return K[None].to(self.device)

    def _get_plan(self, field, B, C, H, W, device):
        spacing, wavelengths = field.spacing, field.wavelengths
        z = self._z.detach().cpu().float().reshape(())
        key = (C, H, W, tuple(spacing.detach().cpu().reshape(-1).tolist()), tuple(wavelengths.detach().cpu().reshape(-1).tolist()),
               float(z), str(device), self.do_padding)
        if key != self._plan_key:
            Hp, Wp = self.compute_padding(H, W)
            if Hp - H < H or Wp - W < W:
                raise ValueError("RSC_prop needs a padded grid of at least twice the field (lower-right submatrix, Props/RSC_Prop.py:207)")
            dx, dy = spacing[0].detach().cpu(), spacing[1].detach().cpu()
            if self.check_Zc:
                self.meshx, self.meshy = self.create_spatial_grid(Hp, Wp, float(dx), float(dy))
                self.check_RS_minimum_z(quality_factor=1, dx=dx, dy=dy, wavelength=torch.min(wavelengths.detach().cpu()))
                self.check_Zc = False
            K = self._kernel_host(Hp, Wp, dx, dy, wavelengths.detach().cpu().float(), z).to(device)
            spec = Fn.fft2_c2c(K) * (dx * dy).to(device)                     # fft2(K_rs) dx dy, natural bin order (our FFT kernels)
            plan = LL.table_plan(B, C, H, W, 0, 0, Hp, Wp, True, device, spec)   # cached table[c][slot_c][slot_r] (split above 16384)
            plan.out_r0, plan.out_c0 = H, W                                  # lower-right submatrix (:207); input sits at (0, 0) (:199)
            self._plan, self._plan_key = plan, key
        self._plan.B = B
        return self._plan

    def forward(self, field):
        data = field.data
        B, C, H, W = self.shape = data.shape
        plan = self._get_plan(field, B, C, H, W, data.device)
        out = Fn.AsmPropagateFn.apply(data, plan)
        return ElectricField(data=out, wavelengths=field.wavelengths, spacing=field.spacing, device=data.device)


class VRS_prop(RSC_prop):
    """Vectorial Rayleigh-Sommerfeld (Props/RSC_Prop.py:218-321): Ez = (Ex x + Ey y) / r, then each of the three
    components is propagated with the scalar kernel; returns a [3,C,H,W] field."""

    def forward(self, field):
        B, C, H, W = self.shape = field.shape
        dx, dy = field.spacing[0], field.spacing[1]
        meshx, meshy = self.create_spatial_grid(H, W, float(dx), float(dy))
        meshx, meshy = meshx.to(field.data.device), meshy.to(field.data.device)
        r = torch.sqrt(meshx ** 2 + meshy ** 2 + self._z.to(field.data.device) ** 2)
        Ex, Ey = field.Ex, field.Ey
        Ez = Ex * meshx / r + Ey * meshy / r                                  # :289 (Ref 1 Eq. 2c)
        vec = torch.cat((Ex, Ey, Ez), dim=0).contiguous()
        plan = self._get_plan(field, 3, C, H, W, vec.device)
        out = Fn.AsmPropagateFn.apply(vec, plan)
        return ElectricField(data=out, wavelengths=field.wavelengths, spacing=field.spacing, device=vec.device)
