"""Thin lens -- drop-in for the reference's Components/Thin_Lens.py.

Same constructor and forward(field) -> ElectricField surface.  The lens kernel exp(-i pi (x^2 + y^2) / (lambda f)) is formed
once per (shape, spacing, wavelengths, f) with the reference's own expressions (Components/Thin_Lens.py:33-64) and applied by
`thz_field_mul` (forward) / its conjugate (backward).
"""
import numpy as np
import torch
import torch.nn as nn

from .. import functional as Fn
from ..DataType.ElectricField import DeferredElements, ElectricField


class Thin_LensElement(nn.Module):

    def __init__(self, focal_length, device=None):
        super().__init__()
        self.device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.focal_length = torch.Tensor([focal_length]).to(self.device)
        self._key, self._ker = None, None

    def create_lens_phase_shift_kernel(self, field):
        """[1,C,H,W] complex64 lens kernel, Components/Thin_Lens.py:33-64 (cached)."""
        pend = DeferredElements.peek(field)
        dev = pend.device if pend is not None else field.data.device       # reading .data would evaluate a pending aperture
        key = (tuple(field.shape[-2:]), tuple(field.spacing.detach().cpu().tolist()), tuple(field.wavelengths.detach().cpu().tolist()),
               float(self.focal_length), str(dev))
        if key != self._key:
            dx, dy = field.spacing[0].detach().cpu(), field.spacing[1].detach().cpu()
            lam = field.wavelengths.detach().cpu().float()[:, None, None]
            height, width = field.height, field.width
            xc = torch.linspace(-((height - 1) // 2), (height - 1) // 2, height)
            yc = torch.linspace(-((width - 1) // 2), (width - 1) // 2, width)
            xg, yg = torch.meshgrid(xc, yc, indexing="ij")
            xg, yg = xg[None, None] * dx, yg[None, None] * dy
            ang = -(np.pi / (lam * self.focal_length.detach().cpu())) * ((xg ** 2) + (yg ** 2))
            self._ker = torch.exp(1j * ang).to(torch.complex64).to(dev).contiguous()
            self._key = key
        return self._ker

    def forward(self, field):
        """Deferred: the multiply is folded into the next propagation's prologue (or evaluated by thz_field_mul when someone
        reads `.data`)."""
        ker = self.create_lens_phase_shift_kernel(field)[0]
        pend = DeferredElements.pending(field)
        if pend is not None:
            mul = ker if pend.mul is None else self._combined(pend.mul, ker)
            d = DeferredElements(pend.x, pend.mask, mul)
        else:
            data = field.data
            Fn.N.require_cuda(data, "field.data")
            d = DeferredElements(data, None, ker)
        return ElectricField._from_deferred(d, field)

    def _combined(self, a, b):
        """Product of two per-wavelength kernels, formed once per pair of tensors (stacked lenses)."""
        key = (id(a), a._version, id(b), b._version)
        if getattr(self, "_comb_key", None) != key:
            self._comb, self._comb_key, self._comb_refs = (a * b).contiguous(), key, (a, b)
        return self._comb
