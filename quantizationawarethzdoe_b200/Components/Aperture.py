"""Aperture -- drop-in for the reference's Components/Aperture.py.

Same constructor and forward(field) -> ElectricField surface ('circ' with a radius, 'rect' with a side length, None = open).
The 0/1 mask is formed once per (shape, spacing) with the reference's expressions (Components/Aperture.py:41-103) and applied
by `thz_field_mul`.
"""
import torch
import torch.nn as nn

from .. import functional as Fn
from ..DataType.ElectricField import DeferredElements, ElectricField


class ApertureElement(nn.Module):

    def __init__(self, aperture_type="circ", aperture_size=None, device=None):
        super().__init__()
        self.device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.aperture_type = aperture_type
        self.aperture_size = aperture_size
        self.aperture = None
        self._key = None

    def add_circ_aperture_to_field(self, input_field, radius=None):
        """Components/Aperture.py:41-72.  (The reference leaves `r` unbound for a radius >= half the field; here: ValueError.)"""
        dx, dy = input_field.spacing[0].detach().cpu(), input_field.spacing[1].detach().cpu()
        height, width = input_field.height, input_field.width
        half = min([dx * height, dy * width]) / 2.0
        if radius is None:
            r = half
        elif torch.tensor(radius) < half:
            r = torch.tensor(radius)
        else:
            raise ValueError("The radius should not larger than the physical length of E-field ")
        x = torch.linspace(-dx * height / 2, dx * height / 2, height, dtype=dx.dtype)
        y = torch.linspace(-dy * width / 2, dy * width / 2, width, dtype=dy.dtype)
        X, Y = torch.meshgrid(x, y, indexing="ij")
        return torch.where(torch.sqrt(X ** 2 + Y ** 2) <= r, 1, 0)[None, None]

    def add_rect_aperture_to_field(self, input_field, rect_width=None, rect_height=None):
        """Components/Aperture.py:74-103."""
        dx, dy = input_field.spacing[0].detach().cpu(), input_field.spacing[1].detach().cpu()
        height, width = input_field.height, input_field.width
        if rect_width is None:
            rect_width = dx * width / 2
        if rect_height is None:
            rect_height = dy * height / 2
        rect_width = min(rect_width, dx * width)
        rect_height = min(rect_height, dy * height)
        x = torch.linspace(-dx * width / 2, dx * width / 2, width, dtype=dx.dtype)
        y = torch.linspace(-dy * height / 2, dy * height / 2, height, dtype=dy.dtype)
        X, Y = torch.meshgrid(x, y, indexing="xy")
        return torch.where((torch.abs(X) <= rect_width / 2) & (torch.abs(Y) <= rect_height / 2), 1, 0)[None, None]

    def forward(self, field):
        pend = DeferredElements.pending(field)
        data = pend.x if pend is not None else field.data
        Fn.N.require_cuda(data, "field.data")
        key = (tuple(data.shape[-2:]), tuple(field.spacing.detach().cpu().tolist()), self.aperture_type,
               None if self.aperture_size is None else float(self.aperture_size), str(data.device))
        if key != self._key:
            if self.aperture_type == "circ":
                mask = self.add_circ_aperture_to_field(field, radius=self.aperture_size)
            elif self.aperture_type == "rect":
                mask = self.add_rect_aperture_to_field(field, rect_height=self.aperture_size, rect_width=self.aperture_size)
            elif self.aperture_type is None:
                mask = torch.ones(1, 1, *data.shape[-2:])
            else:
                raise ValueError("No exisiting aperture shape, please define by yourself")
            self.aperture = mask.to(data.device)
            self._mask_f32 = mask[0, 0].to(torch.float32).to(data.device).contiguous()
            self._key = key
        # deferred: folded into the next propagation's prologue (or evaluated by thz_field_mul when someone reads `.data`)
        if pend is not None:
            mask = self._mask_f32 if pend.mask is None else self._combined(pend.mask, self._mask_f32)
            d = DeferredElements(pend.x, mask, pend.mul)
        else:
            d = DeferredElements(data, self._mask_f32, None)
        return ElectricField._from_deferred(d, field)

    def _combined(self, a, b):
        key = (id(a), a._version, id(b), b._version)
        if getattr(self, "_comb_key", None) != key:
            self._comb, self._comb_key, self._comb_refs = (a * b).contiguous(), key, (a, b)
        return self._comb
