"""The boundary type: a 4-D complex field with its wavelengths and pixel pitch.

Mirror of the reference container (DataType/ElectricField.py:14-195): same constructor, same
validation and exception types (:76-109), same properties; the matplotlib visualisation helpers of
the reference (:199-440) are out of scope.  The propagators and DOE layers of this package accept
this class or the reference's own (duck-typed on .data / .wavelengths / .spacing).

One addition: a field may carry a *deferred DOE modulation* (`_deferred`), set by the DOE layers of
this package.  `.data` materialises it lazily with the stand-alone modulation kernel; ASM_prop
consumes it un-materialised and runs the fused DOE->ASM kernels instead (one HBM pass saved).
"""
from __future__ import annotations

import torch


class DeferredElements:
    """x * mask * mul, not yet evaluated: the output of an aperture / thin lens.  The next consumer fuses it --
    ASM_prop multiplies on load in its row-FFT prologue (and by the conjugate in the adjoint's epilogue), a DOE layer passes
    it on in front of its own phase; reading `.data` evaluates it with the stand-alone kernel.  Pointwise factors commute,
    so further elements simply fold into `mask` (float32 [H,W]) and `mul` (complex64 [C,H,W])."""
    is_elements = True

    def __init__(self, x, mask=None, mul=None):
        self.x, self.mask, self.mul = x, mask, mul
        self.shape = x.shape
        self.device = x.device
        self.uses = 0

    def materialise(self):
        from .. import functional as Fn
        x = self.x
        if self.mask is not None:
            x = Fn.FieldMulFn.apply(x, self.mask)
        if self.mul is not None:
            x = Fn.FieldMulFn.apply(x, self.mul)
        return x

    @staticmethod
    def peek(field):
        """The un-evaluated element chain `field` carries (None if there is none), without counting a use."""
        d = getattr(field, "_deferred", None) if getattr(field, "_data", None) is None else None
        return d if (d is not None and getattr(d, "is_elements", False)) else None

    @staticmethod
    def pending(field):
        """The un-evaluated element chain `field` carries, or None.  A chain that is consumed a SECOND time (a fixed
        field-in-front-of-the-DOE reused by every iteration of an optimisation loop, as the notebooks build it) is evaluated
        once instead and served as a plain tensor from then on: fusing saves a pass per use only if the chain is used once."""
        d = getattr(field, "_deferred", None) if getattr(field, "_data", None) is None else None
        if d is None or not getattr(d, "is_elements", False):
            return None
        d.uses += 1
        if d.uses > 1:
            field.data            # materialise (cached in the field)
            return None
        return d


class ElectricField:
    _BATCH = 0
    _WAVELENGTH = 1
    _HEIGHT = 2
    _WIDTH = 3

    def __init__(self, data, wavelengths=None, spacing=None, requires_grad=None, device=None):
        self.device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self._spacing = self.check_spacing(spacing)
        self._wavelengths = self.check_wavelengths(wavelengths)
        self.field_type = None
        self._deferred = None
        if data is None:
            data = torch.empty(1, len(self._wavelengths), 1, 1)
        self._data = self.check_data(data)

    # -- deferred DOE modulation (package-internal) ------------------------------------------
    @classmethod
    def _from_deferred(cls, deferred, like):
        """deferred: object with .materialise() -> Tensor and .shape; `like` supplies metadata."""
        self = cls.__new__(cls)
        self.device = deferred.device
        self._spacing = like.spacing
        self._wavelengths = like.wavelengths
        self._deferred = deferred
        self._data = None
        if getattr(like, "_row_slab", None) is not None:
            self._row_slab = like._row_slab
        B = deferred.shape[0]
        self.field_type = "scalar" if B == 1 else ("vectorial" if B == 3 else "batch")
        return self

    # -- reference surface -------------------------------------------------------------------
    @property
    def spacing(self):
        return self._spacing

    @spacing.setter
    def spacing(self, spacing):
        self._spacing = self.check_spacing(spacing)

    @property
    def wavelengths(self):
        return self._wavelengths

    @wavelengths.setter
    def wavelengths(self, wavelengths):
        self._wavelengths = self.check_wavelengths(wavelengths)

    @property
    def requires_grad(self):
        return self.data.requires_grad

    @property
    def data(self):
        if self._data is None and self._deferred is not None:
            self._data = self._deferred.materialise()
        return self._data

    @data.setter
    def data(self, data):
        self._deferred = None
        self._data = self.check_data(data)

    def check_spacing(self, spacing):
        if isinstance(spacing, (list, tuple)) and len(spacing) == 2:
            spacing = torch.tensor(spacing, dtype=torch.float32)
        elif isinstance(spacing, (float, int)):
            spacing = torch.tensor([spacing, spacing], dtype=torch.float32)
        if not torch.is_tensor(spacing) or spacing.numel() != 2:
            raise ValueError("Spacing must be a 2-element tensor.")
        return spacing.to(self.device)

    def check_wavelengths(self, wavelengths):
        if isinstance(wavelengths, (list, float, int)):
            wavelengths = torch.tensor([wavelengths] if isinstance(wavelengths, (float, int)) else wavelengths,
                                       dtype=torch.float32)
        if not torch.is_tensor(wavelengths):
            raise ValueError("Wavelengths must be a tensor.")
        return wavelengths.to(self.device)

    def check_data(self, data):
        assert torch.is_tensor(data) and data.ndim == 4, \
            "Data must be a 4D torch tensor with BATCH x Channel (Wavelength) x Height x Width"
        if data.shape[self._WAVELENGTH] != len(self._wavelengths):
            raise ValueError("The number of channels in data should be equal to the number of wavelengths")
        if data.shape[self._BATCH] == 1:
            self.field_type = "scalar"
        elif data.shape[self._BATCH] == 3:
            self.field_type = "vectorial"
        else:
            self.field_type = "batch"
        return data.to(self.device)

    def abs(self):
        return ElectricField(data=self.data.abs(), wavelengths=self._wavelengths, spacing=self._spacing, device=self.device)

    def angle(self):
        return ElectricField(data=self.data.angle(), wavelengths=self._wavelengths, spacing=self._spacing, device=self.device)

    def detach(self):
        return ElectricField(data=self.data.detach(), wavelengths=self._wavelengths.detach(),
                             spacing=self._spacing.detach(), device=self.device)

    def cpu(self):
        return ElectricField(data=self.data.cpu(), wavelengths=self._wavelengths.detach().cpu(),
                             spacing=self._spacing.detach().cpu(), device=torch.device("cpu"))

    @property
    def ndim(self):
        return 4 if self._data is None else self._data.ndim

    @property
    def shape(self):
        return self._deferred.shape if self._data is None and self._deferred is not None else self._data.shape

    @property
    def num_batches(self):
        return self.shape[self._BATCH]

    @property
    def num_wavelengths(self):
        return self.shape[self._WAVELENGTH]

    @property
    def height(self):
        return self.shape[self._HEIGHT]

    @property
    def width(self):
        return self.shape[self._WIDTH]

    @property
    def Ex(self):
        return self.data[[0], ...]

    @property
    def Ey(self):
        return self.data[[1], ...]

    @property
    def Ez(self):
        return self.data[[2], ...]

    def _get_data_for_wavelength(self, wavelength):
        idx = (self._wavelengths == wavelength).nonzero()[0]
        return self.data[:, idx, ...]
