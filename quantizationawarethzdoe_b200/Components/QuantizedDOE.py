"""DOE layers -- drop-in for the hot-path part of the reference's Components/QuantizedDOE.py.

Same class names, constructor dicts (`doe_params`, `optim_params`), parameter names
(`weight_height_map`, `weight_init_phase`, `init_phase`, `height_map`), attributes (`lut`,
`height_map`, `tolerance`, `epsilon`, `tand`, ...) and forward(field[, iter_frac]) -> ElectricField
surface (Components/QuantizedDOE.py:172, 295, 458, 862, 1043, 1218, 1390).  What changes is where the
work happens: level selection, the STE / soft-quantization gradients and the phase modulation run in
the sm_100a kernels behind include/thzdoe.h, and the modulation itself is *deferred*: forward returns
an ElectricField whose data is materialised lazily, so that an ASM_prop placed right after the DOE
(every notebook does that) fuses the multiply into its row-FFT prologue and the DOE adjoint into its
backward epilogue.

Out of scope: the RotationallySymmetric* parameterisations, `visualize`
and `save` (matplotlib / host-side file output); `num_unit` quadrant tiling is supported through the
same torch flip/cat expansion the reference uses (:27-35), which is O(H W) once per step.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from .. import _native as N
from .. import asm_host as AH
from .. import functional as Fn
from ..DataType.ElectricField import ElectricField
from ..utils.units import mm

BASE_PLANE_THICKNESS = 2 * mm
LIGHT_SPEED = 2.998e8


def _copy_quad_to_full(quad_map):
    """Components/QuantizedDOE.py:27-35."""
    if len(quad_map.shape) == 4:
        half = torch.cat([torch.flip(quad_map, dims=[2]), quad_map], dim=2)
        return torch.cat([torch.flip(half, dims=[3]), half], dim=3)
    half = torch.cat([torch.flip(quad_map, dims=[0]), quad_map], dim=0)
    return torch.cat([torch.flip(half, dims=[1]), half], dim=1)


def _phase_to_height_with_material_refractive_idx(_phase, _wavelength, _refractive_index):
    return _phase / (2 * torch.pi / _wavelength) / (_refractive_index - 1)


def _height_to_phase_with_material_refractive_idx(_height, _wavelength, _refractive_index):
    return 2 * torch.pi / _wavelength * (_refractive_index - 1) * _height


def _default_device(device):
    return torch.device("cuda" if torch.cuda.is_available() else "cpu") if device is None else device


class _DeferredModulation:
    """x * [mask * mul] * p(h), not yet evaluated.  ASM_prop consumes (x, height_map, coef, mask, mul) directly."""

    def __init__(self, x, height_map, coef, rows=None, mask=None, mul=None, reducer=None, levels=None):
        self.x, self.height_map, self.coef = x, height_map, coef
        self.levels = levels      # (int32 level map, [C,L] level transmissions) of a quantised map: the propagation's row kernels
                                  # then consume the quantiser's level indices directly (functional.level_phase_table)
        self.reducer = reducer    # parallel.FusedGradReduce: grad_height summed over the ranks inside the fused adjoint
        self.shape = x.shape
        self.device = x.device
        self.rows = rows          # (lo, hi): x holds only these rows of the grid the height map covers (slab-decomposed fields)
        self.mask, self.mul = mask, mul       # pointwise elements in front of the DOE (DataType.ElectricField.DeferredElements)

    def materialise(self):
        hm = self.height_map if self.rows is None else self.height_map[self.rows[0]:self.rows[1]]
        if self.reducer is not None:        # not fused after all (the field was read, or the propagator took another route):
            hm = _SumGradOverRanks.apply(hm, self.reducer.group)        # keep the promise that the gradient arrives summed
        x = self.x
        if self.mask is not None:
            x = Fn.FieldMulFn.apply(x, self.mask)
        if self.mul is not None:
            x = Fn.FieldMulFn.apply(x, self.mul)
        return Fn.DoeModulateFn.apply(x, hm, self.coef)


class _SumGradOverRanks(torch.autograd.Function):
    """Identity whose backward all-reduces (NCCL) the gradient: the unfused twin of parallel.FusedGradReduce."""

    @staticmethod
    def forward(ctx, h, group):
        ctx.group = group
        return h.view_as(h)

    @staticmethod
    def backward(ctx, g):
        import torch.distributed as dist
        g = g.contiguous().clone()
        dist.all_reduce(g, group=ctx.group)
        return g, None


class DOELayer(nn.Module):

    @staticmethod
    def phase_shift_according_to_height(height_map, wavelengths, epsilon, tand):
        """[C,H,W] complex transmission loss(h) * exp(-i phi(h))  (Components/QuantizedDOE.py:46-79),
        evaluated by the modulation kernel on a unit field."""
        wl = torch.as_tensor(wavelengths).reshape(-1)
        coef = AH.doe_coefficients(wl, epsilon, tand).to(height_map.device)
        ones = torch.ones(1, wl.numel(), height_map.shape[0], height_map.shape[1], dtype=torch.complex64,
                          device=height_map.device)
        return Fn.DoeModulateFn.apply(ones, height_map, coef)[0]

    @staticmethod
    def add_height_map_noise(height_map, tolerance=None):
        """Uniform +-tolerance fabrication noise, redrawn every forward (Components/QuantizedDOE.py:81-87)."""
        if tolerance is not None:
            height_map = height_map + (torch.rand_like(height_map) - 0.5) * 2 * tolerance
        return height_map

    def build_height_map(self):
        return NotImplemented

    def _coef(self, wavelengths, epsilon, tand, device):
        """Per-wavelength coefficient table on the device, cached per (tensor objects, in-place versions) so that a
        steady-state forward does no device->host read."""
        def tok(t):
            return (id(t), t._version) if torch.is_tensor(t) else ("py", t)
        fast = (tok(wavelengths), tok(epsilon), tok(tand), str(device))
        cache = self.__dict__.setdefault("_coef_cache", {})
        if cache.get("fast") == fast:
            return cache["coef"]
        key = (tuple(torch.as_tensor(wavelengths).detach().cpu().reshape(-1).tolist()), float(epsilon), float(tand), str(device))
        if cache.get("key") != key:
            cache["coef"] = AH.doe_coefficients(wavelengths, epsilon, tand).to(device)
            cache["key"] = key
        cache["fast"] = fast
        cache["refs"] = (wavelengths, epsilon, tand)
        return cache["coef"]

    def _level_phase(self, lut, coef):
        """[C,L] transmissions of the LUT levels for the wavelengths behind `coef`, cached per (coef, lut) object / version."""
        tok = (id(coef), id(lut), lut._version)
        cache = self.__dict__.setdefault("_lphase_cache", {})
        if cache.get("tok") != tok:
            cache["table"] = Fn.level_phase_table(lut.to(coef.device), coef)
            cache["tok"] = tok
            cache["refs"] = (coef, lut)
        return cache["table"]

    def modulate(self, input_field, preprocessed_height_map, height_tolerance, epsilon, tand, level_index=None, lut=None):
        """Components/QuantizedDOE.py:92-126, deferred (see module docstring).  level_index / lut: the map is quantised,
        preprocessed_height_map == lut[level_index] -- handed on so that the fused propagation can look transmissions up."""
        hm = self.add_height_map_noise(preprocessed_height_map, tolerance=height_tolerance)
        # a row slab of a grid distributed over several GPUs (parallel.shard_rows): the map covers the WHOLE grid
        slab = getattr(input_field, "_row_slab", None)
        full_h = input_field.height * slab[1] if slab is not None else input_field.height
        if full_h != hm.shape[0] or input_field.width != hm.shape[1]:
            hm = nn.functional.interpolate(hm[None, None, :, :], size=[full_h, input_field.width], mode='nearest')
        self._height_map_ = torch.squeeze(hm, (0, 1)) if hm.ndim == 4 else hm
        # an aperture / lens applied just before (still un-evaluated) rides along: x m p(h) in one fused prologue
        from ..DataType.ElectricField import DeferredElements
        pend = DeferredElements.pending(input_field) if slab is None else None
        mask = mul = None
        if pend is not None:
            x, mask, mul = pend.x, pend.mask, pend.mul
        else:
            x = input_field.data
        N.require_cuda(x, "field.data")
        coef = self._coef(input_field.wavelengths, epsilon, tand, x.device)
        rows = (slab[0] * input_field.height, (slab[0] + 1) * input_field.height) if slab is not None else None
        levels = None
        if (level_index is not None and lut is not None and height_tolerance is None and slab is None and
                tuple(level_index.shape) == tuple(self._height_map_.shape) and level_index.dtype == torch.int32):
            levels = (level_index.to(x.device).contiguous(), self._level_phase(lut, coef))
        deferred = _DeferredModulation(x, self._height_map_.to(x.device), coef, rows=rows, mask=mask, mul=mul,
                                       reducer=getattr(self, "grad_reducer", None) if slab is None else None, levels=levels)
        return ElectricField._from_deferred(deferred, input_field)


def _common_init(self, doe_params, device, default_level=6, with_level=True):
    self.device = _default_device(device)
    self.doe_size = doe_params.get('doe_size', None)
    self.doe_dxy = doe_params.get('doe_dxy', None)
    if with_level:
        self.doe_level = doe_params.get('doe_level', default_level)
    self.num_unit = doe_params.get('num_unit', None)
    height_constraint_max = doe_params.get('height_constraint_max', 2 * mm)
    self.height_constraint_max = torch.tensor(height_constraint_max, device=self.device)
    self._hmax = float(torch.tensor(height_constraint_max, dtype=torch.float32))
    self.tolerance = doe_params.get('tolerance', 0.05 * mm)
    material = doe_params.get('material', None)
    self.material = torch.tensor(material, device=self.device)
    self.epsilon = self.material[0]
    self.tand = self.material[1]


def _look_up_table(self, look_up_table):
    """Components/QuantizedDOE.py:1349-1366 (same in every quantized layer)."""
    if look_up_table is None:
        lut = torch.linspace(0, self.height_constraint_max, self.doe_level + 1).to(self.device)
        self.lut = lut[:-1]
    else:
        self.lut = torch.tensor(look_up_table, dtype=torch.float32).to(self.device)
        self.doe_level = len(self.lut)


def _cosine_tau(iter_frac, tau_min, tau_max):
    """Components/QuantizedDOE.py:869-871."""
    return tau_min + 0.5 * (tau_max - tau_min) * (1 + math.cos(iter_frac * math.pi))


def _expand(self, height_map):
    """num_unit quadrant expansion + squeeze, as every preprocessed_height_map ends."""
    if self.num_unit is None:
        return height_map.squeeze(0, 1).to(self.device) if height_map.ndim == 4 else height_map.to(self.device)
    hm = height_map if height_map.ndim == 4 else height_map[None, None]
    return _copy_quad_to_full(hm).squeeze(0, 1).to(self.device)


def _gumbel_noise(shape, device):
    """-log(Exp(1)) noise exactly as F.gumbel_softmax draws it (torch/nn/functional.py)."""
    return -torch.empty(shape, dtype=torch.float32, device=device).exponential_().log()


class FixDOEElement(DOELayer):
    """Components/QuantizedDOE.py:129-177."""

    def __init__(self, height_map, tolerance=0.1 * mm, material=None, device=None):
        super().__init__()
        self.device = _default_device(device)
        self.height_map = nn.parameter.Parameter(torch.as_tensor(height_map).clone().detach().to(self.device))
        self.tolerance = torch.tensor(tolerance, device=self.device) if tolerance is not None else None
        self.material = torch.tensor(material, device=self.device)
        self.epsilon = self.material[0]
        self.tand = self.material[1]

    def forward(self, field, iter_frac=None):
        return self.modulate(input_field=field, preprocessed_height_map=self.height_map, height_tolerance=self.tolerance,
                             epsilon=self.epsilon, tand=self.tand)


class FullPrecisionDOELayer(DOELayer):
    """Components/QuantizedDOE.py:181-301."""

    def __init__(self, doe_params, device=None):
        super().__init__()
        _common_init(self, doe_params, device, with_level=False)
        self.build_weight_height_map()

    def build_weight_height_map(self):
        height, width = self.doe_size[0], self.doe_size[1]
        if self.num_unit is None:
            shape = (1, 1, height, width)
        else:
            shape = (1, 1, int(height / self.num_unit), int(width / self.num_unit))
        self.weight_height_map = nn.parameter.Parameter(
            -torch.pi + 2 * torch.pi * torch.rand(*shape, device=self.device), requires_grad=True)

    def preprocessed_height_map(self):
        height_map = Fn.HeightFromWeightFn.apply(self.weight_height_map, self._hmax, 8.0)
        self.height_map = _expand(self, height_map)
        return self.height_map

    def forward(self, field, iter_frac=None):
        return self.modulate(input_field=field, preprocessed_height_map=self.preprocessed_height_map(),
                             height_tolerance=self.tolerance, epsilon=self.epsilon, tand=self.tand)


class STEQuantizedDOELayer(DOELayer):
    """Components/QuantizedDOE.py:1257-1396: sigmoid height -> nearest LUT level, straight-through gradient."""

    def __init__(self, doe_params, optim_params=None, device=None):
        super().__init__()
        _common_init(self, doe_params, device)
        self.build_weight_height_map()
        self.look_up_table(doe_params.get('look_up_table', None))

    look_up_table = _look_up_table

    def build_weight_height_map(self):
        height, width = self.doe_size[0], self.doe_size[1]
        if self.num_unit is None:
            shape = (1, 1, height, width)
        else:
            u = int(height / self.num_unit)
            shape = (1, 1, u, u)                                   # reference uses unit_size[0] twice (:1376)
        self.weight_height_map = nn.parameter.Parameter(torch.randn(*shape, device=self.device), requires_grad=True)

    def preprocessed_height_map(self):
        q, idx = Fn.SteFromWeightFn.apply(self.weight_height_map, self.lut, self._hmax, 8.0)
        self.level_index = _expand(self, idx)
        self.height_map = _expand(self, q)
        return self.height_map

    def forward(self, field, iter_frac=None):
        hm = self.preprocessed_height_map()
        return self.modulate(input_field=field, preprocessed_height_map=hm, height_tolerance=self.tolerance, epsilon=self.epsilon,
                             tand=self.tand, level_index=self.level_index, lut=self.lut)


class PSQuantizedDOELayer(DOELayer):
    """Components/QuantizedDOE.py:1068-1235: progressive sigmoid quantisation, tau rising linearly."""

    def __init__(self, doe_params, optim_params, device=None):
        super().__init__()
        _common_init(self, doe_params, device)
        self.tau_max = optim_params.get('tau_max', 400)
        self.tau_min = optim_params.get('tau_min', 1)
        self.build_weight_height_map()

    look_up_table = _look_up_table

    def build_weight_height_map(self):
        height, width = self.doe_size[0], self.doe_size[1]
        if self.num_unit is None:
            shape = (height, width)
        else:
            u = int(height / self.num_unit)
            shape = (u, u)
        self.weight_height_map = nn.parameter.Parameter(torch.randn(*shape, device=self.device), requires_grad=True)

    def preprocessed_height_map(self, tau):
        self.height_constraint_min = 0
        height_map = Fn.PsqFn.apply(self.weight_height_map, self._hmax, int(self.doe_level), float(tau))
        self.height_map = _expand(self, height_map)
        return self.height_map

    def forward(self, field, iter_frac=None):
        tau = None
        if iter_frac is not None:
            tau = self.tau_min + (self.tau_max - self.tau_min) * iter_frac          # :1219-1223
        return self.modulate(input_field=field, preprocessed_height_map=self.preprocessed_height_map(tau=tau),
                             height_tolerance=self.tolerance, epsilon=self.epsilon, tand=self.tand)


class _ScoreGumbelBase(DOELayer):
    """Shared plumbing of the three score-Gumbel generations."""

    def __init__(self, doe_params, optim_params, device=None):
        super().__init__()
        _common_init(self, doe_params, device)
        self.c_s = optim_params.get('c_s', 300)
        self.tau_max = optim_params.get('tau_max', 5.5)
        self.tau_min = optim_params.get('tau_min', 2.0)
        self.look_up_table(doe_params.get('look_up_table', None))
        self.build_init_phase()
        self.gumbel_noise = None      # tests may pin the noise: tensor [1,L,H,W]; None = draw every forward

    look_up_table = _look_up_table

    def _kfac(self, wavelengths):
        """2 pi / lambda_min * (sqrt(eps) - 1) in the reference's fp32 op order (:40-41)."""
        lam = torch.as_tensor(wavelengths).detach().to("cpu", torch.float32).min()
        n_idx = torch.sqrt(self.epsilon.detach().to("cpu", torch.float32))
        return float(2 * torch.pi / lam * (n_idx - 1))

    def _noise(self, w):
        L = len(self.lut)
        if self.gumbel_noise is not None:
            return self.gumbel_noise.to(w.device)
        return _gumbel_noise((1, L) + tuple(w.shape[-2:]), w.device)

    def forward(self, field, iter_frac=None):
        tau = None if iter_frac is None else _cosine_tau(iter_frac, self.tau_min, self.tau_max)
        hm = self.preprocessed_height_map(wavelengths=field.wavelengths, tau=tau, iter_frac=iter_frac)
        return self.modulate(input_field=field, preprocessed_height_map=hm, height_tolerance=self.tolerance,
                             epsilon=self.epsilon, tand=self.tand)


class SoftGumbelQuantizedDOELayer(_ScoreGumbelBase):
    """Components/QuantizedDOE.py:303-476 (v1): a free phase parameter scored against the LUT phases."""

    def build_init_phase(self):
        height, width = self.doe_size[0], self.doe_size[1]
        if self.num_unit is None:
            shape = (1, 1, height, width)
        else:
            shape = (1, 1, int(height / self.num_unit), int(width / self.num_unit))
        self.init_phase = nn.parameter.Parameter(
            -torch.pi + 2 * torch.pi * torch.rand(*shape, device=self.device), requires_grad=True)

    def preprocessed_height_map(self, wavelengths, tau, iter_frac=None):
        w = self.init_phase
        q, idx = Fn.GumbelV3Fn.apply(w, self.lut, self._noise(w), self._hmax, self._kfac(wavelengths), float(self.c_s),
                                     float(tau), float(self.tau_max), 1.0, True)
        self.level_index = _expand(self, idx)
        self.height_map = _expand(self, q)
        return self.height_map


class SoftGumbelQuantizedDOELayerv2(_ScoreGumbelBase):
    """Components/QuantizedDOE.py:478-658 (v2): full-precision height until iter_frac 0.5, then quantized."""
    switch = 0.5

    def build_init_phase(self):
        height, width = self.doe_size[0], self.doe_size[1]
        if self.num_unit is None:
            shape = (height, width)
        else:
            shape = (int(height / self.num_unit), int(width / self.num_unit))
        self.weight_init_phase = nn.parameter.Parameter(torch.randn(*shape, device=self.device), requires_grad=True)

    def _beta(self, iter_frac):
        return None if iter_frac <= self.switch else 1.0

    def preprocessed_height_map(self, wavelengths, tau, iter_frac=None):
        w = self.weight_init_phase
        beta = self._beta(iter_frac)
        if beta is None:
            height_map = Fn.HeightFromWeightFn.apply(w, self._hmax, 10.0)
            self.level_index = None
        else:
            height_map, idx = Fn.GumbelV3Fn.apply(w, self.lut, self._noise(w), self._hmax, self._kfac(wavelengths),
                                                  float(self.c_s), float(tau), float(self.tau_max), beta, False)
            self.level_index = _expand(self, idx)
        self.height_map = _expand(self, height_map)
        return self.height_map


class SoftGumbelQuantizedDOELayerv3(SoftGumbelQuantizedDOELayerv2):
    """Components/QuantizedDOE.py:660-890 (v3): full precision for iter_frac <= 0.3, a (1-beta) h + beta q
    blend with beta = (f - 0.3) / 0.5 up to 0.8, pure quantized afterwards (:826-849)."""

    def _beta(self, iter_frac):
        if iter_frac <= 0.3:
            return None
        if iter_frac <= 0.8:
            return (iter_frac - 0.3) / (0.8 - 0.3)
        return 1.0


class NaiveGumbelQuantizedDOELayer(DOELayer):
    """Components/QuantizedDOE.py:892-1065: per-pixel learnable logits [H,W,L], hard Gumbel-softmax."""

    def __init__(self, doe_params, optim_params, device=None):
        super().__init__()
        _common_init(self, doe_params, device)
        self.c_s = optim_params.get('c_s', 300)
        self.tau_max = optim_params.get('tau_max', 5.5)
        self.tau_min = optim_params.get('tau_min', 2.0)
        self.look_up_table(doe_params.get('look_up_table', None))
        self.build_init_logits()
        self.gumbel_noise = None

    look_up_table = _look_up_table

    def build_init_logits(self):
        height, width = self.doe_size[0], self.doe_size[1]
        if self.num_unit is None:
            shape = (height, width, self.doe_level)
        else:
            shape = (int(height / self.num_unit), int(width / self.num_unit), self.doe_level)
        self.weight_height_map = nn.parameter.Parameter(torch.rand(*shape, device=self.device), requires_grad=True)

    def preprocessed_height_map(self, tau):
        w = self.weight_height_map
        noise = self.gumbel_noise.to(w.device) if self.gumbel_noise is not None else _gumbel_noise(w.shape, w.device)
        q, idx = Fn.GumbelNaiveFn.apply(w, self.lut, noise, 1.0 if tau is None else float(tau))
        self.level_index = _expand(self, idx)
        self.height_map = _expand(self, q)
        return self.height_map

    def forward(self, field, iter_frac=None):
        tau = None if iter_frac is None else _cosine_tau(iter_frac, self.tau_min, self.tau_max)
        return self.modulate(input_field=field, preprocessed_height_map=self.preprocessed_height_map(tau=tau),
                             height_tolerance=self.tolerance, epsilon=self.epsilon, tand=self.tand)


class STEQuantizationFunction:
    """Callable twin of the reference autograd.Function (Components/QuantizedDOE.py:1239-1253):
    `STEQuantizationFunction.apply(input, lut)` -> quantized, identity backward."""

    @staticmethod
    def apply(input, lut):
        return Fn.SteQuantizeFn.apply(input, lut)[0]


ste_quan = STEQuantizationFunction.apply
