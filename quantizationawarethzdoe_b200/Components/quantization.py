"""Stand-alone quantizers -- mirror of the reference's Components/quantization.py with the level
selection and surrogate gradients running in the sm_100a kernels (thz_quant_nn_fwd / _bwd).

tau_iter (:12-21), score_thickness (:36-55), nns / nns_poly / nns_sigmoid (:59-126), SoftmaxBasedQuantization (:128-161),
Quantization (:164-207).
The reference module cannot run as written; this one keeps its maths and
call signatures and fixes the plumbing: the LUT lives in DiscreteDOE class attributes set through
DiscreteDOE.set_lut().
"""
import math

import torch
import torch.nn as nn

from .. import functional as Fn
from .discrete_doe import DiscreteDOE


def tau_iter(quan_fn, iter_frac, tau_min, tau_max, r=None):
    if 'softmax' in quan_fn:
        if r is None:
            r = math.log(tau_max / tau_min)
        tau = max(tau_min, tau_max * math.exp(-r * iter_frac))
    elif 'sigmoid' in quan_fn or 'poly' in quan_fn:
        tau = 1 + 10 * iter_frac
    else:
        tau = None
    return tau


def _nn(thickness, s, kind, lut=None, mid=None):
    if lut is None:
        lut, mid = DiscreteDOE.lut, DiscreteDOE.lut_midvals
    if lut is None:
        raise RuntimeError("DiscreteDOE.set_lut(lut) must be called before using the nearest-neighbour quantizers")
    q, _ = Fn.NnQuantizeFn.apply(thickness, lut.to(thickness.device), mid.to(thickness.device), float(s), kind)
    return q


def nns(thickness, s=1.0):
    """NearestNeighborSearch.apply: identity backward (:59-71)."""
    return _nn(thickness, s, 0)


def nns_poly(thickness, s=1.0):
    """NearestNeighborPolyGrad.apply (:73-96)."""
    return _nn(thickness, s, 1)


def nns_sigmoid(thickness, s=1.0):
    """NearestNeighborSigmoidGrad.apply (:98-122)."""
    return _nn(thickness, s, 2)


score_thickness = Fn.score_thickness      # (:36-55) all five scoring functions, one kernel


class SoftmaxBasedQuantization(nn.Module):
    """Components/quantization.py:128-161: score the raw thickness against every level, (Gumbel-)softmax over the levels,
    expectation (soft) or straight-through one-hot (hard) of the LUT.  One fused kernel (thz_quant_softmax_fwd) and one
    explicit backward instead of ~20 pointwise launches over [N,L,H,W] temporaries.

    forward(thickness [N,1,H,W], tau, hard) -> [N,1,H,W].  The reference normalises the scores by the maximum of
    |thickness - lut| over the WHOLE tensor (:41), so the batch is one quantization problem, as there.
    `gumbel_noise` ([1,L,H,W]; tests pin it) replaces the -log(Exp(1)) draw of F.gumbel_softmax when set."""

    def __init__(self, lut, gumbel=True, tau_max=3.0, c=300.):
        super().__init__()
        lut = torch.as_tensor(lut, dtype=torch.float32) if not torch.is_tensor(lut) else lut.detach().to(torch.float32)
        self.lut = lut.reshape(1, len(lut), 1, 1)
        self.c = c
        self.gumbel = gumbel
        self.tau_max = tau_max
        self.gumbel_noise = None
        self.level_index = None

    def forward(self, thickness, tau=1.0, hard=False):
        if thickness.dim() != 4 or thickness.shape[1] != 1:
            raise ValueError("thickness must be [N,1,H,W]")
        dev = thickness.device
        tau32 = torch.as_tensor(tau, dtype=torch.float32).detach().cpu()
        s = float(torch.tensor(self.tau_max, dtype=torch.float32) / tau32)       # fp32 quotient, as (self.tau_max / tau) ** 1 on a tensor tau
        L = self.lut.numel()
        noise = None
        if self.gumbel:
            noise = self.gumbel_noise.to(dev) if self.gumbel_noise is not None else \
                -torch.empty((thickness.shape[0], L) + tuple(thickness.shape[-2:]), dtype=torch.float32, device=dev).exponential_().log()
            if thickness.shape[0] != 1:        # [N,L,H,W] -> level-major [L, N*H*W] for the kernel
                noise = noise.permute(1, 0, 2, 3).contiguous()
        q, idx = Fn.SoftmaxQuantizeFn.apply(thickness, self.lut.reshape(-1).to(dev), noise, float(self.c), float(tau32), s, bool(hard))
        self.level_index = idx
        return q


class Quantization(nn.Module):
    """Method dispatch by name (:164-207): 'nn', 'nn_sigmoid', 'nn_poly', '*softmax*' ('*gumbel*' in the name adds Gumbel
    noise).  The LUT is held per instance (two Quantization objects with different LUTs do not clobber each other);
    DiscreteDOE.set_lut is still called once for code that reads the class attributes the reference's way."""

    def __init__(self, method=None, max_thickness=None, num_bits=4, lut=None, dev=None, tau_min=0.5, tau_max=3.0, r=None, c=300.):
        super().__init__()
        dev = dev or torch.device("cuda" if torch.cuda.is_available() else "cpu")
        if lut is None:
            assert max_thickness is not None
            lut = torch.linspace(0, max_thickness, 2 ** num_bits + 1).to(dev)
        else:
            assert len(lut) == (2 ** num_bits) + 1
            lut = torch.tensor(lut, dtype=torch.float32).to(dev)
        from ..utils.Helper_Functions import lut_mid
        self.lut = lut
        self.lut_midvals = torch.tensor(lut_mid(lut.detach().cpu()), dtype=torch.float32)
        if DiscreteDOE.lut is None:
            DiscreteDOE.set_lut(lut)
        self.quan_fn = None
        m = method.lower()
        self.gumbel = 'gumbel' in m
        self._nn_kind = {'nn': 0, 'nn_poly': 1, 'nn_sigmoid': 2}.get(m)
        if self._nn_kind is not None:
            self.quan_fn = lambda x, s=1.0: _nn(x, s, self._nn_kind, self.lut, self.lut_midvals)
        elif 'softmax' in m:
            self.quan_fn = SoftmaxBasedQuantization(lut[:-1], self.gumbel, tau_max=tau_max, c=c)
        self.method, self.tau_min, self.tau_max, self.r = method, tau_min, tau_max, r

    def forward(self, input_thickness, iter_frac=None, hard=True):
        # the reference leaves tau unbound when iter_frac is None (:196-201, a NameError); tau = 1 is used instead
        tau = tau_iter(self.method, iter_frac, self.tau_min, self.tau_max, self.r) if iter_frac is not None else 1.0
        if self.quan_fn is None:
            return input_thickness
        if self._nn_kind is not None:
            return self.quan_fn(input_thickness, tau if tau is not None else 1.0).squeeze(0, 1)
        return self.quan_fn(input_thickness, tau if tau is not None else 1.0, hard).squeeze(0, 1)
