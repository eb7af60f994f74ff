"""Stand-alone quantizers -- mirror of the reference's Components/quantization.py with the level
selection and surrogate gradients running in the sm_100a kernels (thz_quant_nn_fwd / _bwd).

tau_iter (:12-21), nns / nns_poly / nns_sigmoid (:59-126), Quantization (:164-207).
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


def _nn(thickness, s, kind):
    if DiscreteDOE.lut is None:
        raise RuntimeError("DiscreteDOE.set_lut(lut) must be called before using the nearest-neighbour quantizers")
    q, _ = Fn.NnQuantizeFn.apply(thickness, DiscreteDOE.lut.to(thickness.device), DiscreteDOE.lut_midvals.to(thickness.device),
                                 float(s), kind)
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


class Quantization(nn.Module):
    """Method dispatch by name (:164-207): 'nn', 'nn_sigmoid', 'nn_poly'.  The '*softmax*' / '*gumbel*'
    thickness-space variants (:128-161) are served by the DOE layers' score-Gumbel kernels."""

    def __init__(self, method=None, max_thickness=None, num_bits=4, lut=None, dev=None, tau_min=0.5, tau_max=3.0, r=None, c=300.):
        super().__init__()
        dev = dev or torch.device("cuda" if torch.cuda.is_available() else "cpu")
        if lut is None:
            assert max_thickness is not None
            lut = torch.linspace(0, max_thickness, 2 ** num_bits + 1).to(dev)
        else:
            assert len(lut) == (2 ** num_bits) + 1
            lut = torch.tensor(lut, dtype=torch.float32).to(dev)
        DiscreteDOE.set_lut(lut)
        self.quan_fn = None
        m = method.lower()
        if m == 'nn':
            self.quan_fn = nns
        elif m == 'nn_sigmoid':
            self.quan_fn = nns_sigmoid
        elif m == 'nn_poly':
            self.quan_fn = nns_poly
        elif 'softmax' in m or 'gumbel' in m:
            raise NotImplementedError("thickness-space softmax quantization: use SoftGumbelQuantizedDOELayer*")
        self.method, self.tau_min, self.tau_max, self.r = method, tau_min, tau_max, r

    def forward(self, input_thickness, iter_frac=None, hard=True):
        tau = tau_iter(self.method, iter_frac, self.tau_min, self.tau_max, self.r) if iter_frac is not None else 1.0
        if self.quan_fn is None:
            return input_thickness
        return self.quan_fn(input_thickness, tau if tau is not None else 1.0).squeeze(0, 1)
