"""CPU oracle for DOE phase modulation and quantized level selection.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Restates (torch CPU, fp32, reference op order):
  Components/QuantizedDOE.py:23,46-79   BASE_PLANE_THICKNESS, phase_shift_according_to_height
  Components/QuantizedDOE.py:81-126     add_height_map_noise, modulate
  Components/QuantizedDOE.py:276-283    FullPrecisionDOELayer.preprocessed_height_map
  Components/QuantizedDOE.py:1239-1253  STEQuantizationFunction
  Components/QuantizedDOE.py:1379-1388  STEQuantizedDOELayer.preprocessed_height_map
  Components/QuantizedDOE.py:1193-1216  PSQuantizedDOELayer.preprocessed_height_map (+ tau :1219-1223)
  Components/QuantizedDOE.py:794-860    SoftGumbelQuantizedDOELayerv3 score_phase / preprocessed_height_map
  Components/QuantizedDOE.py:1022-1031  NaiveGumbelQuantizedDOELayer.preprocessed_height_map
  Components/quantization.py:12-21,36-161 tau_iter, score_thickness, NearestNeighbor*, SoftmaxBasedQuantization
  utils/Helper_Functions.py:371-398     lut_mid, nearest_idx
Parity is PINNED by tests/test_oracle_golden.py (reference imported when present) and
tests/golden/*.npz (reference outputs; generator oracle/make_golden.py), plus the known-answer
values of Components/test_all.ipynb cells 13-21.
"""
import math

import torch
import torch.nn.functional as F

BASE_PLANE_THICKNESS = 2 * 1e-3  # QuantizedDOE.py:23 (2 * mm)


# ----------------------------------------------------------------------------- modulation
def phase_shift(height_map, wavelengths, epsilon, tand):
    """QuantizedDOE.py:46-79 -> [C,H,W] complex64 transmission loss(h)*exp(-i phi(h))."""
    h = height_map[None, :, :]
    lam = torch.as_tensor(wavelengths, dtype=torch.float32).view(-1)[:, None, None]
    k = 2 * torch.pi / lam
    epsilon = torch.as_tensor(epsilon, dtype=torch.float32)
    tand = torch.as_tensor(tand, dtype=torch.float32)
    base = torch.tensor(BASE_PLANE_THICKNESS)
    loss = torch.exp(-0.5 * k * (h + base) * tand * torch.sqrt(epsilon))
    delay = torch.exp(-1j * k * (h + base) * (torch.sqrt(epsilon) - 1))
    return loss * delay


def modulate(x, height_map, wavelengths, epsilon, tand, noise=None):
    """QuantizedDOE.py:92-126 on raw tensors.  `noise` is the already-drawn additive height noise
    ((rand-0.5)*2*tol, :81-87) or None; nearest upsampling when the map is smaller than the field (:102-107)."""
    if noise is not None:
        height_map = height_map + noise
    H, W = x.shape[-2], x.shape[-1]
    if height_map.shape[0] != H or height_map.shape[1] != W:
        height_map = F.interpolate(height_map[None, None], size=[H, W], mode="nearest")[0, 0]
    p = phase_shift(height_map, wavelengths, epsilon, tand)
    return x * p[None]


# ----------------------------------------------------------------------------- level selection
def ste_indices(h, lut):
    """QuantizedDOE.py:1243: argmin_j |h - lut_j| (first minimum wins)."""
    return torch.argmin(torch.abs(h.unsqueeze(-1) - lut), dim=-1)


class _STE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, lut):
        return lut[ste_indices(h, lut)]

    @staticmethod
    def backward(ctx, g):
        return g.clone(), None


def ste_quantize(h, lut):
    """QuantizedDOE.py:1239-1253 (identity backward)."""
    return _STE.apply(h, lut)


def sigmoid_height(w, hmax, clamp=8.0):
    """hmax * sigmoid(clamp(w, -c, c)) -- QuantizedDOE.py:277,1381 (c=8), :823 (c=10)."""
    return torch.as_tensor(hmax, dtype=torch.float32) * torch.sigmoid(torch.clamp(w, min=-clamp, max=clamp))


def linear_lut(hmax, levels):
    """QuantizedDOE.py:1352-1355: linspace(0, hmax, L+1)[:-1]."""
    return torch.linspace(0, torch.as_tensor(hmax, dtype=torch.float32), levels + 1)[:-1]


def lut_mid(lut):
    """Helper_Functions.py:373-374."""
    return [(a + b) / 2 for a, b in zip(lut[:-1], lut[1:])]


def nearest_idx(x, midvals):
    """Helper_Functions.py:390-398: bucketize(right=True) % len(midvals) (top bucket wraps to 0)."""
    return torch.bucketize(x.detach(), midvals, right=True) % len(midvals)


def nn_quantize_backward(x, q, idx, lut, s, g, kind):
    """quantization.py:70-71 ('nn'), :80-96 ('nn_poly'), :105-122 ('nn_sigmoid') backward rules.

    The neighbour level is lut[idx + sign(x - q)] evaluated after the wrap, python negative
    indices included, exactly as the reference indexes (:85, :110)."""
    if kind == "nn":
        return g
    dx = x - q
    d_idx = (dx / torch.abs(dx)).int().nan_to_num()
    other = lut[(idx + d_idx)]
    mid = (other + q) / 2
    gap = torch.abs(other - q) + 1e-20
    zz = (x - mid) / gap * 2
    if kind == "nn_poly":
        dout = (0.5 * s * (1 - abs(zz)) ** (s - 1)).nan_to_num()
        return g * (dout * 2.0)
    if kind == "nn_sigmoid":
        zz = zz * s
        dout = torch.sigmoid(zz) * (1 - torch.sigmoid(zz))
        return g * (dout * (4.0 * s))
    raise ValueError(kind)


def tau_schedule(quan_fn, iter_frac, tau_min, tau_max, r=None):
    """quantization.py:12-21."""
    if "softmax" in quan_fn:
        if r is None:
            r = math.log(tau_max / tau_min)
        return max(tau_min, tau_max * math.exp(-r * iter_frac))
    if "sigmoid" in quan_fn or "poly" in quan_fn:
        return 1 + 10 * iter_frac
    return None


def cosine_tau(iter_frac, tau_min, tau_max):
    """QuantizedDOE.py:869-871 (also :1049-1051)."""
    return tau_min + 0.5 * (tau_max - tau_min) * (1 + math.cos(iter_frac * math.pi))


def psq_height(w, hmax, levels, tau):
    """QuantizedDOE.py:1193-1207 progressive-sigmoid quantisation (differentiable)."""
    hmax = torch.as_tensor(hmax, dtype=torch.float32)
    h = hmax * torch.sigmoid(torch.clamp(w, min=-8.0, max=8.0))
    delta = (hmax - 0) / (levels - 1)
    xn = (h - 0) / delta - 0.5
    rng = torch.arange(levels - 1).unsqueeze(0).unsqueeze(2)
    return 0 + delta * torch.sum(torch.sigmoid(tau * (xn.unsqueeze(1) - rng)), dim=1)


def _wrap(a):
    return (a + torch.pi) % (2 * torch.pi) - torch.pi


def score_phase(phase, phase_lut, s):
    """QuantizedDOE.py:794-806 ('sigmoid' scoring)."""
    lut = _wrap(phase_lut[None, :, None, None])
    diff = _wrap(_wrap(phase) - lut)
    diff = diff / torch.pi
    zz = s * diff
    return torch.sigmoid(zz) * (1 - torch.sigmoid(zz)) * 4


def gumbel_hard(scores, tau, gumbel_noise, dim):
    """F.gumbel_softmax(hard=True) with the noise supplied: one_hot(argmax) - sg(soft) + soft."""
    y_soft = ((scores + gumbel_noise) / tau).softmax(dim)
    index = y_soft.max(dim, keepdim=True)[1]
    y_hard = torch.zeros_like(scores).scatter_(dim, index, 1.0)
    return y_hard - y_soft.detach() + y_soft, index


def score_gumbel_v3_height(w, lut, hmax, wavelength_min, epsilon, c_s, tau_min, tau_max, iter_frac, gumbel_noise):
    """SoftGumbelQuantizedDOELayerv3.preprocessed_height_map (QuantizedDOE.py:819-860), noise supplied.

    gumbel_noise: [1,L,H,W] (= -log(Exp(1) samples)).  Returns (height_map [H,W], idx [H,W] or None)."""
    hmax = torch.as_tensor(hmax, dtype=torch.float32)
    epsilon = torch.as_tensor(epsilon, dtype=torch.float32)
    tau = cosine_tau(iter_frac, tau_min, tau_max)
    h = (hmax * torch.sigmoid(torch.clamp(w, min=-10.0, max=10.0)))[None, None]
    idx = None
    if iter_frac > 0.3:
        n_idx = torch.sqrt(epsilon)
        lam = torch.as_tensor(wavelength_min, dtype=torch.float32)
        phase_lut = 2 * torch.pi / lam * (n_idx - 1) * lut             # :40-41
        phase = 2 * torch.pi / lam * (n_idx - 1) * h
        scores = score_phase(phase, phase_lut, (tau_max / tau) ** 1) * c_s * (tau_max / tau) ** 1
        one_hot, idx = gumbel_hard(scores, tau, gumbel_noise, dim=1)
        q = (lut.reshape(1, len(lut), 1, 1) * one_hot).sum(1, keepdim=True)
        if iter_frac <= 0.8:
            beta = (iter_frac - 0.3) / (0.8 - 0.3)
            h = (1 - beta) * h + beta * q
        else:
            h = q
        idx = idx[0, 0]
    return h[0, 0], idx


def naive_gumbel_height(logits, lut, tau, gumbel_noise):
    """NaiveGumbelQuantizedDOELayer.preprocessed_height_map (QuantizedDOE.py:1022-1031); logits [H,W,L]."""
    one_hot, idx = gumbel_hard(logits, 1 if tau is None else tau, gumbel_noise, dim=-1)
    return (lut[None, None, :] * one_hot).sum(dim=-1), idx[..., 0]


def score_thickness(thickness, lut, s):
    """quantization.py:36-46 ('sigmoid' scoring), lut shaped [1,L,1,1]."""
    diff = thickness - lut
    diff = diff / torch.max(torch.abs(diff))
    zz = s * diff
    return torch.sigmoid(zz) * (1 - torch.sigmoid(zz)) * 4


def softmax_quantize(thickness, lut, tau, tau_max, c, gumbel_noise=None, hard=True):
    """quantization.py:128-161 SoftmaxBasedQuantization.forward; thickness [N,1,H,W]."""
    lut4 = lut.reshape(1, -1, 1, 1)
    scores = score_thickness(thickness, lut4, (tau_max / tau) ** 1) * c * (tau_max / tau) ** 1.0
    if gumbel_noise is not None:
        if hard:
            one_hot, _ = gumbel_hard(scores, tau, gumbel_noise, dim=1)
        else:
            one_hot = ((scores + gumbel_noise) / tau).softmax(1)
    else:
        y_soft = F.softmax(scores / tau, dim=1)
        index = y_soft.max(1, keepdim=True)[1]
        y_hard = torch.zeros_like(scores).scatter_(1, index, 1.0)
        one_hot = y_hard + y_soft - y_soft.detach() if hard else y_soft
    return (one_hot * lut4).sum(1, keepdims=True)


# ----------------------------------------------------------------------------- closed-form grads
def modulate_grads(x, height_map, wavelengths, epsilon, tand, g):
    """Closed-form gradients the CUDA backward implements (checked against autograd in tests):
        gx = g * conj(p);  gh = sum_{b,c} Re( conj(g) * x * p * gamma_c ),
        gamma_c = -k_c (0.5 tand sqrt(eps) + i (sqrt(eps) - 1))."""
    p = phase_shift(height_map, wavelengths, epsilon, tand)
    lam = torch.as_tensor(wavelengths, dtype=torch.float32).view(-1)[:, None, None]
    k = 2 * torch.pi / lam
    se = torch.sqrt(torch.as_tensor(epsilon, dtype=torch.float32))
    gamma = -k * (0.5 * torch.as_tensor(tand, dtype=torch.float32) * se + 1j * (se - 1))
    gx = g * torch.conj(p)[None]
    gh = (torch.conj(g) * x * (p * gamma)[None]).real.sum(dim=(0, 1))
    return gx, gh
