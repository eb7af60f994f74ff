"""Hot-path helpers of the reference's utils/Helper_Functions.py, backed by the sm_100a kernels.

ft2 / ift2 (:99-119, :121-160): centred ortho-normalised 2-D transforms
    ft2(x)  = fftshift(fft2(fftshift(x), norm))     ift2(X) = ifftshift(ifft2(ifftshift(X), norm))
computed with the stand-alone FFT kernel (thz_fft2_c2c); the shifts are index rolls on the device
(torch.roll is plumbing, not arithmetic).  ASM_prop does NOT go through these (its shifts cancel).
lut_mid / nearest_idx / nearest_neighbor_search (:371-398): LUT helpers for the NN quantizers.
"""
import torch

from .. import functional as Fn


def perform_ft(input, delta=1, norm='ortho', pad=False, flag_ifft=False):
    if pad:
        raise NotImplementedError("perform_ft(pad=True) (AdaptiveAvgPool2d path, Helper_Functions.py:130-158) is not on the hot path")
    if norm not in ('ortho', 'backward', None):
        raise ValueError("norm must be 'ortho' or 'backward'")
    shift = torch.fft.ifftshift if flag_ifft else torch.fft.fftshift
    x = shift(input, dim=(-2, -1))
    y = Fn.fft2_c2c(x, inverse=flag_ifft, ortho=(norm == 'ortho'))
    return (delta ** 2) * shift(y, dim=(-2, -1))


def ft2(input, delta=1, norm='ortho', pad=False):
    return perform_ft(input=input, delta=delta, norm=norm, pad=pad, flag_ifft=False)


def ift2(input, delta=1, norm='ortho', pad=False):
    return perform_ft(input=input, delta=delta, norm=norm, pad=pad, flag_ifft=True)


def lut_mid(lut):
    return [(a + b) / 2 for a, b in zip(lut[:-1], lut[1:])]


def nearest_idx(input_val, lut_midvals):
    """bucketize(x, mid, right=True) % len(mid) on the device (Helper_Functions.py:390-398)."""
    mid = torch.as_tensor(lut_midvals, dtype=torch.float32).to(input_val.device)
    # lut values are irrelevant for the index; pass the midvals padded by one as a dummy lut
    dummy = torch.cat([mid, mid[-1:]])
    _, idx = Fn.NnQuantizeFn.apply(input_val.detach(), dummy, mid, 1.0, 0)
    return idx.to(torch.int64)


def nearest_neighbor_search(input_val, lut, lut_midvals=None):
    if lut_midvals is None:
        lut_midvals = torch.tensor(lut_mid(lut), dtype=torch.float32)
    idx = nearest_idx(input_val, lut_midvals)
    return lut.to(input_val.device)[idx], idx
