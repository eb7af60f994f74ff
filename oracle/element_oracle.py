"""CPU restatement of the reference's pointwise optical elements (TEST INFRASTRUCTURE ONLY): thin lens
(Components/Thin_Lens.py:33-85) and aperture (Components/Aperture.py:41-135).  Pinned by tests/golden/elem_*.npz from
the unmodified reference; only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file."""
import numpy as np
import torch


def lens_kernel(H, W, spacing, wavelengths, focal_length):
    dx, dy = spacing[0], spacing[1]
    lam = torch.as_tensor(wavelengths, dtype=torch.float32)[:, None, None]
    xc = torch.linspace(-((H - 1) // 2), (H - 1) // 2, H)
    yc = torch.linspace(-((W - 1) // 2), (W - 1) // 2, W)
    xg, yg = torch.meshgrid(xc, yc, indexing="ij")
    xg, yg = xg[None, None] * dx, yg[None, None] * dy
    ang = -(np.pi / (lam * torch.Tensor([focal_length]))) * ((xg ** 2) + (yg ** 2))
    return torch.exp(1j * ang)


def circ_mask(H, W, spacing, radius):
    dx, dy = spacing[0], spacing[1]
    x = torch.linspace(-dx * H / 2, dx * H / 2, H)
    y = torch.linspace(-dy * W / 2, dy * W / 2, W)
    X, Y = torch.meshgrid(x, y, indexing="ij")
    return torch.where(torch.sqrt(X ** 2 + Y ** 2) <= torch.tensor(radius), 1, 0)[None, None]


def rect_mask(H, W, spacing, side):
    dx, dy = spacing[0], spacing[1]
    rw, rh = min(side, dx * W), min(side, dy * H)
    x = torch.linspace(-dx * W / 2, dx * W / 2, W)
    y = torch.linspace(-dy * H / 2, dy * H / 2, H)
    X, Y = torch.meshgrid(x, y, indexing="xy")
    return torch.where((torch.abs(X) <= rw / 2) & (torch.abs(Y) <= rh / 2), 1, 0)[None, None]
