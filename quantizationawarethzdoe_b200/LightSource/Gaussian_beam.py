"""Gaussian-beam source -- mirror of the reference's LightSource/Gaussian_beam.py:Guassian_beam (:10-160).

Not on the hot path (the notebooks evaluate it once, before the optimisation loop): the field is formed with the
reference's own torch expressions on the module's device; it is here so that the notebooks' set-ups
(source -> ASM -> lens -> aperture -> DOE -> ASM, experiment_four_focal_spots.ipynb cell 6) run from this package alone.
"""
import torch
import torch.nn as nn

from ..DataType.ElectricField import ElectricField

LIGHT_SPEED = 2.998e8


class Guassian_beam(nn.Module):

    def __init__(self, height, width, beam_waist_x, beam_waist_y, center=(0, 0), z_w0=(0, 0), alpha=0, wavelengths=None,
                 spacing=None, device=None):
        super().__init__()
        self.device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.height = height
        self.width = width
        self.field = ElectricField(data=None, wavelengths=wavelengths, spacing=spacing, device=self.device)
        if beam_waist_x is None and beam_waist_y is None:
            freqs = LIGHT_SPEED / self.field.wavelengths
            self.beam_waist_x, self.beam_waist_y = self.BeamWaistCorruagtedTK(freqs)
        else:
            self.beam_waist_x = torch.tensor([beam_waist_x], device=self.device)
            self.beam_waist_y = torch.tensor([beam_waist_y], device=self.device)
        self.x0, self.y0 = torch.tensor(center, device=self.device)
        self.z_w0x, self.z_w0y = torch.tensor(z_w0, device=self.device)
        self.alpha = torch.tensor(alpha, device=self.device)

    def BeamWaistCorruagtedTK(self, freqs):
        """Waist fits of the corrugated horn, 220-330 GHz (LightSource/Gaussian_beam.py:67-85)."""
        freqs = freqs / 1e9
        p_E = [2.70171433587848e-13, 3.10350492358753e-10, -6.35088689290759e-07, 0.000322826804965868, -0.0665921902050336, 6.08799187520401]
        p_H = [-1.01507121315420e-11, 1.70791445624058e-08, -1.12281052414283e-05, 0.00360605624858374, -0.564799749943028, 35.5588926870041]
        wx = 1e-3 * (p_E[0] * freqs ** 5 + p_E[1] * freqs ** 4 + p_E[2] * freqs ** 3 + p_E[3] * freqs ** 2 + p_E[4] * freqs + p_E[5])
        wy = 1e-3 * (p_H[0] * freqs ** 5 + p_H[1] * freqs ** 4 + p_H[2] * freqs ** 3 + p_H[3] * freqs ** 2 + p_H[4] * freqs + p_H[5])
        return wx.clone().detach().to(self.device), wy.clone().detach().to(self.device)

    def forward(self):
        # Evaluated on the host (torch's CUDA linspace / exp differ from the CPU ones by ~1e-5 on this field; the parity
        # oracle is the reference's CPU path) and uploaded: this runs once, before the optimisation loop.
        cpu = torch.device("cpu")
        dx, dy = self.field.spacing[0].to(cpu), self.field.spacing[1].to(cpu)
        x = torch.linspace(-dx * self.height / 2, dx * self.height / 2, self.height, device=cpu)
        y = torch.linspace(-dy * self.width / 2, dy * self.width / 2, self.width, device=cpu)
        X, Y = torch.meshgrid(x, y, indexing="ij")
        X, Y = X.unsqueeze(0), Y.unsqueeze(0)
        wavelengths = self.field.wavelengths.to(cpu)[:, None, None]
        k = 2 * torch.pi / wavelengths
        wx, wy = self.beam_waist_x.to(cpu).reshape(-1), self.beam_waist_y.to(cpu).reshape(-1)
        z_w0x, z_w0y, alpha, x0, y0 = (t.to(cpu) for t in (self.z_w0x, self.z_w0y, self.alpha, self.x0, self.y0))
        if len(wx) != len(wavelengths) or len(wy) != len(wavelengths):
            if len(wx) == 1 and len(wy) == 1:
                wx, wy = wx.repeat(len(wavelengths)), wy.repeat(len(wavelengths))
            else:
                raise ValueError('Mismatch between beam waist and wavelength parameters')
        wx, wy = wx[:, None, None], wy[:, None, None]
        Rayleigh_x = torch.pi * wx ** 2 / wavelengths
        Rayleigh_y = torch.pi * wy ** 2 / wavelengths
        Gouy_phase_x = torch.arctan2(z_w0x, Rayleigh_x)
        Gouy_phase_y = torch.arctan2(z_w0y, Rayleigh_y)
        w_x = wx * torch.sqrt(1 + (z_w0x / Rayleigh_x) ** 2)
        w_y = wy * torch.sqrt(1 + (z_w0y / Rayleigh_y) ** 2)
        R_x = 1e12 if z_w0x == 0 else z_w0x * (1 + (Rayleigh_x / z_w0x) ** 2)
        R_y = 1e12 if z_w0y == 0 else z_w0y * (1 + (Rayleigh_y / z_w0y) ** 2)
        x_rot = X * torch.cos(alpha) + Y * torch.sin(alpha)
        y_rot = -X * torch.sin(alpha) + Y * torch.cos(alpha)
        phase = torch.exp(-1j * ((k * z_w0x + k * X ** 2 / (2 * R_x) - Gouy_phase_x) + (k * z_w0y + k * Y ** 2 / (2 * R_y) - Gouy_phase_y)))
        A = (wx / w_x) * (wy / w_y) * torch.exp(-(x_rot - x0) ** 2 / (w_x ** 2) - (y_rot - y0) ** 2 / (w_y ** 2))
        self.field.data = (A * phase).unsqueeze(0).to(self.device)
        return self.field
