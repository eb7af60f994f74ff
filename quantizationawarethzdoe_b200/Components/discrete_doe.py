"""Holder of the discrete look-up table used by the nearest-neighbour quantizers
(reference: Components/discrete_doe.py:6-35).  The reference reads `DiscreteDOE.lut` /
`DiscreteDOE.lut_midvals` on the CLASS (Components/quantization.py:64-65), which only works when they
are plain class attributes; `set_lut` below assigns them that way."""
import torch

from ..utils.Helper_Functions import lut_mid


class DiscreteDOE:
    lut_midvals = None
    lut = None
    prev_idx = 0.

    @classmethod
    def set_lut(cls, new_lut):
        if new_lut is None:
            cls.lut = None
            cls.lut_midvals = None
            return
        lut = new_lut.clone().detach() if torch.is_tensor(new_lut) else torch.tensor(new_lut)
        cls.lut = lut.to(torch.float32)
        cls.lut_midvals = torch.tensor(lut_mid(cls.lut), dtype=torch.float32)
