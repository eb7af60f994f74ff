"""Loss and optimizer of the optimisation loop around the propagation, as fused sm_100a kernels.

The reference notebooks (experiment_four_focal_spots.ipynb cell 8 and its siblings) run

    out_amp = normalize(torch.abs(out_field.data) ** 2)        # utils/Helper_Functions.py:185-193
    loss = nn.MSELoss()(out_amp, target)
    optimizer.zero_grad(); loss.backward(); optimizer.step()   # torch.optim.Adam / AdamW, lr = 0.02

`normalized_intensity_mse(out_field.data, target)` replaces the first two lines (one pass for the per-batch maximum,
one for the loss and its gradient wrt the complex field) and `FusedAdam` the optimizer; neither keeps host-side state
that changes per iteration, so forward + backward + update can be captured once in a CUDA graph and replayed.
"""
import ctypes

import torch

from . import _native as N


class NormalizedIntensityMSEFn(torch.autograd.Function):
    """loss = mean((|y|^2 / max_b |y|^2 - target)^2); backward returns the stored d loss / d y."""

    @staticmethod
    def forward(ctx, y, target):
        N.require_cuda(y, "field")
        if y.dtype != torch.complex64 or y.dim() != 4:
            raise ValueError("field must be a complex64 [B,C,H,W] tensor")
        y = y.contiguous()
        B = y.shape[0]
        n_per_b = y[0].numel()
        t = torch.broadcast_to(target.to(device=y.device, dtype=torch.float32), y.shape).contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=y.device)
        need = ctx.needs_input_grad[0]
        gy = torch.empty_like(y) if need else None
        scratch = torch.empty(max(1, 3 * B), dtype=torch.int32, device=y.device)
        N.check(N.lib().thz_normmse_loss(N.ptr(y), N.ptr(t), B, n_per_b, N.ptr(scratch), N.ptr(loss), N.ptr(gy),
                                         N.current_stream_ptr(y.device)), "thz_normmse_loss")
        ctx.gy = gy
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        gy = ctx.gy
        ctx.gy = None
        return (gy * g if gy is not None else None), None


def normalized_intensity_mse(y, target):
    """nn.MSELoss()(normalize(torch.abs(y) ** 2), target) of the reference loop, fused (forward and gradient)."""
    return NormalizedIntensityMSEFn.apply(y, target)


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam (decoupled_weight_decay=False) / AdamW (True) for float32 CUDA parameters, one kernel per
    parameter tensor, step counter on the device (CUDA-graph capturable).  Same defaults as torch."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled_weight_decay=False):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                      decoupled_weight_decay=decoupled_weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                N.require_cuda(p, "parameter")
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise ValueError("FusedAdam needs contiguous float32 parameters")
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                    st["step"] = torch.zeros(1, dtype=torch.int32, device=p.device)
                g = p.grad.contiguous()
                N.check(N.lib().thz_adam_step(N.ptr(p), N.ptr(g), N.ptr(st["exp_avg"]), N.ptr(st["exp_avg_sq"]), N.ptr(st["step"]),
                                              p.numel(), group["lr"], b1, b2, group["eps"], group["weight_decay"],
                                              1 if group["decoupled_weight_decay"] else 0, 1, N.current_stream_ptr(p.device)),
                        "thz_adam_step")
        return loss
