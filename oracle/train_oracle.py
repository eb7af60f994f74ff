"""CPU restatement of the loss / optimizer lines of the reference's optimisation loop (TEST INFRASTRUCTURE ONLY).

    out_amp = normalize(torch.abs(out_field.data) ** 2)      utils/Helper_Functions.py:185-193
    loss = nn.MSELoss()(out_amp, target)                      experiment_four_focal_spots.ipynb cell 8
    torch.optim.Adam / AdamW (lr = 0.02)                      same cell

Pinned by tests/golden/train_*.npz, which oracle/make_golden.py produces with the reference's own `normalize`
and torch's own optimizers; only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.
"""
import torch


def normalize(x):
    """utils/Helper_Functions.py:185-193: divide every batch entry by its maximum (in place in the reference)."""
    b = x.shape[0]
    flat = x.reshape(b, -1)
    return (flat / flat.max(1, keepdim=True)[0]).reshape(x.shape)


def normalized_intensity_mse(y, target):
    """loss and d loss / d y by torch autograd on the CPU, fp32."""
    y = y.detach().clone().requires_grad_(True)
    loss = torch.nn.functional.mse_loss(normalize(torch.abs(y) ** 2), target.expand(y.shape))
    (gy,) = torch.autograd.grad(loss, y)
    return loss.detach(), gy


def adam_reference(p0, grads, lr=0.02, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False):
    """Apply torch.optim.Adam / AdamW for len(grads) steps with the given gradients; returns the parameter history."""
    p = torch.nn.Parameter(p0.detach().clone())
    cls = torch.optim.AdamW if decoupled else torch.optim.Adam
    opt = cls([p], lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
    hist = []
    for g in grads:
        p.grad = g.clone()
        opt.step()
        hist.append(p.detach().clone())
    return hist
