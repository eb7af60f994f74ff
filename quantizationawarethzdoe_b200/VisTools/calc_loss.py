"""Loss landscape of a single-DOE system: losses on an xnum x ynum grid of weight perturbations
theta* + a delta + b eta (reference: VisTools/calc_loss.py:8-55, surface file :67-87, index order :90-98, perturbation :101-108).

The reference evaluates the grid one point at a time: overwrite the weights, run the model, normalise the intensity
by its maximum, nn.MSELoss, write the whole loss array back to the HDF5 file after EVERY point.  It is an embarrassingly
parallel caller of the propagation hot path, so here

  * grid points are evaluated `batch` at a time: the DOE layer turns each perturbed weight set into a height map (its
    own small kernels, incl. the per-forward tolerance noise / Gumbel noise the layer draws), the maps are stacked and
    ONE fused DOE -> propagation pass (thz_asm_propagate with per-entry height maps, doe_hmap_bstride) carries all of
    them over the shared input field; the normalised-intensity MSE of every entry comes from one more kernel pair
    (thz_normmse_loss_each);
  * with torch.distributed initialised the grid points are dealt round-robin over the ranks (no data-path collective)
    and the loss vector is summed once at the end;
  * the surface file is written once, at the end, with the reference's layout: datasets `xcoordinates`, `ycoordinates`,
    `loss`.  h5py is not part of this image: without it the same three arrays go to `3d_surface_file.npz` (np.load
    gives the same keys); with h5py installed the `.h5` file is written as the reference does.

A model is batched when it exposes its three parts -- `landscape_parts() -> (input_field, doe_layer, propagator)` or the
notebook's attribute names `input_field`, `doe`, `asm_prop3` -- and its forward is propagator(doe(input_field, iter_frac));
any other nn.Module is evaluated point by point through `model.forward(iter_frac=1)` exactly as the reference does.
"""
import os

import numpy as np
import torch
import torch.nn as nn

from .. import _native as N
from .. import functional as Fn

try:                                    # optional: absent in this image
    import h5py
except Exception:                       # pragma: no cover
    h5py = None


def setup_surface_file(args, save_path):
    """:67-87 -- coordinates + a loss array of -1 (= not evaluated yet)."""
    xcoordinates = np.linspace(args.xmin, args.xmax, args.xnum)
    ycoordinates = np.linspace(args.ymin, args.ymax, args.ynum)
    losses = -np.ones(shape=(len(xcoordinates), len(ycoordinates)))
    if h5py is not None:
        surface_path = f"{save_path}/3d_surface_file.h5"
        with h5py.File(surface_path, 'w') as f:
            f['xcoordinates'], f['ycoordinates'], f['loss'] = xcoordinates, ycoordinates, losses
    else:
        surface_path = f"{save_path}/3d_surface_file.npz"
        np.savez(surface_path, xcoordinates=xcoordinates, ycoordinates=ycoordinates, loss=losses)
    return surface_path


def read_surface_file(surface_path):
    """-> dict(xcoordinates, ycoordinates, loss) from either container."""
    if surface_path.endswith(".npz"):
        with np.load(surface_path) as z:
            return {k: z[k] for k in ("xcoordinates", "ycoordinates", "loss")}
    with h5py.File(surface_path, 'r') as f:
        return {k: f[k][:] for k in ("xcoordinates", "ycoordinates", "loss")}


def _write_losses(surface_path, losses):
    if surface_path.endswith(".npz"):
        d = read_surface_file(surface_path)
        np.savez(surface_path, xcoordinates=d["xcoordinates"], ycoordinates=d["ycoordinates"], loss=losses)
    else:
        with h5py.File(surface_path, 'r+') as f:
            f["loss"][:] = losses
            f.flush()


def get_indices(vals, xcoordinates, ycoordinates):
    """:90-98.  Note the reference's index convention, kept as is: `vals` has shape (xnum, ynum) but the coordinate of flat
    index i comes from np.meshgrid(x, y), which is (ynum, xnum) -- flat index i = iy * xnum + ix is evaluated at (x[ix], y[iy])
    and stored at vals.ravel()[i]."""
    inds = np.array(range(vals.size))
    inds = inds[vals.ravel() <= 0]
    xcoord_mesh, ycoord_mesh = np.meshgrid(xcoordinates, ycoordinates)
    return inds, np.c_[xcoord_mesh.ravel()[inds], ycoord_mesh.ravel()[inds]]


def overwrite_weights(model, init_weights, directions, step):
    """:101-108  p = theta* + a delta + b eta."""
    dx, dy = directions[0], directions[1]
    changes = [d0 * step[0] + d1 * step[1] for (d0, d1) in zip(dx, dy)]
    for (p, w, d) in zip(model.parameters(), init_weights, changes):
        p.data = w + d


def _parts(model):
    if hasattr(model, "landscape_parts"):
        return model.landscape_parts()
    if all(hasattr(model, a) for a in ("input_field", "doe", "asm_prop3")):
        return model.input_field, model.doe, model.asm_prop3
    return None


def _is_plain_mse(loss_f):
    return loss_f is None or (isinstance(loss_f, nn.MSELoss) and loss_f.reduction == "mean")


def _losses_of(out, target, loss_f):
    """out complex64 [Bc,C,H,W] -> float32 [Bc]: loss_f(|out|^2 / max |out|^2, target) per entry (:35-38)."""
    Bc = out.shape[0]
    if _is_plain_mse(loss_f):
        n = out[0].numel()
        t = target.to(device=out.device, dtype=torch.float32)
        if t.numel() != n:
            t = torch.broadcast_to(t, out.shape[1:]).contiguous()
        t = t.contiguous()
        losses = torch.empty(Bc, dtype=torch.float32, device=out.device)
        scratch = torch.empty(2 * Bc, dtype=torch.int32, device=out.device)
        N.check(N.lib().thz_normmse_loss_each(N.ptr(out), N.ptr(t), 1, Bc, n, N.ptr(scratch), N.ptr(losses),
                                              N.current_stream_ptr(out.device)), "thz_normmse_loss_each")
        return losses
    vals = []                           # a user-supplied loss: called per entry on the normalised intensity, like the reference
    for b in range(Bc):
        o = torch.abs(out[b:b + 1]) ** 2
        vals.append(loss_f(o / torch.max(o), target).reshape(()))
    return torch.stack(vals).float()


@torch.no_grad()
def calulate_single_element_loss_landscape(args, model, target, loss_f=nn.MSELoss(), directions=None, save_path=None, batch=64,
                                           group=None):
    """Mirror of VisTools/calc_loss.py:8-55 (same arguments, same returned surface path, same file contents); `batch` grid
    points per fused pass, `group` the torch.distributed group to shard over (default: the world, if initialised).
    The model's weights are restored to theta* afterwards (the reference leaves them at the last grid point)."""
    import torch.distributed as dist
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_available() and dist.is_initialized() else (0, 1)
    if rank == 0:
        surface_path = setup_surface_file(args, save_path)
    if world > 1:
        dist.barrier(group)
    if rank != 0:
        surface_path = f"{save_path}/3d_surface_file." + ("h5" if h5py is not None else "npz")
    init_weights = [p.data.clone() for p in model.parameters()]
    surf = read_surface_file(surface_path)
    losses = surf["loss"].copy()
    inds, coords = get_indices(losses, surf["xcoordinates"], surf["ycoordinates"])
    mine = list(range(rank, len(inds), world))
    dev = next(model.parameters()).device
    vec = torch.zeros(len(inds), dtype=torch.float32, device=dev)
    parts = _parts(model)
    if parts is not None:
        field, doe, prop = parts
        x = field.data.contiguous()                     # a fixed field in front of the DOE: evaluated once, reused by every grid point
        if x.shape[0] != 1 or not hasattr(prop, "_get_plan") or not hasattr(prop, "kernel_mode"):
            parts = None                # a batch of input fields / not an ASM propagator: evaluate point by point
    if parts is not None:
        coef = None
        for s in range(0, len(mine), batch):
            sel = mine[s:s + batch]
            maps = []
            for k in sel:
                overwrite_weights(model, init_weights, directions, coords[k])
                u = doe(field, iter_frac=1)                        # deferred modulation: level selection only, no field pass
                d = getattr(u, "_deferred", None)
                if d is None:
                    raise RuntimeError("the DOE layer did not return a deferred modulation; use the point-by-point path")
                maps.append(d.height_map.to(torch.float32))
                coef = d.coef
            hm = torch.stack(maps).contiguous()                  # [Bc, H, W]
            out = Fn.doe_asm_sweep(x, hm, prop, coef, field.spacing, field.wavelengths)
            vec[sel] = _losses_of(out, target, loss_f)
    else:
        for k in mine:
            overwrite_weights(model, init_weights, directions, coords[k])
            out = model.forward(iter_frac=1).data
            vec[k] = _losses_of(out if out.shape[0] == 1 else out.reshape((1, -1) + tuple(out.shape[-2:])), target, loss_f)[0]
    if world > 1:
        dist.all_reduce(vec, group=group)
    for p, w in zip(model.parameters(), init_weights):
        p.data = w
    losses.ravel()[inds] = vec.cpu().numpy().astype(losses.dtype)
    if rank == 0:
        _write_losses(surface_path, losses)
    if world > 1:
        dist.barrier(group)
    return surface_path


def calulate_DONN_loss_landscape(args, model, directions=None, save_path=None):
    """:58-64 -- a stub in the reference too ("Implement it if possible and needed")."""
    pass
