"""Multi-GPU execution of the hot path: one process per GPU, torch.distributed (NCCL) for the plumbing.

Two patterns:

1. Data parallel over independent units (wavelengths, depths, DONN batch entries).  Fields of different
   (batch, wavelength, z) never interact; only the DOE parameters are shared.  `shard_field` gives each
   rank a contiguous chunk, every rank runs the ordinary single-GPU modules, and `allreduce_gradients`
   sums the DOE parameter gradients once per step (one flat fp32 bucket: 16 MB at 2048^2).

2. Slab-decomposed propagation of ONE grid too large for a GPU (config 5, 16384^2).  Rank g owns rows
   [g H/G, (g+1) H/G) of the field.  `SlabAsm` runs the three kernels of the fused pipeline separately
   (thz_asm_desc.stages) around two all-to-all transposes:
        row FFT (local rows) -> all-to-all -> column FFT . H . column iFFT (local columns)
                             -> all-to-all -> row iFFT + crop (local rows)
   Only live rows travel: zero-pad pruning halves the first exchange and crop pruning the second.
   Backward is the same schedule with conj(H) and the pad / crop regions swapped.
"""
import torch
import torch.distributed as dist

from . import _native as N
from . import asm_host as AH
from . import functional as Fn
from .DataType.ElectricField import ElectricField


# ----------------------------------------------------------------------------- data parallel
def shard_range(n_units, rank, world):
    """Contiguous, balanced chunk [lo, hi) of n_units for `rank` (first n % world ranks get one extra)."""
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_field(field, rank, world, axis="wavelength"):
    """The slice of an ElectricField this rank works on: along wavelengths (C) or batch (B)."""
    data = field.data
    if axis == "wavelength":
        lo, hi = shard_range(data.shape[1], rank, world)
        return ElectricField(data[:, lo:hi].contiguous(), wavelengths=field.wavelengths[lo:hi], spacing=field.spacing,
                             device=data.device)
    if axis == "batch":
        lo, hi = shard_range(data.shape[0], rank, world)
        return ElectricField(data[lo:hi].contiguous(), wavelengths=field.wavelengths, spacing=field.spacing, device=data.device)
    raise ValueError("axis must be 'wavelength' or 'batch'")


def allreduce_gradients(params, group=None, average=False):
    """Sum (or average) .grad of the given parameters over the process group with ONE all-reduce of a
    flat fp32 bucket.  Parameters without a gradient contribute zeros (ranks must agree on the list)."""
    params = [p for p in params if p.requires_grad]
    if not params or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(torch.float32) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].reshape(p.shape).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n


# ----------------------------------------------------------------------------- slab-decomposed ASM
def _all_to_all(send, group):
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    return recv


class _SlabPlan:
    def __init__(self, G, rank, C, H, W, pad_h, pad_w, Hp, Wp, unpad, device, rowvec, colvec, scal, table, tf_mode):
        if H % G or Wp % G:
            raise ValueError("slab FFT: rows (%d) and padded width (%d) must be divisible by the world size %d" % (H, Wp, G))
        self.G, self.rank, self.C = G, rank, C
        self.H, self.W, self.pad_h, self.pad_w, self.Hp, self.Wp = H, W, pad_h, pad_w, Hp, Wp
        self.outH, self.outW, self.out_r0, self.out_c0 = (H, W, pad_h, pad_w) if unpad else (Hp, Wp, 0, 0)
        if self.outH % G:
            raise ValueError("slab FFT: output rows (%d) must be divisible by the world size %d" % (self.outH, G))
        self.Wc = Wp // G
        c0 = rank * self.Wc
        self.rowvec = rowvec.to(device) if rowvec is not None else None
        self.colvec = colvec[:, c0:c0 + self.Wc].contiguous().to(device) if colvec is not None else None
        self.scal = scal.to(device) if scal is not None else None
        self.table = table[:, :, c0:c0 + self.Wc].contiguous().to(device) if table is not None else None
        self.tf_mode = tf_mode
        self.tw_h, self.tw_w, self.tw_c = N.twiddles(Hp, device), N.twiddles(Wp, device), N.twiddles(self.Wc, device)


def _slab_run(x_local, p, conj, group):
    """One slab-decomposed propagation.  x_local [B,C,rows_local,cols] -> y_local [B,C,out_rows_local,out_cols]."""
    G, B, C = p.G, x_local.shape[0], p.C
    nbc = B * C
    dev = x_local.device
    if not conj:
        inH, inW, in_r0, in_c0, outH, outW, out_r0, out_c0 = p.H, p.W, p.pad_h, p.pad_w, p.outH, p.outW, p.out_r0, p.out_c0
    else:
        inH, inW, in_r0, in_c0, outH, outW, out_r0, out_c0 = p.outH, p.outW, p.out_r0, p.out_c0, p.H, p.W, p.pad_h, p.pad_w
    Hl, Ol, Wc, Wp = inH // G, outH // G, p.Wc, p.Wp
    assert x_local.shape[2] == Hl and x_local.shape[3] == inW, "local slab has the wrong shape"
    common = dict(B=B, C=C, Hp=p.Hp, tf_mode=p.tf_mode, tf_conj=1 if conj else 0, rowvec=p.rowvec, scal=p.scal,
                  doe_mode=0, doe_base=0.0, hmap=None, coef=None, xsaved=None, gh=None, tw_h=p.tw_h)
    # ---- stage 1: row FFT of the local rows -> t1 [nbc, Hl, Wp]
    t1 = torch.empty(nbc * Hl * Wp, dtype=torch.complex64, device=dev)
    Fn._asm_call(AH.build_desc(x=x_local, y=None, inH=Hl, inW=inW, Wp=Wp, in_r0=0, in_c0=in_c0, outH=Hl, outW=outW, out_r0=0,
                               out_c0=out_c0, colvec=p.colvec, table=p.table, tw_w=p.tw_w, ws=t1, stages=1, **common), dev)
    # ---- transpose 1: rank j receives my rows of ITS column block
    send = t1.view(nbc, Hl, G, Wc).permute(2, 0, 1, 3).contiguous()               # [G, nbc, Hl, Wc]
    recv = _all_to_all(send, group)                                               # recv[i] = rows of rank i, my columns
    rowsT = max(inH, outH)
    t2 = torch.zeros(nbc, rowsT, Wc, dtype=torch.complex64, device=dev) if rowsT > inH else torch.empty(nbc, inH, Wc, dtype=torch.complex64, device=dev)
    t2[:, :inH] = recv.permute(1, 0, 2, 3).reshape(nbc, inH, Wc)
    # ---- stage 2: column FFT . H . column iFFT on the local columns (in place)
    Fn._asm_call(AH.build_desc(x=None, y=None, inH=inH, inW=min(inW, Wc), Wp=Wc, in_r0=in_r0, in_c0=0, outH=outH, outW=min(outW, Wc),
                               out_r0=out_r0, out_c0=0, colvec=p.colvec, table=p.table, tw_w=p.tw_c, ws=t2, stages=2, **common), dev)
    # ---- transpose 2: rank j receives its output rows of my column block
    send = t2[:, :outH].reshape(nbc, G, Ol, Wc).permute(1, 0, 2, 3).contiguous()  # [G, nbc, Ol, Wc]
    recv = _all_to_all(send, group)                                               # recv[i] = my rows, columns of rank i
    t3 = recv.permute(1, 2, 0, 3).reshape(nbc * Ol * Wp).contiguous()             # [nbc, Ol, Wp]
    # ---- stage 3: row iFFT + crop of the local output rows
    y = torch.empty(B, C, Ol, outW, dtype=torch.complex64, device=dev)
    Fn._asm_call(AH.build_desc(x=None, y=y, inH=Ol, inW=inW, Wp=Wp, in_r0=0, in_c0=in_c0, outH=Ol, outW=outW, out_r0=0,
                               out_c0=out_c0, colvec=p.colvec, table=p.table, tw_w=p.tw_w, ws=t3, stages=4, **common), dev)
    return y


class _SlabFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_local, plan, group):
        ctx.plan, ctx.group = plan, group
        return _slab_run(Fn._c64(x_local, "field.data"), plan, False, group)

    @staticmethod
    def backward(ctx, g):
        return _slab_run(Fn._c64(g, "grad_output"), ctx.plan, True, ctx.group), None, None


class SlabAsm(torch.nn.Module):
    """Band-limited ASM of a field whose ROWS are distributed over the ranks of a process group.

    forward(local_field) takes this rank's row slab as an ElectricField [B, C, H/G, W] and returns this rank's
    slab of the propagated field.  Constructor arguments as ASM_prop (Props/ASM_Prop.py:19-27)."""

    def __init__(self, z_distance=0.0, do_padding=True, do_unpad_after_pad=True, padding_scale=None, bandlimit_kernel=True,
                 bandlimit_type="exact", group=None, kernel_mode="inregister"):
        super().__init__()
        self.z = torch.as_tensor(z_distance, dtype=torch.float32)
        self.do_padding, self.do_unpad_after_pad = do_padding, do_unpad_after_pad
        self.padding_scale = AH.normalise_padding_scale(padding_scale, do_padding)
        self.bandlimit_kernel, self.bandlimit_type = bandlimit_kernel, bandlimit_type
        self.group, self.kernel_mode = group, kernel_mode
        self._key, self._plan = None, None

    def forward(self, field):
        G, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        data = field.data
        B, C, Hl, W = data.shape
        H = Hl * G
        key = (C, H, W, tuple(field.spacing.detach().cpu().tolist()), tuple(field.wavelengths.detach().cpu().tolist()), float(self.z), G, rank)
        if key != self._key:
            pad_h, pad_w, Hp, Wp = AH.compute_padding(H, W, self.padding_scale, self.do_padding)
            rowvec, colvec, scal = AH.tf_vectors(Hp, Wp, field.spacing, field.wavelengths, self.z, self.bandlimit_kernel, self.bandlimit_type)
            table, mode = None, 0
            dv = AH.tf_device_vectors(rowvec, colvec, scal) if self.kernel_mode == "inregister" else None
            if dv is not None:
                rowvec, colvec, scal = dv
            else:
                Hc = AH.tf_centred_reference_order(Hp, Wp, field.spacing, field.wavelengths, self.z, self.bandlimit_kernel, self.bandlimit_type)
                table, mode = AH.tf_table_slot_order(Hc), 1
            self._plan = _SlabPlan(G, rank, C, H, W, pad_h, pad_w, Hp, Wp, bool(self.do_padding and self.do_unpad_after_pad),
                                   data.device, rowvec, colvec, scal, table, mode)
            self._key = key
        out = _SlabFn.apply(data, self._plan, self.group)
        return ElectricField(out, wavelengths=field.wavelengths, spacing=field.spacing, device=data.device)
