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
import os

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


def shard_rows(field, rank, world):
    """This rank's row slab of a field whose grid is distributed over `world` GPUs (what SlabAsm consumes).  The returned
    field remembers that it is a slab, so a DOE layer applied to it keeps its height map at the size of the WHOLE grid and
    SlabAsm fuses the modulation (and its adjoint) into the slab pipeline."""
    data = field.data
    H = data.shape[2]
    if H % world:
        raise ValueError("rows (%d) must be divisible by the number of ranks (%d)" % (H, world))
    lo, hi = shard_range(H, rank, world)
    f = ElectricField(data[:, :, lo:hi].contiguous(), wavelengths=field.wavelengths, spacing=field.spacing, device=data.device)
    f._row_slab = (rank, world)
    return f


def allreduce_gradients(params, group=None, average=False):
    """Sum (or average) .grad of the given parameters over the process group with ONE all-reduce of a
    flat fp32 bucket.  Parameters without a gradient contribute zeros (ranks must agree on the list)."""
    params = [p for p in params if p.requires_grad and not getattr(p, "_thz_grad_is_reduced", False)]
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


class FusedGradReduce:
    """The data-parallel sum of a DOE's grad_height formed INSIDE the adjoint's last kernel instead of by an all-reduce after it.

    grad_height [H,W] float32 lives in symmetric memory with an NVLS multicast mapping; the row-iFFT epilogue of every rank adds
    its partial sums to the multicast address (multimem.red, thz_asm_desc.doe_gh_mode = 1), the NVSwitch applies each add to
    all replicas, and when the kernels of all ranks have finished every rank holds the complete sum -- the transfer rides under
    the kernel, what remains of the collective is ONE symmetric-memory barrier (~10 us against ~60-100 us for the 16 MiB NCCL
    all-reduce at 2048^2).  Two buffers alternate: the idle one is zeroed before the barrier of the current step, so a rank
    that runs ahead into the next step finds every replica clean.

        red = parallel.fuse_gradient_allreduce(doe)        # once; afterwards doe's weight gradient is already the SUM over ranks
        loss.backward()                                    # no allreduce_gradients for that parameter

    Needs NVLS multicast (NVSwitch systems; `FusedGradReduce.available()`); all ranks must run the same sequence of backward
    passes.  The weights must be replicated (plain data parallelism): grad_weight = grad_height . dh/dw is then linear in the
    summed grad_height."""

    @staticmethod
    def available(device=None):
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return False
        try:
            import torch.distributed._symmetric_memory as symm
            dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
            probe = symm.empty(64, dtype=torch.float32, device=dev)
            return bool(symm.rendezvous(probe, dist.group.WORLD.group_name).multicast_ptr)
        except Exception:
            return False

    def __init__(self, H, W, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.H, self.W = int(H), int(W)
        n = self.H * self.W
        self.buf = symm.empty(2 * n, dtype=torch.float32, device=device)
        self.hdl = symm.rendezvous(self.buf, self.group.group_name)
        if not self.hdl.multicast_ptr:
            raise RuntimeError("FusedGradReduce needs NVLS multicast support (symmetric memory has no multicast mapping here)")
        self.buf.zero_()
        self.hdl.barrier()
        self.views = [self.buf[:n].view(self.H, self.W), self.buf[n:].view(self.H, self.W)]
        self.mc = [int(self.hdl.multicast_ptr), int(self.hdl.multicast_ptr) + 4 * n]
        self.cur = 0

    def target(self):
        """Multicast address the adjoint kernel of THIS backward pass adds into."""
        return self.mc[self.cur]

    def finish(self):
        """After the adjoint kernel was enqueued: zero the idle buffer, barrier, return this rank's replica (= the sum)."""
        out = self.views[self.cur]
        self.cur ^= 1
        self.views[self.cur].zero_()
        self.hdl.barrier()
        return out


def fuse_gradient_allreduce(doe_layer, group=None):
    """Attach a FusedGradReduce to a DOE layer whose modulation is fused into ASM_prop: its height-map gradient (hence the weight
    gradient) comes out of backward already summed over the ranks.  `allreduce_gradients` skips such parameters."""
    H, W = doe_layer.doe_size
    red = FusedGradReduce(H, W, doe_layer.device, group)
    doe_layer.grad_reducer = red
    for p in doe_layer.parameters():
        p._thz_grad_is_reduced = True
    return red


# ----------------------------------------------------------------------------- slab-decomposed ASM
def _all_to_all(send, group):
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    return recv


class _SlabPlan:
    def __init__(self, G, rank, C, H, W, pad_h, pad_w, Hp, Wp, unpad, device, rowvec, colvec, scal, table, tf_mode, row_chunked=False):
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
        self.table = table[:, c0:c0 + self.Wc].contiguous().to(device) if table is not None else None      # [C, slot_c, slot_r]
        self.tf_mode = tf_mode
        self.row_chunked = 1 if (row_chunked and tf_mode == 0) else 0
        self.tw_h, self.tw_w, self.tw_c = N.twiddles(Hp, device), N.twiddles(Wp, device), N.twiddles(self.Wc, device)


SLAB_TIMINGS = None     # debug: set to a list to collect (label, cuda event) marks of the next _slab_run calls


def _mark(label):
    if SLAB_TIMINGS is not None:
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        SLAB_TIMINGS.append((label, e))


def _doe_kw(doe, stage):
    """Descriptor fields of the fused DOE for one stage of a slab run.  doe = None or dict(mode=1|2, hmap=<this rank's rows of
    the height map>, coef, xsaved=<this rank's rows of the saved field>, gh=<this rank's rows of grad_height>): the forward
    multiplies by p(h) in the row-FFT prologue (stage 1), the adjoint runs conj(p) / grad_height in the row-iFFT epilogue
    (stage 4) -- exactly where the single-GPU pipeline fuses them (thz_asm_desc.doe_mode)."""
    if doe is not None and ((doe["mode"] == 1 and stage == 1) or (doe["mode"] == 2 and stage == 4)):
        return dict(doe_mode=doe["mode"], doe_base=Fn.BASE_PLANE_THICKNESS, hmap=doe["hmap"], coef=doe["coef"],
                    xsaved=doe.get("xsaved"), gh=doe.get("gh"))
    return dict(doe_mode=0, doe_base=0.0, hmap=None, coef=None, xsaved=None, gh=None)


def _slab_run(x_local, p, conj, group, doe=None):
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
                  tw_h=p.tw_h, tf_row_chunked=p.row_chunked)
    _mark("start")
    # ---- stage 1: row FFT of the local rows -> t1 [nbc, Hl, Wp]
    t1 = torch.empty(nbc * Hl * Wp, dtype=torch.complex64, device=dev)
    Fn._asm_call(AH.build_desc(x=x_local, y=None, inH=Hl, inW=inW, Wp=Wp, in_r0=0, in_c0=in_c0, outH=Hl, outW=outW, out_r0=0,
                               out_c0=out_c0, colvec=p.colvec, table=p.table, tw_w=p.tw_w, ws=t1, stages=1, **common, **_doe_kw(doe, 1)), dev)
    _mark("row fft")
    # ---- transpose 1: rank j receives my rows of ITS column block
    send = t1.view(nbc, Hl, G, Wc).permute(2, 0, 1, 3).contiguous()               # [G, nbc, Hl, Wc]
    recv = _all_to_all(send, group)                                               # recv[i] = rows of rank i, my columns
    rowsT = max(inH, outH)
    t2 = torch.zeros(nbc, rowsT, Wc, dtype=torch.complex64, device=dev) if rowsT > inH else torch.empty(nbc, inH, Wc, dtype=torch.complex64, device=dev)
    t2[:, :inH] = recv.permute(1, 0, 2, 3).reshape(nbc, inH, Wc)
    _mark("transpose 1")
    # ---- stage 2: column FFT . H . column iFFT on the local columns (in place)
    Fn._asm_call(AH.build_desc(x=None, y=None, inH=inH, inW=min(inW, Wc), Wp=Wc, in_r0=in_r0, in_c0=0, outH=outH, outW=min(outW, Wc),
                               out_r0=out_r0, out_c0=0, colvec=p.colvec, table=p.table, tw_w=p.tw_c, ws=t2, stages=2, **common, **_doe_kw(doe, 2)), dev)
    _mark("column pass")
    # ---- transpose 2: rank j receives its output rows of my column block
    send = t2[:, :outH].reshape(nbc, G, Ol, Wc).permute(1, 0, 2, 3).contiguous()  # [G, nbc, Ol, Wc]
    recv = _all_to_all(send, group)                                               # recv[i] = my rows, columns of rank i
    t3 = recv.permute(1, 2, 0, 3).reshape(nbc * Ol * Wp).contiguous()             # [nbc, Ol, Wp]
    _mark("transpose 2")
    # ---- stage 3: row iFFT + crop of the local output rows
    want_y = doe is None or doe["mode"] != 2 or doe.get("want_gx", True)
    y = torch.empty(B, C, Ol, outW, dtype=torch.complex64, device=dev) if want_y else None
    Fn._asm_call(AH.build_desc(x=None, y=y, inH=Ol, inW=inW, Wp=Wp, in_r0=0, in_c0=in_c0, outH=Ol, outW=outW, out_r0=0,
                               out_c0=out_c0, colvec=p.colvec, table=p.table, tw_w=p.tw_w, ws=t3, stages=4, **common, **_doe_kw(doe, 4)), dev)
    _mark("row ifft")
    return y


# ----------------------------------------------------------------------------- slab FFT over peer memory (NVLink P2P)
class _PeerSlabs:
    """Two column slabs per rank in symmetric memory -- S1 [nbc, Wc/4, rows, 4] (row spectra, blocked for the column kernel)
    and S2 [nbc, rows, Wc] (column pass output) -- with peer-mapped pointers to all of them on every rank, so the row-FFT
    kernel can store each row segment straight into its owner's S1 and the row-iFFT kernel can read its rows straight
    out of the peers' S2: the two transposes of the slab FFT happen inside the kernels' stores / loads and overlap with the
    butterflies; no pack / all-to-all / unpack passes.  torch's symmetric-memory allocator provides the mapping and a
    stream-ordered cross-GPU barrier (plumbing); the data path is ours."""

    def __init__(self, numel, device, group):
        import torch.distributed._symmetric_memory as symm
        self.buf = symm.empty(4 * numel, dtype=torch.float32, device=device)      # 2 x numel complex64 as float pairs
        self.hdl = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        base = [int(q) for q in self.hdl.buffer_ptrs]
        self.ptrs1 = base
        self.ptrs2 = [q + 8 * numel for q in base]
        self.numel = numel

    def local(self, which):
        c = torch.view_as_complex(self.buf.view(-1, 2))
        return c[:self.numel] if which == 1 else c[self.numel:]

    def barrier(self):
        self.hdl.barrier()


SLAB_BLOCKED = int(os.environ.get("THZ_SLAB_BLOCKED", "0"))   # 1: S1 slabs in 4-column blocks (32-byte remote stores: slower on NVLink)


def _slab_stage_descs(p, x_local, y_local, rank, ptrs1, ptrs2, s1_local, conj, doe=None):
    """The three descriptors of one rank's part of a peer-memory slab propagation (stage 1 scatter into the S1 slabs, stage 2
    from the local S1 to the local S2, stage 4 gather from the S2 slabs).  ptrsK[d] = address of rank d's slab SK."""
    G, B, C = p.G, x_local.shape[0], p.C
    if not conj:
        inH, inW, in_r0, in_c0, outH, outW, out_r0, out_c0 = p.H, p.W, p.pad_h, p.pad_w, p.outH, p.outW, p.out_r0, p.out_c0
    else:
        inH, inW, in_r0, in_c0, outH, outW, out_r0, out_c0 = p.outH, p.outW, p.out_r0, p.out_c0, p.H, p.W, p.pad_h, p.pad_w
    Hl, Ol, Wc, Wp = inH // G, outH // G, p.Wc, p.Wp
    rowsT = max(inH, outH)
    common = dict(B=B, C=C, Hp=p.Hp, tf_mode=p.tf_mode, tf_conj=1 if conj else 0, rowvec=p.rowvec, scal=p.scal,
                  tw_h=p.tw_h, colvec=p.colvec, table=p.table, tf_row_chunked=p.row_chunked)
    d1 = AH.build_desc(x=x_local, y=None, inH=Hl, inW=inW, Wp=Wp, in_r0=0, in_c0=in_c0, outH=Hl, outW=outW, out_r0=0, out_c0=out_c0,
                       tw_w=p.tw_w, ws=None, stages=1, slab=(G, rank * Hl, rowsT, ptrs1, SLAB_BLOCKED), **common, **_doe_kw(doe, 1))
    d2 = AH.build_desc(x=None, y=None, inH=inH, inW=min(inW, Wc), Wp=Wc, in_r0=in_r0, in_c0=0, outH=outH, outW=min(outW, Wc),
                       out_r0=out_r0, out_c0=0, tw_w=p.tw_c, ws=s1_local, stages=2, slab=(G, 0, rowsT, [ptrs2[rank]], SLAB_BLOCKED),
                       **common, **_doe_kw(doe, 2))
    d3 = AH.build_desc(x=None, y=y_local, inH=Ol, inW=inW, Wp=Wp, in_r0=0, in_c0=in_c0, outH=Ol, outW=outW, out_r0=0, out_c0=out_c0,
                       tw_w=p.tw_w, ws=None, stages=4, slab=(G, rank * Ol, rowsT, ptrs2), **common, **_doe_kw(doe, 4))
    return d1, d2, d3


def _slab_run_peer(x_local, p, conj, slabs, doe=None):
    """Slab-decomposed propagation with the transposes fused into the row kernels (see _PeerSlabs)."""
    B, C, dev = x_local.shape[0], p.C, x_local.device
    outH, outW = (p.outH, p.outW) if not conj else (p.H, p.W)
    want_y = doe is None or doe["mode"] != 2 or doe.get("want_gx", True)
    y = torch.empty(B, C, outH // p.G, outW, dtype=torch.complex64, device=dev) if want_y else None
    d1, d2, d3 = _slab_stage_descs(p, x_local, y, p.rank, slabs.ptrs1, slabs.ptrs2, slabs.local(1), conj, doe)
    _mark("start")
    Fn._asm_call(d1, dev)                # row FFT; stores go to the owners of the column blocks (their S1)
    _mark("row fft + scatter")
    slabs.barrier()                      # all row segments have landed in my S1
    Fn._asm_call(d2, dev)                # column FFT . H . column iFFT: S1 -> S2
    _mark("column pass")
    slabs.barrier()                      # every S2 is final
    Fn._asm_call(d3, dev)                # row iFFT; loads come from the owners of the column blocks (their S2)
    _mark("gather + row ifft")
    return y


def slab_emulate_ranks(x_full, plans, conj=False, doe=None):
    """Test helper: run the G ranks of a peer-memory slab propagation one after the other in THIS process (all column
    slabs on one device), exercising exactly the descriptors and kernels of `_slab_run_peer`.  x_full [B,C,H,W].
    doe = dict(mode, hmap [H,W], coef[, xsaved [B,C,H,W], gh [H,W]]) with FULL tensors: each emulated rank gets its rows."""
    G, p0 = len(plans), plans[0]
    B, dev = x_full.shape[0], x_full.device
    inH = p0.H if not conj else p0.outH
    outH, outW = (p0.outH, p0.outW) if not conj else (p0.H, p0.W)
    rowsT = max(p0.H, p0.outH)
    n = B * p0.C * rowsT * p0.Wc
    s1 = [torch.zeros(n, dtype=torch.complex64, device=dev) for _ in range(G)]
    s2 = [torch.zeros(n, dtype=torch.complex64, device=dev) for _ in range(G)]
    ptrs1, ptrs2 = [t.data_ptr() for t in s1], [t.data_ptr() for t in s2]
    Hl = inH // G
    ys = [torch.empty(B, p0.C, outH // G, outW, dtype=torch.complex64, device=dev) for _ in range(G)]
    xs = [x_full[:, :, r * Hl:(r + 1) * Hl].contiguous() for r in range(G)]
    def rank_doe(r):
        if doe is None:
            return None
        rows = (p0.H // G)
        d = dict(mode=doe["mode"], coef=doe["coef"], hmap=doe["hmap"][r * rows:(r + 1) * rows].contiguous())
        if doe["mode"] == 2:
            d["xsaved"] = doe["xsaved"][:, :, r * rows:(r + 1) * rows].contiguous()
            d["gh"] = doe["gh_parts"][r]
        return d

    if doe is not None and doe["mode"] == 2:
        doe["gh_parts"] = [torch.empty(p0.H // G, p0.W, dtype=torch.float32, device=dev) for _ in range(G)]
    descs = [_slab_stage_descs(plans[r], xs[r], ys[r], r, ptrs1, ptrs2, s1[r], conj, rank_doe(r)) for r in range(G)]
    for k in range(3):
        for r in range(G):
            Fn._asm_call(descs[r][k], dev)
    return torch.cat(ys, dim=2)


def _peer_transport_possible(plan, device):
    from . import _native
    return (device.type == "cuda" and dist.get_backend() == "nccl" and plan.G <= 8 and plan.Wc % 4 == 0 and
            bool(_native.lib().thz_fft_is_static(plan.Wp)) and bool(_native.lib().thz_fft_is_static(plan.Hp)))


class _SlabFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_local, plan, group, slabs):
        ctx.plan, ctx.group, ctx.slabs = plan, group, slabs
        x_local = Fn._c64(x_local, "field.data")
        return _slab_run_peer(x_local, plan, False, slabs) if slabs is not None else _slab_run(x_local, plan, False, group)

    @staticmethod
    def backward(ctx, g):
        g = Fn._c64(g, "grad_output")
        gx = _slab_run_peer(g, ctx.plan, True, ctx.slabs) if ctx.slabs is not None else _slab_run(g, ctx.plan, True, ctx.group)
        return gx, None, None, None


class _SlabDoeFn(torch.autograd.Function):
    """y_local = slab-ASM(x_local * p(h[local rows])) with the DOE fused into the slab pipeline's own kernels, as on one GPU:
    forward = DOE phase in the row-FFT prologue; backward = conj(p) multiply and grad_height in the row-iFFT epilogue.
    Every rank returns grad_height for ITS rows (zeros elsewhere); the caller's gradient all-reduce sums the slabs."""

    @staticmethod
    def forward(ctx, x_local, hmap, plan, group, slabs, coef):
        x_local = Fn._c64(x_local, "field.data")
        N.require_cuda(hmap, "height_map")
        hmap = hmap.to(torch.float32).contiguous()
        rows = plan.H // plan.G
        h_loc = hmap[plan.rank * rows:(plan.rank + 1) * rows].contiguous()
        doe = dict(mode=1, hmap=h_loc, coef=coef)
        y = _slab_run_peer(x_local, plan, False, slabs, doe) if slabs is not None else _slab_run(x_local, plan, False, group, doe)
        ctx.plan, ctx.group, ctx.slabs, ctx.coef, ctx.hshape = plan, group, slabs, coef, hmap.shape
        ctx.save_for_backward(x_local, h_loc)
        return y

    @staticmethod
    def backward(ctx, g):
        plan = ctx.plan
        x_local, h_loc = ctx.saved_tensors
        g = Fn._c64(g, "grad_output")
        rows = plan.H // plan.G
        need_x, need_h = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gh_loc = torch.empty(rows, plan.W, dtype=torch.float32, device=g.device)
        doe = dict(mode=2, hmap=h_loc, coef=ctx.coef, xsaved=x_local, gh=gh_loc, want_gx=need_x)
        gx = _slab_run_peer(g, plan, True, ctx.slabs, doe) if ctx.slabs is not None else _slab_run(g, plan, True, ctx.group, doe)
        gh = None
        if need_h:
            gh = torch.zeros(ctx.hshape, dtype=torch.float32, device=g.device)
            gh[plan.rank * rows:(plan.rank + 1) * rows] = gh_loc
        return (gx if need_x else None), gh, None, None, None, None


class SlabAsm(torch.nn.Module):
    """Band-limited ASM of a field whose ROWS are distributed over the ranks of a process group.

    forward(local_field) takes this rank's row slab as an ElectricField [B, C, H/G, W] and returns this rank's
    slab of the propagated field.  Constructor arguments as ASM_prop (Props/ASM_Prop.py:19-27)."""

    def __init__(self, z_distance=0.0, do_padding=True, do_unpad_after_pad=True, padding_scale=None, bandlimit_kernel=True,
                 bandlimit_type="exact", group=None, kernel_mode="auto", transport="auto"):
        super().__init__()
        if transport not in ("auto", "peer", "nccl"):
            raise ValueError("transport must be 'auto', 'peer' or 'nccl'")
        # 'peer': transposes fused into the row kernels over NVLink peer memory (_PeerSlabs); 'nccl': pack + all_to_all +
        # unpack around the kernels (also the CPU / gloo test path); 'auto': peer when it can be set up, else nccl.
        self.transport = transport
        self._slabs = None
        self.z = torch.as_tensor(z_distance, dtype=torch.float32)
        self.do_padding, self.do_unpad_after_pad = do_padding, do_unpad_after_pad
        self.padding_scale = AH.normalise_padding_scale(padding_scale, do_padding)
        self.bandlimit_kernel, self.bandlimit_type = bandlimit_kernel, bandlimit_type
        self.group, self.kernel_mode = group, kernel_mode
        self._key, self._plan, self._fast = None, None, None

    def _make_slabs(self, nbc, device):
        """Symmetric-memory column slabs for the peer transport (collective: every rank calls it with the same sizes)."""
        p = self._plan
        if self.transport == "nccl" or not _peer_transport_possible(p, device):
            if self.transport == "peer":
                raise RuntimeError("SlabAsm(transport='peer') needs CUDA + NCCL, <= 8 ranks and a static-path padded width")
            return None
        try:
            return _PeerSlabs(nbc * max(p.H, p.outH) * p.Wc, device, self.group)
        except Exception as e:          # no P2P mapping on this system: the NCCL transport still works
            if self.transport == "peer":
                raise
            import warnings
            warnings.warn("SlabAsm: peer-memory transport unavailable (%s); using all_to_all" % (e,))
            return None

    def forward(self, field):
        G, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        # a DOE layer's output arrives un-materialised (ElectricField._deferred): the slab pipeline fuses it like ASM_prop does.
        # The layer holds the FULL height map [H, W] (replicated parameters), the field is this rank's row slab.
        deferred = getattr(field, "_deferred", None) if getattr(field, "_data", None) is None else None
        if deferred is not None and (getattr(deferred, 'rows', None) is None or deferred.mask is not None or deferred.mul is not None or
                                     deferred.height_map.shape[0] != deferred.x.shape[2] * G):
            deferred = None                     # not a row slab under a global map: materialise and propagate plainly
        data = deferred.x if deferred is not None else field.data
        B, C, Hl, W = data.shape
        H = Hl * G
        # same tensor objects (unchanged in place) as last time -> same plan, no device->host reads (they would serialise
        # the host with the GPU on every call)
        fast = (id(field.spacing), field.spacing._version, id(field.wavelengths), field.wavelengths._version, C, H, W,
                float(self.z), G, rank, str(data.device))
        if self._plan is not None and self._fast == fast:
            key = self._key
        else:
            key = (C, H, W, tuple(field.spacing.detach().cpu().tolist()), tuple(field.wavelengths.detach().cpu().tolist()),
                   float(self.z), G, rank)
            self._fast, self._fast_refs = fast, (field.spacing, field.wavelengths)
        if key != self._key:
            pad_h, pad_w, Hp, Wp = AH.compute_padding(H, W, self.padding_scale, self.do_padding)
            rowvec, colvec, scal = AH.tf_vectors(Hp, Wp, field.spacing, field.wavelengths, self.z, self.bandlimit_kernel, self.bandlimit_type)
            table, mode = None, 0
            chunked = AH.row_vectors_chunked(Hp)
            mode_ = self.kernel_mode                 # as ASM_prop: 'auto' generates H in registers only where that keeps parity
            if mode_ == "auto":
                est = AH.inregister_estimate_for(Hp, Wp, field.spacing, field.wavelengths, self.z, self.bandlimit_kernel, self.bandlimit_type)
                mode_ = "inregister" if est <= AH.INREGISTER_BUDGET else "cached"
            dv = AH.tf_device_vectors(rowvec, colvec, scal, chunked=chunked) if mode_ == "inregister" else None
            self.resolved_kernel_mode = "inregister" if dv is not None else "cached"
            if dv is not None:
                rowvec, colvec, scal = dv
            else:
                table = AH.tf_table_device(Hp, Wp, field.spacing, field.wavelengths, self.z, self.bandlimit_kernel, self.bandlimit_type,
                                           data.device) if data.device.type == "cuda" else None
                if table is None:
                    Hc = AH.tf_centred_reference_order(Hp, Wp, field.spacing, field.wavelengths, self.z, self.bandlimit_kernel, self.bandlimit_type)
                    table = AH.tf_table_slot_order(Hc)
                mode = 1
            self._plan = _SlabPlan(G, rank, C, H, W, pad_h, pad_w, Hp, Wp, bool(self.do_padding and self.do_unpad_after_pad),
                                   data.device, rowvec, colvec, scal, table, mode, row_chunked=chunked)
            self._key = key
            self._slabs = self._make_slabs(B * C, data.device)
        elif self._slabs is not None and self._slabs.numel < B * C * max(self._plan.H, self._plan.outH) * self._plan.Wc:
            self._slabs = self._make_slabs(B * C, data.device)
        if deferred is not None:
            out = _SlabDoeFn.apply(data, deferred.height_map, self._plan, self.group, self._slabs, deferred.coef)
        else:
            out = _SlabFn.apply(data, self._plan, self.group, self._slabs)
        res = ElectricField(out, wavelengths=field.wavelengths, spacing=field.spacing, device=data.device)
        if self._plan.outH == self._plan.H:         # still a row slab of the same grid: a following DOE layer may fuse again
            res._row_slab = (rank, G)
        return res
