"""World-size-2 `gloo` tests of the multi-GPU host logic, run on CPU.

The ranks execute the REAL kernel bodies through the CPU replay harness (tests/emul) -- injected at the
single call site `functional._asm_call` -- so what is exercised here is exactly what runs on N GPUs:
sharding, the flat-bucket gradient all-reduce, and the slab FFT's pack / all-to-all / unpack schedule."""
import ctypes
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import rel_l2

mm = 1e-3


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _install_cpu_replay():
    from helpers import emul_lib
    from quantizationawarethzdoe_b200 import _native as N, functional as Fn
    E = emul_lib()

    def call(desc, device):
        rc = E.thz_emul_asm_propagate(ctypes.byref(desc), 148)
        assert rc == 0, rc

    Fn._asm_call = call
    N.require_cuda = lambda t, name="tensor": None


def _worker(rank, world, port, which, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        _install_cpu_replay()
        out[rank] = {"dp": _dp_case, "slab": _slab_case, "slab_doe": _slab_doe_case}[which](rank, world)
    finally:
        dist.destroy_process_group()


def _dp_case(rank, world):
    """4 wavelengths over 2 ranks; replicated STE DOE; all-reduced weight gradient == full-batch oracle gradient."""
    from oracle import asm_oracle as AO, doe_oracle as DO
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField, FullPrecisionDOELayer
    from quantizationawarethzdoe_b200 import parallel as P
    from quantizationawarethzdoe_b200 import functional as Fn
    cpu = torch.device("cpu")
    n, lams = 32, [1 * mm, 1.02 * mm, 1.04 * mm, 1.06 * mm]
    torch.manual_seed(0)
    x = torch.randn(1, 4, n, n, dtype=torch.complex64)
    torch.manual_seed(1)
    w = torch.randn(1, 1, n, n)
    # product modules on CPU tensors (kernels replayed); the height construction kernel is replaced by its torch twin
    Fn.HeightFromWeightFn.apply = staticmethod(lambda ww, hmax, c: DO.sigmoid_height(ww, hmax, c))
    doe = FullPrecisionDOELayer(dict(doe_size=[n, n], doe_dxy=0.5 * mm, height_constraint_max=1 * mm, tolerance=None,
                                     material=[2.66, 0.003]), device=cpu)
    with torch.no_grad():
        doe.weight_height_map.copy_(w)
    asm = ASM_prop(z_distance=0.1, device=cpu, kernel_mode="cached")
    asm.check_Zc = False
    full = ElectricField(x, wavelengths=lams, spacing=0.5 * mm, device=cpu)
    mine = P.shard_field(full, rank, world, axis="wavelength")
    assert mine.data.shape[1] == 2 and P.shard_range(5, 0, 2) == (0, 3) and P.shard_range(5, 1, 2) == (3, 5)
    y = asm(doe(mine)).data
    loss = 0.5 * (y.real ** 2 + y.imag ** 2).sum()
    loss.backward()
    P.allreduce_gradients(doe.parameters())
    # oracle: all four wavelengths at once
    wo = w.clone().requires_grad_(True)
    yo = AO.asm_forward(DO.modulate(x, DO.sigmoid_height(wo[0, 0], 1 * mm), lams, 2.66, 0.003), lams, 0.5 * mm, 0.1)
    (go,) = torch.autograd.grad(0.5 * (yo.real ** 2 + yo.imag ** 2).sum(), wo)
    return rel_l2(doe.weight_height_map.grad, go)


def _slab_case(rank, world):
    """One 64x96 field (2 wavelengths) split by rows over 2 ranks: slab forward and adjoint == oracle."""
    from oracle import asm_oracle as AO
    from quantizationawarethzdoe_b200 import ElectricField
    from quantizationawarethzdoe_b200 import parallel as P
    cpu = torch.device("cpu")
    H, W, lams = 64, 96, [1 * mm, 1.05 * mm]
    torch.manual_seed(0)
    x = torch.randn(2, 2, H, W, dtype=torch.complex64)
    g = torch.randn(2, 2, H, W, dtype=torch.complex64)
    lo, hi = P.shard_range(H, rank, world)
    errs = []
    for mode in ("cached", "inregister"):
        slab = P.SlabAsm(z_distance=0.1, kernel_mode=mode)
        xl = x[:, :, lo:hi].contiguous().requires_grad_(True)
        yl = slab(ElectricField(xl, wavelengths=lams, spacing=0.5 * mm, device=cpu)).data
        (gxl,) = torch.autograd.grad(yl, xl, g[:, :, lo:hi].contiguous())
        xo = x.clone().requires_grad_(True)
        yo = AO.asm_forward(xo, lams, 0.5 * mm, 0.1)
        (gxo,) = torch.autograd.grad(yo, xo, g)
        errs += [rel_l2(yl.detach(), yo.detach()[:, :, lo:hi]), rel_l2(gxl, gxo[:, :, lo:hi])]
    return max(errs)


def _slab_doe_case(rank, world):
    """DOE layer + slab-decomposed ASM: the rank's row slab goes through a full-size (replicated) DOE whose modulation is
    fused into the slab pipeline's row-FFT prologue, its adjoint (conj(p), grad_height of the local rows) into the row-iFFT
    epilogue; the all-reduced weight gradient and the local input gradient == autograd through the oracle on the full grid."""
    from oracle import asm_oracle as AO, doe_oracle as DO
    from quantizationawarethzdoe_b200 import ElectricField, FullPrecisionDOELayer
    from quantizationawarethzdoe_b200 import functional as Fn
    from quantizationawarethzdoe_b200 import parallel as P
    cpu = torch.device("cpu")
    H, W, lams = 64, 48, [1 * mm, 1.05 * mm]
    torch.manual_seed(0)
    x = torch.randn(1, 2, H, W, dtype=torch.complex64)
    g = torch.randn(1, 2, H, W, dtype=torch.complex64)
    torch.manual_seed(1)
    w = torch.randn(1, 1, H, W)
    Fn.HeightFromWeightFn.apply = staticmethod(lambda ww, hmax, c: DO.sigmoid_height(ww, hmax, c))
    doe = FullPrecisionDOELayer(dict(doe_size=[H, W], doe_dxy=0.5 * mm, height_constraint_max=1 * mm, tolerance=None,
                                     material=[2.66, 0.003]), device=cpu)
    with torch.no_grad():
        doe.weight_height_map.copy_(w)
    lo, hi = P.shard_range(H, rank, world)
    slab = P.SlabAsm(z_distance=0.1, kernel_mode="cached")
    xf = x.clone().requires_grad_(True)
    mine = P.shard_rows(ElectricField(xf, wavelengths=lams, spacing=0.5 * mm, device=cpu), rank, world)
    u = doe(mine)
    assert u._data is None and u._deferred.rows == (lo, hi)            # modulation deferred, map kept at full size
    yl = slab(u).data
    gxf, gw = torch.autograd.grad(yl, (xf, doe.weight_height_map), g[:, :, lo:hi].contiguous())
    doe.weight_height_map.grad = gw
    P.allreduce_gradients(doe.parameters())
    xo, wo = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    yo = AO.asm_forward(DO.modulate(xo, DO.sigmoid_height(wo[0, 0], 1 * mm), lams, 2.66, 0.003), lams, 0.5 * mm, 0.1)
    gxo, gwo = torch.autograd.grad(yo, (xo, wo), g)
    return max(rel_l2(yl.detach(), yo.detach()[:, :, lo:hi]), rel_l2(gxf[:, :, lo:hi], gxo[:, :, lo:hi]),
               rel_l2(doe.weight_height_map.grad, gwo))


@pytest.mark.parametrize("which,tol", [("dp", 2e-6), ("slab", 2e-5), ("slab_doe", 2e-6)])
def test_two_rank_gloo(which, tol):
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, which, out), nprocs=2, join=True)
        assert len(out) == 2
        for r in range(2):
            assert out[r] < tol, (which, r, out[r])
