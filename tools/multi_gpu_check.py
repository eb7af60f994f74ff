#!/usr/bin/env python3
"""Multi-GPU functional check (launch with torchrun, one rank per GPU):
   * slab-decomposed ASM over all ranks == single-GPU ASM_prop on rank 0 (forward and adjoint);
   * data-parallel DOE step: all-reduced weight gradient == single-GPU full-batch gradient.
   Prints one line per check on rank 0 and exits non-zero on failure."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from quantizationawarethzdoe_b200 import ASM_prop, ElectricField, STEQuantizedDOELayer, parallel as P  # noqa: E402

mm = 1e-3
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
n = int(os.environ.get("THZ_SLAB_N", "2048"))


def rel(a, b):
    return float((a - b).norm() / b.norm())


# ---- slab vs single GPU
lams = [1 * mm, 1.03 * mm][:int(os.environ.get("THZ_SLAB_C", "2"))]
torch.manual_seed(0)
x = torch.randn(1, len(lams), n, n, dtype=torch.complex64, device=dev)
g = torch.randn(1, len(lams), n, n, dtype=torch.complex64, device=dev)
lo, hi = P.shard_range(n, rank, world)
transport = os.environ.get("THZ_SLAB_TRANSPORT", "auto")     # auto / peer / nccl
slab = P.SlabAsm(z_distance=0.1, transport=transport)
xl = x[:, :, lo:hi].contiguous().requires_grad_(True)
yl = slab(ElectricField(xl, wavelengths=lams, spacing=0.5 * mm, device=dev)).data
(gxl,) = torch.autograd.grad(yl, xl, g[:, :, lo:hi].contiguous())
# timing of the slab forward (device events, max over ranks)
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
fl = ElectricField(xl.detach(), wavelengths=lams, spacing=0.5 * mm, device=dev)   # one field object: plan lookups stay on the fast path
slab(fl)
torch.cuda.synchronize()
dist.barrier()
e0.record()
for _ in range(10):
    slab(fl)
e1.record()
torch.cuda.synchronize()
tms = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
dist.all_reduce(tms, op=dist.ReduceOp.MAX)
P.SLAB_TIMINGS = []
slab(fl)
torch.cuda.synchronize()
marks, P.SLAB_TIMINGS = P.SLAB_TIMINGS, None
stage_ms = ", ".join("%s %.2f" % (marks[i][0], marks[i - 1][1].elapsed_time(marks[i][1])) for i in range(1, len(marks)))
asm = ASM_prop(z_distance=0.1, device=dev)
asm.check_Zc = False
ff = ElectricField(x, wavelengths=lams, spacing=0.5 * mm, device=dev)
asm(ff)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    asm(ff)
e1.record()
torch.cuda.synchronize()
single_ms = e0.elapsed_time(e1) / 10
xf = x.clone().requires_grad_(True)
yf = asm(ElectricField(xf, wavelengths=lams, spacing=0.5 * mm, device=dev)).data
(gxf,) = torch.autograd.grad(yf, xf, g)
e = torch.tensor([rel(yl.detach(), yf.detach()[:, :, lo:hi]), rel(gxl, gxf[:, :, lo:hi])], device=dev)
dist.all_reduce(e, op=dist.ReduceOp.MAX)
if rank == 0:
    print("slab %d^2 (padded %d^2) over %d GPUs [%s transport] vs single GPU: fwd %.2e adjoint %.2e; slab forward %.2f ms, single-GPU forward %.2f ms" % (
        n, 2 * n, world, "peer" if slab._slabs is not None else "nccl", e[0], e[1], float(tms), single_ms))
    print("  rank 0 stages (ms):", stage_ms)
ok &= bool(e.max() < 2e-6)

# ---- DOE fused into the slab pipeline (prologue of the row-FFT kernel / epilogue of the row-iFFT kernel) vs one GPU
torch.manual_seed(7)
doe_s = STEQuantizedDOELayer(dict(doe_size=[n, n], doe_dxy=0.5 * mm, doe_level=4, height_constraint_max=1 * mm, tolerance=None,
                                  material=[2.66, 0.003]), {}, device=dev)
xs = x.clone().requires_grad_(True)
mine = P.shard_rows(ElectricField(xs, wavelengths=lams, spacing=0.5 * mm, device=dev), rank, world)
ys = slab(doe_s(mine)).data
gxs, gws = torch.autograd.grad(ys, (xs, doe_s.weight_height_map), g[:, :, lo:hi].contiguous())
dist.all_reduce(gws)                        # every rank holds the rows it owns; the sum is the full weight gradient
xf2 = x.clone().requires_grad_(True)
yf2 = asm(doe_s(ElectricField(xf2, wavelengths=lams, spacing=0.5 * mm, device=dev))).data
gxf2, gwf2 = torch.autograd.grad(yf2, (xf2, doe_s.weight_height_map), g)
e = torch.tensor([rel(ys.detach(), yf2.detach()[:, :, lo:hi]), rel(gxs[:, :, lo:hi], gxf2[:, :, lo:hi]), rel(gws, gwf2)], device=dev)
dist.all_reduce(e, op=dist.ReduceOp.MAX)
if rank == 0:
    print("DOE-fused slab over %d GPUs vs single GPU: fwd %.2e grad_x %.2e grad_w %.2e" % (world, e[0], e[1], e[2]))
ok &= bool(e.max() < 2e-6)

# ---- chirp-z propagation sharded over wavelengths (independent units, no collective) vs all wavelengths on one GPU
from quantizationawarethzdoe_b200 import CZT_prop  # noqa: E402
Cz = 2 * world
lam_z = [1 * mm * (1 + 0.02 * c) for c in range(Cz)]
torch.manual_seed(11)
xz = torch.randn(1, Cz, 256, 256, dtype=torch.complex64, device=dev)
czt = CZT_prop(z_distance=0.4, device=dev)
mine_z = P.shard_field(ElectricField(xz, wavelengths=lam_z, spacing=0.5 * mm, device=dev), rank, world)
yz = czt(mine_z, 128, 128, 0.2 * mm, 0.2 * mm).data
czt_full = CZT_prop(z_distance=0.4, device=dev)
yz_full = czt_full(ElectricField(xz, wavelengths=lam_z, spacing=0.5 * mm, device=dev), 128, 128, 0.2 * mm, 0.2 * mm).data
zlo, zhi = P.shard_range(Cz, rank, world)
e = torch.tensor([rel(yz, yz_full[:, zlo:zhi])], device=dev)
dist.all_reduce(e, op=dist.ReduceOp.MAX)
if rank == 0:
    print("CZT sharded over %d x %d wavelengths vs one GPU: %.2e" % (world, Cz // world, e[0]))
ok &= bool(e.max() < 1e-6)

# ---- data parallel over wavelengths
C = 2 * world
lam_all = [1 * mm * (1 + 0.01 * c) for c in range(C)]
torch.manual_seed(1)
xa = torch.randn(1, C, 512, 512, dtype=torch.complex64, device=dev)
torch.manual_seed(2)
doe = STEQuantizedDOELayer(dict(doe_size=[512, 512], doe_dxy=0.5 * mm, doe_level=4, height_constraint_max=1 * mm, tolerance=None,
                                material=[2.66, 0.003]), {}, device=dev)
full = ElectricField(xa, wavelengths=lam_all, spacing=0.5 * mm, device=dev)
y = asm(doe(P.shard_field(full, rank, world))).data
(0.5 * (y.real ** 2 + y.imag ** 2).sum()).backward()
P.allreduce_gradients(doe.parameters())
g_dp = doe.weight_height_map.grad.clone()
doe.weight_height_map.grad = None
y = asm(doe(full)).data
(0.5 * (y.real ** 2 + y.imag ** 2).sum()).backward()
e = torch.tensor([rel(g_dp, doe.weight_height_map.grad)], device=dev)
dist.all_reduce(e, op=dist.ReduceOp.MAX)
if rank == 0:
    print("data-parallel grad over %d GPUs vs full batch: %.2e" % (world, e[0]))
ok &= bool(e.max() < 1e-5)

# ---- the same sum formed inside the adjoint's last kernel (multimem.red over the NVSwitch, parallel.FusedGradReduce) instead of
# by the NCCL all-reduce after it; also through the unfused fall-back (modulation materialised by reading .data)
if P.FusedGradReduce.available(dev):
    g_full = doe.weight_height_map.grad.clone()
    P.fuse_gradient_allreduce(doe)
    errs = []
    for it in range(3):                                   # three passes: both buffers of the double buffer get re-used
        doe.weight_height_map.grad = None
        y = asm(doe(P.shard_field(full, rank, world))).data
        (0.5 * (y.real ** 2 + y.imag ** 2).sum()).backward()
        P.allreduce_gradients(doe.parameters())           # must be a no-op for the fused parameter
        errs.append(rel(doe.weight_height_map.grad, g_full))
    doe.weight_height_map.grad = None
    f = doe(P.shard_field(full, rank, world))
    _ = f.data                                            # evaluates the modulation: the propagation can no longer fuse it
    y = asm(f).data
    (0.5 * (y.real ** 2 + y.imag ** 2).sum()).backward()
    errs.append(rel(doe.weight_height_map.grad, g_full))
    e = torch.tensor(errs, device=dev)
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("in-kernel gradient sum (multimem.red) over %d GPUs vs full batch: %s; unfused fall-back %.2e"
              % (world, " ".join("%.2e" % v for v in e[:3].tolist()), e[3]))
    ok &= bool(e.max() < 1e-5)
elif rank == 0:
    print("in-kernel gradient sum: NVLS multicast not available here, skipped")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
