"""Real multi-rank runs of the N > 1 paths, one process per GPU under torchrun: skipped on boxes with fewer
than two GPUs (the driver's single-GPU test tier), run by `gpurun --gpus 2 -- python -m pytest tests -m gpu -k multi_gpu`.
The same schedules run on the CPU under gloo in tests/test_distributed_cpu.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_ngpus() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("transport", ["peer", "nccl"])
def test_slab_and_data_parallel_equal_single_gpu(transport):
    """tools/multi_gpu_check.py on min(#GPUs, 4) ranks: slab-decomposed ASM (forward, adjoint, DOE fused) == ASM_prop on one
    GPU to 2e-6 (same arithmetic), data-parallel weight gradient == full-batch gradient."""
    n = min(_ngpus(), 4)
    env = dict(os.environ, THZ_SLAB_TRANSPORT=transport, THZ_SLAB_N="1024")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(29511 + (transport == "nccl")), os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert ("[%s transport]" % transport) in r.stdout


@pytest.mark.skipif(_ngpus() < 2, reason="needs at least 2 GPUs")
def test_modules_on_a_non_current_device():
    """the library used the calling thread's CURRENT device for launches, attributes and the SM count.  Every entry
    point now makes the device that owns its pointers current for the call: a pipeline built on cuda:1 runs, and matches the
    same pipeline on cuda:0 bit for bit, while the current device stays 0."""
    from quantizationawarethzdoe_b200 import ASM_prop, CZT_prop, ElectricField, STEQuantizedDOELayer
    mm = 1e-3
    torch.cuda.set_device(0)
    outs = []
    for dev in (torch.device("cuda:0"), torch.device("cuda:1")):
        torch.manual_seed(0)
        x = torch.randn(1, 2, 256, 256, dtype=torch.complex64)
        torch.manual_seed(1)
        doe = STEQuantizedDOELayer(dict(doe_size=[256, 256], doe_dxy=0.5 * mm, doe_level=4, height_constraint_max=1 * mm, tolerance=None,
                                        material=[2.66, 0.003]), {}, device=dev)
        asm = ASM_prop(z_distance=0.1, device=dev)
        asm.check_Zc = False
        xd = x.to(dev).requires_grad_(True)
        y = asm(doe(ElectricField(xd, wavelengths=[1 * mm, 1.03 * mm], spacing=0.5 * mm, device=dev))).data
        gx, gw = torch.autograd.grad(y, (xd, doe.weight_height_map), y.detach())
        c = CZT_prop(z_distance=0.4, device=dev)(ElectricField(xd.detach(), wavelengths=[1 * mm, 1.03 * mm], spacing=0.5 * mm, device=dev),
                                                128, 128, 0.2 * mm, 0.2 * mm).data
        torch.cuda.synchronize(dev)
        assert y.device == dev and c.device == dev
        outs.append((y.detach().cpu(), gx.cpu(), gw.cpu(), c.cpu()))
    assert torch.cuda.current_device() == 0
    for a, b in zip(*outs):
        assert torch.equal(a, b)
