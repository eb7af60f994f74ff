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
