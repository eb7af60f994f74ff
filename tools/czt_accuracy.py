#!/usr/bin/env python3
"""Accuracy and speed of the two Toeplitz-GEMM implementations against a float64 evaluation of the same
Toeplitz product (test/diagnostic tool; torch.matmul in float64 is only the yardstick here).
usage: tools/czt_accuracy.py [H] [M] [C]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from quantizationawarethzdoe_b200 import CZT_prop, ElectricField  # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
M = int(sys.argv[2]) if len(sys.argv) > 2 else H // 2
C = int(sys.argv[3]) if len(sys.argv) > 3 else 2
mm = 1e-3
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(1, C, H, H, dtype=torch.complex64, device=dev)
lams = [1 * mm * (1 + 0.01 * c) for c in range(C)]
outs, times = {}, {}
for impl in ("tc", "simt"):
    os.environ["THZ_CZT_IMPL"] = impl
    czt = CZT_prop(z_distance=0.5, device=dev)
    f = ElectricField(x, wavelengths=lams, spacing=0.5 * mm, device=dev)
    y = czt(f, M, M, 0.1 * mm, 0.1 * mm).data
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5 if impl == "tc" else 2
    e0.record()
    for _ in range(reps):
        y = czt(f, M, M, 0.1 * mm, 0.1 * mm).data
    e1.record()
    torch.cuda.synchronize()
    outs[impl], times[impl] = y, e0.elapsed_time(e1) / reps
p = czt._plan
k1 = torch.arange(p.M1, device=dev)[:, None]
h = torch.arange(H, device=dev)[None, :]
Ty = p.gy.to(torch.complex128)[:, (H + k1 - h) % p.Ly]
Tx = p.gx.to(torch.complex128)[:, (H + k1 - h) % p.Lx]
ref = p.Q.to(torch.complex128) * torch.matmul(torch.matmul(Ty, x[0].to(torch.complex128) * p.P.to(torch.complex128)), Tx.transpose(-2, -1))
flops = 8.0 * (M * H * H + M * H * M) * C
for impl in ("tc", "simt"):
    err = float((outs[impl][0].to(torch.complex128) - ref).norm() / ref.norm())
    print("%-5s H=%d M=%d C=%d  rel-L2 vs float64: %.2e   %.3f ms  (%.1f TFLOP/s complex-as-4-real, x3 = %.1f tensor TF/s)" % (
        impl, H, M, C, err, times[impl], flops / times[impl] / 1e9, 3 * flops / times[impl] / 1e9))
