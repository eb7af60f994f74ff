#!/usr/bin/env python3
"""Config 5 on one GPU (one 8192^2 field on a 16384^2 canvas, forward + adjoint) with per-kernel event timing -- for A/B runs of
the layout switches (THZ_T2_LOG2, THZ_NO_K2FAST, ...).  python tools/c5_ab.py [--n 8192]"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from quantizationawarethzdoe_b200 import ASM_prop, ElectricField, _native as N  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=8192)
ap.add_argument("--tag", default="")
a = ap.parse_args()
mm = 1e-3
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(1, 1, a.n, a.n, dtype=torch.complex64, device=dev).requires_grad_(True)
asm = ASM_prop(z_distance=0.1, device=dev, kernel_mode="inregister")
asm.check_Zc = False
f = ElectricField(x, wavelengths=[1 * mm], spacing=0.5 * mm, device=dev)


def step():
    y = asm(f).data
    (g,) = torch.autograd.grad(y, x, y.detach())
    return y, g


for _ in range(2):
    y, g = step()
torch.cuda.synchronize()
chk = (float(y.detach().abs().double().sum()), float(g.abs().double().sum()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    step()
e1.record()
torch.cuda.synchronize()
plain = e0.elapsed_time(e1) / reps
lib = N.lib()
lib.thz_profile_enable(1)
for _ in range(reps):
    step()
torch.cuda.synchronize()
ms, cnt = (ctypes.c_float * 10)(), (ctypes.c_int32 * 10)()
lib.thz_profile_read(10, ms, cnt)
lib.thz_profile_enable(0)
names = ["row_fft", "column_pass", "row_ifft"]
print("%-14s n=%d fwd+adjoint %.3f ms" % (a.tag, a.n, plain), {nm: round(ms[i] / reps, 3) for i, nm in enumerate(names)}, "checksums %.6e %.6e" % chk)
