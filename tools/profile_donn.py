#!/usr/bin/env python3
"""Config 4 (3-layer DONN, 200 -> 400 pad) step for ncu / per-kernel event timing.  python tools/profile_donn.py [--b 256] [--steps 2] [--events]"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from quantizationawarethzdoe_b200 import ASM_prop, ElectricField, STEQuantizedDOELayer, _native as N  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--b", type=int, default=256)
ap.add_argument("--n", type=int, default=200)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--events", action="store_true")
ap.add_argument("--mode", default="auto")
a = ap.parse_args()
mm = 1e-3
dev = torch.device("cuda:0")
n, B, layers = a.n, a.b, 3
torch.manual_seed(0)
does = [STEQuantizedDOELayer(dict(doe_size=[n, n], doe_dxy=0.5 * mm, doe_level=4, height_constraint_max=1 * mm, tolerance=None,
                                  material=[2.66, 0.003]), {}, device=dev) for _ in range(layers)]
asms = [ASM_prop(z_distance=0.05, device=dev, kernel_mode=a.mode) for _ in range(layers)]
for p in asms:
    p.check_Zc = False
x = torch.randn(B, 1, n, n, dtype=torch.complex64, device=dev)
lam_t, sp_t = torch.tensor([1 * mm], device=dev), torch.tensor([0.5 * mm, 0.5 * mm], device=dev)


def step():
    f = ElectricField(x, wavelengths=lam_t, spacing=sp_t, device=dev)
    for d, p in zip(does, asms):
        f = p(d(f))
    y = f.data
    return torch.autograd.grad(y, [d.weight_height_map for d in does], y.detach())


for _ in range(a.steps):
    g = step()
torch.cuda.synchronize()
if a.events:
    lib = N.lib()
    lib.thz_profile_enable(1)
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms, cnt = (ctypes.c_float * 10)(), (ctypes.c_int32 * 10)()
    lib.thz_profile_read(10, ms, cnt)
    lib.thz_profile_enable(0)
    names = ["row_fft_fwd", "column_pass", "row_ifft", "fft2_col", "doe", "quant", "czt", "train", "czt_tc", "reserved"]
    print("B=%d step %.3f ms (with event overhead)" % (B, e0.elapsed_time(e1) / reps),
          {nm: (round(ms[i] / reps, 4), cnt[i] // reps) for i, nm in enumerate(names) if cnt[i]}, "modes", [p.resolved_kernel_mode for p in asms])
print("ok", float(g[0].abs().mean()))
