#!/usr/bin/env python3
"""A short, fixed sequence of hot-path steps for ncu (launch list / --set full captures).

    python tools/profile_step.py [--c 4] [--n 2048] [--steps 2] [--mode inregister|cached]

Same step as bench.py (STE DOE -> fused ASM forward -> g = y -> adjoint), smaller lambda batch so that a
full-metrics capture stays short.  Prints nothing a bench would report: numbers under a profiler are not bench values."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from quantizationawarethzdoe_b200 import ASM_prop, ElectricField, STEQuantizedDOELayer, functional as Fn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--c", type=int, default=4)
ap.add_argument("--n", type=int, default=2048)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--mode", default="inregister")
a = ap.parse_args()
for k, env in (("bc_chunk", "THZ_BC_CHUNK"), ("k2_cols", "THZ_K2_COLS"), ("lines", "THZ_LINES")):
    if os.environ.get(env):
        Fn.TUNE[k] = int(os.environ[env])
mm = 1e-3
dev = torch.device("cuda:0")
lams = [1 * mm * (1 + 0.01 * c) for c in range(a.c)]
torch.manual_seed(0)
x = torch.randn(1, a.c, a.n, a.n, dtype=torch.complex64, device=dev).requires_grad_(True)
doe = STEQuantizedDOELayer(dict(doe_size=[a.n, a.n], doe_dxy=0.5 * mm, doe_level=4, height_constraint_max=1 * mm,
                                tolerance=None, material=[2.66, 0.003]), {}, device=dev)
asm = ASM_prop(z_distance=0.1, device=dev, kernel_mode=a.mode)
asm.check_Zc = False
for _ in range(a.steps):
    y = asm(doe(ElectricField(x, wavelengths=lams, spacing=0.5 * mm, device=dev))).data
    gx, gw = torch.autograd.grad(y, (x, doe.weight_height_map), y.detach())
torch.cuda.synchronize()
print("ok", float(gw.abs().mean()))
