#!/usr/bin/env python3
"""The optimisation loop of the reference's experiment_four_focal_spots.ipynb (cells 2-8) on this package.

    Gaussian beam -> ASM 127 mm -> thin lens -> rect aperture   (evaluated once: the field in front of the DOE)
    loop:  quantized DOE -> ASM 200 mm -> normalize(|y|^2) -> MSE against a four-spot target -> Adam (lr 0.02)

usage: python examples/four_focal_spots.py [--iters 600] [--layer ste|gumbel|psq|full] [--graph]
--graph captures one iteration (forward, loss, backward, optimizer) in a CUDA graph and replays it; only for layers whose
forward does not depend on host-side state (ste, full)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from quantizationawarethzdoe_b200 import (ASM_prop, ApertureElement, FullPrecisionDOELayer, FusedAdam, Guassian_beam,  # noqa: E402
                                          PSQuantizedDOELayer, SoftGumbelQuantizedDOELayerv3, STEQuantizedDOELayer,
                                          Thin_LensElement, normalized_intensity_mse)

mm = 1e-3


def four_spot_target(n, dev):
    t = torch.zeros(1, 1, n, n, device=dev)
    for cy, cx in ((n // 4, n // 4), (n // 4, 3 * n // 4), (3 * n // 4, n // 4), (3 * n // 4, 3 * n // 4)):
        yy, xx = torch.meshgrid(torch.arange(n, device=dev), torch.arange(n, device=dev), indexing="ij")
        t[0, 0] += torch.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * 2.0 ** 2))
    return t / t.max()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=600)
    ap.add_argument("--layer", default="ste", choices=["ste", "gumbel", "psq", "full"])
    ap.add_argument("--graph", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    n, dxy, lam = 100, 1 * mm, 2.998e8 / 300e9
    doe_params = dict(doe_size=[n, n], doe_dxy=dxy, doe_level=4, look_up_table=None, num_unit=None, height_constraint_max=1 * mm,
                      tolerance=None if args.graph else 10e-6, material=[2.66, 0.03])
    layer = {"ste": lambda: STEQuantizedDOELayer(doe_params, {}, device=dev),
             "gumbel": lambda: SoftGumbelQuantizedDOELayerv3(doe_params, dict(c_s=100, tau_max=2.5, tau_min=1.5), device=dev),
             "psq": lambda: PSQuantizedDOELayer(doe_params, dict(c_s=300, tau_max=400, tau_min=1), device=dev),
             "full": lambda: FullPrecisionDOELayer(doe_params, device=dev)}[args.layer]()
    src = Guassian_beam(height=n, width=n, beam_waist_x=None, beam_waist_y=None, wavelengths=lam, spacing=dxy, device=dev)
    pre = [ASM_prop(z_distance=0.127, bandlimit_type="exact", padding_scale=2, device=dev), Thin_LensElement(0.127, device=dev),
           ApertureElement("rect", 0.08, device=dev)]
    asm = ASM_prop(z_distance=200 * mm, bandlimit_type="exact", padding_scale=2, device=dev)
    field = src()
    for m in pre:
        field = m(field)
    target = four_spot_target(n, dev)
    opt = FusedAdam(layer.parameters(), lr=0.02)

    def iteration(frac):
        out = asm(layer(field, frac))
        loss = normalized_intensity_mse(out.data, target)
        opt.zero_grad(set_to_none=False)
        loss.backward()
        opt.step()
        return loss

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        first = float(iteration(0.0))
        if args.graph:
            if args.layer not in ("ste", "full"):
                raise SystemExit("--graph needs a layer without host-side schedules (ste, full)")
            iteration(None)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if args.graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            loss = iteration(None)
        for _ in range(args.iters):
            g.replay()
    else:
        with torch.cuda.stream(side):
            for it in range(args.iters):
                loss = iteration(it / args.iters)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("layer %s: loss %.5f -> %.5f after %d iterations, %.2f ms/iteration (%s), levels used: %s mm" % (
        args.layer, first, float(loss), args.iters, dt / args.iters * 1e3, "CUDA graph" if args.graph else "eager",
        [round(float(v) / mm, 3) for v in layer.height_map.detach().unique()[:8]]))


if __name__ == "__main__":
    main()
