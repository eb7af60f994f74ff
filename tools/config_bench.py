#!/usr/bin/env python3
"""Timings of the other BASELINE.json configs (bench.py stays on the headline metric).

    python tools/config_bench.py donn    # config 4: 3 x (STE DOE + ASM 200 -> 400), batch 1024, fwd + bwd
    python tools/config_bench.py c2      # config 2: 1000 -> 2000 pad, 8-level DOE, one Adam-style step (fwd + adjoint)
    python tools/config_bench.py czt     # config 3: CZT 2048^2 -> 1024^2, 16 wavelengths
    python tools/config_bench.py iteration  # configs 1 / 2 as a whole optimisation iteration (loss + Adam), eager and CUDA graph
    python tools/config_bench.py zsweep  # depth sweep: the z setter is moved between forwards (experiment_extend_depth_of_focus)
Prints one JSON line per config with device-event timings (warm, 10 repetitions)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from quantizationawarethzdoe_b200 import ASM_prop, CZT_prop, ElectricField, STEQuantizedDOELayer  # noqa: E402

mm = 1e-3
dev = torch.device("cuda:0")


def timeit(fn, reps=10, warm=3):
    st = torch.cuda.current_stream()
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def donn():
    n, B, layers = 200, 1024, 3
    torch.manual_seed(0)
    does = [STEQuantizedDOELayer(dict(doe_size=[n, n], doe_dxy=0.5 * mm, doe_level=4, height_constraint_max=1 * mm, tolerance=None,
                                      material=[2.66, 0.003]), {}, device=dev) for _ in range(layers)]
    asms = [ASM_prop(z_distance=0.05, device=dev) for _ in range(layers)]
    for a in asms:
        a.check_Zc = False
    x = torch.randn(B, 1, n, n, dtype=torch.complex64, device=dev)
    lam_t, sp_t = torch.tensor([1 * mm], device=dev), torch.tensor([0.5 * mm, 0.5 * mm], device=dev)

    def step():
        f = ElectricField(x, wavelengths=lam_t, spacing=sp_t, device=dev)
        for d, a in zip(does, asms):
            f = a(d(f))
        y = f.data
        torch.autograd.grad(y, [d.weight_height_map for d in does], y.detach())

    ms = timeit(step)
    print(json.dumps({"config": "C4 DONN 3x(STE DOE + ASM 200->400), batch 1024, fwd+bwd", "ms_per_step": ms, "samples_per_s": B / ms * 1e3}))


def c2():
    n = 1000
    torch.manual_seed(0)
    doe = STEQuantizedDOELayer(dict(doe_size=[n, n], doe_dxy=0.5 * mm, doe_level=8, height_constraint_max=1 * mm, tolerance=None,
                                    material=[2.66, 0.003]), {}, device=dev)
    asm = ASM_prop(z_distance=0.1, device=dev)
    asm.check_Zc = False
    x = torch.randn(1, 1, n, n, dtype=torch.complex64, device=dev)
    lam_t, sp_t = torch.tensor([1 * mm], device=dev), torch.tensor([0.5 * mm, 0.5 * mm], device=dev)

    def step():
        y = asm(doe(ElectricField(x, wavelengths=lam_t, spacing=sp_t, device=dev))).data
        torch.autograd.grad(y, doe.weight_height_map, y.detach())

    # Everything runs on a side stream: the autograd leaf's AccumulateGrad node is bound to the stream of the first
    # forward, and a node bound to the legacy default stream cannot take part in a later graph capture.
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ms = timeit(step, reps=50)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):         # the same step captured once and replayed: no Python / launch overhead
        step()
    ms_graph = timeit(graph.replay, reps=200)
    print(json.dumps({"config": "C2 1000->2000 pad, 8-level STE DOE, fwd+adjoint", "ms_per_step": ms, "Msamples_per_s": 4e6 / ms / 1e3,
                      "cuda_graph_ms_per_step": ms_graph, "cuda_graph_Msamples_per_s": 4e6 / ms_graph / 1e3}))


def czt():
    H, M, C = 2048, 1024, 16
    torch.manual_seed(0)
    x = torch.randn(1, C, H, H, dtype=torch.complex64, device=dev)
    lams = [1 * mm * (1 + 0.01 * c) for c in range(C)]
    out = {}
    for impl in ("tc", "simt"):
        os.environ["THZ_CZT_IMPL"] = impl
        prop = CZT_prop(z_distance=0.5, device=dev)
        f = ElectricField(x, wavelengths=lams, spacing=0.5 * mm, device=dev)
        ms = timeit(lambda: prop(f, M, M, 0.1 * mm, 0.1 * mm), reps=5, warm=2)
        flops = 8.0 * (M * H * H + M * H * M) * C
        out[impl] = {"ms": ms, "complex_gemm_tflops": flops / ms / 1e9, "tensor_tflops_3xtf32": 3 * flops / ms / 1e9 if impl == "tc" else None}
    print(json.dumps({"config": "C3 CZT 2048^2 -> 1024^2, 16 wavelengths, forward", **out}))


def iteration():
    """One full optimisation iteration of the notebooks' loop (config 1: 512 -> 1024 pad, 4-level STE DOE): DOE -> ASM ->
    normalize(|y|^2) + MSE -> backward -> Adam, eager and as one replayed CUDA graph."""
    from quantizationawarethzdoe_b200 import FusedAdam, normalized_intensity_mse
    out = []
    for n, levels in ((512, 4), (1000, 8)):
        torch.manual_seed(0)
        doe = STEQuantizedDOELayer(dict(doe_size=[n, n], doe_dxy=0.5 * mm, doe_level=levels, height_constraint_max=1 * mm, tolerance=None,
                                        material=[2.66, 0.003]), {}, device=dev)
        asm = ASM_prop(z_distance=0.1, device=dev)
        asm.check_Zc = False
        opt = FusedAdam(doe.parameters(), lr=0.02)
        x = torch.randn(1, 1, n, n, dtype=torch.complex64, device=dev)
        target = torch.rand(1, 1, n, n, device=dev)
        lam_t, sp_t = torch.tensor([1 * mm], device=dev), torch.tensor([0.5 * mm, 0.5 * mm], device=dev)

        def it():
            y = asm(doe(ElectricField(x, wavelengths=lam_t, spacing=sp_t, device=dev))).data
            loss = normalized_intensity_mse(y, target)
            opt.zero_grad(set_to_none=False)
            loss.backward()
            opt.step()

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            ms = timeit(it, reps=50)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            it()
        ms_graph = timeit(graph.replay, reps=200)
        out.append({"n": n, "levels": levels, "eager_ms": ms, "cuda_graph_ms": ms_graph, "iterations_per_s_graph": 1e3 / ms_graph})
    print(json.dumps({"config": "full iteration: STE DOE + ASM + normalized-intensity MSE + backward + FusedAdam", "cases": out}))


def zsweep():
    """200 propagation distances through one ASM_prop (z setter, experiment_extend_depth_of_focus.ipynb cell 5): every
    forward rebuilds the transfer-function vectors on the host, so this measures host + device per z (wall clock)."""
    import time
    out = []
    for n in (100, 1000):
        torch.manual_seed(0)
        asm = ASM_prop(z_distance=0.1, device=dev, padding_scale=2 if n == 100 else None)
        asm.check_Zc = False
        x = torch.randn(1, 1, n, n, dtype=torch.complex64, device=dev)
        f = ElectricField(x, wavelengths=torch.tensor([1 * mm], device=dev), spacing=torch.tensor([1 * mm, 1 * mm], device=dev), device=dev)
        zs = [0.05 + 0.001 * i for i in range(200)]
        for z in zs[:5]:
            asm.z = z
            asm(f)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for z in zs:
            asm.z = z
            y = asm(f).data
        torch.cuda.synchronize()
        out.append({"n": n, "padded": asm.compute_padding(n, n)[0], "ms_per_z_wall": (time.perf_counter() - t0) / len(zs) * 1e3})
    print(json.dumps({"config": "depth sweep, 200 z through the z setter, forward only", "cases": out}))


if __name__ == "__main__":
    for w in (sys.argv[1:] or ["donn", "c2", "czt", "zsweep"]):
        {"donn": donn, "c2": c2, "czt": czt, "zsweep": zsweep, "iteration": iteration}[w]()
