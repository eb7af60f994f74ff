"""GPU parity tests proper: the sm_100a kernels, called through the reference-shaped modules and the
C ABI, against (a) the golden vectors produced by the reference and (b) the oracle on seeded inputs.

Tolerances (north_star): rel-L2 <= 1e-5 on complex64 fields and gradients; level indices bit-exact.
"""
import numpy as np
import pytest
import torch

from helpers import asm_case_kwargs, golden, golden_names, record, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-5
mm = 1e-3


@pytest.fixture(scope="module")
def dev():
    from quantizationawarethzdoe_b200 import _native as N
    N.lib()           # fail loudly if the native library is missing
    return torch.device("cuda:0")


def _asm(g, dev, mode):
    from quantizationawarethzdoe_b200 import ASM_prop
    a = ASM_prop(z_distance=g["z"], device=dev, kernel_mode=mode, **asm_case_kwargs(g))
    a.check_Zc = False
    return a


def _inregister_estimate(g):
    """Host-side estimate of how far the reference's own H (MKL sqrt) is from the correctly rounded one for a fixture."""
    from quantizationawarethzdoe_b200 import asm_host as AH
    kw = asm_case_kwargs(g)
    ps = AH.normalise_padding_scale(kw["padding_scale"], kw["do_padding"])
    _, _, Hp, Wp = AH.compute_padding(g["x"].shape[-2], g["x"].shape[-1], ps, kw["do_padding"])
    return AH.inregister_deviation_estimate(Hp, Wp, g["spacing"].float(), g["wavelengths"].float(), torch.tensor(g["z"]), True,
                                            kw["bandlimit_type"])


@pytest.mark.parametrize("mode", ["auto", "cached", "inregister"])
@pytest.mark.parametrize("name", golden_names("asm_"))
def test_asm_matches_reference_vectors(name, mode, dev):
    """'auto' is what ASM_prop ships as the default and 'cached' is its fallback: both are held to 1e-5 on every
    reference vector.  Forced 'inregister' is held to 1e-5 wherever the plan-time estimate says the reference's own
    non-IEEE sqrt leaves room for it (that is where 'auto' picks it); elsewhere the measured distance must be explained
    by that estimate (<= 2.5 x), and is recorded, not hidden."""
    from quantizationawarethzdoe_b200 import ElectricField, asm_host as AH
    g = golden(name)
    x = g["x"].to(dev).requires_grad_(True)
    f = ElectricField(x, wavelengths=g["wavelengths"].float(), spacing=g["spacing"].float(), device=dev)
    a = _asm(g, dev, mode)
    out = a(f)
    y = out.data
    assert y.shape == g["y"].shape and y.dtype == torch.complex64
    (gx,) = torch.autograd.grad(y, x, g["g"].to(dev))
    ey, eg = rel_l2(y.detach().cpu(), g["y"]), rel_l2(gx.cpu(), g["gx"])
    est = _inregister_estimate(g)
    record("asm_golden", fixture=name, mode=mode, resolved=a.resolved_kernel_mode, y=ey, gx=eg, inregister_estimate=est)
    tol = TOL
    if mode == "inregister" and est > AH.INREGISTER_BUDGET:
        tol = max(TOL, 2.5 * est)
    if mode == "auto":
        assert a.resolved_kernel_mode == ("inregister" if est <= AH.INREGISTER_BUDGET else "cached")
    assert ey < tol and eg < tol
    assert torch.equal(out.spacing.cpu(), f.spacing.cpu()) and torch.equal(out.wavelengths.cpu(), f.wavelengths.cpu())


@pytest.mark.parametrize("N_,C,scale", [(512, 1, None), (1000, 1, None), (256, 4, 2), (200, 3, None)])
def test_asm_matches_oracle_at_config_sizes(N_, C, scale, dev):
    """BASELINE configs[0] (512 -> 1024) and configs[1] (1000 -> 2000) grids, + pad 3x, + 400 (config 4 layer)."""
    from oracle import asm_oracle as AO
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField
    torch.manual_seed(0)
    lams = [1 * mm * (1 + 0.01 * c) for c in range(C)]
    x = torch.randn(1, C, N_, N_, dtype=torch.complex64)
    g = torch.randn(1, C, N_, N_, dtype=torch.complex64)
    xo = x.clone().requires_grad_(True)
    yo = AO.asm_forward(xo, lams, 0.5 * mm, 0.1, padding_scale=scale)
    (gxo,) = torch.autograd.grad(yo, xo, g)
    for mode, tol in (("auto", TOL), ("cached", TOL), ("inregister", TOL)):
        a = ASM_prop(z_distance=0.1, padding_scale=scale, device=dev, kernel_mode=mode)
        a.check_Zc = False
        xd = x.to(dev).requires_grad_(True)
        y = a(ElectricField(xd, wavelengths=lams, spacing=0.5 * mm, device=dev)).data
        (gx,) = torch.autograd.grad(y, xd, g.to(dev))
        ey, eg = rel_l2(y.detach().cpu(), yo.detach()), rel_l2(gx.cpu(), gxo)
        record("asm_config_sizes", N=N_, C=C, mode=mode, resolved=a.resolved_kernel_mode, y=ey, gx=eg)
        assert ey < tol and eg < tol, mode


@pytest.mark.parametrize("mode", ["auto", "cached"])
def test_metric_shape_matches_oracle(mode, dev):
    """The benchmarked configuration itself (bench.py: 4-level STE DOE fused into band-limited ASM, 2048^2 -> 4096^2 pad,
    z = 100 mm), ONE wavelength, against the oracle on the CPU: field, gradient wrt the input field, gradient wrt the
    DOE weights <= 1e-5 rel-L2; the level map bit-exact.  'auto' must resolve to the in-register H here -- it is what
    bench.py times."""
    from oracle import asm_oracle as AO, doe_oracle as DO
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField, STEQuantizedDOELayer
    n, lam, z, hmax, L = 2048, [1 * mm], 0.1, 1 * mm, 4
    torch.manual_seed(0)
    x = torch.randn(1, 1, n, n, dtype=torch.complex64)
    torch.manual_seed(1)
    doe = STEQuantizedDOELayer(dict(doe_size=[n, n], doe_dxy=0.5 * mm, doe_level=L, height_constraint_max=hmax, tolerance=None,
                                    material=[2.66, 0.003]), {}, device=dev)
    asm = ASM_prop(z_distance=z, device=dev, kernel_mode=mode)
    asm.check_Zc = False
    xd = x.to(dev).requires_grad_(True)
    y = asm(doe(ElectricField(xd, wavelengths=lam, spacing=0.5 * mm, device=dev))).data
    gx, gw = torch.autograd.grad(y, (xd, doe.weight_height_map), y.detach())        # the bench's loss gradient g = y
    torch.cuda.synchronize()
    if mode == "auto":
        assert asm.resolved_kernel_mode == "inregister", asm.inregister_estimate
    # oracle, same weights, fed the SAME upstream gradient (the GPU's y) so that gx / gw isolate the adjoint path
    xr = x.clone().requires_grad_(True)
    wr = doe.weight_height_map.detach().cpu().clone().requires_grad_(True)
    h = DO.ste_quantize(DO.sigmoid_height(wr[0, 0], hmax), DO.linear_lut(hmax, L))
    yo = AO.asm_forward(DO.modulate(xr, h, lam, 2.66, 0.003), lam, 0.5 * mm, z)
    gxo, gwo = torch.autograd.grad(yo, (xr, wr), y.detach().cpu())
    flips = int((doe.height_map.detach().cpu() != h.detach()).sum())
    ey, egx, egw = rel_l2(y.detach().cpu(), yo.detach()), rel_l2(gx.cpu(), gxo), rel_l2(gw.cpu(), gwo)
    record("metric_shape", mode=mode, resolved=asm.resolved_kernel_mode, y=ey, gx=egx, gw=egw, level_flips=flips,
           inregister_estimate=asm.inregister_estimate)
    assert flips == 0
    assert ey < TOL and egx < TOL and egw < TOL


def test_c3_czt_matches_oracle_on_the_tensor_cores(dev):
    """BASELINE config 3 geometry, one wavelength: 2048^2 -> 1024^2 zoomed chirp-z propagation against the oracle's FFT
    (Bluestein) form, forward and input gradient -- with proof that the tcgen05 kernel produced it: impl='tc' turns
    ineligibility into an error, and the launch counters show tcgen05 launches and no CUDA-core GEMM launch."""
    from oracle import czt_oracle as CO
    from quantizationawarethzdoe_b200 import CZT_prop, ElectricField, _native as N, functional as Fn
    torch.manual_seed(7)
    n, M, lam = 2048, 1024, [1 * mm]
    x = torch.randn(1, 1, n, n, dtype=torch.complex64)
    gy = torch.randn(1, 1, M, M, dtype=torch.complex64)
    xo = x.clone().requires_grad_(True)
    yo = CO.czt_forward(xo, torch.tensor(lam), torch.tensor([0.5 * mm, 0.5 * mm]), torch.tensor(0.5), M, M, 0.1 * mm, 0.1 * mm)
    (gxo,) = torch.autograd.grad(yo, xo, gy)
    lib = N.lib()
    tc0, simt0 = lib.thz_launch_count_class(8), lib.thz_launch_count_class(6)
    old = Fn.TUNE["czt_impl"]
    Fn.TUNE["czt_impl"] = "tc"
    try:
        czt = CZT_prop(z_distance=0.5, device=dev)
        xd = x.to(dev).requires_grad_(True)
        y = czt(ElectricField(xd, wavelengths=lam, spacing=0.5 * mm, device=dev), M, M, 0.1 * mm, 0.1 * mm).data
        (gx,) = torch.autograd.grad(y, xd, gy.to(dev))
        torch.cuda.synchronize()
    finally:
        Fn.TUNE["czt_impl"] = old
    assert lib.thz_launch_count_class(8) - tc0 >= 4 and lib.thz_launch_count_class(6) == simt0
    ey, eg = rel_l2(y.detach().cpu(), yo.detach()), rel_l2(gx.cpu(), gxo)
    record("c3_czt", y=ey, gx=eg, tc_launches=lib.thz_launch_count_class(8) - tc0)
    assert ey < TOL and eg < TOL


def test_czt_impl_tc_is_never_silently_replaced(dev):
    """impl='tc' on a call the tcgen05 kernel cannot serve (chirp filter shorter than 64) raises instead of switching."""
    from quantizationawarethzdoe_b200 import CZT_prop, ElectricField, functional as Fn
    x = torch.randn(1, 1, 12, 12, dtype=torch.complex64, device=dev)
    old = Fn.TUNE["czt_impl"]
    Fn.TUNE["czt_impl"] = "tc"
    try:
        with pytest.raises(NotImplementedError, match="not eligible"):
            CZT_prop(z_distance=0.5, device=dev)(ElectricField(x, wavelengths=[1 * mm], spacing=0.5 * mm, device=dev), 8, 8, 0.1 * mm, 0.1 * mm)
    finally:
        Fn.TUNE["czt_impl"] = old


def test_c4_three_layer_donn_chain_matches_oracle(dev):
    """BASELINE config 4 layer stack at a small batch: 3 x (4-level STE DOE -> ASM 200 -> 400 pad), batch 8, loss
    gradient g = y: output field and the weight gradients of ALL THREE layers against autograd through the oracle."""
    from oracle import asm_oracle as AO, doe_oracle as DO
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField, STEQuantizedDOELayer
    n, B, lam, z, hmax, L = 200, 8, [1 * mm], 0.05, 1 * mm, 4
    torch.manual_seed(3)
    x = torch.randn(B, 1, n, n, dtype=torch.complex64)
    g = torch.randn(B, 1, n, n, dtype=torch.complex64)
    does, asms = [], []
    for i in range(3):
        torch.manual_seed(20 + i)
        does.append(STEQuantizedDOELayer(dict(doe_size=[n, n], doe_dxy=0.5 * mm, doe_level=L, height_constraint_max=hmax,
                                              tolerance=None, material=[2.66, 0.003]), {}, device=dev))
        a = ASM_prop(z_distance=z, device=dev)
        a.check_Zc = False
        asms.append(a)
    f = ElectricField(x.to(dev), wavelengths=lam, spacing=0.5 * mm, device=dev)
    for d, a in zip(does, asms):
        f = a(d(f))
    y = f.data
    gws = torch.autograd.grad(y, [d.weight_height_map for d in does], g.to(dev))
    ws = [d.weight_height_map.detach().cpu().clone().requires_grad_(True) for d in does]
    u = x
    flips = 0
    for d, w in zip(does, ws):
        h = DO.ste_quantize(DO.sigmoid_height(w[0, 0], hmax), DO.linear_lut(hmax, L))
        flips += int((d.height_map.detach().cpu() != h.detach()).sum())
        u = AO.asm_forward(DO.modulate(u, h, lam, 2.66, 0.003), lam, 0.5 * mm, z)
    gwo = torch.autograd.grad(u, ws, g)
    errs = [rel_l2(a_.cpu(), b_) for a_, b_ in zip(gws, gwo)]
    ey = rel_l2(y.detach().cpu(), u.detach())
    record("c4_donn_chain", y=ey, gw1=errs[0], gw2=errs[1], gw3=errs[2], level_flips=flips, modes=[a.resolved_kernel_mode for a in asms])
    assert flips == 0 and ey < TOL and max(errs) < TOL


def test_workspace_survives_growth_after_graph_capture(dev):
    """A captured step keeps the address of its scratch buffer; a later, larger call on the same stream must not free
    that buffer: replay after growing the workspace, compare with eager."""
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField
    n = 256
    torch.manual_seed(0)
    x = torch.randn(1, 1, n, n, dtype=torch.complex64, device=dev)
    lam_t, sp_t = torch.tensor([1 * mm], device=dev), torch.tensor([0.5 * mm, 0.5 * mm], device=dev)
    asm = ASM_prop(z_distance=0.1, device=dev)
    asm.check_Zc = False
    F = lambda: asm(ElectricField(x, wavelengths=lam_t, spacing=sp_t, device=dev)).data
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            eager = F().clone()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            out = F()
        # a much larger call on the SAME stream regrows the cached workspace ...
        big = ASM_prop(z_distance=0.1, device=dev)
        big.check_Zc = False
        xb = torch.randn(2, 1, 1024, 1024, dtype=torch.complex64, device=dev)
        big(ElectricField(xb, wavelengths=lam_t, spacing=sp_t, device=dev))
        junk = [torch.full((1 << 20,), 7.0, device=dev) for _ in range(8)]     # ... and the allocator gets a chance to reuse freed blocks
        g.replay()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    assert torch.equal(out, eager)
    del junk


def test_full_size_properties(dev):
    """Metric shape (2048 -> 4096 pad): size-independent properties instead of a CPU comparison:
    adjoint identity <A x, y> = <x, A^H y>, linearity, and energy non-increase (|H| <= 1)."""
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField
    N_, C = 2048, 2
    lams = [1 * mm, 1.01 * mm]
    torch.manual_seed(3)
    a = ASM_prop(z_distance=0.1, device=dev)
    a.check_Zc = False
    x1 = torch.randn(1, C, N_, N_, dtype=torch.complex64, device=dev)
    x2 = torch.randn(1, C, N_, N_, dtype=torch.complex64, device=dev)
    yv = torch.randn(1, C, N_, N_, dtype=torch.complex64, device=dev)
    F = lambda t: a(ElectricField(t, wavelengths=lams, spacing=0.5 * mm, device=dev)).data
    xg = x1.clone().requires_grad_(True)
    Ax = F(xg)
    (AHy,) = torch.autograd.grad(Ax, xg, yv)
    lhs = torch.sum(Ax.detach().conj() * yv)
    rhs = torch.sum(x1.conj() * AHy)
    assert abs(lhs - rhs) / abs(lhs) < 1e-5
    lin = F(x1 + 2 * x2) - (Ax.detach() + 2 * F(x2))
    assert lin.norm() / Ax.detach().norm() < 1e-5
    assert Ax.detach().norm() <= x1.norm() * (1 + 1e-5)


@pytest.mark.parametrize("shape,modes", [((1024, 1024), ("inregister", "cached")), ((2048, 2048), ("inregister", "cached")),
                                         ((4096, 4096), ("inregister",)), ((2048, 1024), ("inregister",)), ((1024, 4096), ("inregister",))],
                         ids=lambda v: "x".join(map(str, v)) if isinstance(v[0], int) else None)
def test_tma_stores_are_bit_identical(dev, monkeypatch, shape, modes):
    """Transform lengths 2048 / 4096 / 8192 under centred 2x padding (square and non-square canvases): the row-FFT kernel and
    the column kernel hand their output to the TMA (cp.async.bulk.tensor stores: thz_p2_k1t into the blocked intermediate,
    thz_p2_k2ft into the row-major one), and the row-iFFT kernel stages its rows with one bulk copy per field from a
    column-permuted intermediate (thz_p2_k3t; the permuting store is exercised with the column kernel's TMA and plain stores).
    Launch class 9 counts the TMA variants.  Forward field and adjoint must equal the plain kernels (THZ_NO_K1TMA=1,
    THZ_NO_K2TMA=1, THZ_NO_K3TMA=1) bit for bit, for both transfer-function modes, and the TMA kernels must be the ones that ran."""
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField, _native as N
    lib = N.lib()
    h, w = shape
    torch.manual_seed(11)
    x = torch.randn(1, 1, h, w, dtype=torch.complex64, device=dev)
    gy = torch.randn(1, 1, h, w, dtype=torch.complex64, device=dev)
    served = (2048, 4096, 8192)
    k1 = 2 if 2 * w in served else 0                                      # row lines
    k2 = 2 if 2 * h in served else 0                                      # column lines
    k3 = 2 if (2 * h in (4096, 8192) and 2 * w in served) else 0          # needs the column kernel's 2-column tiles
    switches = ("THZ_NO_K1TMA", "THZ_NO_K2TMA", "THZ_NO_K3TMA")
    variants = (("all", (), k1 + k2 + k3), ("k1_only", ("THZ_NO_K2TMA", "THZ_NO_K3TMA"), k1), ("k2_only", ("THZ_NO_K1TMA", "THZ_NO_K3TMA"), k2),
                ("k3_plain_k2", ("THZ_NO_K1TMA", "THZ_NO_K2TMA"), k3), ("plain", switches, 0))
    assert k1 + k2 + k3 >= 4
    for mode in modes:
        out = {}
        for name, off, expect in variants:
            for v in switches:
                if v in off:
                    monkeypatch.setenv(v, "1")
                else:
                    monkeypatch.delenv(v, raising=False)
            a = ASM_prop(z_distance=0.1, device=dev, kernel_mode=mode)
            a.check_Zc = False
            xg = x.clone().requires_grad_(True)
            c9, c1 = lib.thz_launch_count_class(9), lib.thz_launch_count_class(1)
            y = a(ElectricField(xg, wavelengths=[1 * mm], spacing=0.5 * mm, device=dev)).data
            (gx,) = torch.autograd.grad(y, xg, gy)
            torch.cuda.synchronize()
            d9, d1 = lib.thz_launch_count_class(9) - c9, lib.thz_launch_count_class(1) - c1
            assert d1 == 2 and d9 == expect, (shape, mode, name, d1, d9)
            out[name] = (y.detach().clone(), gx.clone())
        for name in ("all", "k1_only", "k2_only", "k3_plain_k2"):
            assert torch.equal(out[name][0], out["plain"][0]) and torch.equal(out[name][1], out["plain"][1]), (shape, mode, name)
        assert float(out["all"][0].abs().max()) > 0
        del out
    for v in switches:
        monkeypatch.delenv(v, raising=False)


def _doe_layer(name, g, dev):
    import quantizationawarethzdoe_b200 as Q
    N_ = g["x"].shape[-1]
    dp = dict(doe_size=[N_, N_], doe_dxy=0.5 * mm, doe_level=int(g["levels"]), height_constraint_max=g["hmax"], tolerance=None,
              material=g["material"].tolist())
    op = dict(c_s=300, tau_max=5.5, tau_min=2.0)
    kw = {}
    if name == "doe_ste":
        layer, pname = Q.STEQuantizedDOELayer(dp, op, device=dev), "weight_height_map"
    elif name == "doe_fullprecision":
        layer, pname = Q.FullPrecisionDOELayer(dp, device=dev), "weight_height_map"
    elif name == "doe_psq":
        layer, pname = Q.PSQuantizedDOELayer(dp, dict(tau_max=400, tau_min=1), device=dev), "weight_height_map"
    elif name.startswith("doe_gumbel_v3"):
        layer, pname = Q.SoftGumbelQuantizedDOELayerv3(dp, op, device=dev), "weight_init_phase"
    elif name == "doe_gumbel_v2":
        layer, pname = Q.SoftGumbelQuantizedDOELayerv2(dp, op, device=dev), "weight_init_phase"
    elif name == "doe_gumbel_v1":
        layer, pname = Q.SoftGumbelQuantizedDOELayer(dp, op, device=dev), "init_phase"
    elif name == "doe_gumbel_naive":
        layer, pname = Q.NaiveGumbelQuantizedDOELayer(dp, op, device=dev), "weight_height_map"
    else:
        raise KeyError(name)
    if "iter_frac" in g:
        kw["iter_frac"] = g["iter_frac"]
    if "noise" in g:
        layer.gumbel_noise = g["noise"].to(dev)
    with torch.no_grad():
        getattr(layer, pname).copy_(g["w"].to(dev))
    return layer, getattr(layer, pname), kw


def _fp32_gradient_budget(name, g):
    """Tolerance for a gradient that runs through sigmoid' x softmax chains (PSQ, score-Gumbel): the reference's own fp32
    autograd result sits up to 1.5e-5 from the float64 evaluation of the same formulas (measured here, on the CPU, with the
    oracle in float64), so the bar is max(1e-5, 2 x that distance) -- 1e-5 wherever fp32 itself is that good."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from test_oracle_golden import _height_for
    from oracle import asm_oracle as AO, doe_oracle as DO
    if name in ("doe_gumbel_v1",) or g["gw"].abs().max() == 0:       # v1 has no oracle restatement; measured 3e-6
        return TOL
    g64 = dict(g)
    g64["w"] = g["w"].double()
    for k in ("noise", "lut"):
        if k in g:
            g64[k] = g[k].double()
    w, h = _height_for(name, g64)
    u = DO.modulate(g["x"].to(torch.complex128), h, g["wavelengths"].float(), g["material"][0].float(), g["material"][1].float())
    y = AO.asm_forward(u, g["wavelengths"].float(), g["spacing"].float(), g["z"])
    (gw64,) = torch.autograd.grad(y, w, g["g"].to(y.dtype))
    return max(TOL, 2.0 * rel_l2(g["gw"].double(), gw64))


@pytest.mark.parametrize("name", [n for n in golden_names("doe_") if n != "doe_fix_edoe4"])
def test_doe_layers_match_reference_vectors(name, dev):
    """Every DOE layer: height map (levels bit-exact up to sigmoid ulps, counted), modulated field,
    fused DOE->ASM output and the gradient wrt the layer's parameter, vs the reference's autograd."""
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField
    g = golden(name)
    layer, param, kw = _doe_layer(name, g, dev)
    lam, sp = g["wavelengths"].float(), g["spacing"].float()
    f = ElectricField(g["x"].to(dev), wavelengths=lam, spacing=sp, device=dev)
    asm = ASM_prop(z_distance=g["z"], device=dev, kernel_mode="cached")
    asm.check_Zc = False
    # (1) fused path: DOE output consumed un-materialised by ASM_prop
    u = layer(f, **kw)
    assert u._data is None, "modulation should be deferred until someone reads .data"
    y = asm(u).data
    (gw,) = torch.autograd.grad(y, param, g["g"].to(dev))
    hm = layer.height_map.detach().cpu()
    mism = (hm != g["height_map"])
    continuous = name in ("doe_fullprecision", "doe_psq", "doe_gumbel_v3_02", "doe_gumbel_v3_05")    # no level selection / blended
    # level selection is integer work: every discrete map must equal the reference's bit for bit
    if continuous:
        assert rel_l2(hm, g["height_map"]) < 1e-6
    else:
        assert int(mism.sum()) == 0, "level flips: %d" % int(mism.sum())
    ey = rel_l2(y.detach().cpu(), g["y"])
    egw = rel_l2(gw.cpu(), g["gw"]) if g["gw"].abs().max() > 0 else 0.0
    budget = _fp32_gradient_budget(name, g)
    record("doe_golden", fixture=name, level_flips=(-1 if continuous else int(mism.sum())), y=ey, gw=egw, gw_budget=budget)
    assert ey < TOL
    assert egw < budget
    # (2) stand-alone modulation kernel (materialised .data) and its own backward
    u2 = layer(f, **kw)
    ud = u2.data
    assert rel_l2(ud.detach().cpu(), g["u"]) < TOL
    (gw2,) = torch.autograd.grad(ud, param, g["g"].to(dev))
    if g["gw_modulate_only"].abs().max() > 0:
        e2 = rel_l2(gw2.cpu(), g["gw_modulate_only"])
        record("doe_golden_modulate_only", fixture=name, gw=e2, gw_budget=budget)
        assert e2 < budget


def test_fix_doe_element_on_reference_height_map(dev):
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField, FixDOEElement
    g = golden("doe_fix_edoe4")
    fix = FixDOEElement(height_map=g["height_map"].numpy(), tolerance=None, material=g["material"].tolist(), device=dev)
    f = ElectricField(g["x"].to(dev), wavelengths=g["wavelengths"].float(), spacing=g["spacing"].float(), device=dev)
    asm = ASM_prop(z_distance=g["z"], padding_scale=2, device=dev, kernel_mode="cached")
    asm.check_Zc = False
    u = fix(f)
    y = asm(u).data
    (gh,) = torch.autograd.grad(y, fix.height_map, g["g"].to(dev))
    assert rel_l2(y.detach().cpu(), g["y"]) < TOL
    assert rel_l2(gh.cpu(), g["gh"]) < TOL
    assert rel_l2(fix(f).data.detach().cpu(), g["u"]) < TOL
    p = FixDOEElement.phase_shift_according_to_height(g["height_map"].to(dev), g["wavelengths"].float(), fix.epsilon, fix.tand)
    assert rel_l2((g["x"].to(dev) * p[None]).cpu(), g["u"]) < TOL


def test_level_indices_bit_exact_given_identical_inputs(dev):
    """Indices are integer work: bit-exact given the same fp32 height map (SURVEY 8 a-8 contract)."""
    from quantizationawarethzdoe_b200 import functional as Fn
    from quantizationawarethzdoe_b200.Components import quantization as QZ
    from quantizationawarethzdoe_b200.Components.discrete_doe import DiscreteDOE
    g = golden("quant_ste")
    q, idx = Fn.SteQuantizeFn.apply(g["h"].to(dev), g["lut"].to(dev))
    assert torch.equal(idx.cpu().long(), g["idx"]) and torch.equal(q.cpu(), g["q"])
    x = g["kat_x"].to(dev).requires_grad_(True)
    qk, _ = Fn.SteQuantizeFn.apply(x, g["kat_lut"].to(dev))
    assert qk.tolist() == [0.0, 0.5, 0.5, 1.0]
    (gr,) = torch.autograd.grad(qk.sum(), x)
    assert torch.equal(gr.cpu(), torch.ones(4))
    # ties and maximum sizes: every midpoint between levels, and a 4096^2 map against torch on the device
    lut = torch.linspace(0, 1e-3, 9)[:-1]
    mids = (lut[:-1] + lut[1:]) / 2
    edge = torch.cat([mids, lut, torch.tensor([-1.0, 2.0])])
    _, ie = Fn.SteQuantizeFn.apply(edge.to(dev), lut.to(dev))
    assert torch.equal(ie.cpu().long(), torch.argmin(torch.abs(edge.unsqueeze(-1) - lut), dim=-1))
    big = torch.rand(4096, 4096, device=dev) * 1e-3
    _, ib = Fn.SteQuantizeFn.apply(big, lut.to(dev))
    ref = torch.argmin(torch.abs(big.unsqueeze(-1) - lut.to(dev)), dim=-1)
    assert torch.equal(ib.long(), ref)
    # empty input
    _, i0 = Fn.SteQuantizeFn.apply(torch.empty(0, device=dev), lut.to(dev))
    assert i0.numel() == 0
    n = golden("quant_nn")
    DiscreteDOE.set_lut(n["lut"])
    for kind, fn in (("nn", QZ.nns), ("nn_poly", QZ.nns_poly), ("nn_sigmoid", QZ.nns_sigmoid)):
        xx = n["x"].to(dev).requires_grad_(True)
        qq = fn(xx, float(n["s"]))
        assert torch.equal(qq.detach().cpu(), n["q_" + kind])
        (gr,) = torch.autograd.grad(qq, xx, torch.ones_like(qq))
        assert rel_l2(gr.cpu(), n["grad_" + kind]) < 1e-5
    from quantizationawarethzdoe_b200.utils.Helper_Functions import nearest_idx
    assert nearest_idx(n["kat_x"].to(dev), n["mid"]).tolist() == [0, 0, 0, 1, 1, 1, 2, 3, 0, 0, 0, 0]


@pytest.mark.parametrize("n,C,levels", [(512, 3, 4), (2048, 2, 4), (200, 1, 8), (1000, 1, 8)])
def test_level_tables_are_bit_identical_to_per_pixel_transmission(n, C, levels, dev):
    """Level selection fused into the propagation (north_star (3)): an STE layer hands ASM_prop its int32 level map and the
    [C, L] transmissions of the levels; the static row kernels look p up instead of evaluating exp / sincos per pixel and
    wavelength.  Same numbers, bit for bit, for the field and the input gradient; the weight gradient to summation order."""
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField, STEQuantizedDOELayer
    lams = [1 * mm * (1 + 0.02 * c) for c in range(C)]
    torch.manual_seed(0)
    x = torch.randn(2, C, n, n, dtype=torch.complex64, device=dev)
    torch.manual_seed(1)
    doe = STEQuantizedDOELayer(dict(doe_size=[n, n], doe_dxy=0.5 * mm, doe_level=levels, height_constraint_max=1 * mm, tolerance=None,
                                    material=[2.66, 0.003]), {}, device=dev)
    asm = ASM_prop(z_distance=0.1, device=dev)
    asm.check_Zc = False
    res = []
    for use_tables in (True, False):
        xd = x.clone().requires_grad_(True)
        f = doe(ElectricField(xd, wavelengths=lams, spacing=0.5 * mm, device=dev))
        lev = f._deferred.levels
        assert lev is not None and lev[0].dtype == torch.int32 and tuple(lev[1].shape) == (C, levels)
        assert torch.equal(doe.lut[lev[0].long()], doe.height_map)            # the map IS lut[level map]
        if not use_tables:
            f._deferred.levels = None
        y = asm(f).data
        gx, gw = torch.autograd.grad(y, (xd, doe.weight_height_map), y.detach())
        res.append((y.detach(), gx, gw))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    # grad_height is accumulated with atomics when the fields of a row are split over several CTAs: same terms, any order
    assert rel_l2(res[0][2], res[1][2]) < 1e-6


@pytest.mark.parametrize("name", ["gumbel_hard", "gumbel_soft", "plain_hard", "plain_soft"])
def test_softmax_quantization_matches_reference(name, dev):
    """Quantization(method='*softmax*' | '*gumbel*') -> SoftmaxBasedQuantization + score_thickness
    (Components/quantization.py:36-55, 128-161, 164-207): forward vs the reference's own output, gradient vs the oracle's
    out-of-place restatement (the reference's backward raises), budgeted by the fp32-vs-float64 distance of that gradient."""
    from quantizationawarethzdoe_b200.Components.quantization import Quantization
    g = golden("quant_softmax")
    method = "gumbel_softmax" if "gumbel" in name else "softmax"
    qz = Quantization(method=method, max_thickness=g["hmax"], num_bits=2, dev=dev, tau_min=g["tau_min"], tau_max=g["tau_max"], c=g["c"])
    if "gumbel" in name:
        qz.quan_fn.gumbel_noise = g["noise_" + name].to(dev)
    t = g["thickness"].to(dev).requires_grad_(True)
    q = qz(t, iter_frac=g["frac_" + name], hard="hard" in name)
    (gt,) = torch.autograd.grad(q, t, g["gq"].to(dev))
    ref_g, g64 = g["gt_" + name], g["gt64_" + name]
    budget = max(TOL, 2.0 * rel_l2(ref_g.double(), g64))
    eq, eg = rel_l2(q.detach().cpu(), g["q_" + name]), rel_l2(gt.cpu(), ref_g)
    record("quant_softmax", case=name, q=eq, gt=eg, gt_budget=budget, level_flips=int((q.detach().cpu() != g["q_" + name]).sum()) if "hard" in name else -1)
    assert q.shape == g["q_" + name].shape
    assert eq < 1e-6 and eg < budget
    if "hard" in name:      # the selected LEVEL is exact; the value carries the ulps of (onehot + y) - y, which depend on libm's exp
        lut = g["lut"][:-1]
        lv = lambda v: torch.argmin((v.reshape(-1, 1) - lut[None, :]).abs(), dim=1)
        assert torch.equal(lv(q.detach().cpu()), lv(g["q_" + name]))


def test_score_thickness_and_per_instance_luts(dev):
    from quantizationawarethzdoe_b200.Components.quantization import Quantization, score_thickness
    torch.manual_seed(3)
    t = (torch.rand(2, 1, 33, 17) * 1.2e-3 - 1e-4)
    lut = torch.linspace(0, 1e-3, 5)[:-1]
    diff = t - lut.reshape(1, -1, 1, 1)
    diff = diff / torch.max(torch.abs(diff))
    ref = torch.sigmoid(2.5 * diff) * (1 - torch.sigmoid(2.5 * diff)) * 4
    assert rel_l2(score_thickness(t.to(dev), lut.to(dev), 2.5).cpu(), ref) < 2e-6
    # two quantizers with different LUTs coexist (the LUT used to be process-global)
    a = Quantization(method="nn", max_thickness=1.0, num_bits=2, dev=dev)
    b = Quantization(method="nn", max_thickness=2.0, num_bits=2, dev=dev)
    x = torch.tensor([0.3, 0.9], device=dev).reshape(1, 1, 1, 2)
    assert a(x, iter_frac=0.1).reshape(-1).tolist() == [0.25, 0.0] and b(x, iter_frac=0.1).reshape(-1).tolist() == [0.5, 1.0]


@pytest.mark.parametrize("H,W", [(64, 64), (60, 100), (1000, 1000), (2048, 1024)])
def test_fft2_and_shifted_helpers(H, W, dev):
    from quantizationawarethzdoe_b200 import functional as Fn
    from quantizationawarethzdoe_b200.utils.Helper_Functions import ft2, ift2
    from oracle.asm_oracle import shifted_fft2
    torch.manual_seed(2)
    x = torch.randn(2, 1, H, W, dtype=torch.complex64)
    y = Fn.fft2_c2c(x.to(dev))
    assert rel_l2(y.cpu(), torch.fft.fft2(x)) < 2e-6
    assert rel_l2(Fn.fft2_c2c(y, inverse=True).cpu(), x) < 2e-6          # round trip
    assert rel_l2(ft2(x.to(dev)).cpu(), shifted_fft2(x)) < 2e-6
    assert rel_l2(ift2(x.to(dev)).cpu(), shifted_fft2(x, inverse=True)) < 2e-6


def test_errors_cross_the_abi_as_python_exceptions(dev):
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField
    a = ASM_prop(z_distance=0.1, device=dev)
    a.check_Zc = False
    nopad = ASM_prop(z_distance=0.1, do_padding=False, device=dev)
    nopad.check_Zc = False
    # 13 x 13 (no padding) has no radix plan: served through the chirp-z path now, like any size the reference's torch.fft takes
    y13 = nopad(ElectricField(torch.ones(1, 1, 13, 13, dtype=torch.complex64, device=dev), 1e-3, 1e-3, device=dev)).data
    assert y13.shape == (1, 1, 13, 13) and "chirp-z" in nopad.resolved_kernel_mode
    with pytest.raises(TypeError, match="complex64"):
        a(ElectricField(torch.zeros(1, 1, 16, 16, dtype=torch.complex128, device=dev), 1e-3, 1e-3, device=dev))


@pytest.mark.parametrize("N_,C,scale", [(101, 2, None),       # 202 = 2 x 101
                                        (134, 1, None),       # 268 = 4 x 67
                                        (67, 1, 2),           # padding_scale 2: 201 = 3 x 67
                                        (1031, 1, None)])     # 2062 = 2 x 1031: 4096-point chirp convolutions
def test_any_grid_size_matches_oracle(N_, C, scale, dev):
    """Edge lengths with a prime factor > 7 (VERDICT r1 missing #3): ASM_prop accepts what torch.fft accepts -- chirp-z on the
    fused pipeline (bluestein.py) -- forward, input gradient and the DOE weight gradient through a materialised modulation."""
    from oracle import asm_oracle as AO, doe_oracle as DO
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField, STEQuantizedDOELayer
    lams = [1 * mm * (1 + 0.03 * c) for c in range(C)]
    torch.manual_seed(0)
    x = torch.randn(1, C, N_, N_, dtype=torch.complex64)
    g = torch.randn(1, C, N_, N_, dtype=torch.complex64)
    torch.manual_seed(1)
    doe = STEQuantizedDOELayer(dict(doe_size=[N_, N_], doe_dxy=0.5 * mm, doe_level=4, height_constraint_max=1 * mm, tolerance=None,
                                    material=[2.66, 0.003]), {}, device=dev)
    asm = ASM_prop(z_distance=0.1, padding_scale=scale, device=dev)
    asm.check_Zc = False
    xd = x.to(dev).requires_grad_(True)
    y = asm(doe(ElectricField(xd, wavelengths=lams, spacing=0.5 * mm, device=dev))).data
    assert "chirp-z" in asm.resolved_kernel_mode
    gx, gw = torch.autograd.grad(y, (xd, doe.weight_height_map), g.to(dev))
    xo = x.clone().requires_grad_(True)
    wo = doe.weight_height_map.detach().cpu().clone().requires_grad_(True)
    h = DO.ste_quantize(DO.sigmoid_height(wo[0, 0], 1 * mm), DO.linear_lut(1 * mm, 4))
    yo = AO.asm_forward(DO.modulate(xo, h, lams, 2.66, 0.003), lams, 0.5 * mm, 0.1, padding_scale=scale)
    gxo, gwo = torch.autograd.grad(yo, (xo, wo), g)
    ey, egx, egw = rel_l2(y.detach().cpu(), yo.detach()), rel_l2(gx.cpu(), gxo), rel_l2(gw.cpu(), gwo)
    record("any_grid_size", N=N_, C=C, padded=asm.compute_padding(N_, N_)[0], y=ey, gx=egx, gw=egw)
    assert ey < TOL and egx < TOL and egw < TOL


@pytest.mark.parametrize("H,W,C,split", [(24, 16384, 2, "1 x 2"),        # 48 x 32768 canvas: the long edge in two
                                         (16, 32768, 1, "1 x 4"),        # 32 x 65536: in four
                                         (10000, 12, 1, "2 x 1")])       # 20000 x 24: mixed-radix 10000-point sub-lines
@pytest.mark.parametrize("mode", ["auto", "cached"])
def test_long_lines_match_oracle(H, W, C, split, mode, dev):
    """Canvas edges above 16384 points (VERDICT r1 missing #3, second half): one outer decimation step around the fused pipeline
    (longline.py) -- ASM_prop forward, input gradient and DOE weight gradient (materialised modulation) against the oracle."""
    from oracle import asm_oracle as AO, doe_oracle as DO
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField, STEQuantizedDOELayer
    lams = [1 * mm * (1 + 0.03 * c) for c in range(C)]
    torch.manual_seed(0)
    x = torch.randn(1, C, H, W, dtype=torch.complex64)
    g = torch.randn(1, C, H, W, dtype=torch.complex64)
    torch.manual_seed(1)
    doe = STEQuantizedDOELayer(dict(doe_size=[H, W], doe_dxy=0.5 * mm, doe_level=4, height_constraint_max=1 * mm, tolerance=None,
                                    material=[2.66, 0.003]), {}, device=dev)
    asm = ASM_prop(z_distance=0.05, kernel_mode=mode, device=dev)
    asm.check_Zc = False
    xd = x.to(dev).requires_grad_(True)
    y = asm(doe(ElectricField(xd, wavelengths=lams, spacing=0.5 * mm, device=dev))).data
    assert asm.resolved_kernel_mode.endswith("(split %s)" % split)
    gx, gw = torch.autograd.grad(y, (xd, doe.weight_height_map), g.to(dev))
    xo = x.clone().requires_grad_(True)
    wo = doe.weight_height_map.detach().cpu().clone().requires_grad_(True)
    h = DO.ste_quantize(DO.sigmoid_height(wo[0, 0], 1 * mm), DO.linear_lut(1 * mm, 4))
    yo = AO.asm_forward(DO.modulate(xo, h, lams, 2.66, 0.003), lams, 0.5 * mm, 0.05)
    gxo, gwo = torch.autograd.grad(yo, (xo, wo), g)
    ey, egx, egw = rel_l2(y.detach().cpu(), yo.detach()), rel_l2(gx.cpu(), gxo), rel_l2(gw.cpu(), gwo)
    record("long_lines", H=H, W=W, C=C, mode=asm.resolved_kernel_mode, y=ey, gx=egx, gw=egw)
    assert ey < TOL and egx < TOL and egw < TOL


@pytest.mark.parametrize("mode", ["inregister", "cached"])
def test_forced_split_equals_the_direct_pipeline(mode, dev, monkeypatch):
    """Same operator three ways: a 1024 x 2048 canvas run directly, as 2 x 4 (512-point sub-lines) and as 1 x 2 -- the
    split path exercised on the static kernels at sizes where the direct answer exists."""
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField, functional as Fn
    lams = [1 * mm, 1.05 * mm]
    torch.manual_seed(4)
    x = torch.randn(2, 2, 512, 1024, dtype=torch.complex64, device=dev)
    g = torch.randn(2, 2, 512, 1024, dtype=torch.complex64, device=dev)
    res = []
    for ml, want in ((16384, None), (512, "2 x 4"), (1024, "1 x 2")):
        monkeypatch.setitem(Fn.TUNE, "max_line", ml)
        asm = ASM_prop(z_distance=0.08, kernel_mode=mode, device=dev)
        asm.check_Zc = False
        xd = x.clone().requires_grad_(True)
        y = asm(ElectricField(xd, wavelengths=lams, spacing=0.5 * mm, device=dev)).data
        (gx,) = torch.autograd.grad(y, xd, g)
        assert asm.resolved_kernel_mode == (mode if want is None else "%s (split %s)" % (mode, want))
        res.append((y.detach(), gx))
    for y, gx in res[1:]:
        ey, egx = rel_l2(y, res[0][0]), rel_l2(gx, res[0][1])
        record("forced_split", mode=mode, y=ey, gx=egx)
        assert ey < 1e-6 and egx < 1e-6


@pytest.mark.parametrize("H,W", [(6, 32768), (20000, 10), (3, 10007)])      # 10007 is prime: 32768-point chirp convolution, split
def test_fft2_of_long_lines(H, W, dev):
    from quantizationawarethzdoe_b200 import functional as Fn
    torch.manual_seed(2)
    x = torch.randn(2, 1, H, W, dtype=torch.complex64)
    assert rel_l2(Fn.fft2_c2c(x.to(dev)).cpu(), torch.fft.fft2(x)) < 3e-6
    assert rel_l2(Fn.fft2_c2c(x.to(dev), inverse=True).cpu(), torch.fft.ifft2(x)) < 3e-6
    assert rel_l2(Fn.fft2_c2c(x.to(dev), ortho=True).cpu(), torch.fft.fft2(x, norm="ortho")) < 3e-6


@pytest.mark.parametrize("H,W", [(101, 67), (13, 64), (202, 268)])
def test_fft2_of_any_size(H, W, dev):
    from quantizationawarethzdoe_b200 import functional as Fn
    from quantizationawarethzdoe_b200.utils.Helper_Functions import ft2, ift2
    from oracle.asm_oracle import shifted_fft2
    torch.manual_seed(2)
    x = torch.randn(2, 1, H, W, dtype=torch.complex64)
    assert rel_l2(Fn.fft2_c2c(x.to(dev)).cpu(), torch.fft.fft2(x)) < 3e-6
    assert rel_l2(Fn.fft2_c2c(x.to(dev), inverse=True).cpu(), torch.fft.ifft2(x)) < 3e-6
    assert rel_l2(ft2(x.to(dev)).cpu(), shifted_fft2(x)) < 3e-6
    assert rel_l2(ift2(x.to(dev)).cpu(), shifted_fft2(x, inverse=True)) < 3e-6


@pytest.mark.parametrize("H,W,scale", [(512, 1024, None),      # 1024 x 2048
                                       (256, 512, 2),          # padding_scale 2 (the notebooks): 768 x 1536
                                       (1024, 200, 2)])        # 3072 x 600 (600 = generic engine)
def test_static_fast_path_agrees_with_generic_engine(H, W, scale, dev, monkeypatch):
    """The compile-time specialised kernels and the runtime-planned engine must give the same field and the same
    adjoint (same algorithm, same twiddles): run both on one input (THZ_NO_P2=1 selects the generic engine)."""
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField
    torch.manual_seed(5)
    x = torch.randn(2, 2, H, W, dtype=torch.complex64, device=dev)
    lams = [1 * mm, 1.02 * mm]
    outs = []
    for flag in ("0", "1"):
        monkeypatch.setenv("THZ_NO_P2", flag)
        a = ASM_prop(z_distance=0.1, device=dev, padding_scale=scale)
        a.check_Zc = False
        xr = x.clone().requires_grad_(True)
        y = a(ElectricField(xr, wavelengths=lams, spacing=0.5 * mm, device=dev)).data
        (gx,) = torch.autograd.grad(y, xr, y.detach())
        outs.append((y.detach().cpu(), gx.cpu()))
    assert rel_l2(outs[0][0], outs[1][0]) < 1e-6
    assert rel_l2(outs[0][1], outs[1][1]) < 1e-6


@pytest.fixture(params=["auto", "simt"])
def czt_impl(request, monkeypatch):
    """Both Toeplitz-GEMM implementations: 'auto' = the tcgen05 3xTF32 kernel wherever it is eligible (the small reference
    vectors include chirp filters shorter than 64, which only the CUDA-core kernel serves) and the CUDA-core baseline."""
    from quantizationawarethzdoe_b200 import functional as Fn
    monkeypatch.setitem(Fn.TUNE, "czt_impl", request.param)
    return request.param


@pytest.mark.parametrize("name", golden_names("czt_"))
def test_czt_matches_reference_vectors(name, dev, czt_impl):
    """CZT_prop forward against the reference's own output; tolerance 1e-5 rel-L2 (north_star)."""
    from quantizationawarethzdoe_b200 import CZT_prop, ElectricField
    g = golden(name)
    M = int(g["M"])
    f = ElectricField(g["x"].to(dev), wavelengths=g["wavelengths"].float(), spacing=g["spacing"].float(), device=dev)
    czt = CZT_prop(z_distance=g["z"], device=dev)
    out = czt(f, outputHeight=M, outputWidth=M, outputPixel_dx=g["out_dx"], outputPixel_dy=g["out_dx"])
    assert out.data.shape == g["y"].shape
    assert rel_l2(out.data.cpu(), g["y"]) < TOL
    assert torch.allclose(out.spacing.cpu(), g["out_spacing"].float())


def test_czt_adjoint_and_oracle(dev, czt_impl):
    """Gradient wrt the input field: explicit adjoint GEMMs vs autograd through the oracle's dense form,
    plus the adjoint identity <A x, y> = <x, A^H y> at a larger size."""
    from oracle import czt_oracle as CO
    from quantizationawarethzdoe_b200 import CZT_prop, ElectricField
    torch.manual_seed(4)
    H, W, M, lams = 72, 56, 40, [1 * mm, 1.07 * mm]
    x = torch.randn(1, 2, H, W, dtype=torch.complex64)
    gy = torch.randn(1, 2, M, M, dtype=torch.complex64)
    xo = x.clone().requires_grad_(True)
    yo = CO.czt_forward_dense(xo, torch.tensor(lams), torch.tensor([0.5 * mm, 0.5 * mm]), torch.tensor(0.4), M, M, 0.2 * mm, 0.2 * mm)
    (gxo,) = torch.autograd.grad(yo, xo, gy)
    czt = CZT_prop(z_distance=0.4, device=dev)
    xd = x.to(dev).requires_grad_(True)
    y = czt(ElectricField(xd, wavelengths=lams, spacing=0.5 * mm, device=dev), M, M, 0.2 * mm, 0.2 * mm).data
    (gx,) = torch.autograd.grad(y, xd, gy.to(dev))
    assert rel_l2(y.detach().cpu(), yo.detach()) < TOL
    assert rel_l2(gx.cpu(), gxo) < TOL
    # larger, non-multiple-of-tile sizes: adjoint identity
    H, W, M = 300, 260, 150
    czt2 = CZT_prop(z_distance=0.5, device=dev)
    xa = torch.randn(1, 1, H, W, dtype=torch.complex64, device=dev).requires_grad_(True)
    yv = torch.randn(1, 1, M, M, dtype=torch.complex64, device=dev)
    Ax = czt2(ElectricField(xa, wavelengths=[1 * mm], spacing=0.5 * mm, device=dev), M, M, 0.1 * mm, 0.1 * mm).data
    (AHy,) = torch.autograd.grad(Ax, xa, yv)
    lhs, rhs = torch.sum(Ax.detach().conj() * yv), torch.sum(xa.detach().conj() * AHy)
    record("czt_adjoint_identity", impl=czt_impl, rel=float(abs(lhs - rhs) / abs(lhs)),
           rel_to_norms=float(abs(lhs - rhs) / (Ax.detach().norm() * yv.norm())))
    assert abs(lhs - rhs) / abs(lhs) < 1e-5        # measured 3-5e-7 (relative to |<Ax, y>|, itself ~1e-2 of |Ax| |y|)


def test_czt_rejects_non_square_output_like_the_reference(dev):
    from quantizationawarethzdoe_b200 import CZT_prop, ElectricField
    f = ElectricField(torch.zeros(1, 1, 32, 32, dtype=torch.complex64, device=dev), 1e-3, 1e-3, device=dev)
    with pytest.raises(RuntimeError, match="outputHeight must equal outputWidth"):
        CZT_prop(z_distance=0.5, device=dev)(f, outputHeight=16, outputWidth=24)


def test_czt_tensor_core_kernel_agrees_with_cuda_core_kernel(dev, monkeypatch):
    """Config-3-like shape (reduced): 1024^2 -> 512^2, 2 wavelengths; tcgen05 3xTF32 vs fp32 CUDA cores."""
    from quantizationawarethzdoe_b200 import CZT_prop, ElectricField
    torch.manual_seed(9)
    x = torch.randn(1, 2, 1024, 1024, dtype=torch.complex64, device=dev)
    outs = {}
    from quantizationawarethzdoe_b200 import functional as Fn
    for impl in ("tc", "simt"):
        monkeypatch.setitem(Fn.TUNE, "czt_impl", impl)
        czt = CZT_prop(z_distance=0.5, device=dev)
        f = ElectricField(x, wavelengths=[1 * mm, 1.05 * mm], spacing=0.5 * mm, device=dev)
        outs[impl] = czt(f, 512, 512, 0.1 * mm, 0.1 * mm).data
    assert rel_l2(outs["tc"].cpu(), outs["simt"].cpu()) < 5e-6     # each is ~2e-6 from a float64 evaluation


def test_step_is_cuda_graph_capturable(dev):
    """The whole optimisation step (DOE level selection -> fused DOE+ASM forward -> loss gradient -> adjoint ->
    weight gradient) makes no host synchronisation and allocates only through torch, so it can be captured
    in a CUDA graph and replayed: replay reproduces the eager gradient bit for bit."""
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField, STEQuantizedDOELayer
    n = 256
    torch.manual_seed(0)
    doe = STEQuantizedDOELayer(dict(doe_size=[n, n], doe_dxy=0.5 * mm, doe_level=8, height_constraint_max=1 * mm, tolerance=None,
                                    material=[2.66, 0.003]), {}, device=dev)
    asm = ASM_prop(z_distance=0.1, device=dev)
    asm.check_Zc = False
    x = torch.randn(1, 1, n, n, dtype=torch.complex64, device=dev)
    lam_t, sp_t = torch.tensor([1 * mm], device=dev), torch.tensor([0.5 * mm, 0.5 * mm], device=dev)

    def step():
        y = asm(doe(ElectricField(x, wavelengths=lam_t, spacing=sp_t, device=dev))).data
        (gw,) = torch.autograd.grad(y, doe.weight_height_map, y.detach())
        return gw

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            eager = step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        static_gw = step()
    with torch.no_grad():
        doe.weight_height_map.add_(0.25)          # new weights: the replay must see them
    g.replay()
    torch.cuda.synchronize()
    fresh = step()
    assert torch.equal(static_gw, fresh)
    assert not torch.equal(static_gw, eager)


# ------------------------------------------------------------------------------- loss + optimizer
@pytest.mark.parametrize("name", golden_names("train_loss"))
def test_normalized_intensity_mse_matches_reference(name, dev):
    """Fused normalize(|y|^2) + MSE loss and its gradient wrt the complex field vs the reference's own lines
    (utils/Helper_Functions.py:185-193 + nn.MSELoss + autograd, incl. the gradient through max())."""
    from quantizationawarethzdoe_b200 import normalized_intensity_mse
    g = golden(name)
    y = g["y"].to(dev).requires_grad_(True)
    loss = normalized_intensity_mse(y, g["target"].to(dev))
    (gy,) = torch.autograd.grad(loss, y)
    assert abs(float(loss) - g["loss"]) <= 1e-6 * abs(g["loss"])
    assert rel_l2(gy.cpu(), g["gy"]) <= 1e-6


def test_normalized_intensity_mse_large_vs_oracle(dev):
    from oracle import train_oracle as TO
    from quantizationawarethzdoe_b200 import normalized_intensity_mse
    torch.manual_seed(3)
    y = torch.randn(2, 3, 300, 257, dtype=torch.complex64)
    t = torch.rand(2, 3, 300, 257)
    lo, go = TO.normalized_intensity_mse(y, t)
    yd = y.to(dev).requires_grad_(True)
    loss = normalized_intensity_mse(yd, t.to(dev))
    (gy,) = torch.autograd.grad(loss, yd)
    assert abs(float(loss) - float(lo)) <= 2e-6 * float(lo)
    assert rel_l2(gy.cpu(), go) <= 2e-6
    # the same call without a gradient request, and an upstream factor
    assert abs(float(normalized_intensity_mse(y.to(dev), t.to(dev))) - float(lo)) <= 2e-6 * float(lo)
    (g3,) = torch.autograd.grad(3.0 * normalized_intensity_mse(yd, t.to(dev)), yd)
    assert rel_l2(g3.cpu(), 3.0 * go) <= 2e-6


@pytest.mark.parametrize("name", golden_names("train_adam"))
def test_fused_adam_matches_torch(name, dev):
    """FusedAdam (device-side step counter) vs torch.optim.Adam / AdamW over 6 steps (fp32 bias corrections on the
    device instead of torch's float64 host scalars: <= 1e-6 relative)."""
    from quantizationawarethzdoe_b200 import FusedAdam
    g = golden(name)
    p = torch.nn.Parameter(g["p0"].to(dev))
    opt = FusedAdam([p], lr=0.02, weight_decay=g["weight_decay"], decoupled_weight_decay=bool(g["decoupled"]))
    for i in range(g["grads"].shape[0]):
        p.grad = g["grads"][i].to(dev)
        opt.step()
        assert rel_l2(p.detach().cpu(), g["hist"][i]) <= 1e-6, i
    assert int(opt.state[p]["step"].item()) == g["grads"].shape[0]


def test_whole_iteration_replays_as_one_cuda_graph(dev):
    """DOE -> ASM -> normalized-intensity MSE -> backward -> FusedAdam captured once and replayed: the weights after
    three replays equal three eager iterations (same kernels, same order)."""
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField, FusedAdam, STEQuantizedDOELayer, normalized_intensity_mse
    n = 64
    params = dict(doe_size=[n, n], doe_dxy=1 * mm, doe_level=4, height_constraint_max=1 * mm, tolerance=None, material=[2.66, 0.003])
    torch.manual_seed(5)
    x = torch.randn(1, 1, n, n, dtype=torch.complex64, device=dev)
    target = torch.rand(1, 1, n, n, device=dev)
    lam, sp = torch.tensor([1 * mm], device=dev), torch.tensor([1 * mm, 1 * mm], device=dev)

    def build():
        torch.manual_seed(11)
        doe = STEQuantizedDOELayer(params, {}, device=dev)
        asm = ASM_prop(z_distance=0.05, device=dev)
        asm.check_Zc = False
        opt = FusedAdam(doe.parameters(), lr=0.02)
        return doe, asm, opt

    def iteration(doe, asm, opt):
        out = asm(doe(ElectricField(x, wavelengths=lam, spacing=sp, device=dev)))
        loss = normalized_intensity_mse(out.data, target)
        opt.zero_grad(set_to_none=False)
        loss.backward()
        opt.step()
        return loss

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        doe_e, asm_e, opt_e = build()
        for _ in range(4):
            iteration(doe_e, asm_e, opt_e)
        doe_g, asm_g, opt_g = build()
        iteration(doe_g, asm_g, opt_g)                       # warm-up: allocates state, builds the plan
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        iteration(doe_g, asm_g, opt_g)                       # capture records, does not execute
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    assert int(opt_g.state[doe_g.weight_height_map]["step"].item()) == 4
    assert rel_l2(doe_g.weight_height_map.detach().cpu(), doe_e.weight_height_map.detach().cpu()) <= 1e-6


@pytest.mark.parametrize("G,n", [(2, 512), (4, 1024)])
def test_peer_memory_slab_kernels_on_one_gpu(G, n, dev):
    """The scatter / gather row kernels of the peer-memory slab FFT, with the G ranks run one after the other on this
    GPU (all column slabs local): identical arithmetic to the single-GPU pipeline, so the result is bit-identical."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from test_emul_kernels import _slab_plans
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField, parallel as P
    lams = [1 * mm, 1.04 * mm]
    torch.manual_seed(0)
    x = torch.randn(1, 2, n, n, dtype=torch.complex64, device=dev)
    g = torch.randn(1, 2, n, n, dtype=torch.complex64, device=dev)
    plans = _slab_plans(G, 2, n, n, lams, 0.5 * mm, 0.1, dev)
    y = P.slab_emulate_ranks(x, plans)
    gx = P.slab_emulate_ranks(g, plans, conj=True)
    asm = ASM_prop(z_distance=0.1, device=dev)
    asm.check_Zc = False
    xr = x.clone().requires_grad_(True)
    yr = asm(ElectricField(xr, wavelengths=lams, spacing=0.5 * mm, device=dev)).data
    (gxr,) = torch.autograd.grad(yr, xr, g)
    assert torch.equal(y, yr.detach())
    assert torch.equal(gx, gxr)


# ------------------------------------------------------------------------------- Rayleigh-Sommerfeld
@pytest.mark.parametrize("name", golden_names("rsc_"))
def test_rsc_matches_reference_vectors(name, dev):
    """RSC_prop / VRS_prop on the fused FFT pipeline (input in the upper-left corner, lower-right crop, cached spectrum of
    the impulse response) against the reference's own output and input gradient."""
    from quantizationawarethzdoe_b200 import ElectricField, RSC_prop, VRS_prop
    g = golden(name)
    cls = VRS_prop if name == "rsc_vectorial" else RSC_prop
    prop = cls(z_distance=g["z"], device=dev)
    prop.check_Zc = False
    x = g["x"].to(dev).requires_grad_(True)
    y = prop(ElectricField(x, wavelengths=g["wavelengths"].float(), spacing=g["spacing"].float(), device=dev)).data
    assert y.shape == g["y"].shape
    assert rel_l2(y.detach().cpu(), g["y"]) < TOL
    if "gx" in g:
        (gx,) = torch.autograd.grad(y, x, g["g"].to(dev))
        assert rel_l2(gx.cpu(), g["gx"]) < TOL


def test_rsc_large_vs_oracle(dev):
    """512 -> 1024 grid (static kernels, blocked intermediate) and a batch of fields vs the oracle."""
    from oracle import rsc_oracle as RO
    from quantizationawarethzdoe_b200 import ElectricField, RSC_prop
    torch.manual_seed(2)
    x = torch.randn(2, 2, 512, 512, dtype=torch.complex64)
    lams = [1 * mm, 1.02 * mm]
    yo = RO.rsc_forward(x, lams, 0.5 * mm, 0.25)
    prop = RSC_prop(z_distance=0.25, device=dev)
    prop.check_Zc = False
    y = prop(ElectricField(x.to(dev), wavelengths=lams, spacing=0.5 * mm, device=dev)).data
    assert rel_l2(y.cpu(), yo) < TOL


def test_lens_and_apertures_match_reference(dev):
    """Thin_LensElement / ApertureElement through thz_field_mul: outputs and input gradients vs the reference."""
    from quantizationawarethzdoe_b200 import ApertureElement, ElectricField, Thin_LensElement
    g = golden("elem_lens_aperture")
    els = {"lens": Thin_LensElement(focal_length=g["focal"], device=dev), "circ": ApertureElement("circ", g["radius"], device=dev),
           "rect": ApertureElement("rect", g["side"], device=dev)}
    for name, el in els.items():
        x = g["x"].to(dev).requires_grad_(True)
        y = el(ElectricField(x, wavelengths=g["wavelengths"].float(), spacing=g["spacing"].float(), device=dev)).data
        (gx,) = torch.autograd.grad(y, x, g["g"].to(dev))
        tol = 2e-6 if name == "lens" else 0.0
        assert rel_l2(y.detach().cpu(), g["y_" + name]) <= tol, name
        assert rel_l2(gx.cpu(), g["gx_" + name]) <= tol, name
    with pytest.raises(ValueError):
        ApertureElement("circ", 1.0, device=dev)(ElectricField(g["x"].to(dev), wavelengths=g["wavelengths"].float(),
                                                               spacing=g["spacing"].float(), device=dev))


def test_lens_and_aperture_are_fused_into_the_next_propagation(dev):
    """aperture -> lens -> DOE -> ASM runs as ONE fused pipeline (elements multiplied on load in the row-FFT
    kernel, their conjugates in the adjoint's epilogue): no stand-alone pointwise launch, same field / gradients as the
    oracle chain; and aperture -> ASM without a DOE."""
    from oracle import asm_oracle as AO, doe_oracle as DO, element_oracle as EO
    from quantizationawarethzdoe_b200 import (ASM_prop, ApertureElement, ElectricField, STEQuantizedDOELayer, Thin_LensElement,
                                              _native as N)
    n, C, lams, dxy, z = 256, 2, [1 * mm, 1.04 * mm], 0.5 * mm, 0.1
    torch.manual_seed(0)
    x = torch.randn(2, C, n, n, dtype=torch.complex64)
    g = torch.randn(2, C, n, n, dtype=torch.complex64)
    torch.manual_seed(1)
    doe = STEQuantizedDOELayer(dict(doe_size=[n, n], doe_dxy=dxy, doe_level=4, height_constraint_max=1 * mm, tolerance=None,
                                    material=[2.66, 0.003]), {}, device=dev)
    ap, lens = ApertureElement("circ", 0.05, device=dev), Thin_LensElement(focal_length=0.2, device=dev)
    asm = ASM_prop(z_distance=z, device=dev, kernel_mode="cached")
    asm.check_Zc = False
    lib = N.lib()
    xd = x.to(dev).requires_grad_(True)
    f0 = ElectricField(xd, wavelengths=lams, spacing=dxy, device=dev)
    asm(doe(lens(ap(f0))))                                 # warm-up: builds masks, kernels, plans
    l4 = lib.thz_launch_count_class(4)
    u = doe(lens(ap(f0)))
    assert u._data is None and u._deferred.mask is not None and u._deferred.mul is not None
    y = asm(u).data
    gx, gw = torch.autograd.grad(y, (xd, doe.weight_height_map), g.to(dev))
    torch.cuda.synchronize()
    assert lib.thz_launch_count_class(4) == l4, "a stand-alone pointwise kernel ran: the elements were not fused"
    mask = EO.circ_mask(n, n, torch.tensor([dxy, dxy]), 0.05).float()
    ker = EO.lens_kernel(n, n, torch.tensor([dxy, dxy]), torch.tensor(lams), 0.2)
    xo = x.clone().requires_grad_(True)
    wo = doe.weight_height_map.detach().cpu().clone().requires_grad_(True)
    h = DO.ste_quantize(DO.sigmoid_height(wo[0, 0], 1 * mm), DO.linear_lut(1 * mm, 4))
    yo = AO.asm_forward(DO.modulate(xo * mask.reshape(1, 1, n, n) * ker.reshape(1, C, n, n), h, lams, 2.66, 0.003), lams, dxy, z)
    gxo, gwo = torch.autograd.grad(yo, (xo, wo), g)
    e = [rel_l2(y.detach().cpu(), yo.detach()), rel_l2(gx.cpu(), gxo), rel_l2(gw.cpu(), gwo)]
    record("fused_elements", y=e[0], gx=e[1], gw=e[2])
    assert max(e) < TOL
    # aperture only, then propagation (the DONN's encode_object): fused too; materialised path gives the same numbers
    xd2 = x.to(dev).requires_grad_(True)
    y2 = asm(ap(ElectricField(xd2, wavelengths=lams, spacing=dxy, device=dev))).data
    (gx2,) = torch.autograd.grad(y2, xd2, g.to(dev))
    xo2 = x.clone().requires_grad_(True)
    yo2 = AO.asm_forward(xo2 * mask.reshape(1, 1, n, n), lams, dxy, z)
    (gxo2,) = torch.autograd.grad(yo2, xo2, g)
    assert rel_l2(y2.detach().cpu(), yo2.detach()) < TOL and rel_l2(gx2.cpu(), gxo2) < TOL
    assert lib.thz_launch_count_class(4) == l4
    mat = ap(ElectricField(x.to(dev), wavelengths=lams, spacing=dxy, device=dev)).data       # someone reads .data: stand-alone kernel
    assert lib.thz_launch_count_class(4) == l4 + 1 and rel_l2(mat.cpu(), x * mask.reshape(1, 1, n, n)) == 0.0
    # a fixed field in front of the DOE that every iteration of a loop re-uses (the notebooks' field_before_DOE): fused on its first
    # use, evaluated ONCE on the second, a plain tensor from then on
    fixed = lens(ap(ElectricField(x.to(dev), wavelengths=lams, spacing=dxy, device=dev)))
    outs = [asm(doe(fixed)).data.detach().clone() for _ in range(3)]
    assert lib.thz_launch_count_class(4) == l4 + 3                   # + mask multiply + lens multiply, once
    assert torch.equal(outs[1], outs[2]) and rel_l2(outs[0], outs[1]) < 1e-6 and fixed._data is not None


def test_notebook_setup_end_to_end(dev):
    """The whole set-up of experiment_four_focal_spots.ipynb (cells 2-8, STE layer) from this package alone:
    Gaussian beam -> ASM 127 mm (padding_scale 2) -> thin lens -> rect aperture -> 4-level STE DOE -> ASM 200 mm ->
    normalize(|y|^2) + MSE -> gradient of the DOE weights, against the reference run on the CPU."""
    from quantizationawarethzdoe_b200 import (ASM_prop, ApertureElement, Guassian_beam, STEQuantizedDOELayer, Thin_LensElement,
                                              normalized_intensity_mse)
    g = golden("setup_four_focal_spots")
    n, dxy, lam = 100, g["spacing"], g["wavelength"]
    src = Guassian_beam(height=n, width=n, beam_waist_x=None, beam_waist_y=None, wavelengths=lam, spacing=dxy, device=dev)
    asm1 = ASM_prop(z_distance=0.127, bandlimit_type="exact", padding_scale=2, bandlimit_kernel=True, device=dev)
    asm3 = ASM_prop(z_distance=0.2, bandlimit_type="exact", padding_scale=2, bandlimit_kernel=True, device=dev)
    asm1.check_Zc = asm3.check_Zc = False
    lens, ap = Thin_LensElement(focal_length=0.127, device=dev), ApertureElement("rect", 0.08, device=dev)
    doe = STEQuantizedDOELayer(dict(doe_size=[n, n], doe_dxy=dxy, doe_level=4, look_up_table=None, num_unit=None,
                                    height_constraint_max=1 * mm, tolerance=None, material=[2.66, 0.03]), {}, device=dev)
    with torch.no_grad():
        doe.weight_height_map.copy_(g["w0"].to(dev))
    s0 = src()
    # the waist comes from a degree-5 polynomial fit evaluated in fp32 with heavy cancellation (LightSource/Gaussian_beam.py:
    # 67-85): its value depends on the host's libm at the 1e-5 level, so the source is checked loosely and the pipeline
    # below is fed the reference's own source field
    assert rel_l2(s0.data.cpu(), g["source"]) < 1e-4
    from quantizationawarethzdoe_b200 import ElectricField
    s0 = ElectricField(g["source"].to(dev), wavelengths=s0.wavelengths, spacing=s0.spacing, device=dev)
    fin = ap(lens(asm1(s0)))
    assert rel_l2(fin.data.cpu(), g["field_before_doe"]) < TOL
    out = asm3(doe(fin, None))
    assert rel_l2(out.data.detach().cpu(), g["out"]) < TOL
    assert torch.equal(doe.height_map.detach().cpu(), g["height_map"])           # level selection bit-exact
    loss = normalized_intensity_mse(out.data, g["target"].to(dev))
    (gw,) = torch.autograd.grad(loss, doe.weight_height_map)
    assert abs(float(loss) - g["loss"]) <= 1e-5 * abs(g["loss"])
    assert rel_l2(gw.cpu(), g["gw"]) < 2e-5


# ------------------------------------------------------------------------------- loss-landscape sweep
@pytest.mark.parametrize("batched", [True, False])
def test_loss_landscape_matches_the_reference_run(batched, dev, tmp_path):
    """calulate_single_element_loss_landscape on the small single-DOE system of tests/golden/landscape_single_doe.npz, whose
    `loss` array was produced by the reference's own function: batched (7 grid points per fused pass, ragged last chunk)
    and point by point, same surface file contents."""
    import types
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField, STEQuantizedDOELayer
    from quantizationawarethzdoe_b200.VisTools import calulate_single_element_loss_landscape, read_surface_file
    g = golden("landscape_single_doe")
    n = g["x"].shape[-1]

    class Setup(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.input_field = ElectricField(data=g["x"].to(dev), wavelengths=g["wavelength"], spacing=g["spacing"], device=dev)
            self.doe = STEQuantizedDOELayer(dict(doe_size=[n, n], doe_dxy=g["spacing"], doe_level=4, look_up_table=None, num_unit=None,
                                                 height_constraint_max=1 * mm, tolerance=None, material=[2.66, 0.03]), {}, device=dev)
            prop = ASM_prop(z_distance=g["z"], bandlimit_type='exact', padding_scale=None, bandlimit_kernel=True, device=dev)
            prop.check_Zc = False
            if batched:
                self.asm_prop3 = prop          # the notebook's attribute name: picked up by the batched path
            else:
                self.prop = prop

        def forward(self, iter_frac):
            return (self.asm_prop3 if batched else self.prop)(self.doe(self.input_field, iter_frac))

    model = Setup()
    with torch.no_grad():
        model.doe.weight_height_map.copy_(g["w0"].to(dev))
    a = g["args"].tolist()
    args = types.SimpleNamespace(xmin=a[0], xmax=a[1], xnum=int(a[2]), ymin=a[3], ymax=a[4], ynum=int(a[5]))
    directions = [[g["dx"].to(dev)], [g["dy"].to(dev)]]
    lib = __import__("quantizationawarethzdoe_b200")._native.lib()
    l0 = lib.thz_launch_count_class(0)
    path = calulate_single_element_loss_landscape(args, model, g["target"].to(dev), loss_f=torch.nn.MSELoss(), directions=directions,
                                                  save_path=str(tmp_path), batch=7)
    row_launches = lib.thz_launch_count_class(0) - l0
    assert row_launches == (3 if batched else 20)          # 20 grid points: ceil(20 / 7) fused passes, or one per point
    out = read_surface_file(path)
    assert np.array_equal(out["xcoordinates"], g["xcoordinates"].numpy()) and np.array_equal(out["ycoordinates"], g["ycoordinates"].numpy())
    err = float(np.abs(out["loss"] - g["loss"].numpy()).max() / np.abs(g["loss"].numpy()).max())
    record("loss_landscape", batched=batched, max_rel_err=err)
    assert out["loss"].shape == (5, 4) and err < 1e-5
    assert torch.equal(model.doe.weight_height_map.detach().cpu(), g["w0"])          # weights restored
