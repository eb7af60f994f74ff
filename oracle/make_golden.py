#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference).

TEST INFRASTRUCTURE.  Run in the build container only (the reference does not travel to the GPU box):

    python oracle/make_golden.py

Every fixture stores its inputs and the reference's outputs (forward values, gradients obtained by
the reference's own autograd, height maps, level indices), so the GPU parity tests and the oracle
self-checks need nothing but the .npz files.  Sizes are kept small (< 300 KB per file).
"""
import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference, quiet  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
mm = 1e-3


def save(name, **arrs):
    conv = {}
    for k, v in arrs.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        conv[k] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **conv)
    print("wrote", name, {k: tuple(v.shape) for k, v in conv.items() if v.ndim})


ASM_CASES = {
    # name: (B, C, H, W, padding_scale, wavelengths, spacing, z, bandlimit_type, do_padding, do_unpad)
    "asm_pow2": (1, 1, 64, 64, None, [1 * mm], 0.5 * mm, 0.1, "exact", True, True),
    "asm_multi_lambda": (2, 3, 32, 48, None, [0.95 * mm, 1 * mm, 1.05 * mm], 0.5 * mm, 0.08, "exact", True, True),
    "asm_pad3_aniso": (1, 1, 50, 50, 2, [1 * mm], [1 * mm, 0.7 * mm], 0.26, "exact", True, True),
    "asm_approx_mixedpad": (1, 2, 30, 36, [1, 2], [0.9 * mm, 1.2 * mm], 0.5 * mm, 0.05, "approx", True, True),
    "asm_nounpad": (1, 1, 40, 24, None, [1 * mm], 0.5 * mm, 0.1, "exact", True, False),
    "asm_nopad": (1, 1, 48, 48, None, [1 * mm], 0.5 * mm, 0.1, "exact", False, True),
    "asm_far": (1, 1, 100, 100, 2, [1 * mm], 1 * mm, 0.3, "exact", True, True),   # notebook geometry: 100 -> 300, 300 GHz
}


def gen_asm(R):
    for name, (B, C, H, W, scale, lams, dxy, z, bt, do_pad, do_unpad) in ASM_CASES.items():
        torch.manual_seed(zlib.crc32(name.encode()) % 1000)
        x = torch.randn(B, C, H, W, dtype=torch.complex64).requires_grad_(True)
        f = R.ElectricField(x, wavelengths=lams, spacing=dxy, device=torch.device("cpu"))
        asm = R.ASM_prop(z_distance=z, padding_scale=scale, bandlimit_type=bt, do_padding=do_pad,
                         do_unpad_after_pad=do_unpad, device=torch.device("cpu"))
        with quiet():
            y = asm(f).data
            kern = asm.create_kernel(f)
        g = torch.randn_like(y)
        (gx,) = torch.autograd.grad(y, x, g)
        extra = {}
        if name in ("asm_pow2", "asm_approx_mixedpad"):
            extra["kernel"] = kern
        save(name, x=x, g=g, y=y, gx=gx, wavelengths=np.array(lams, np.float64), spacing=np.array(np.broadcast_to(dxy, (2,)), np.float64),
             z=np.float64(z), padding_scale=np.array([-1.0] if scale is None else np.broadcast_to(scale, (2,)), np.float64),
             bandlimit_type=np.array(bt), do_padding=np.array(do_pad), do_unpad=np.array(do_unpad), **extra)


def gen_doe(R):
    QD = R.QD
    dev = torch.device("cpu")
    N, lams = 48, [1 * mm, 1.05 * mm]
    dp = dict(doe_size=[N, N], doe_dxy=0.5 * mm, doe_level=4, height_constraint_max=1 * mm, tolerance=None, material=[2.66, 0.003])
    op = dict(c_s=300, tau_max=5.5, tau_min=2.0)
    torch.manual_seed(7)
    x = torch.randn(2, 2, N, N, dtype=torch.complex64)
    g = torch.randn(2, 2, N, N, dtype=torch.complex64)
    f = R.ElectricField(x, wavelengths=lams, spacing=0.5 * mm, device=dev)
    asm = R.ASM_prop(z_distance=0.1, device=dev)
    common = dict(x=x, g=g, wavelengths=np.array(lams), spacing=np.array([0.5 * mm] * 2), z=np.float64(0.1),
                  hmax=np.float64(1 * mm), levels=np.int64(4), material=np.array([2.66, 0.003]))

    def run(layer, param, fwd_kwargs, name, **extra):
        with quiet():
            u = layer(f, **fwd_kwargs)
            y = asm(u).data
        (gw,) = torch.autograd.grad(y, param, g, retain_graph=True)
        (gw_mod,) = torch.autograd.grad(u.data, param, g)
        save(name, w=param, height_map=layer.height_map, u=u.data, y=y, gw=gw, gw_modulate_only=gw_mod,
             lut=getattr(layer, "lut", torch.zeros(0)), **common, **extra)

    torch.manual_seed(11)
    ste = QD.STEQuantizedDOELayer(dp, op, device=dev)
    run(ste, ste.weight_height_map, {}, "doe_ste")
    torch.manual_seed(12)
    fp = QD.FullPrecisionDOELayer(dp, device=dev)
    run(fp, fp.weight_height_map, {}, "doe_fullprecision")
    torch.manual_seed(13)
    psq = QD.PSQuantizedDOELayer(dp, dict(tau_max=400, tau_min=1), device=dev)
    run(psq, psq.weight_height_map, dict(iter_frac=0.3), "doe_psq", iter_frac=np.float64(0.3))
    torch.manual_seed(14)
    v3 = QD.SoftGumbelQuantizedDOELayerv3(dp, op, device=dev)
    for fr in (0.2, 0.5, 0.9):
        torch.manual_seed(100)
        noise = -torch.empty(1, 4, N, N).exponential_().log()       # what F.gumbel_softmax draws under this seed
        torch.manual_seed(100)
        run(v3, v3.weight_init_phase, dict(iter_frac=fr), "doe_gumbel_v3_%02d" % int(fr * 10), iter_frac=np.float64(fr),
            noise=noise, c_s=np.float64(300), tau_max=np.float64(5.5), tau_min=np.float64(2.0))
    torch.manual_seed(15)
    v2 = QD.SoftGumbelQuantizedDOELayerv2(dp, op, device=dev)
    torch.manual_seed(101)
    noise = -torch.empty(1, 4, N, N).exponential_().log()
    torch.manual_seed(101)
    run(v2, v2.weight_init_phase, dict(iter_frac=0.7), "doe_gumbel_v2", iter_frac=np.float64(0.7), noise=noise,
        c_s=np.float64(300), tau_max=np.float64(5.5), tau_min=np.float64(2.0))
    torch.manual_seed(16)
    v1 = QD.SoftGumbelQuantizedDOELayer(dp, op, device=dev)
    torch.manual_seed(102)
    noise = -torch.empty(1, 4, N, N).exponential_().log()
    torch.manual_seed(102)
    run(v1, v1.init_phase, dict(iter_frac=0.4), "doe_gumbel_v1", iter_frac=np.float64(0.4), noise=noise,
        c_s=np.float64(300), tau_max=np.float64(5.5), tau_min=np.float64(2.0))
    torch.manual_seed(17)
    ng = QD.NaiveGumbelQuantizedDOELayer(dp, op, device=dev)
    torch.manual_seed(103)
    noise = -torch.empty(N, N, 4).exponential_().log()
    torch.manual_seed(103)
    run(ng, ng.weight_height_map, dict(iter_frac=0.4), "doe_gumbel_naive", iter_frac=np.float64(0.4), noise=noise,
        tau_max=np.float64(5.5), tau_min=np.float64(2.0))

    # FixDOEElement on the reference's own 80x80 4-level height map (edoe_4levels.npy is a pickled dict)
    try:
        hm = np.load(os.path.join("/root/reference", "edoe_4levels.npy"), allow_pickle=True)
        hm = hm.item()["thickness"] if hm.dtype == object else hm
        hm = torch.tensor(np.asarray(hm, dtype=np.float32))
        Hh, Ww = hm.shape
        torch.manual_seed(21)
        xx = torch.randn(1, 1, Hh, Ww, dtype=torch.complex64)
        gg = torch.randn(1, 1, Hh, Ww, dtype=torch.complex64)
        fix = QD.FixDOEElement(height_map=hm.numpy(), tolerance=0.0, material=[2.66, 0.003], device=dev)
        ff = R.ElectricField(xx, wavelengths=[1 * mm], spacing=1 * mm, device=dev)
        asm2 = R.ASM_prop(z_distance=0.1, padding_scale=2, device=dev)
        with quiet():
            u = fix(ff)
            y = asm2(u).data
        (gh,) = torch.autograd.grad(y, fix.height_map, gg)
        save("doe_fix_edoe4", x=xx, g=gg, height_map=hm, u=u.data, y=y, gh=gh, wavelengths=np.array([1 * mm]),
             spacing=np.array([1 * mm] * 2), z=np.float64(0.1), material=np.array([2.66, 0.003]))
    except Exception as e:  # pragma: no cover
        print("edoe fixture skipped:", e)


def gen_quant(R):
    # nearest-neighbour quantizers, patched the way SURVEY a-9 describes
    lut = torch.linspace(0, 1, 5)
    R.DiscreteDOE.lut = lut
    R.DiscreteDOE.lut_midvals = torch.tensor(R.HF.lut_mid(lut))
    xs = torch.tensor([0, .1, .124, .125, .126, .3, .6, .874, .875, .9, 1.0, 1.2])
    idx = R.HF.nearest_idx(xs, R.DiscreteDOE.lut_midvals)
    torch.manual_seed(5)
    xr = (torch.rand(4096) * 1.3 - 0.1).requires_grad_(True)
    s = torch.tensor(2.0)
    out = {}
    for kind, fn in (("nn", R.QZ.nns), ("nn_poly", R.QZ.nns_poly), ("nn_sigmoid", R.QZ.nns_sigmoid)):
        q = fn(xr, s)
        (gr,) = torch.autograd.grad(q, xr, torch.ones_like(q))
        out["q_" + kind] = q
        out["grad_" + kind] = gr
    save("quant_nn", lut=lut, mid=R.DiscreteDOE.lut_midvals, kat_x=xs, kat_idx=idx, x=xr, s=s, **out)
    # STE known answers (Components/test_all.ipynb cells 13-21)
    ste = R.QD.STEQuantizationFunction.apply
    a = torch.tensor([0.1, 0.4, 0.7, 1.2], requires_grad=True)
    l3 = torch.tensor([0, 0.5, 1.0])
    q = ste(a, l3)
    (ga,) = torch.autograd.grad(q.sum(), a)
    torch.manual_seed(6)
    hh = torch.rand(64, 64) * 1e-3
    l8 = torch.linspace(0, 1e-3, 9)[:-1]
    save("quant_ste", kat_x=a, kat_lut=l3, kat_q=q, kat_grad=ga, h=hh, lut=l8, q=ste(hh, l8),
         idx=torch.argmin(torch.abs(hh.unsqueeze(-1) - l8), dim=-1))


def gen_quant_softmax(R):
    """SoftmaxBasedQuantization + score_thickness through the reference's own Quantization dispatch
    (Components/quantization.py:36-55, 128-161, 164-207): '*softmax*' with and without Gumbel noise, hard and soft."""
    H, W, bits = 48, 40, 2
    hmax = 1 * mm
    torch.manual_seed(41)
    t0 = torch.rand(1, 1, H, W) * 1.2 * hmax - 0.1 * hmax
    gq = torch.randn(H, W)
    cases = {"gumbel_hard": ("gumbel_softmax", 0.5, True, 201), "gumbel_soft": ("gumbel_softmax", 0.2, False, 202),
             "plain_hard": ("softmax", 0.3, True, None), "plain_soft": ("softmax", 0.8, False, None)}
    # NOTE: the reference's own backward through score_thickness raises ("modified by an inplace operation": `diff /=
    # torch.max(torch.abs(diff))`, quantization.py:41, invalidates what abs() saved), so the reference pins only the
    # FORWARD values here; the gradients stored next to them come from the oracle's out-of-place restatement
    # (doe_oracle.softmax_quantize), i.e. autograd through diff / max|diff| including the path through the max.
    from oracle import doe_oracle as DO
    lut_full = torch.linspace(0, hmax, 2 ** bits + 1)
    out = {}
    for name, (method, frac, hard, seed) in cases.items():
        qz = R.QZ.Quantization(method=method, max_thickness=hmax, num_bits=bits, dev=torch.device("cpu"), tau_min=0.5, tau_max=3.0, c=300.)
        noise = None
        if seed is not None:
            torch.manual_seed(seed)
            noise = -torch.empty(1, 2 ** bits, H, W).exponential_().log()   # what F.gumbel_softmax draws under this seed
            torch.manual_seed(seed)
            out["noise_" + name] = noise
        with torch.no_grad():
            q = qz(t0.clone(), iter_frac=frac, hard=hard)
        try:
            qz(t0.clone().requires_grad_(True), iter_frac=frac, hard=hard).sum().backward()
            raise SystemExit("the reference's backward unexpectedly works: regenerate the gradients from it")
        except RuntimeError as e:
            assert "inplace" in str(e)
        tau = R.QZ.tau_iter(method, frac, 0.5, 3.0, None)
        t = t0.clone().requires_grad_(True)
        qo = DO.softmax_quantize(t, lut_full[:-1], torch.tensor(tau, dtype=torch.float32), 3.0, 300., gumbel_noise=noise, hard=hard).squeeze(0, 1)
        assert torch.equal(qo.detach(), q), name            # the restatement reproduces the reference forward bit for bit
        (gt,) = torch.autograd.grad(qo, t, gq)
        # the same restatement in float64: how far fp32 itself is from the exact gradient (the soft branches at small tau are
        # ill-conditioned -- logits of O(1000) -- and the tests budget the GPU's distance with this number)
        t64 = t0.double().requires_grad_(True)
        q64 = DO.softmax_quantize(t64, lut_full[:-1].double(), float(np.float32(tau)), 3.0, 300.,
                                  gumbel_noise=None if noise is None else noise.double(), hard=hard).squeeze(0, 1)
        (gt64,) = torch.autograd.grad(q64, t64, gq.double())
        out["gt64_" + name] = gt64
        out["q_" + name], out["gt_" + name] = q, gt
        out["tau_" + name] = np.float64(tau)
        out["frac_" + name] = np.float64(frac)
    # a map with many equal values: the global max |diff| of score_thickness has ties, autograd splits its gradient evenly
    tt = torch.zeros(1, 1, 8, 8)
    tt[0, 0, 2:5, 3:6] = 0.4 * hmax
    qz = R.QZ.Quantization(method="softmax", max_thickness=hmax, num_bits=bits, dev=torch.device("cpu"), tau_min=0.5, tau_max=3.0, c=300.)
    gq2 = torch.randn(8, 8)
    with torch.no_grad():
        q2 = qz(tt.clone(), iter_frac=0.6, hard=True)
    tau2 = R.QZ.tau_iter("softmax", 0.6, 0.5, 3.0, None)
    tg = tt.clone().requires_grad_(True)
    qo2 = DO.softmax_quantize(tg, lut_full[:-1], torch.tensor(tau2, dtype=torch.float32), 3.0, 300., hard=True).squeeze(0, 1)
    assert torch.equal(qo2.detach(), q2)
    (gt2,) = torch.autograd.grad(qo2, tg, gq2)
    save("quant_softmax", thickness=t0, gq=gq, lut=lut_full, hmax=np.float64(hmax), tau_max=np.float64(3.0),
         tau_min=np.float64(0.5), c=np.float64(300.), ties_t=tt, ties_gq=gq2, ties_q=q2, ties_gt=gt2,
         ties_tau=np.float64(tau2), **out)


def gen_czt(R):
    cases = {
        "czt_small": (64, 64, 32, [1 * mm, 1.1 * mm], 0.5 * mm, 0.1 * mm, 0.5),
        "czt_rect_in": (96, 80, 48, [1 * mm], 1 * mm, 0.25 * mm, 0.3),
        "czt_zoom_out": (40, 40, 64, [0.9 * mm, 1 * mm, 1.05 * mm], 0.5 * mm, 0.2 * mm, 0.2),
        "czt_testscript": (200, 200, 200, [1 * mm], 1 * mm, 1 * mm, 0.5),       # test_czt.py:17-37 geometry, fp32
    }
    for name, (H, W, M, lams, dx, dxo, z) in cases.items():
        torch.manual_seed(zlib.crc32(name.encode()) % 1000)
        if name == "czt_testscript":
            # Gaussian beam, waist 2 mm, as test_czt.py builds it
            xs = (torch.arange(H) - H / 2 + 0.5) * dx
            X, Y = torch.meshgrid(xs, xs, indexing="ij")
            x = torch.exp(-(X ** 2 + Y ** 2) / (2 * mm) ** 2).to(torch.complex64)[None, None]
        else:
            x = torch.randn(1, len(lams), H, W, dtype=torch.complex64)
        f = R.ElectricField(x, wavelengths=lams, spacing=dx, device=torch.device("cpu"))
        czt = R.CZT_prop(z_distance=z, device=torch.device("cpu"))
        with quiet():
            y = czt(f, outputHeight=M, outputWidth=M, outputPixel_dx=dxo, outputPixel_dy=dxo)
        save(name, x=x, y=y.data, wavelengths=np.array(lams), spacing=np.array([dx, dx]), z=np.float64(z), M=np.int64(M),
             out_dx=np.float64(dxo), out_spacing=y.spacing)


def gen_train(R):
    """Loss / optimizer lines of the notebook loop: the reference's own normalize + nn.MSELoss, torch's own Adam / AdamW."""
    torch.manual_seed(7)
    for name, (B, C, H, W) in {"train_loss_single": (1, 1, 40, 40), "train_loss_batch": (3, 2, 24, 16)}.items():
        y = torch.randn(B, C, H, W, dtype=torch.complex64, requires_grad=True)
        target = torch.rand(B, C, H, W)
        out_amp = R.HF.normalize(torch.abs(y) ** 2)
        loss = torch.nn.MSELoss()(out_amp, target)
        (gy,) = torch.autograd.grad(loss, y)
        save(name, y=y, target=target, loss=loss, gy=gy)
    p0 = torch.randn(50, 30)
    grads = [torch.randn(50, 30) * (0.5 + 0.1 * i) for i in range(6)]
    for name, (wd, dec) in {"train_adam": (0.0, False), "train_adam_wd": (0.05, False), "train_adamw": (0.01, True)}.items():
        p = torch.nn.Parameter(p0.clone())
        opt = (torch.optim.AdamW if dec else torch.optim.Adam)([p], lr=0.02, weight_decay=wd)
        hist = []
        for g in grads:
            p.grad = g.clone()
            opt.step()
            hist.append(p.detach().clone())
        save(name, p0=p0, grads=torch.stack(grads), hist=torch.stack(hist), weight_decay=np.float64(wd), decoupled=np.int64(dec))


def gen_rsc(R):
    """Rayleigh-Sommerfeld convolution (Props/RSC_Prop.py): scalar and vectorial, forward output + input gradient."""
    cpu = torch.device("cpu")
    cases = {"rsc_pow2": (1, 2, 64, 64, [1 * mm, 1.05 * mm], 0.5 * mm, 0.1), "rsc_rect": (1, 1, 48, 40, [1 * mm], [1 * mm, 0.7 * mm], 0.2),
             "rsc_notebook": (1, 1, 100, 100, [1 * mm], 1 * mm, 0.3)}
    for name, (B, C, H, W, lams, dxy, z) in cases.items():
        torch.manual_seed(zlib.crc32(name.encode()) % 1000)
        x = torch.randn(B, C, H, W, dtype=torch.complex64, requires_grad=True)
        g = torch.randn(B, C, H, W, dtype=torch.complex64)
        prop = R.RSC_prop(z_distance=z, device=cpu)
        prop.check_Zc = False      # check_RS_minimum_z (Props/RSC_Prop.py:117) crashes when its energy-conservation bound is <= 0
        with quiet():
            y = prop(R.ElectricField(x, wavelengths=lams, spacing=dxy, device=cpu)).data
        (gx,) = torch.autograd.grad(y, x, g)
        save(name, x=x, y=y, g=g, gx=gx, wavelengths=np.array(lams), spacing=np.array(dxy if isinstance(dxy, list) else [dxy, dxy]), z=np.float64(z))
    torch.manual_seed(9)
    x = torch.randn(3, 2, 32, 32, dtype=torch.complex64)
    prop = R.VRS_prop(z_distance=0.08, device=cpu)
    prop.check_Zc = False
    with quiet():
        y = prop(R.ElectricField(x, wavelengths=[1 * mm, 1.1 * mm], spacing=0.5 * mm, device=cpu)).data
    save("rsc_vectorial", x=x, y=y, wavelengths=np.array([1 * mm, 1.1 * mm]), spacing=np.array([0.5 * mm, 0.5 * mm]), z=np.float64(0.08))


def gen_elements(R):
    """Thin lens and apertures (Components/Thin_Lens.py, Components/Aperture.py): forward output + input gradient."""
    cpu = torch.device("cpu")
    torch.manual_seed(21)
    x = torch.randn(2, 2, 33, 40, dtype=torch.complex64, requires_grad=True)
    g = torch.randn(2, 2, 33, 40, dtype=torch.complex64)
    lams, sp = [1 * mm, 1.2 * mm], [0.5 * mm, 0.4 * mm]
    f = R.ElectricField(x, wavelengths=lams, spacing=sp, device=cpu)
    out = {}
    for name, el in {"lens": R.Thin_LensElement(focal_length=0.127), "circ": R.ApertureElement("circ", 6 * mm),
                     "rect": R.ApertureElement("rect", 8 * mm)}.items():
        el.device = cpu
        if name == "lens":
            el.focal_length = el.focal_length.cpu()
        y = el(f).data
        (gx,) = torch.autograd.grad(y, x, g)
        out["y_" + name], out["gx_" + name] = y, gx
    save("elem_lens_aperture", x=x, g=g, wavelengths=np.array(lams), spacing=np.array(sp), focal=np.float64(0.127),
         radius=np.float64(6 * mm), side=np.float64(8 * mm), **out)


def gen_setup(R):
    """The notebooks' whole set-up (experiment_four_focal_spots.ipynb cells 2-8, `Submm_Setupv2` with the STE layer):
    Gaussian beam -> ASM 127 mm -> thin lens -> rect aperture -> 4-level STE DOE -> ASM 200 mm -> normalize(|y|^2) -> MSE."""
    cpu = torch.device("cpu")
    n, dxy, lam = 100, 1 * mm, 2.998e8 / 300e9
    src = R.Guassian_beam(height=n, width=n, beam_waist_x=None, beam_waist_y=None, wavelengths=lam, spacing=dxy, device=cpu)
    asm1 = R.ASM_prop(z_distance=0.127, bandlimit_type="exact", padding_scale=2, bandlimit_kernel=True, device=cpu)
    lens = R.Thin_LensElement(focal_length=0.127)
    lens.device, lens.focal_length = cpu, lens.focal_length.cpu()
    ap = R.ApertureElement(aperture_type="rect", aperture_size=0.08)
    ap.device = cpu
    torch.manual_seed(31)
    doe = R.QD.STEQuantizedDOELayer(dict(doe_size=[n, n], doe_dxy=dxy, doe_level=4, look_up_table=None, num_unit=None,
                                         height_constraint_max=1 * mm, tolerance=None, material=[2.66, 0.03]), {}, device=cpu)
    asm3 = R.ASM_prop(z_distance=0.2, bandlimit_type="exact", padding_scale=2, bandlimit_kernel=True, device=cpu)
    target = torch.rand(1, 1, n, n)
    with quiet():
        s0 = src()               # (the reference's source mutates its waists in forward: call it once)
        source = s0.data.clone()
        fin = ap(lens(asm1(s0)))
        out = asm3(doe(fin, None))
        loss = torch.nn.MSELoss()(R.HF.normalize(torch.abs(out.data) ** 2), target)
        (gw,) = torch.autograd.grad(loss, doe.weight_height_map)
    save("setup_four_focal_spots", source=source, field_before_doe=fin.data, w0=doe.weight_height_map.detach(), out=out.data, target=target,
         loss=loss, gw=gw, height_map=doe.height_map.detach(), wavelength=np.float64(lam), spacing=np.float64(dxy))


class _FakeH5File:
    """Just enough of h5py.File for the reference's VisTools/calc_loss.py (h5py is not installed here): datasets are numpy
    arrays in a per-path dict; f[k] = arr, f[k][:], f[k][:] = arr, flush(), context manager."""
    store = {}

    def __init__(self, path, mode="r"):
        self.d = _FakeH5File.store.setdefault(path, {}) if mode != "w" else _FakeH5File.store.__setitem__(path, {}) or _FakeH5File.store[path]

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __setitem__(self, k, v):
        self.d[k] = np.array(v)

    def __getitem__(self, k):
        return self.d[k]

    def flush(self):
        pass


def gen_landscape(R):
    """The reference's own calulate_single_element_loss_landscape (VisTools/calc_loss.py:8-55), unmodified, on a small
    single-DOE system built from the reference's modules -- run over an in-memory stand-in for h5py.  A non-square grid
    (5 x 4) pins the reference's index convention (meshgrid(x, y) coordinates written into an (xnum, ynum) array)."""
    import importlib
    import types
    fake = types.ModuleType("h5py")
    fake.File = _FakeH5File
    sys.modules["h5py"] = fake
    CL = importlib.import_module("VisTools.calc_loss")
    importlib.reload(CL)
    n, lam, dxy = 24, 1 * mm, 0.5 * mm
    torch.manual_seed(51)
    x = torch.randn(1, 1, n, n, dtype=torch.complex64)

    class Setup(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.input_field = R.ElectricField(data=x, wavelengths=lam, spacing=dxy)
            with quiet():
                self.doe = R.QD.STEQuantizedDOELayer(dict(doe_size=[n, n], doe_dxy=dxy, doe_level=4, look_up_table=None, num_unit=None,
                                                          height_constraint_max=1 * mm, tolerance=None, material=[2.66, 0.03]), {})
                self.asm_prop3 = R.ASM_prop(z_distance=0.05, bandlimit_type='exact', padding_scale=None, bandlimit_kernel=True)

        def forward(self, iter_frac):
            return self.asm_prop3(self.doe(self.input_field, iter_frac))

    torch.manual_seed(52)
    model = Setup()
    w0 = [p.data.clone() for p in model.parameters()]
    torch.manual_seed(53)
    directions = []
    for _ in range(2):
        d = [torch.randn(w.size()) for w in w0]
        for di, wi in zip(d, w0):
            di.mul_(wi.norm() / (di.norm() + 1e-10))
        directions.append(d)
    target = torch.rand(1, 1, n, n)
    args = types.SimpleNamespace(xmin=-1.0, xmax=1.0, xnum=5, ymin=-0.5, ymax=0.5, ynum=4)
    with quiet():
        path = CL.calulate_single_element_loss_landscape(args, model, target, loss_f=torch.nn.MSELoss(), directions=directions,
                                                         save_path="/tmp/thz_landscape_golden")
    d = _FakeH5File.store[path]
    save("landscape_single_doe", x=x, w0=w0[0], dx=directions[0][0], dy=directions[1][0], target=target, wavelength=np.float64(lam),
         spacing=np.float64(dxy), z=np.float64(0.05), xcoordinates=d["xcoordinates"], ycoordinates=d["ycoordinates"], loss=d["loss"],
         args=np.array([args.xmin, args.xmax, args.xnum, args.ymin, args.ymax, args.ynum]), nparams=np.int64(len(w0)))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    R = import_reference()
    gen_asm(R)
    gen_doe(R)
    gen_quant(R)
    gen_quant_softmax(R)
    gen_czt(R)
    gen_train(R)
    gen_rsc(R)
    gen_elements(R)
    gen_setup(R)
    gen_landscape(R)
