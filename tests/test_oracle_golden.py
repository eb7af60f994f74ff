"""The oracle (oracle/*.py, a torch-CPU restatement of the reference) against the committed golden
vectors produced by the reference itself (oracle/make_golden.py).  Runs anywhere, no GPU."""
import pytest
import torch

from helpers import asm_case_kwargs, golden, golden_names, rel_l2
from oracle import asm_oracle as AO
from oracle import czt_oracle as CO
from oracle import doe_oracle as DO


@pytest.mark.parametrize("name", golden_names("asm_"))
def test_asm_oracle_matches_reference_vectors(name):
    g = golden(name)
    kw = asm_case_kwargs(g)
    x = g["x"].clone().requires_grad_(True)
    y = AO.asm_forward(x, g["wavelengths"].float(), g["spacing"].float(), g["z"], **kw)
    assert torch.equal(torch.view_as_real(y.detach()), torch.view_as_real(g["y"]))     # bit-exact
    (gx,) = torch.autograd.grad(y, x, g["g"])
    assert rel_l2(gx, g["gx"]) == 0.0
    if "kernel" in g:
        _, _, Hp, Wp = AO.compute_padding(x.shape[-2], x.shape[-1], kw["padding_scale"], kw["do_padding"])
        k = AO.centred_transfer_function(Hp, Wp, g["spacing"].float(), g["wavelengths"].float(), g["z"], True, kw["bandlimit_type"])
        assert torch.equal(torch.view_as_real(k), torch.view_as_real(g["kernel"]))


def test_critical_distance_print_value():
    # experiment_four_focal_spots.ipynb cell 7 prints 0.26003873 m for 100^2, pitch 1 mm, padding_scale 2, 300 GHz
    lam = 2.998e8 / 300e9
    assert abs(AO.critical_distance(300, [1e-3, 1e-3], [lam]) - 0.26003873) < 1e-6


def _height_for(name, g):
    w = g["w"].clone().requires_grad_(True)
    hmax, L = g["hmax"], int(g["levels"])
    idx = None
    if name == "doe_ste":
        h = DO.ste_quantize(DO.sigmoid_height(w[0, 0], hmax), g["lut"])
    elif name == "doe_fullprecision":
        h = DO.sigmoid_height(w[0, 0], hmax)
    elif name == "doe_psq":
        tau = 1 + (400 - 1) * g["iter_frac"]
        h = DO.psq_height(w, hmax, L, tau)
    elif name.startswith("doe_gumbel_v3") or name == "doe_gumbel_v2":
        fr = g["iter_frac"]
        if name == "doe_gumbel_v2":   # v2 = v3 with a hard switch at 0.5 (QuantizedDOE.py:608-618)
            fr3 = 0.9 if fr > 0.5 else 0.2
            tau = DO.cosine_tau(fr, g["tau_min"], g["tau_max"])
            h, idx = _v3_with_tau(w, g, tau, fr3)
        else:
            h, idx = DO.score_gumbel_v3_height(w, g["lut"], hmax, g["wavelengths"].float().min(), g["material"][0].float(),
                                               g["c_s"], g["tau_min"], g["tau_max"], fr, g["noise"])
    elif name == "doe_gumbel_naive":
        tau = DO.cosine_tau(g["iter_frac"], g["tau_min"], g["tau_max"])
        h, idx = DO.naive_gumbel_height(w, g["lut"], tau, g["noise"])
    else:
        raise KeyError(name)
    return w, h


def _v3_with_tau(w, g, tau, fr3):
    """score_gumbel_v3_height with an externally supplied tau (v2 uses the same maths, other schedule)."""
    import math
    # invert the cosine schedule so that the oracle's internal cosine_tau(fr') returns `tau`
    tmin, tmax = g["tau_min"], g["tau_max"]
    orig = DO.cosine_tau
    DO.cosine_tau = lambda *_a, **_k: tau
    try:
        return DO.score_gumbel_v3_height(w, g["lut"], g["hmax"], g["wavelengths"].float().min(), g["material"][0].float(),
                                         g["c_s"], tmin, tmax, fr3, g["noise"])
    finally:
        DO.cosine_tau = orig


@pytest.mark.parametrize("name", [n for n in golden_names("doe_") if n not in ("doe_fix_edoe4", "doe_gumbel_v1")])
def test_doe_oracle_matches_reference_vectors(name):
    g = golden(name)
    w, h = _height_for(name, g)
    assert torch.equal(h.detach(), g["height_map"])
    eps, tand = g["material"][0].float(), g["material"][1].float()
    lam, sp = g["wavelengths"].float(), g["spacing"].float()
    u = DO.modulate(g["x"], h, lam, eps, tand)
    assert rel_l2(u.detach(), g["u"]) == 0.0
    y = AO.asm_forward(u, lam, sp, g["z"])
    assert rel_l2(y.detach(), g["y"]) == 0.0
    (gw,) = torch.autograd.grad(y, w, g["g"])
    if g["gw"].abs().max() > 0:
        assert rel_l2(gw, g["gw"]) < 1e-6
    else:
        assert gw.abs().max() == 0


def test_fix_doe_on_reference_height_map():
    g = golden("doe_fix_edoe4")
    h = g["height_map"].clone().requires_grad_(True)
    eps, tand = g["material"][0].float(), g["material"][1].float()
    u = DO.modulate(g["x"], h, g["wavelengths"].float(), eps, tand)
    y = AO.asm_forward(u, g["wavelengths"].float(), g["spacing"].float(), g["z"], padding_scale=2)
    assert rel_l2(y.detach(), g["y"]) == 0.0
    (gh,) = torch.autograd.grad(y, h, g["g"])
    assert rel_l2(gh, g["gh"]) < 1e-6
    # closed-form gradients the CUDA adjoint implements
    gprime = AO.asm_adjoint(g["g"], g["wavelengths"].float(), g["spacing"].float(), g["z"], padding_scale=2)
    _, gh_cf = DO.modulate_grads(g["x"], g["height_map"], g["wavelengths"].float(), eps, tand, gprime)
    assert rel_l2(gh_cf, g["gh"]) < 2e-6


def test_quantizer_known_answers():
    g = golden("quant_ste")
    # Components/test_all.ipynb cell 21: input [0.1,0.4,0.7,1.2], lut [0,0.5,1.0] -> [0,0.5,0.5,1.0], grad ones
    x = g["kat_x"].clone().requires_grad_(True)
    q = DO.ste_quantize(x, g["kat_lut"])
    assert q.tolist() == [0.0, 0.5, 0.5, 1.0] == g["kat_q"].tolist()
    (gr,) = torch.autograd.grad(q.sum(), x)
    assert torch.equal(gr, g["kat_grad"]) and torch.equal(gr, torch.ones(4))
    assert torch.equal(DO.ste_indices(g["h"], g["lut"]), g["idx"])
    n = golden("quant_nn")
    assert DO.nearest_idx(n["kat_x"], n["mid"]).tolist() == n["kat_idx"].tolist() == [0, 0, 0, 1, 1, 1, 2, 3, 0, 0, 0, 0]
    idx = DO.nearest_idx(n["x"], n["mid"])
    q = n["lut"][idx]
    for kind in ("nn", "nn_poly", "nn_sigmoid"):
        assert torch.equal(q, n["q_" + kind])
        gr = DO.nn_quantize_backward(n["x"], q, idx, n["lut"], n["s"], torch.ones_like(q), kind)
        assert torch.equal(gr, n["grad_" + kind])
    assert torch.equal(DO.linear_lut(4.0, 3), torch.linspace(0, 4, 4)[:-1])


@pytest.mark.parametrize("name", golden_names("czt_"))
def test_czt_oracle_matches_reference_vectors(name):
    g = golden(name)
    M = int(g["M"])
    args = (g["x"], g["wavelengths"].float(), g["spacing"].float(), torch.tensor(g["z"], dtype=torch.float32), M, M, g["out_dx"], g["out_dx"])
    y = CO.czt_forward(*args)
    assert torch.equal(torch.view_as_real(y), torch.view_as_real(g["y"]))            # FFT restatement: bit-exact
    yd = CO.czt_forward_dense(*args)
    assert rel_l2(yd, g["y"]) < 1e-5                                                   # dense separable form the GEMM path uses


@pytest.mark.parametrize("name", golden_names("train_loss"))
def test_loss_oracle_matches_reference_vectors(name):
    """normalize + MSELoss restatement vs the reference's own normalize (utils/Helper_Functions.py:185-193)."""
    from oracle import train_oracle as TO
    g = golden(name)
    loss, gy = TO.normalized_intensity_mse(g["y"], g["target"])
    assert abs(float(loss) - g["loss"]) <= 1e-7 * abs(g["loss"])
    assert rel_l2(gy, g["gy"]) <= 1e-7


@pytest.mark.parametrize("name", golden_names("train_adam"))
def test_adam_oracle_matches_torch_optimizers(name):
    from oracle import train_oracle as TO
    g = golden(name)
    hist = TO.adam_reference(g["p0"], list(g["grads"]), lr=0.02, weight_decay=g["weight_decay"], decoupled=bool(g["decoupled"]))
    assert torch.equal(torch.stack(hist), g["hist"])


@pytest.mark.parametrize("name", golden_names("rsc_"))
def test_rsc_oracle_matches_reference_vectors(name):
    """Rayleigh-Sommerfeld convolution restatement vs the reference's RSC_prop / VRS_prop (Props/RSC_Prop.py)."""
    from oracle import rsc_oracle as RO
    g = golden(name)
    fwd = RO.vrs_forward if name == "rsc_vectorial" else RO.rsc_forward
    x = g["x"].clone().requires_grad_(True)
    y = fwd(x, g["wavelengths"].float(), g["spacing"].float(), g["z"])
    assert rel_l2(y.detach(), g["y"]) <= 2e-6
    if "gx" in g:
        (gx,) = torch.autograd.grad(y, x, g["g"])
        assert rel_l2(gx, g["gx"]) <= 2e-6


def test_element_oracle_matches_reference_vectors():
    """Thin lens kernel and aperture masks vs the reference's Thin_LensElement / ApertureElement outputs."""
    from oracle import element_oracle as EO
    g = golden("elem_lens_aperture")
    x, sp = g["x"], g["spacing"].float()
    H, W = x.shape[-2:]
    assert rel_l2(x * EO.lens_kernel(H, W, sp, g["wavelengths"].float(), g["focal"]), g["y_lens"]) <= 1e-6
    assert torch.equal(x * EO.circ_mask(H, W, sp, g["radius"]), g["y_circ"])
    assert torch.equal(x * EO.rect_mask(H, W, sp, g["side"]), g["y_rect"])


def test_gaussian_source_mirror_matches_reference_field():
    """LightSource mirror (pure torch, evaluated once before the loop) vs the field the reference's Guassian_beam produced."""
    from quantizationawarethzdoe_b200 import Guassian_beam
    g = golden("setup_four_focal_spots")
    src = Guassian_beam(height=100, width=100, beam_waist_x=None, beam_waist_y=None, wavelengths=g["wavelength"], spacing=g["spacing"],
                        device=torch.device("cpu"))
    # fitted waist: a degree-5 polynomial in fp32 with heavy cancellation, host-libm dependent at the 1e-5 level
    assert rel_l2(src().data, g["source"]) <= 1e-4
    assert torch.equal(src().data, src().data)               # idempotent (the reference's forward is not)


@pytest.mark.parametrize("name", ["gumbel_hard", "gumbel_soft", "plain_hard", "plain_soft"])
def test_softmax_quantizer_oracle_matches_reference_forward(name):
    """SoftmaxBasedQuantization / score_thickness restatement vs the reference's own forward output, bit for bit (the
    reference's backward raises -- in-place normalisation, quantization.py:41 -- so only the forward is reference-pinned)."""
    g = golden("quant_softmax")
    noise = g.get("noise_" + name)
    tau = torch.tensor(g["tau_" + name], dtype=torch.float32)
    q = DO.softmax_quantize(g["thickness"], g["lut"][:-1], tau, g["tau_max"], g["c"], gumbel_noise=noise, hard="hard" in name)
    assert torch.equal(q.squeeze(0, 1), g["q_" + name])
