"""CPU replay of the CUDA kernel bodies (tests/emul/emul.cpp) against the oracle.

This is the `not gpu` check of everything that can go wrong in the kernels short of the hardware:
plan factorisation, digit-reversal bookkeeping, zero-pad / crop pruning, anisotropic and odd paddings,
chunking over fields, the fused DOE prologue and the DOE adjoint epilogue with its register
accumulators, and the transfer function in both modes.  The same source compiles for sm_100a."""
import ctypes

import pytest
import torch

from helpers import emul_lib, rel_l2
from oracle import asm_oracle as AO
from oracle import doe_oracle as DO
from quantizationawarethzdoe_b200 import _native as N
from quantizationawarethzdoe_b200 import asm_host as AH

E = emul_lib()
TOL = 1e-5       # north_star tolerance on complex64 fields and gradients
TOL_TABLE = 2e-6  # cached-H mode differs from the reference only by FFT round-off


def _run(x, desc_kwargs, sm=148):
    d = AH.build_desc(**desc_kwargs)
    rc = E.thz_emul_asm_propagate(ctypes.byref(d), sm)
    assert rc == 0, rc


def _setup(B, C, H, W, scale, lams, dxy, z, bt="exact", do_pad=True, unpad=True, mode=0, chunk=0):
    ps = AH.normalise_padding_scale(scale, do_pad)
    ph, pw, Hp, Wp = AH.compute_padding(H, W, ps, do_pad)
    outH, outW, or0, oc0 = (H, W, ph, pw) if (do_pad and unpad) else (Hp, Wp, 0, 0)
    chunked = AH.row_vectors_chunked(Hp)      # the static column kernels read the row vectors in the chunked layout
    rv, cv, sc = AH.tf_device_vectors(*AH.tf_vectors(Hp, Wp, dxy, lams, z, True, bt), E.thz_emul_slot_to_bin, chunked=chunked)
    table = None
    if mode == 1:
        table = AH.tf_table_slot_order(AH.tf_centred_reference_order(Hp, Wp, dxy, lams, z, True, bt), E.thz_emul_slot_to_bin)
    base = dict(B=B, C=C, inH=H, inW=W, Hp=Hp, Wp=Wp, in_r0=ph, in_c0=pw, outH=outH, outW=outW, out_r0=or0, out_c0=oc0,
                tf_mode=mode, tf_conj=0, rowvec=rv, colvec=cv, scal=sc, table=table, doe_mode=0, doe_base=0.0, hmap=None,
                coef=None, xsaved=None, gh=None, tw_h=N.twiddles_host(Hp), tw_w=N.twiddles_host(Wp),
                ws=torch.zeros(AH.workspace_elems(B, C, max(H, outH), max(H, outH), Wp, chunk, Hp=Hp), dtype=torch.complex64), bc_chunk=chunk,
                tf_row_chunked=1 if (chunked and mode == 0) else 0)
    return base, (outH, outW)


CASES = [
    # B, C, H, W, scale, wavelengths, spacing, z, kwargs
    (1, 1, 32, 32, None, [1e-3], 0.5e-3, 0.1, {}),
    (2, 2, 64, 64, None, [1e-3, 1.01e-3], 0.5e-3, 0.1, {}),
    (1, 1, 50, 50, 2, [1e-3], [1e-3, 0.7e-3], 0.26, dict(mode=1)),                      # 150 = 25*6
    (1, 3, 30, 36, [1, 2], [0.9e-3, 1e-3, 1.2e-3], 0.5e-3, 0.05, dict(bt="approx")),     # 60 x 108
    (1, 1, 100, 100, 2, [1e-3], 1e-3, 0.3, dict(mode=1)),                                # notebook geometry 100 -> 300
    (3, 2, 40, 24, None, [1e-3, 1.1e-3], 0.5e-3, 0.1, dict(unpad=False, chunk=4)),       # padded output, ragged chunks
    (1, 1, 48, 48, None, [1e-3], 0.5e-3, 0.1, dict(do_pad=False, mode=1)),
    (1, 1, 200, 200, None, [1e-3], 0.5e-3, 0.1, dict(mode=1)),                           # 400 = 25*16 (config 4 layer size)
    (1, 1, 256, 256, None, [1e-3], 0.5e-3, 0.1, {}),
    (1, 1, 7, 21, 2, [1e-3], 0.5e-3, 0.02, dict(mode=1)),                                # odd sizes: 21 x 63 (radix 7, 3)
    # power-of-two fast path (compile-time specialised kernels, Hp/Wp in 256..16384)
    (2, 2, 128, 128, None, [1e-3, 1.01e-3], 0.5e-3, 0.1, dict(mode=1)),                  # 256 x 256: two stages
    (1, 1, 256, 512, None, [1e-3], 0.5e-3, 0.1, dict(mode=1)),                           # 512 x 1024: three stages, 8-col tiles
    (1, 1, 128, 1024, None, [1e-3], 0.5e-3, 0.1, {}),                                    # 256 x 2048, in-register H
    (1, 1, 100, 128, None, [1e-3], 0.5e-3, 0.1, dict(mode=1)),                           # mixed: generic rows (200) + p2 cols
    (3, 1, 256, 256, None, [1e-3], 0.5e-3, 0.1, dict(mode=1, unpad=False, chunk=2)),     # 512 x 512 padded output, chunks
    # static path for mixed-radix lengths whose slot offsets stay compile-time constants
    (2, 1, 200, 200, None, [1e-3], 0.5e-3, 0.05, {}),                                    # 400 x 400 = (25*16)^2, in-register H
    (1, 1, 1000, 128, None, [1e-3], 0.5e-3, 0.1, dict(mode=1)),                          # 2000 (25*20*4) x 256
    (1, 2, 128, 400, None, [1e-3, 1.1e-3], 0.5e-3, 0.1, dict(mode=1)),                   # 256 x 800 (25*16*2)
    (1, 1, 256, 128, 2, [1e-3], 0.5e-3, 0.1, dict(mode=1)),                              # padding_scale 2: 768 (12*8*8) x 384 (generic)
    (1, 1, 512, 256, 2, [1e-3], 0.5e-3, 0.1, {}),                                        # 1536 (12*16*8) x 768, in-register H
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d_s%s" % (c[2], c[3], c[4]))
def test_asm_forward_replay_matches_oracle(case):
    B, C, H, W, scale, lams, dxy, z, kw = case
    torch.manual_seed(0)
    x = torch.randn(B, C, H, W, dtype=torch.complex64)
    base, (outH, outW) = _setup(B, C, H, W, scale, lams, dxy, z, **kw)
    y = torch.zeros(B, C, outH, outW, dtype=torch.complex64)
    _run(x, dict(base, x=x, y=y))
    yo = AO.asm_forward(x, lams, dxy, z, padding_scale=scale, do_padding=kw.get("do_pad", True),
                        do_unpad_after_pad=kw.get("unpad", True), bandlimit_type=kw.get("bt", "exact"))
    # in-register H on tiny grids is dominated by the handful of bins where torch's CPU sqrt is 1 ulp off
    tol = TOL_TABLE if kw.get("mode", 0) == 1 else 2e-5
    assert rel_l2(y, yo) < tol


def test_in_register_mask_is_bit_exact():
    """Where the kernel zeroes H must be exactly where the reference zeroes it."""
    H = W = 64
    lams, dxy, z = [1e-3, 1.3e-3], 0.5e-3, 0.4
    base, _ = _setup(1, 2, H, W, None, lams, dxy, z)
    # propagate a delta at the canvas origin region: spectrum is flat, so |fft2(y_padded)| shows the mask
    Hp = Wp = 128
    base.update(inH=Hp, inW=Wp, in_r0=0, in_c0=0, outH=Hp, outW=Wp, out_r0=0, out_c0=0,
                ws=torch.zeros(2 * Hp * Wp, dtype=torch.complex64))
    x = torch.zeros(1, 2, Hp, Wp, dtype=torch.complex64)
    x[..., 0, 0] = 1
    y = torch.zeros_like(x)
    _run(x, dict(base, x=x, y=y))
    spec = torch.fft.fft2(y)
    Hc = AO.centred_transfer_function(Hp, Wp, dxy, lams, z)
    Hn = torch.fft.ifftshift(Hc, dim=(-2, -1))
    assert torch.equal(spec.abs() > 0.5, Hn.abs() > 0.5)
    assert rel_l2(spec, Hn) < TOL


@pytest.mark.parametrize("B,C,H,W,scale,mode,sm,chunk", [
    (1, 1, 32, 32, None, 1, 148, 0),
    (2, 3, 48, 40, None, 1, 148, 0),
    (2, 3, 48, 40, None, 1, 148, 4),      # gh accumulated across chunks (atomic path)
    (5, 1, 50, 50, 2, 1, 4, 0),           # few SMs: one CTA walks all five fields, register accumulators
    (1, 2, 64, 64, None, 0, 148, 0),
    (3, 2, 128, 256, None, 1, 148, 0),    # power-of-two fast path: 256 x 512
    (3, 2, 128, 256, None, 1, 2, 4),      # same, few SMs (one CTA walks several fields) and chunks (atomic gh)
    (1, 1, 512, 128, None, 0, 148, 0),    # 1024 x 256
    (3, 1, 200, 200, None, 1, 148, 0),    # 400 x 400 static mixed-radix path (DONN layer size), DOE fwd + adjoint
])
def test_doe_fused_forward_and_adjoint_replay(B, C, H, W, scale, mode, sm, chunk):
    lams = [1e-3, 1.02e-3, 1.05e-3][:C]
    dxy, z, eps, tand = 0.5e-3, 0.1, 2.66, 0.003
    torch.manual_seed(0)
    x = torch.randn(B, C, H, W, dtype=torch.complex64)
    h = torch.rand(H, W) * 1e-3
    g = torch.randn(B, C, H, W, dtype=torch.complex64)
    base, _ = _setup(B, C, H, W, scale, lams, dxy, z, mode=mode, chunk=chunk)
    coef = AH.doe_coefficients(lams, eps, tand)
    y = torch.zeros_like(x)
    _run(x, dict(base, x=x, y=y, doe_mode=1, doe_base=DO.BASE_PLANE_THICKNESS, hmap=h, coef=coef), sm)
    xr, hr = x.clone().requires_grad_(True), h.clone().requires_grad_(True)
    yo = AO.asm_forward(DO.modulate(xr, hr, lams, eps, tand), lams, dxy, z, padding_scale=scale)
    gxo, gho = torch.autograd.grad(yo, (xr, hr), g)
    gx, gh = torch.zeros_like(x), torch.full((H, W), 7.0)
    _run(g, dict(base, x=g, y=gx, tf_conj=1, doe_mode=2, doe_base=DO.BASE_PLANE_THICKNESS, hmap=h, coef=coef, xsaved=x, gh=gh), sm)
    tol = TOL_TABLE if mode == 1 else TOL
    assert rel_l2(y, yo.detach()) < tol
    assert rel_l2(gx, gxo) < tol
    assert rel_l2(gh, gho) < tol


@pytest.mark.parametrize("H,W", [(16, 16), (64, 48), (60, 100), (8, 250), (35, 27)])
@pytest.mark.parametrize("inverse", [0, 1])
def test_fft2_replay_matches_torch(H, W, inverse):
    torch.manual_seed(1)
    x = torch.randn(3, H, W, dtype=torch.complex64)
    y = torch.zeros_like(x)
    ws = torch.zeros_like(x)
    rc = E.thz_emul_fft2_c2c(N.ptr(x), N.ptr(y), 3, H, W, inverse, 1, N.ptr(N.twiddles_host(H)), N.ptr(N.twiddles_host(W)), N.ptr(ws))
    assert rc == 0
    ref = (torch.fft.ifft2 if inverse else torch.fft.fft2)(x, norm="ortho")
    assert rel_l2(y, ref) < 2e-6


def test_unsupported_length_is_reported():
    radices = (ctypes.c_int32 * 16)()
    ns = ctypes.c_int32()
    assert E.thz_emul_plan_info(26, radices, ctypes.byref(ns)) == -3      # 13 is not a supported radix
    assert E.thz_emul_plan_info(4096, radices, ctypes.byref(ns)) == 0 and list(radices)[:3] == [16, 16, 16]
    assert E.thz_emul_plan_info(2000, radices, ctypes.byref(ns)) == 0 and list(radices)[:3] == [25, 20, 4]


# ------------------------------------------------------------------------------- slab FFT over peer memory
def _slab_plans(G, C, H, W, lams, dxy, z, device, mode="inregister"):
    from quantizationawarethzdoe_b200 import parallel as P
    pad_h, pad_w, Hp, Wp = AH.compute_padding(H, W, AH.normalise_padding_scale(None, True), True)
    sp, wl = torch.tensor([dxy, dxy]), torch.tensor(lams)
    rowvec, colvec, scal = AH.tf_vectors(Hp, Wp, sp, wl, torch.tensor(z), True, "exact")
    table, tf_mode = None, 0
    chunked = AH.row_vectors_chunked(Hp)
    if mode == "inregister":
        rowvec, colvec, scal = AH.tf_device_vectors(rowvec, colvec, scal, chunked=chunked)
    else:
        table, tf_mode = AH.tf_table_slot_order(AH.tf_centred_reference_order(Hp, Wp, sp, wl, torch.tensor(z), True, "exact")), 1
    return [P._SlabPlan(G, r, C, H, W, pad_h, pad_w, Hp, Wp, True, device, rowvec, colvec, scal, table, tf_mode, row_chunked=chunked)
            for r in range(G)]


@pytest.mark.parametrize("G,mode", [(2, "cached"), (4, "inregister")])
def test_peer_memory_slab_schedule_replay(G, mode, monkeypatch):
    """The peer-memory slab FFT (row kernels scatter / gather column slabs, thz_asm_desc.slab_*): the G ranks are
    replayed one after the other in this process, with all column slabs in host memory, and must reproduce the
    oracle's full-grid propagation and adjoint."""
    from quantizationawarethzdoe_b200 import functional as Fn, parallel as P

    def call(desc, device):
        rc = E.thz_emul_asm_propagate(ctypes.byref(desc), 148)
        assert rc == 0, rc

    monkeypatch.setattr(Fn, "_asm_call", call)
    H = W = 128
    lams, dxy, z = [1e-3, 1.04e-3], 0.5e-3, 0.1
    torch.manual_seed(0)
    x = torch.randn(1, 2, H, W, dtype=torch.complex64)
    g = torch.randn(1, 2, H, W, dtype=torch.complex64)
    plans = _slab_plans(G, 2, H, W, lams, dxy, z, torch.device("cpu"), mode)
    y = P.slab_emulate_ranks(x, plans)
    gx = P.slab_emulate_ranks(g, plans, conj=True)
    xo = x.clone().requires_grad_(True)
    yo = AO.asm_forward(xo, lams, dxy, z)
    (gxo,) = torch.autograd.grad(yo, xo, g)
    tol = TOL_TABLE if mode == "cached" else TOL
    assert rel_l2(y, yo.detach()) < tol
    assert rel_l2(gx, gxo) < tol


def test_slab_descriptor_is_validated():
    """slab_parts > 1 needs stages 1 or 4, a width divisible by the parts, rows inside the slab and all pointers."""
    base, _ = _setup(1, 1, 128, 128, None, [1e-3], 0.5e-3, 0.1)
    x = torch.zeros(1, 1, 64, 128, dtype=torch.complex64)
    slab = torch.zeros(128 * 128, dtype=torch.complex64)
    kw = dict(base, x=x, y=None, inH=64, outH=64, in_r0=0, out_r0=0, ws=None)
    ok = AH.build_desc(**dict(kw, stages=1, slab=(2, 64, 128, [slab.data_ptr(), slab.data_ptr()])))
    assert E.thz_emul_asm_propagate(ctypes.byref(ok), 148) == 0
    for bad in (dict(stages=3), dict(stages=2), dict(stages=1, slab=(3, 0, 128, [slab.data_ptr()] * 3)), dict(stages=1, slab=(2, 100, 128, [slab.data_ptr()] * 2)),
                dict(stages=1, slab=(2, 0, 128, [slab.data_ptr(), 0]))):
        d = AH.build_desc(**dict(dict(kw, stages=1, slab=(2, 64, 128, [slab.data_ptr(), slab.data_ptr()])), **bad))
        assert E.thz_emul_asm_propagate(ctypes.byref(d), 148) in (-1, -2), bad


@pytest.mark.parametrize("env", [dict(THZ_NO_TILED="1"), dict(THZ_NO_PRUNE="1"), dict(THZ_T1_LOG2="3"), dict(THZ_T1_LOG2="2", THZ_T2_LOG2="3"),
                                 dict(THZ_T2_LOG2="2"), dict(THZ_NO_P2="1"), dict(THZ_NO_K2FAST="1"), dict(THZ_NO_K2FAST="1", MODE0="1")], ids=lambda e: ",".join("%s=%s" % kv for kv in e.items()))
def test_layout_and_pruning_switches_do_not_change_the_result(env, monkeypatch):
    """The A/B switches of DESIGN 3.5 (row-major / blocked intermediates of either width, full instead of pruned stages,
    runtime-planned engine) select different code for the same arithmetic: forward and DOE adjoint must agree with the default
    path to round-off."""
    B, C, H, W = 2, 2, 128, 256                      # 256 x 512 padded: static kernels, centred 2x padding (pruned stages)
    lams, dxy, z = [1e-3, 1.03e-3], 0.5e-3, 0.1
    torch.manual_seed(3)
    x = torch.randn(B, C, H, W, dtype=torch.complex64)
    g = torch.randn(B, C, H, W, dtype=torch.complex64)
    h = torch.rand(H, W) * 1e-3
    coef = AH.doe_coefficients(lams, 2.66, 0.003)

    mode = 0 if env.get("MODE0") else 1          # THZ_NO_K2FAST: the general column kernel instead of thz_p2_k2f, table and vectors

    def run():
        base, _ = _setup(B, C, H, W, None, lams, dxy, z, mode=mode)
        y = torch.zeros_like(x)
        _run(x, dict(base, x=x, y=y, doe_mode=1, doe_base=DO.BASE_PLANE_THICKNESS, hmap=h, coef=coef))
        gx, gh = torch.zeros_like(x), torch.zeros(H, W)
        adj = dict(base, x=g, y=gx, tf_conj=1, doe_mode=2, doe_base=DO.BASE_PLANE_THICKNESS, hmap=h, coef=coef, xsaved=x, gh=gh,
                   inH=base["outH"], inW=base["outW"], in_r0=base["out_r0"], in_c0=base["out_c0"],
                   outH=base["inH"], outW=base["inW"], out_r0=base["in_r0"], out_c0=base["in_c0"])
        _run(g, adj)
        return y, gx, gh

    ref = run()
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    got = run()
    for a, b in zip(got, ref):
        assert rel_l2(a, b) < 2e-6


def test_column_permutation_is_a_pair_preserving_bijection():
    """thz_t2_perm_col: for every line length the TMA-staged row-iFFT kernel serves, the slot -> column map of the permuted
    intermediate is a bijection, keeps even / odd slot pairs adjacent on an even column (the column kernel's 2-column tiles write
    them as one 16-byte piece), and puts the slots of one butterfly where its thread reads them: [t / 2][u][t % 2]."""
    served = {}
    for W in (256, 400, 1024, 2000, 2048, 4096, 8192, 16384):
        R = E.thz_emul_k3_tma_radix(W)
        if R:
            served[W] = R
    assert served == {2048: 8, 4096: 16, 8192: 4}
    for W, R in served.items():
        cols = [E.thz_emul_t2_perm_col(p, R, W) for p in range(W)]
        assert sorted(cols) == list(range(W))
        for p in range(0, W, 2):
            assert cols[p] % 2 == 0 and cols[p + 1] == cols[p] + 1
        nbu = W // R
        for u in (0, 1, nbu // 2 + 3, nbu - 1):
            for t in range(R):
                assert cols[R * u + t] == (t // 2) * (2 * nbu) + 2 * u + t % 2


def test_column_permuted_intermediate_replay(monkeypatch):
    """The TMA-staged row-iFFT kernel (thz_p2_k3t) reads a K2 -> K3 intermediate whose columns the column kernel permutes
    (thz_t2_perm_col; it needs 2-column tiles, i.e. 4096-point columns).  CPU replay of exactly that index arithmetic --
    permuting store of thz_p2_k2f, dense staged rows, first butterfly from the dense copy -- against the default path:
    forward and DOE adjoint must be bit-identical (same values, same operations, another place in memory)."""
    B, C, H, W = 1, 1, 2048, 1024                    # 4096 x 2048 padded
    lams, dxy, z = [1.03e-3], 0.5e-3, 0.1
    torch.manual_seed(5)
    x = torch.randn(B, C, H, W, dtype=torch.complex64)
    g = torch.randn(B, C, H, W, dtype=torch.complex64)
    h = torch.rand(H, W) * 1e-3
    coef = AH.doe_coefficients(lams, 2.66, 0.003)

    def run():
        base, _ = _setup(B, C, H, W, None, lams, dxy, z, mode=0)
        y = torch.zeros_like(x)
        _run(x, dict(base, x=x, y=y, doe_mode=1, doe_base=DO.BASE_PLANE_THICKNESS, hmap=h, coef=coef))
        gx, gh = torch.zeros_like(x), torch.zeros(H, W)
        adj = dict(base, x=g, y=gx, tf_conj=1, doe_mode=2, doe_base=DO.BASE_PLANE_THICKNESS, hmap=h, coef=coef, xsaved=x, gh=gh,
                   inH=base["outH"], inW=base["outW"], in_r0=base["out_r0"], in_c0=base["out_c0"],
                   outH=base["inH"], outW=base["inW"], out_r0=base["in_r0"], out_c0=base["in_c0"])
        _run(g, adj)
        return y, gx, gh

    ref = run()
    assert float(ref[0].abs().max()) > 0 and float(ref[2].abs().max()) > 0 and E.thz_emul_last_t2_perm() == 0
    monkeypatch.setenv("THZ_EMUL_T2_PERM", "1")
    got = run()
    assert E.thz_emul_last_t2_perm() == 8            # 2048-point rows: 16 x 16 x 8, last radix 8 -- the permuted path really ran
    for a, b in zip(got, ref):
        assert torch.equal(a, b)


SOFTMAX_CASES = ("gumbel_hard", "gumbel_soft", "plain_hard", "plain_soft")


def _softmaxq_replay(E, t, lut, noise, tau, hard, gq):
    import ctypes
    P = lambda v: ctypes.c_void_p(v.data_ptr()) if v is not None else None
    n, L = t.numel(), lut.numel()
    s = float(torch.tensor(3.0) / torch.tensor(tau, dtype=torch.float32))
    q, A, Bm, Ee, gt = (torch.empty(n) for _ in range(5))
    idx, st = torch.empty(n, dtype=torch.int32), torch.zeros(4)
    gq = gq.reshape(-1).contiguous()
    assert E.thz_emul_softmaxq(P(t), P(lut), L, P(noise), 300., float(torch.tensor(tau, dtype=torch.float32)), s, int(hard), P(q),
                               P(idx), P(A), P(Bm), P(Ee), P(st), n) == 0
    assert E.thz_emul_softmaxq_bwd(P(gq), P(A), P(Bm), P(Ee), P(st), P(gt), n) == 0
    return q, gt, st


@pytest.mark.parametrize("name", SOFTMAX_CASES)
def test_softmax_quantizer_body_replay(name):
    """thz_softmaxq_pixel (the body of thz_quant_softmax_fwd / _bwd) replayed on the CPU against the reference's forward
    values and the oracle's gradients (tests/golden/quant_softmax.npz).  The soft branches at small tau are ill-conditioned
    in fp32 (logits of O(1000)): the gradient is held to the distance the fp32 reference itself has from float64."""
    from helpers import golden
    E = emul_lib()
    g = golden("quant_softmax")
    t = g["thickness"].reshape(-1).contiguous()
    lut = g["lut"][:-1].contiguous()
    noise = g["noise_" + name].reshape(lut.numel(), -1).contiguous() if ("noise_" + name) in g else None
    q, gt, _ = _softmaxq_replay(E, t, lut, noise, g["tau_" + name], "hard" in name, g["gq"])
    ref_q, ref_g, g64 = g["q_" + name].reshape(-1), g["gt_" + name].reshape(-1), g["gt64_" + name].reshape(-1)
    assert rel_l2(q, ref_q) < 1e-6
    if "hard" in name:      # the selected LEVEL is exact; the value carries the ulps of (onehot + y) - y, which depend on libm's exp
        lv = lambda v: torch.argmin((v[:, None] - lut[None, :]).abs(), dim=1)
        assert torch.equal(lv(q), lv(ref_q))
    budget = max(1e-5, 2.0 * rel_l2(ref_g.double(), g64))
    assert rel_l2(gt, ref_g) < budget, (rel_l2(gt, ref_g), budget)


def test_softmax_quantizer_ties_in_the_global_max():
    """A map with repeated values: many (pixel, level) pairs attain max|diff|, torch.max splits its gradient evenly."""
    from helpers import golden
    E = emul_lib()
    g = golden("quant_softmax")
    q, gt, st = _softmaxq_replay(E, g["ties_t"].reshape(-1).contiguous(), g["lut"][:-1].contiguous(), None, g["ties_tau"], True,
                                 g["ties_gq"])
    assert st[1] == 55
    assert torch.equal(q, g["ties_q"].reshape(-1))
    assert (gt - g["ties_gt"].reshape(-1)).abs().max() <= 1e-6 * g["ties_gt"].abs().max() + 1e-20


@pytest.mark.parametrize("func", ["sigmoid", "log", "poly", "sine", "chirp"])
def test_score_thickness_body_replay(func):
    """All five scoring functions of the reference's score_thickness (quantization.py:36-55), out-of-place restatement."""
    import ctypes, math
    E = emul_lib()
    torch.manual_seed(3)
    t = torch.rand(2, 1, 9, 7) * 1.2e-3 - 1e-4
    lut = torch.linspace(0, 1e-3, 5)[:-1].contiguous()
    s = 2.5
    diff = t - lut.reshape(1, -1, 1, 1)
    diff = diff / torch.max(torch.abs(diff))
    ref = {"sigmoid": lambda: torch.sigmoid(s * diff) * (1 - torch.sigmoid(s * diff)) * 4,
           "log": lambda: -torch.log(diff.abs() + 1e-20) * s,
           "poly": lambda: (1 - torch.abs(diff) ** s),
           "sine": lambda: torch.cos(math.pi * (s * diff).clamp(-1., 1.)),
           "chirp": lambda: 1 - torch.cos(math.pi * (1 - diff.abs()) ** s)}[func]()
    out = torch.empty(2, 4, 9, 7)
    P = lambda v: ctypes.c_void_p(v.data_ptr())
    assert E.thz_emul_score_thickness(P(t.contiguous()), P(lut), 4, s, ["sigmoid", "log", "poly", "sine", "chirp"].index(func), P(out),
                                      2, 63) == 0
    assert rel_l2(out, ref) < 2e-6


def _install_bluestein_cpu(monkeypatch):
    """Kernel replay for thz_asm_propagate + a torch twin of the pointwise multiply kernel (thz_field_mul has no replay)."""
    import ctypes
    from quantizationawarethzdoe_b200 import _native as Nn, functional as Fn
    E = emul_lib()

    def call(desc, device):
        rc = E.thz_emul_asm_propagate(ctypes.byref(desc), 148)
        assert rc == 0, rc

    def field_mul(x, m, real, per_channel, conj):
        B, C, H, W = x.shape
        mm_ = m.reshape((C if per_channel else 1), H, W)
        return x * (mm_.conj() if (conj and not real) else mm_)[None]

    monkeypatch.setattr(Fn, "_asm_call", call)
    monkeypatch.setattr(Fn.FieldMulFn, "_run", staticmethod(field_mul))
    monkeypatch.setattr(Nn, "require_cuda", lambda t, name="tensor": None)
    monkeypatch.setattr(Nn, "twiddles", lambda n, device: Nn.twiddles_host(n))
    monkeypatch.setattr(Nn, "slot_to_bin", lambda n, fn=None, _orig=Nn.slot_to_bin: _orig(n, E.thz_emul_slot_to_bin))


@pytest.mark.parametrize("H,W,scale", [(26, 17, None),       # 52 x 34: 13 and 17 are beyond the radix plans
                                       (11, 32, 2)])         # 33 x 96: one unsupported edge is enough
def test_any_length_asm_through_chirp_z(H, W, scale, monkeypatch):
    """Grids whose padded edge has a prime factor > 7 (the reference's torch.fft takes any size): the chirp-z path
    (bluestein.py: pointwise chirp -> fused convolution -> transfer function -> fused convolution -> chirp) against the
    oracle, forward and adjoint, with the real kernel bodies replayed on the CPU."""
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField, bluestein as BL
    _install_bluestein_cpu(monkeypatch)
    monkeypatch.setattr(BL, "length_supported", lambda n: E_plan_ok(n))
    lams, dxy, z = [1e-3, 1.04e-3], 0.5e-3, 0.05
    torch.manual_seed(0)
    x = torch.randn(2, 2, H, W, dtype=torch.complex64)
    g = torch.randn(2, 2, H, W, dtype=torch.complex64)
    asm = ASM_prop(z_distance=z, padding_scale=scale, device=torch.device("cpu"))
    asm.check_Zc = False
    xr = x.clone().requires_grad_(True)
    y = asm(ElectricField(xr, wavelengths=lams, spacing=dxy, device=torch.device("cpu"))).data
    assert "chirp-z" in asm.resolved_kernel_mode
    (gx,) = torch.autograd.grad(y, xr, g)
    xo = x.clone().requires_grad_(True)
    yo = AO.asm_forward(xo, lams, dxy, z, padding_scale=scale)
    (gxo,) = torch.autograd.grad(yo, xo, g)
    assert rel_l2(y.detach(), yo.detach()) < TOL and rel_l2(gx, gxo) < TOL


def E_plan_ok(n):
    import ctypes
    E = emul_lib()
    rad, ns = (ctypes.c_int32 * 16)(), ctypes.c_int32(0)
    return E.thz_emul_plan_info(int(n), rad, ctypes.byref(ns)) == 0


@pytest.mark.parametrize("H,W", [(13, 22), (34, 19)])
def test_any_length_fft2_through_chirp_z(H, W, monkeypatch):
    from quantizationawarethzdoe_b200 import bluestein as BL
    _install_bluestein_cpu(monkeypatch)
    torch.manual_seed(1)
    x = torch.randn(3, 1, H, W, dtype=torch.complex64)
    for inverse, ortho in ((False, False), (True, False), (False, True), (True, True)):
        ref = (torch.fft.ifft2 if inverse else torch.fft.fft2)(x, norm="ortho" if ortho else "backward")
        got = BL.BluesteinFft2(H, W, inverse, ortho, torch.device("cpu"))(x)
        assert rel_l2(got, ref) < 3e-6, (inverse, ortho)


@pytest.mark.parametrize("H,W,scale,with_doe", [(48, 40, None, True),        # runtime-planned engine (96 x 80)
                                                (128, 256, None, True),      # static kernels, pruned stages (256 x 512)
                                                (128, 256, None, False),     # aperture + lens only (no DOE): adjoint = conj multiply
                                                (200, 200, None, True)])     # 400 x 400 mixed radix
def test_pointwise_elements_fused_into_the_row_kernels(H, W, scale, with_doe):
    """Aperture mask + thin-lens kernel in front of the (DOE +) propagation, fused into the row-FFT prologue
    (forward) and the row-iFFT epilogue (adjoint): replayed kernel bodies vs autograd through the oracle."""
    B, C = 2, 2
    lams = [1e-3, 1.03e-3]
    dxy, z, eps, tand = 0.5e-3, 0.1, 2.66, 0.003
    torch.manual_seed(0)
    x = torch.randn(B, C, H, W, dtype=torch.complex64)
    g = torch.randn(B, C, H, W, dtype=torch.complex64)
    h = torch.rand(H, W) * 1e-3
    mask = (torch.rand(H, W) > 0.3).float()
    mul = torch.exp(1j * torch.randn(C, H, W)).to(torch.complex64).contiguous()
    base, _ = _setup(B, C, H, W, scale, lams, dxy, z, mode=1)
    coef = AH.doe_coefficients(lams, eps, tand)
    y = torch.zeros_like(x)
    fwd = dict(base, x=x, y=y, elem_mode=1, elem_mask=mask, elem_mul=mul)
    if with_doe:
        fwd.update(doe_mode=1, doe_base=DO.BASE_PLANE_THICKNESS, hmap=h, coef=coef)
    _run(x, fwd)
    xr, hr = x.clone().requires_grad_(True), h.clone().requires_grad_(True)
    u = xr * mask[None, None] * mul[None]
    if with_doe:
        u = DO.modulate(u, hr, lams, eps, tand)
    yo = AO.asm_forward(u, lams, dxy, z, padding_scale=scale)
    assert rel_l2(y, yo.detach()) < TOL_TABLE
    gx, gh = torch.zeros_like(x), torch.full((H, W), 7.0)
    adj = dict(base, x=g, y=gx, tf_conj=1, elem_mode=2, elem_mask=mask, elem_mul=mul)
    if with_doe:
        adj.update(doe_mode=2, doe_base=DO.BASE_PLANE_THICKNESS, hmap=h, coef=coef, xsaved=x, gh=gh)
        gxo, gho = torch.autograd.grad(yo, (xr, hr), g)
    else:
        (gxo,) = torch.autograd.grad(yo, xr, g)
    _run(g, adj)
    assert rel_l2(gx, gxo) < TOL_TABLE
    if with_doe:
        assert rel_l2(gh, gho) < TOL_TABLE


# ------------------------------------------------------------------------------- lines above max_line(): the outer split
def _install_longline_cpu(monkeypatch, max_line):
    """Replay of thz_split_pre / thz_split_post + the fused pipeline, with the direct-transform limit lowered to `max_line` so
    that small grids take the split path (longline.py) -- the same code that serves edges above 16384 points on the device."""
    from quantizationawarethzdoe_b200 import functional as Fn, longline as LL
    _install_bluestein_cpu(monkeypatch)
    E = emul_lib()
    monkeypatch.setitem(Fn.TUNE, "max_line", max_line)
    monkeypatch.setattr(LL, "_plan_ok", E_plan_ok)

    def pre(x, region, Hp, Wp, Pr, Pc, conj_tw=False, scale=1.0):
        H, W, r0, c0 = region
        x = x.contiguous()
        u = torch.zeros(tuple(x.shape[:-2]) + (Pr * Pc, Hp // Pr, Wp // Pc), dtype=torch.complex64)
        rc = E.thz_emul_split_pre(N.ptr(x), N.ptr(u), x.numel() // (H * W), H, W, r0, c0, Hp, Wp, Pr, Pc,
                                  N.ptr(LL.line_twiddles(Hp, "cpu")), N.ptr(LL.line_twiddles(Wp, "cpu")), int(conj_tw), float(scale))
        assert rc == 0
        return u

    def post(v, y, region, Hp, Wp, Pr, Pc, conj_tw=False, scale=1.0):
        H, W, r0, c0 = region
        v = v.contiguous()
        out = torch.zeros(y.shape, dtype=torch.complex64)
        rc = E.thz_emul_split_post(N.ptr(v), N.ptr(out), y.numel() // (H * W), H, W, r0, c0, Hp, Wp, Pr, Pc,
                                   N.ptr(LL.line_twiddles(Hp, "cpu")), N.ptr(LL.line_twiddles(Wp, "cpu")), int(conj_tw), float(scale))
        assert rc == 0
        y.copy_(out)
        return y

    def fft2(x, inverse=False, ortho=False):
        x = x.contiguous()
        H, W = x.shape[-2:]
        y, ws = torch.zeros_like(x), torch.zeros_like(x)
        rc = E.thz_emul_fft2_c2c(N.ptr(x), N.ptr(y), x.numel() // (H * W), H, W, int(inverse), int(ortho), N.ptr(N.twiddles_host(H)),
                                 N.ptr(N.twiddles_host(W)), N.ptr(ws))
        assert rc == 0
        return y

    monkeypatch.setattr(LL, "split_pre", pre)
    monkeypatch.setattr(LL, "split_post", post)
    monkeypatch.setattr(Fn, "fft2_c2c", fft2)
    return LL


@pytest.mark.parametrize("H,W,scale,max_line,mode,split", [
    (64, 64, None, 64, "inregister", "2 x 2"),          # 128 x 128 canvas, both edges split in two
    (32, 128, None, 64, "cached", "1 x 4"),             # 64 x 256: only the long edge, in four
    (60, 50, [3, 1], 100, "inregister", "4 x 1"),       # 240 x 100: mixed radix sub-lengths (60, 100), unpad crop off-centre
    (96, 40, None, 48, "cached", "4 x 2")])             # 192 x 80 -> 48 x 40 sub-problems on the runtime-planned engine
def test_long_line_asm_through_the_outer_split(H, W, scale, max_line, mode, split, monkeypatch):
    """ASM_prop on a canvas with an edge above the direct-transform limit: split -> un-padded fused propagation of Pr Pc
    decimated sub-problems -> merge, against the oracle, forward and adjoint, both transfer-function modes."""
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField
    _install_longline_cpu(monkeypatch, max_line)
    lams, dxy, z = [1e-3, 1.04e-3], 0.5e-3, 0.03
    torch.manual_seed(0)
    x = torch.randn(2, 2, H, W, dtype=torch.complex64)
    g = torch.randn(2, 2, H, W, dtype=torch.complex64)
    asm = ASM_prop(z_distance=z, padding_scale=scale, kernel_mode=mode, device=torch.device("cpu"))
    asm.check_Zc = False
    xr = x.clone().requires_grad_(True)
    y = asm(ElectricField(xr, wavelengths=lams, spacing=dxy, device=torch.device("cpu"))).data
    assert asm.resolved_kernel_mode == "%s (split %s)" % (mode, split)
    (gx,) = torch.autograd.grad(y, xr, g)
    xo = x.clone().requires_grad_(True)
    yo = AO.asm_forward(xo, lams, dxy, z, padding_scale=scale)
    (gxo,) = torch.autograd.grad(yo, xo, g)
    tol = TOL if mode == "cached" else 2e-5            # in-register H: the sqrt deviation of DESIGN.md section 4, same as unsplit
    assert rel_l2(y.detach(), yo.detach()) < tol and rel_l2(gx, gxo) < tol


def test_long_line_split_equals_the_direct_plan(monkeypatch):
    """The split path and the direct path are the same operator: 256 x 128 canvas run directly and as 2 x 2 / 4 x 1 splits."""
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField, functional as Fn
    _install_longline_cpu(monkeypatch, 16384)
    lams, dxy, z = [1e-3], 0.5e-3, 0.02
    torch.manual_seed(3)
    x = torch.randn(1, 1, 128, 64, dtype=torch.complex64)
    outs = []
    for ml in (16384, 128, 64):
        monkeypatch.setitem(Fn.TUNE, "max_line", ml)
        asm = ASM_prop(z_distance=z, kernel_mode="cached", do_unpad_after_pad=False, device=torch.device("cpu"))
        asm.check_Zc = False
        outs.append(asm(ElectricField(x, wavelengths=lams, spacing=dxy, device=torch.device("cpu"))).data)
        assert ("split" in asm.resolved_kernel_mode) == (ml < 256)
    assert outs[0].shape == (1, 1, 256, 128)
    assert rel_l2(outs[1], outs[0]) < 1e-6 and rel_l2(outs[2], outs[0]) < 1e-6


@pytest.mark.parametrize("H,W,max_line", [(128, 64, 64), (40, 256, 64), (200, 120, 60)])
def test_long_line_fft2_through_the_outer_split(H, W, max_line, monkeypatch):
    LL = _install_longline_cpu(monkeypatch, max_line)
    torch.manual_seed(2)
    x = torch.randn(2, 3, H, W, dtype=torch.complex64)
    for inverse, ortho in ((False, False), (True, False), (False, True), (True, True)):
        ref = (torch.fft.ifft2 if inverse else torch.fft.fft2)(x, norm="ortho" if ortho else "backward")
        assert rel_l2(LL.fft2_split(x, inverse, ortho), ref) < 2e-6, (inverse, ortho)


def test_long_chirp_convolution_uses_the_split(monkeypatch):
    """Chirp-z lengths whose convolution canvas exceeds the direct limit (n > max_line / 2) go through the split plan."""
    from quantizationawarethzdoe_b200 import bluestein as BL, longline as LL
    _install_longline_cpu(monkeypatch, 64)
    monkeypatch.setattr(BL, "length_supported", lambda n: n <= 64 and E_plan_ok(n))
    torch.manual_seed(5)
    x = torch.randn(2, 1, 13, 47, dtype=torch.complex64)            # 47 -> 128-point convolution = 2 x 64
    plan = BL.BluesteinFft2(13, 47, False, False, torch.device("cpu"))
    assert isinstance(plan.conv.plan, LL.SplitAsmPlan) and (plan.conv.plan.Pr, plan.conv.plan.Pc) == (1, 2)
    assert rel_l2(plan(x), torch.fft.fft2(x)) < 3e-6


@pytest.mark.parametrize("B,C,H,W,sm,chunk", [(2, 3, 128, 256, 148, 0),      # static row kernels, 256 x 512 canvas
                                              (3, 2, 128, 256, 2, 4),        # few SMs + chunks (atomic grad_height)
                                              (3, 1, 200, 200, 148, 0),      # 400-point lines, several per CTA
                                              (2, 2, 48, 40, 148, 0)])       # run-time planned lengths: the tables are ignored
def test_quantised_doe_level_tables_equal_per_pixel_evaluation(B, C, H, W, sm, chunk):
    """Quantised height map h = lut[idx]: handing the row kernels the level map + per-level transmissions (thz_asm_desc.doe_level_*)
    gives what evaluating the transmission per pixel gives, forward and adjoint (replayed kernel bodies)."""
    lams = [1e-3, 1.02e-3, 1.05e-3][:C]
    dxy, z, eps, tand = 0.5e-3, 0.1, 2.66, 0.003
    torch.manual_seed(0)
    x = torch.randn(B, C, H, W, dtype=torch.complex64)
    g = torch.randn(B, C, H, W, dtype=torch.complex64)
    lut = torch.linspace(0, 1e-3, 5)[:-1].contiguous()
    idx = torch.randint(0, 4, (H, W), dtype=torch.int32)
    h = lut[idx.long()].contiguous()
    coef = AH.doe_coefficients(lams, eps, tand)
    # per-level transmissions through the replayed fused prologue itself: a 1 x 4 "field" of ones over the LUT as height map
    lphase = DO.modulate(torch.ones(1, C, 1, 4, dtype=torch.complex64), lut.reshape(1, 4), lams, eps, tand).reshape(C, 4).contiguous()
    base, _ = _setup(B, C, H, W, None, lams, dxy, z, mode=1, chunk=chunk)
    outs = []
    for lev in (False, True):
        kw = dict(level_idx=idx, level_phase=lphase) if lev else {}
        y = torch.zeros_like(x)
        _run(x, dict(base, x=x, y=y, doe_mode=1, doe_base=DO.BASE_PLANE_THICKNESS, hmap=h, coef=coef, **kw), sm)
        gx, gh = torch.zeros_like(x), torch.full((H, W), 7.0)
        _run(g, dict(base, x=g, y=gx, tf_conj=1, doe_mode=2, doe_base=DO.BASE_PLANE_THICKNESS, hmap=h, coef=coef, xsaved=x, gh=gh, **kw), sm)
        outs.append((y, gx, gh))
    for a, b in zip(*outs):
        assert rel_l2(b, a) < 1e-6            # the oracle's torch exp / sincos vs the kernels' own: a few 1e-7
    if H == 48:
        assert all(torch.equal(a, b) for a, b in zip(*outs))      # tables not used off the static path: identical
    else:
        assert not torch.equal(outs[0][0], outs[1][0])            # ... and used on it (the table came from another evaluation)
