"""`not gpu` checks of the C-ABI library (loads, exports everything include/thzdoe.h declares, host-only
planning helpers) and of the host-side logic mirrored from the reference (padding, validation, errors).
No compute call is made: there is no GPU here."""
import ctypes
import subprocess
import os
import re

import numpy as np
import pytest
import torch

from helpers import ROOT, golden, rel_l2
from oracle import asm_oracle as AO
from quantizationawarethzdoe_b200 import _native as N
from quantizationawarethzdoe_b200 import asm_host as AH


@pytest.fixture(scope="module")
def lib():
    from quantizationawarethzdoe_b200 import build
    build.build()
    return N.load_library()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "thzdoe.h")).read()
    declared = set(re.findall(r"\b(thz_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("thz_asm_desc")
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), "libthzdoe.so does not export %s" % name
    assert declared == set(N.EXPORTS), "python binding list out of sync with include/thzdoe.h"
    assert lib.thz_version() >= 100


def test_descriptor_layout_matches_header(tmp_path):
    """ctypes mirrors vs the C header: sizes and field offsets as gcc lays the structs out."""
    # thz_asm_desc: 12 int32, 2 ptr, 2 int32, 4 ptr, int32+float, 4 ptr, 2 ptr, ptr, u64, 4 int32, 4 int32, 8 ptr
    assert ctypes.sizeof(N.AsmDesc) == 12 * 4 + 2 * 8 + 2 * 4 + 4 * 8 + 8 + 4 * 8 + 2 * 8 + 8 + 8 + 4 * 4 + 4 * 4 + 8 * 8 + 2 * 4 + 8 + 2 * 4 + 2 * 8 + 2 * 8 + 2 * 4
    fields = ["x", "tf_mode", "tf_table", "doe_base", "doe_gh", "ws_bytes", "stages", "slab_parts", "slab_ptrs", "tf_row_chunked", "doe_hmap_bstride", "elem_mode", "elem_mul", "doe_level_idx", "doe_levels"]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "thzdoe.h"\nint main(void) {\n'
                   '  printf("%zu %zu", sizeof(thz_asm_desc), sizeof(thz_toeplitz_gemm_desc));\n' +
                   "".join('  printf(" %%zu", offsetof(thz_asm_desc, %s));\n' % f for f in fields) + "  return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert got[0] == ctypes.sizeof(N.AsmDesc) and got[1] == ctypes.sizeof(N.ToeplitzGemmDesc)
    assert got[2:] == [getattr(N.AsmDesc, f).offset for f in fields]


def test_plan_helpers(lib):
    radices = (ctypes.c_int32 * 16)()
    ns = ctypes.c_int32()
    for n, want in ((4096, [16, 16, 16]), (1024, [16, 8, 8]), (2000, [25, 20, 4]), (400, [25, 16]), (3000, [25, 20, 6]), (300, [25, 12])):
        assert lib.thz_fft_plan_info(n, radices, ctypes.byref(ns)) == 0
        assert list(radices)[:ns.value] == want and int(np.prod(want)) == n
    assert lib.thz_fft_plan_info(2 * 13, radices, ctypes.byref(ns)) == -3
    assert b"prime factor" in lib.thz_last_error()
    for n in (8, 60, 400, 4096):
        perm = N.slot_to_bin(n, lib.thz_fft_slot_to_bin)
        assert sorted(perm.tolist()) == list(range(n))
    tw = (ctypes.c_float * (2 * 12))()
    assert lib.thz_fft_twiddles(12, tw) == 0
    t = torch.tensor(list(tw)).reshape(12, 2)
    assert rel_l2(torch.view_as_complex(t), N.twiddles_host(12)) < 1e-7


def test_null_descriptor_is_rejected_without_touching_cuda(lib):
    d = N.AsmDesc()
    assert lib.thz_asm_propagate(ctypes.byref(d), None) == -1      # THZ_E_NULL
    assert lib.thz_asm_workspace_bytes(ctypes.byref(d)) == 0
    assert lib.thz_quant_ste_fwd(None, 0, 0.0, 0.0, None, 4, None, None, None, 10, None) == -1


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(N.ThzError, match="no CPU fallback"):
        N.load_library(str(tmp_path / "nope.so"))


# ----------------------------------------------------------------------------- host logic
def test_padding_rules():
    assert AH.compute_padding(50, 100, AH.normalise_padding_scale(1), True) == (25, 50, 100, 200)   # ASM_Prop.py:50-54
    assert AH.compute_padding(50, 100, AH.normalise_padding_scale(torch.tensor([1, 2])), True) == (25, 100, 100, 300)
    assert AH.compute_padding(100, 100, AH.normalise_padding_scale(2), True) == (100, 100, 300, 300)
    assert AH.compute_padding(33, 36, AH.normalise_padding_scale([1, 2]), True) == (16, 36, 65, 108)
    assert AH.compute_padding(7, 9, None, False) == (0, 0, 7, 9)
    with pytest.raises(Exception, match="padding_scale"):
        AH.normalise_padding_scale([1, 2, 3])
    for ps in (None, 2, [1, 2], torch.tensor(3.0), torch.tensor([1.0, 2.0])):
        a = AH.compute_padding(40, 24, AH.normalise_padding_scale(ps), True)
        b = AO.compute_padding(40, 24, ps, True)
        assert a == b


@pytest.mark.parametrize("bt", ["exact", "approx"])
def test_separable_vectors_rebuild_the_reference_kernel(bt):
    """tf_vectors -> full kernel must equal the reference kernel bit for bit (same CPU ops)."""
    Hp, Wp, lams, sp, z = 60, 108, [0.9e-3, 1.2e-3], [0.5e-3, 0.4e-3], 0.05
    Hc = AH.tf_centred_reference_order(Hp, Wp, sp, lams, z, True, bt)
    ref = AO.centred_transfer_function(Hp, Wp, sp, lams, z, True, bt)[0]
    assert torch.equal(torch.view_as_real(Hc), torch.view_as_real(ref))
    g = golden("asm_approx_mixedpad" if bt == "approx" else "asm_pow2")
    Hp, Wp = g["kernel"].shape[-2:]
    Hc = AH.tf_centred_reference_order(Hp, Wp, g["spacing"].float(), g["wavelengths"].float(), g["z"], True, g["bandlimit_type"])
    assert torch.equal(torch.view_as_real(Hc), torch.view_as_real(g["kernel"][0]))


def test_electric_field_validation_matches_reference():
    from quantizationawarethzdoe_b200 import ElectricField
    cpu = torch.device("cpu")
    x = torch.zeros(1, 2, 4, 4, dtype=torch.complex64)
    f = ElectricField(x, wavelengths=[1e-3, 2e-3], spacing=1e-3, device=cpu)
    assert f.field_type == "scalar" and f.spacing.tolist() == pytest.approx([1e-3, 1e-3]) and f.wavelengths.dtype == torch.float32
    assert (f.height, f.width, f.num_wavelengths, f.num_batches) == (4, 4, 2, 1)
    assert ElectricField(torch.zeros(3, 1, 2, 2), 1e-3, 1e-3, device=cpu).field_type == "vectorial"
    assert ElectricField(torch.zeros(5, 1, 2, 2), 1e-3, 1e-3, device=cpu).field_type == "batch"
    with pytest.raises(ValueError, match="channels"):
        ElectricField(x, wavelengths=[1e-3], spacing=1e-3, device=cpu)
    with pytest.raises(ValueError, match="Spacing"):
        ElectricField(x, wavelengths=[1e-3, 2e-3], spacing=[1e-3, 1e-3, 1e-3], device=cpu)
    with pytest.raises(ValueError, match="Wavelengths"):
        ElectricField(x, wavelengths=None, spacing=1e-3, device=cpu)
    with pytest.raises(AssertionError):
        ElectricField(torch.zeros(2, 4, 4), wavelengths=[1e-3, 2e-3], spacing=1e-3, device=cpu)


def test_asm_prop_surface_and_loud_failure_without_cuda():
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField
    cpu = torch.device("cpu")
    with pytest.raises(Exception, match="padding_scale"):
        ASM_prop(padding_scale=[1, 2, 3], device=cpu)
    a = ASM_prop(z_distance=0.1, padding_scale=2, device=cpu)
    assert a.compute_padding(100, 100) == (300, 300) and a.compute_padding(100, 100, True) == (100, 100)
    a.z = np.float64(0.2)                      # depth-sweep loops assign python / numpy scalars (ASM_Prop.py:190-195)
    assert float(a._z_f32()) == pytest.approx(0.2)
    a.z = torch.tensor(0.3)
    assert float(a.z) == pytest.approx(0.3)
    f = ElectricField(torch.zeros(1, 1, 8, 8, dtype=torch.complex64), 1e-3, 1e-3, device=cpu)
    assert a.create_kernel(f).shape == (1, 1, 24, 24)
    assert set(a.state_dict().keys()) == {"_Kx", "_Ky"}
    with pytest.raises(N.ThzError, match="no CPU path"):   # the product never falls back to the CPU
        a.check_Zc = False
        a(f)


def test_row_thresholds_binary_search_equals_dense_evaluation(lib):
    """The O((Hp+Wp) log Wp) threshold builder a depth sweep calls per z must equal the dense O(Hp Wp) evaluation
    of the reference's keep mask (Props/ASM_Prop.py:262, :290-301) bit for bit, incl. rows with no kept bin."""
    from quantizationawarethzdoe_b200 import asm_host as AH
    cases = [(256, 256, .5e-3, .5e-3, 1e-3, .1, True, 'exact'), (400, 600, 1e-3, .7e-3, 1e-3, .2, True, 'approx'),
             (300, 300, .3e-3, .3e-3, 1e-3, .05, True, 'exact'), (512, 512, 2e-3, 2e-3, 1e-3, 5.0, True, 'exact'),
             (64, 96, .5e-3, .5e-3, .9e-3, .01, False, 'exact'), (25, 25, 1e-3, 1e-3, 1e-3, 0.3, True, 'exact'),
             (1000, 1000, .2e-3, .2e-3, 1e-3, 2.0, True, 'exact')]
    for Hp, Wp, dx, dy, lam, z, bl, bt in cases:
        wl = torch.tensor([lam, lam * 1.07, lam * 0.8])
        rv, cv, sc = AH.tf_vectors(Hp, Wp, torch.tensor([dx, dy]), wl, torch.tensor(z), bl, bt)
        fast, dense = AH.tf_row_thresholds(rv, cv, sc), AH._tf_row_thresholds_dense(rv, cv, sc)
        assert fast is not None and dense is not None
        assert torch.equal(fast, dense), (Hp, Wp, bt)                                    # C helper (host code of the .so)
        assert torch.equal(AH._tf_row_thresholds_numpy(rv, cv, sc), dense), (Hp, Wp, bt)  # numpy restatement


def test_loss_landscape_host_helpers(tmp_path):
    """Surface-file layout and the reference's index convention (VisTools/calc_loss.py:67-98), no GPU involved."""
    import types
    from quantizationawarethzdoe_b200.VisTools import calc_loss as CL
    args = types.SimpleNamespace(xmin=-1.0, xmax=1.0, xnum=5, ymin=-0.5, ymax=0.5, ynum=4)
    path = CL.setup_surface_file(args, str(tmp_path))
    d = CL.read_surface_file(path)
    assert d["loss"].shape == (5, 4) and (d["loss"] == -1).all()
    assert np.allclose(d["xcoordinates"], np.linspace(-1, 1, 5)) and np.allclose(d["ycoordinates"], np.linspace(-0.5, 0.5, 4))
    inds, coords = CL.get_indices(d["loss"], d["xcoordinates"], d["ycoordinates"])
    assert inds.tolist() == list(range(20))
    # flat index i = iy * xnum + ix is evaluated at (x[ix], y[iy]) -- np.meshgrid(x, y) order -- whatever the loss array's shape
    assert np.allclose(coords[7], [d["xcoordinates"][2], d["ycoordinates"][1]])
    d["loss"].ravel()[inds[:3]] = [0.5, 0.25, 0.125]
    CL._write_losses(path, d["loss"])
    d2 = CL.read_surface_file(path)
    assert d2["loss"][0, :3].tolist() == [0.5, 0.25, 0.125]
    inds2, _ = CL.get_indices(d2["loss"], d2["xcoordinates"], d2["ycoordinates"])
    assert inds2.tolist() == list(range(3, 20))                 # evaluated points are skipped on a re-run
    m = torch.nn.Linear(2, 2, bias=False)
    w0 = [p.data.clone() for p in m.parameters()]
    CL.overwrite_weights(m, w0, [[torch.ones(2, 2)], [2 * torch.ones(2, 2)]], (0.5, -0.25))
    assert torch.allclose(m.weight.data, w0[0])                 # 0.5 * 1 - 0.25 * 2 = 0
    CL.overwrite_weights(m, w0, [[torch.ones(2, 2)], [2 * torch.ones(2, 2)]], (1.0, 1.0))
    assert torch.allclose(m.weight.data, w0[0] + 3)


def test_chirp_z_host_tables():
    """bluestein.py host math, no GPU: exact chirp phases, convolution length, and the kernel spectrum reproduces a DFT when the
    convolution is carried out with numpy."""
    from quantizationawarethzdoe_b200 import bluestein as BL
    assert BL.conv_length(202) == 512 and BL.conv_length(8) == 16 and BL.conv_length(8192) == 16384
    assert BL.conv_length(8209) == 32768                     # above 16384: one outer split (longline.py) serves it
    with pytest.raises(NotImplementedError, match="chirp convolution"):
        BL.conv_length(32769)
    n = 101
    w = BL.chirp(n)
    j = np.arange(n)
    assert np.allclose(w, np.exp(-1j * np.pi * (j.astype(np.float64) ** 2) / n), atol=1e-12)
    rng = np.random.default_rng(0)
    H, W = 13, 10
    x = rng.standard_normal((H, W)) + 1j * rng.standard_normal((H, W))
    L1, L2 = BL.conv_length(H), BL.conv_length(W)
    a = np.zeros((L1, L2), complex)
    a[:H, :W] = x * np.outer(BL.chirp(H), BL.chirp(W))
    b = np.fft.ifft2(np.fft.fft2(a) * BL.kernel_spectrum(H, W, L1, L2))[:H, :W]
    assert np.allclose(b * np.outer(BL.chirp(H), BL.chirp(W)), np.fft.fft2(x), atol=1e-10)


@pytest.mark.parametrize("Hp,Wp,bt,z", [(64, 96, "exact", 0.1), (60, 45, "exact", 0.26), (50, 36, "approx", 0.05), (35, 27, "exact", 0.3)])
def test_quarter_angles_rebuild_the_reference_table(Hp, Wp, bt, z):
    """kernel_mode 'cached': the host evaluates the reference's phase angles on the unique quarter of the frequency grid only
    and the device expands them (thz_tf_table_from_angles).  Restated here with torch on the CPU -- same index maps, same
    threshold mask -- the expansion must equal the table made from the reference's full-grid kernel bit for bit."""
    from quantizationawarethzdoe_b200 import asm_host as AH
    sp, lam, zt = torch.tensor([0.5e-3, 0.7e-3]), torch.tensor([1e-3, 1.06e-3]), torch.tensor(z)
    fn = None            # the library's own host-side planning helper (no GPU involved)
    ref = AH.tf_table_slot_order(AH.tf_centred_reference_order(Hp, Wp, sp, lam, zt, True, bt), fn)        # [C, Wp, Hp]
    rowtau, colk2, _ = AH.tf_device_vectors(*AH.tf_vectors(Hp, Wp, sp, lam, zt, True, bt), fn, chunked=False)
    angq = AH.tf_angle_quarter(Hp, Wp, sp, lam, zt)
    ra, ca = AH.abs_bin_of_slots(Hp, fn).long(), AH.abs_bin_of_slots(Wp, fn).long()
    ang = angq[:, ra][:, :, ca].transpose(1, 2)                         # [C, slot_c, slot_r]
    keep = colk2[:, :, None] <= rowtau[:, None, :, 1]
    tab = torch.where(keep, torch.exp(1j * ang), torch.zeros((), dtype=torch.complex64))
    assert torch.equal(torch.view_as_real(tab), torch.view_as_real(ref))
