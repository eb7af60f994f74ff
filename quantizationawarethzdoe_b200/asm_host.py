"""Host-side preparation for the fused ASM kernels: geometry, separable transfer-function vectors,
cached transfer-function tables and descriptor assembly.

Everything here is O(Hp + Wp) per wavelength (or a one-off table build) and runs once per
(shape, spacing, wavelengths, z); results are cached by the modules.  The fp32 operation order of
the reference's create_kernel (Props/ASM_Prop.py:212-311) is kept so that the band-limit mask the
kernel evaluates is bit-identical to the reference's.
"""
import ctypes

import numpy as np
import torch

from . import _native as N


def normalise_padding_scale(padding_scale, do_padding=True):
    """ASM_prop.__init__ normalisation (Props/ASM_Prop.py:75-98) -> 2-element tensor or None."""
    if not do_padding:
        return None
    err = False
    if not torch.is_tensor(padding_scale):
        if padding_scale is None:
            padding_scale = torch.tensor([1, 1])
        elif np.isscalar(padding_scale):
            padding_scale = torch.tensor([padding_scale, padding_scale])
        else:
            padding_scale = torch.tensor(padding_scale)
            if padding_scale.numel() != 2:
                err = True
    elif padding_scale.numel() == 1:
        v = padding_scale.reshape(-1)[0]
        padding_scale = torch.stack([v, v])
    elif padding_scale.numel() == 2:
        padding_scale = padding_scale.squeeze()
    else:
        err = True
    if err:
        raise Exception("Invalid value for argument 'padding_scale'.  Should be a real-valued non-negative scalar "
                        "number or a two-element tuple/tensor containing real-valued non-negative scalar numbers.")
    return padding_scale


def compute_padding(H, W, padding_scale, do_padding):
    """Props/ASM_Prop.py:119-136 -> (padH, padW, Hp, Wp)."""
    if not do_padding:
        return 0, 0, int(H), int(W)
    pad_h = int(np.floor((float(padding_scale[0]) * H) / 2))
    pad_w = int(np.floor((float(padding_scale[1]) * W) / 2))
    return pad_h, pad_w, H + 2 * pad_h, W + 2 * pad_w


def _as_f32_cpu(v):
    return torch.as_tensor(v).detach().to(device="cpu", dtype=torch.float32)


def _spacing2(spacing):
    s = _as_f32_cpu(spacing).reshape(-1)
    if s.numel() == 1:
        s = s.repeat(2)
    if s.numel() != 2:
        raise ValueError("Spacing must be a 2-element tensor.")
    return s


def tf_vectors(Hp, Wp, spacing, wavelengths, z, bandlimit=True, bandlimit_type="exact"):
    """Separable pieces of the band-limited transfer function in FFT-bin order (CPU fp32 tensors).

    rowvec [C,Hp,4] = {Kx^2, Kx^2/(2 pi u_lim)^2, Kx^2/klam^2, 0}
    colvec [C,Wp,4] = {Ky^2, Ky^2/klam^2,        Ky^2/(2 pi v_lim)^2, 0}
    scal   [C,2]    = {klam^2, z}
    The kernel keeps a bin iff rowvec.y+colvec.y <= 1 and rowvec.z+colvec.z <= 1 and klam^2-(Kx^2+Ky^2) >= 0
    (Props/ASM_Prop.py:262, 297-301); for 'approx' (:303-306) the two flags are folded into .y."""
    spacing = _spacing2(spacing)
    lam = _as_f32_cpu(wavelengths).reshape(-1)
    z = _as_f32_cpu(z).reshape(())
    dx, dy = spacing[0:1], spacing[1:2]
    C = lam.numel()
    kx = (torch.linspace(0, Hp - 1, Hp) - (Hp // 2)) / Hp            # ASM_Prop.py:142
    ky = (torch.linspace(0, Wp - 1, Wp) - (Wp // 2)) / Wp            # :143
    Kx = 2 * torch.pi * kx / dx                                      # :245
    Ky = 2 * torch.pi * ky / dy                                      # :246
    Kx2, Ky2 = Kx ** 2, Ky ** 2                                      # :249
    k_lam = 2 * torch.tensor(np.pi) / lam[:, None]                   # :253  [C,1]
    k_lam2 = k_lam ** 2                                              # :254
    rowvec = torch.zeros(C, Hp, 4)
    colvec = torch.zeros(C, Wp, 4)
    rowvec[:, :, 0] = Kx2[None, :]
    colvec[:, :, 0] = Ky2[None, :]
    if bandlimit:
        if bandlimit_type == "exact":
            du = ((2 * np.pi / dx) / (2 * Hp)) / (2 * np.pi)         # :290
            dv = ((2 * np.pi / dy) / (2 * Hp)) / (2 * np.pi)         # :291 (Hp for both axes: reference quirk)
            u_lim = 1 / torch.sqrt(((2 * du * z) ** 2) + 1) / lam[:, None]     # :293
            v_lim = 1 / torch.sqrt(((2 * dv * z) ** 2) + 1) / lam[:, None]     # :294
            rowvec[:, :, 1] = (Kx2[None, :]) / ((2 * np.pi * u_lim) ** 2)      # :297 first term
            colvec[:, :, 1] = (Ky2[None, :]) / (k_lam ** 2)                    # :297 second term
            rowvec[:, :, 2] = (Kx2[None, :]) / (k_lam ** 2)                    # :298 first term
            colvec[:, :, 2] = (Ky2[None, :]) / ((2 * np.pi * v_lim) ** 2)      # :298 second term
        elif bandlimit_type == "approx":
            len_x = Hp * dx                                          # :274
            len_y = Hp * dy                                          # :275
            kx_max = 2 * np.pi / torch.sqrt(((2 * (1 / len_x) * z) ** 2) + 1) / lam[:, None]   # :303
            ky_max = 2 * np.pi / torch.sqrt(((2 * (1 / len_y) * z) ** 2) + 1) / lam[:, None]   # :304
            rowvec[:, :, 1] = (torch.abs(Kx)[None, :] > kx_max).float() * 2.0   # :306
            colvec[:, :, 1] = (torch.abs(Ky)[None, :] > ky_max).float() * 2.0
        else:
            raise Exception("Should not be in this state.")
    rowvec = torch.fft.ifftshift(rowvec, dim=1).contiguous()
    colvec = torch.fft.ifftshift(colvec, dim=1).contiguous()
    scal = torch.stack([k_lam2[:, 0], z.expand(C)], dim=1).contiguous()
    return rowvec, colvec, scal


def tf_row_thresholds(rowvec, colvec, scal):
    """tau [C,Hp] through the library's host helper `thz_tf_row_thresholds` (include/thzdoe.h); None if the
    band-limit quotients are not monotone (the caller then uses the cached-table mode).  Same result as
    `_tf_row_thresholds_numpy` / `_tf_row_thresholds_dense` below, which the tests compare it with."""
    import ctypes
    rv, cv, sc = rowvec.contiguous(), colvec.contiguous(), scal.contiguous()
    C, Hp, Wp = rv.shape[0], rv.shape[1], cv.shape[1]
    tau = torch.empty(C, Hp)
    rc = N.lib().thz_tf_row_thresholds(C, Hp, Wp, ctypes.c_void_p(rv.data_ptr()), ctypes.c_void_p(cv.data_ptr()),
                                       ctypes.c_void_p(sc.data_ptr()), ctypes.c_void_p(tau.data_ptr()))
    if rc == N.THZ_E_UNSUPPORTED:
        return None
    N.check(rc, "thz_tf_row_thresholds")
    return tau


def _tf_row_thresholds_numpy(rowvec, colvec, scal):
    """Fold the per-bin keep conditions into one threshold per row (natural bin order in and out).

    keep(r, c) = (rowvec.y[r] + colvec.y[c] <= 1) & (rowvec.z[r] + colvec.z[c] <= 1) & !(klam^2 - (Kx^2[r] + Ky^2[c]) < 0)
    (Props/ASM_Prop.py:262, :290-301, fp32 op order) is, for a fixed row, monotone non-increasing along the columns
    sorted by Ky^2: the two band-limit quotients are non-decreasing in Ky^2 (checked below) and a rounded fp32 add /
    subtract with one operand fixed is monotone.  tau[r] = the largest Ky^2 that is kept (-1 if none), so that
    keep(r, c) == (Ky^2[c] <= tau[r]) exactly; found by a binary search over the sorted columns, vectorised over
    wavelengths and rows (numpy float32: same IEEE single arithmetic, less per-call overhead than torch) --
    O((Hp + Wp) log Wp), a depth sweep rebuilds this for every z.  Returns tau [C,Hp], or None if the quotients are
    not monotone (never observed; the caller then falls back to the cached-table mode)."""
    import numpy as np
    rv = rowvec.numpy()
    cv = colvec.numpy()
    klam2 = scal[:, 0].numpy().reshape(-1, 1)
    C, Hp, _ = rv.shape
    Wp = cv.shape[1]
    order = np.argsort(cv[:, :, 0], axis=1, kind="stable")
    ky2 = np.take_along_axis(cv[:, :, 0], order, axis=1)
    b1 = np.take_along_axis(cv[:, :, 1], order, axis=1)
    b2 = np.take_along_axis(cv[:, :, 2], order, axis=1)
    if (b1[:, 1:] < b1[:, :-1]).any() or (b2[:, 1:] < b2[:, :-1]).any():
        return None
    r0, r1, r2 = rv[:, :, 0], rv[:, :, 1], rv[:, :, 2]
    one = np.float32(1)
    zero = np.float32(0)
    lo = np.zeros((C, Hp), dtype=np.int64)                # invariant: sorted columns [0, lo) kept, [hi, Wp) dropped
    hi = np.full((C, Hp), Wp, dtype=np.int64)
    for _ in range(int(Wp).bit_length() + 1):
        mid = (lo + hi) >> 1
        m = np.minimum(mid, Wp - 1)
        k = np.take_along_axis(ky2, m, axis=1)
        ok = ((r1 + np.take_along_axis(b1, m, axis=1)) <= one) & ((r2 + np.take_along_axis(b2, m, axis=1)) <= one)
        ok &= ~((klam2 - (r0 + k)) < zero)
        ok &= mid < hi
        lo = np.where(ok, mid + 1, lo)
        hi = np.where(ok, hi, np.minimum(hi, np.maximum(mid, lo)))
    t = np.take_along_axis(ky2, np.maximum(lo - 1, 0), axis=1)
    tau = np.where(lo > 0, t, np.float32(-1.0)).astype(np.float32)
    return torch.from_numpy(tau)


def _tf_row_thresholds_dense(rowvec, colvec, scal, chunk_rows=512):
    """O(Hp Wp) evaluation of the same thresholds with an explicit monotonicity check (tests compare the two)."""
    C, Hp, _ = rowvec.shape
    tau = torch.empty(C, Hp)
    for c in range(C):
        order = torch.argsort(colvec[c, :, 0], stable=True)
        ky2 = colvec[c, order, 0]
        b1, b2 = colvec[c, order, 1], colvec[c, order, 2]
        for r0 in range(0, Hp, chunk_rows):
            rv = rowvec[c, r0:r0 + chunk_rows]
            keep = ((rv[:, None, 1] + b1[None, :]) <= 1) & ((rv[:, None, 2] + b2[None, :]) <= 1)
            keep &= ~((scal[c, 0] - (rv[:, None, 0] + ky2[None, :])) < 0)
            if bool((keep[:, 1:] & ~keep[:, :-1]).any()):
                return None
            cnt = keep.sum(dim=1)
            t = ky2[(cnt - 1).clamp_min(0)]
            tau[c, r0:r0 + chunk_rows] = torch.where(cnt > 0, t, torch.full_like(t, -1.0))
    return tau


def row_vectors_chunked(Hp):
    """True if the column kernels for length Hp read the row vectors in the chunked layout (thz_asm_desc.tf_row_chunked):
    the static kernels do, the runtime-planned engine reads plain slot order."""
    return bool(N.lib().thz_fft_is_static(int(Hp)))


def chunk_row_vectors(rowtau, last_radix):
    """[C,Hp,2] slot order -> [C, R/2, Hp/R, 4]: entry (q, u) = slots R u + 2 q and R u + 2 q + 1 (R = last radix)."""
    C, Hp, _ = rowtau.shape
    R = int(last_radix)
    return rowtau.reshape(C, Hp // R, R // 2, 4).permute(0, 2, 1, 3).contiguous()


def tf_device_vectors(rowvec, colvec, scal, slot_to_bin_fn=None, chunked=False):
    """What the kernels consume (tf_mode 0): rowtau [C,Hp,2] = {Kx^2, tau} and colk2 [C,Wp] = Ky^2, both in the
    plans' slot order, plus scal.  chunked=True re-lays rowtau out for the static column kernels (pass
    tf_row_chunked=1 in the descriptor).  Returns None if the thresholds cannot be formed."""
    tau = tf_row_thresholds(rowvec, colvec, scal)
    if tau is None:
        return None
    pr = N.slot_to_bin(rowvec.shape[1], slot_to_bin_fn)
    pc = N.slot_to_bin(colvec.shape[1], slot_to_bin_fn)
    rowtau = torch.stack([rowvec[:, :, 0], tau], dim=2)[:, pr].contiguous()
    if chunked:
        rowtau = chunk_row_vectors(rowtau, N.plan_radices(rowvec.shape[1])[-1])
    colk2 = colvec[:, pc, 0].contiguous()
    return rowtau, colk2, scal


def tf_centred_reference_order(Hp, Wp, spacing, wavelengths, z, bandlimit=True, bandlimit_type="exact"):
    """Full centred transfer function [C,Hp,Wp] complex64 assembled on the host from the separable
    vectors with the same torch CPU ops the reference uses per pixel (sqrt, exp(1j .)), so that the
    cached-table mode reproduces the reference kernel bit for bit (incl. torch's CPU sqrt/exp)."""
    rowvec, colvec, scal = tf_vectors(Hp, Wp, spacing, wavelengths, z, bandlimit, bandlimit_type)
    rv = torch.fft.fftshift(rowvec, dim=1)
    cv = torch.fft.fftshift(colvec, dim=1)
    K2 = rv[:, :, None, 0] + cv[:, None, :, 0]
    d = scal[:, 0, None, None] - K2
    ang = scal[:, 1, None, None] * torch.sqrt(d)
    out = torch.exp(1j * ang)
    out[d < 0] = 0
    keep = ((rv[:, :, None, 1] + cv[:, None, :, 1]) <= 1) & ((rv[:, :, None, 2] + cv[:, None, :, 2]) <= 1)
    out[~keep] = 0
    return out


def tf_angle_quarter(Hp, Wp, spacing, wavelengths, z):
    """Phase angles z * sqrt(klam^2 - (Kx^2 + Ky^2)) of the reference's transfer function (Props/ASM_Prop.py:245-257, same
    fp32 ops, same torch CPU sqrt) on the unique quarter of the frequency grid: [C, Hp//2+1, Wp//2+1], entry (i, j) = the
    bins at centred distance i, j from DC.  Kx^2 / Ky^2 are exactly even in the centred index (k and -k are exact negatives),
    so expanding the quarter reproduces the full-grid angles bit for bit (tests) at a quarter of the host work."""
    rowvec, colvec, scal = tf_vectors(Hp, Wp, spacing, wavelengths, z, False, "exact")
    rv = torch.fft.fftshift(rowvec[:, :, 0], dim=1)
    cv = torch.fft.fftshift(colvec[:, :, 0], dim=1)
    ru = torch.flip(rv[:, :Hp // 2 + 1], dims=[1]).contiguous()        # centred indices Hp//2 .. 0  ->  distance 0 .. Hp//2
    cu = torch.flip(cv[:, :Wp // 2 + 1], dims=[1]).contiguous()
    K2 = ru[:, :, None] + cu[:, None, :]
    d = scal[:, 0, None, None] - K2
    return (scal[:, 1, None, None] * torch.sqrt(d)).contiguous()


def abs_bin_of_slots(n, slot_to_bin_fn=None):
    """int32 [n]: distance from DC of the frequency bin every slot of the length-n plan holds (min(b, n - b), b = bin)."""
    b = N.slot_to_bin(n, slot_to_bin_fn)
    ch = n // 2
    return torch.where(b < n - ch, b, n - b).to(torch.int32)


def tf_table_device(Hp, Wp, spacing, wavelengths, z, bandlimit, bandlimit_type, device):
    """The tf_mode 1 table [C, Wp, Hp] built ON THE DEVICE from host-evaluated quarter angles (thz_tf_table_from_angles);
    None if the band-limit mask cannot be folded into row thresholds (the caller then builds the whole table on the host).
    ~4x less host arithmetic and ~8x less upload than tf_table_slot_order(tf_centred_reference_order(...))."""
    rowvec, colvec, scal = tf_vectors(Hp, Wp, spacing, wavelengths, z, bandlimit, bandlimit_type)
    dv = tf_device_vectors(rowvec, colvec, scal, chunked=False)
    if dv is None:
        return None
    rowtau, colk2, _ = dv
    angq = tf_angle_quarter(Hp, Wp, spacing, wavelengths, z)
    C = angq.shape[0]
    table = torch.empty(C, Wp, Hp, dtype=torch.complex64, device=device)
    d_ang, d_rt, d_ck = angq.to(device), rowtau.to(device), colk2.to(device)
    d_ra, d_ca = abs_bin_of_slots(Hp).to(device), abs_bin_of_slots(Wp).to(device)
    N.check(N.lib().thz_tf_table_from_angles(N.ptr(d_ang), C, Hp // 2 + 1, Wp // 2 + 1, N.ptr(d_rt), N.ptr(d_ck), N.ptr(d_ra), N.ptr(d_ca),
                                             Hp, Wp, N.ptr(table), N.current_stream_ptr(device)), "thz_tf_table_from_angles")
    return table


_unit_estimates = {}


def inregister_estimate_for(Hp, Wp, spacing, wavelengths, z, bandlimit=True, bandlimit_type="exact"):
    """inregister_deviation_estimate, amortised over depth sweeps: the set of bins where the two square roots disagree does
    not depend on z (d = klam^2 - K^2 does not), only the angle error scales with it, so the estimate is |z| times a
    per-geometry constant that is evaluated once (at the first z asked for) and cached."""
    sp = _spacing2(spacing)
    lam = _as_f32_cpu(wavelengths).reshape(-1)
    key = (Hp, Wp, tuple(sp.tolist()), tuple(lam.tolist()), bool(bandlimit), bandlimit_type)
    zf = abs(float(_as_f32_cpu(z).reshape(())))
    unit = _unit_estimates.get(key)
    if unit is None:
        z0 = zf if zf > 0 else 1.0
        unit = inregister_deviation_estimate(Hp, Wp, sp, lam, torch.tensor(z0), bandlimit, bandlimit_type) / z0
        if len(_unit_estimates) > 64:
            _unit_estimates.clear()
        _unit_estimates[key] = unit
    return unit * zf


INREGISTER_BUDGET = 5e-6     # kernel_mode='auto' generates H in registers only if the estimate below stays under this


def inregister_deviation_estimate(Hp, Wp, spacing, wavelengths, z, bandlimit=True, bandlimit_type="exact", max_rows=64):
    """rel-L2 distance between the reference's transfer function and the one the column kernel generates in registers,
    estimated on `max_rows` evenly spaced rows of the centred grid.

    The mask is bit-identical by construction; the phase z*sqrt(klam^2 - K^2) is not: torch's CPU sqrt (MKL VML, what the
    reference runs: Props/ASM_Prop.py:257) is off by one ulp from the correctly rounded value on ~0.7 % of the elements,
    and the GPU (like numpy) rounds correctly.  One ulp of sqrt(d) ~ k is 2^-11 .. 2^-10 at THz wavelengths, times z:
    3e-6 at z = 0.1 m, 1.5e-5 at z = 0.3 m.  Both square roots are evaluated here with the libraries in question, so the
    estimate needs no model of either."""
    rowvec, colvec, scal = tf_vectors(Hp, Wp, spacing, wavelengths, z, bandlimit, bandlimit_type)
    ridx = torch.unique(torch.linspace(0, Hp - 1, min(Hp, max_rows)).round().long())
    rv, cv = rowvec[:, ridx], colvec
    K2 = rv[:, :, None, 0] + cv[:, None, :, 0]
    d = scal[:, 0, None, None] - K2
    keep = ((rv[:, :, None, 1] + cv[:, None, :, 1]) <= 1) & ((rv[:, :, None, 2] + cv[:, None, :, 2]) <= 1) & ~(d < 0)
    if not bool(keep.any()):
        return 0.0
    zc = scal[:, 1, None, None]
    ang_ref = zc * torch.sqrt(d)                                              # the reference's library
    with np.errstate(invalid="ignore"):
        ang_gpu = zc * torch.from_numpy(np.sqrt(d.numpy()))                    # IEEE-correct, as sqrt.rn on the GPU
    delta = (ang_ref - ang_gpu)[keep].double()
    return float(torch.sqrt(torch.mean(delta * delta)))


def tf_table_slot_order(Hc, slot_to_bin_fn=None):
    """Centred kernel [C,Hp,Wp] -> table[c][slot_c][slot_r] = ifftshift(Hc)[c][bin(slot_r)][bin(slot_c)]  ([C,Wp,Hp]: column-major,
    the order the column pass consumes it -- each thread multiplies R consecutive rows of one column with 16-byte loads)."""
    Hn = torch.fft.ifftshift(Hc, dim=(-2, -1))
    pr = N.slot_to_bin(Hn.shape[-2], slot_to_bin_fn)
    pc = N.slot_to_bin(Hn.shape[-1], slot_to_bin_fn)
    return Hn[:, pr][:, :, pc].transpose(1, 2).contiguous()


def critical_distance(Hp, spacing, wavelengths):
    """Zc of Props/ASM_Prop.py:280."""
    spacing = _spacing2(spacing)
    lam = _as_f32_cpu(wavelengths).reshape(-1)
    dx = spacing[0:1]
    return (Hp * dx ** 2) * torch.sqrt(1 - (torch.max(lam) / (2 * dx)) ** 2) / torch.max(lam)


def doe_coefficients(wavelengths, epsilon, tand):
    """[C,4] = {k_c, tand, sqrt(eps), sqrt(eps)-1} in fp32 (Components/QuantizedDOE.py:67-74)."""
    lam = _as_f32_cpu(wavelengths).reshape(-1)
    k = 2 * torch.pi / lam
    eps = _as_f32_cpu(epsilon).reshape(())
    td = _as_f32_cpu(tand).reshape(())
    se = torch.sqrt(eps)
    return torch.stack([k, td.expand_as(k), se.expand_as(k), (se - 1).expand_as(k)], dim=1).contiguous()


def build_desc(x, y, B, C, inH, inW, Hp, Wp, in_r0, in_c0, outH, outW, out_r0, out_c0, tf_mode, tf_conj,
               rowvec, colvec, scal, table, doe_mode, doe_base, hmap, coef, xsaved, gh, tw_h, tw_w, ws,
               bc_chunk=0, tune_k2_cols=0, tune_lines=0, stages=0, slab=None, tf_row_chunked=0, elem_mode=0, elem_mask=None, elem_mul=None,
               level_idx=None, level_phase=None):
    """Fill a thz_asm_desc from tensors (device or, in the CPU replay tests, host tensors)."""
    d = N.AsmDesc()
    d.B, d.C, d.inH, d.inW, d.Hp, d.Wp = B, C, inH, inW, Hp, Wp
    d.in_r0, d.in_c0, d.outH, d.outW, d.out_r0, d.out_c0 = in_r0, in_c0, outH, outW, out_r0, out_c0
    d.x, d.y = N.ptr(x), N.ptr(y)
    d.tf_mode, d.tf_conj = tf_mode, tf_conj
    d.tf_rowvec, d.tf_colvec, d.tf_scal, d.tf_table = N.ptr(rowvec), N.ptr(colvec), N.ptr(scal), N.ptr(table)
    d.doe_mode, d.doe_base = doe_mode, doe_base
    d.doe_hmap, d.doe_coef, d.doe_xsaved, d.doe_gh = N.ptr(hmap), N.ptr(coef), N.ptr(xsaved), N.ptr(gh)
    d.tw_h, d.tw_w = N.ptr(tw_h), N.ptr(tw_w)
    d.ws = N.ptr(ws)
    d.ws_bytes = ws.numel() * ws.element_size() if ws is not None else 0
    d.tf_row_chunked = int(tf_row_chunked)
    d.elem_mode, d.elem_mask, d.elem_mul = int(elem_mode), N.ptr(elem_mask), N.ptr(elem_mul)
    d.doe_level_idx, d.doe_level_phase = N.ptr(level_idx), N.ptr(level_phase)
    d.doe_levels = int(level_phase.shape[-1]) if level_phase is not None else 0
    if slab is not None:          # (parts, row0, rows, [pointer of every rank's column slab][, blocked]); see thz_asm_desc.slab_*
        d.slab_parts, d.slab_row0, d.slab_rows = int(slab[0]), int(slab[1]), int(slab[2])
        d.slab_blocked = int(slab[4]) if len(slab) > 4 else 0
        for i, q in enumerate(slab[3]):
            d.slab_ptrs[i] = int(q)
    d.bc_chunk, d.tune_k2_cols, d.tune_lines, d.stages = bc_chunk, tune_k2_cols, tune_lines, stages
    return d


def workspace_elems(B, C, inH, outH, Wp, bc_chunk=0, Hp=None, stages=0):
    """complex64 elements of workspace thz_asm_propagate needs (thz_asm_workspace_bytes): one intermediate of the live rows,
    two when both transform lengths are served by the static kernels (blocked row spectra + row-major column-pass output)."""
    d = N.AsmDesc()
    d.B, d.C, d.inH, d.outH, d.Wp, d.Hp = B, C, inH, outH, Wp, (Hp if Hp is not None else 0)
    d.bc_chunk, d.stages = bc_chunk, stages
    return int(N.lib().thz_asm_workspace_bytes(ctypes.byref(d))) // 8
