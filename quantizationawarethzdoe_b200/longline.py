"""Canvases with an edge longer than the in-shared-memory plans hold (> 16384 points, e.g. a 16384^2 field padded 2x).

The reference hands any size to torch.fft (utils/Helper_Functions.py:141-150, Props/ASM_Prop.py:329-341).  Here one outer
decimation-in-frequency step per long axis (thz_split_pre / thz_split_post, include/thzdoe.h) turns an Hp x Wp problem,
Hp = Pr Mr, Wp = Pc Mc with P in {1, 2, 4}, into Pr Pc independent Mr x Mc problems that the fused pipeline runs as extra
"channels" of ONE un-padded thz_asm_propagate call:

    ifft2(H . fft2(pad(x)))[i, j] = 1/(Pr Pc) sum_ab conj(w_r^{a i} w_c^{b j}) . ifft2_M( H[Pr . + a, Pc . + b] . fft2_M(u_ab) )[i mod Mr, j mod Mc]
    u_ab[n, m] = sum_st pad(x)[n + s Mr, m + t Mc] w_r^{a (n + s Mr)} w_c^{b (m + t Mc)}

The transfer function of sub-problem (a, b) is the decimated one: in the in-register mode its separable vectors and row
thresholds are the full grid's, sub-sampled (the keep test Ky^2 <= tau[row] is exact for any subset of columns); in the cached
mode the table is built per sub-problem from the same host-evaluated quarter angles.  The padded canvas itself is never stored:
the split reads the live region, the merge writes the cropped region.  Lines up to 4 x 16384 = 65536 points are served.
"""
import numpy as np
import torch

from . import _native as N
from . import asm_host as AH
from . import functional as Fn

SPLIT_FACTORS = (1, 2, 4)
SCRATCH_BYTES = 16 << 30          # the stacks u, v of one pass are kept below this by walking the batch in groups


def max_line():
    """Longest line transformed directly (16384 = what fits shared memory); tests lower it to exercise the split at small sizes."""
    return int(Fn.TUNE.get("max_line", 16384))


def _plan_ok(n):
    import ctypes
    rad, ns = (ctypes.c_int32 * 16)(), ctypes.c_int32(0)
    return N.lib().thz_fft_plan_info(int(n), rad, ctypes.byref(ns)) == 0


def split_factor(n):
    """Smallest P in {1, 2, 4} such that the length-n/P line is transformed directly; None if there is none."""
    for P in SPLIT_FACTORS:
        if n % P == 0 and n // P <= max_line() and _plan_ok(n // P):
            return P
    return None


def needs_split(Hp, Wp):
    return max(Hp, Wp) > max_line()


_line_tw = {}


def line_twiddles(n, device):
    """w^j = exp(-2 pi i j / n), j < n, complex64 on `device` (evaluated in float64, once per length and device)."""
    key = (int(n), str(device))
    t = _line_tw.get(key)
    if t is None:
        if len(_line_tw) > 32:
            _line_tw.clear()
        j = np.arange(n, dtype=np.float64)
        t = _line_tw[key] = torch.from_numpy(np.exp(-2j * np.pi * j / n).astype(np.complex64)).to(device)
    return t


def split_pre(x, region, Hp, Wp, Pr, Pc, conj_tw=False, scale=1.0):
    """x complex64 [F..., H, W] = the region (H, W, r0, c0) of the canvas -> u [F..., Pr Pc, Hp/Pr, Wp/Pc]."""
    H, W, r0, c0 = region
    F = x.numel() // max(H * W, 1)
    u = torch.empty(tuple(x.shape[:-2]) + (Pr * Pc, Hp // Pr, Wp // Pc), dtype=torch.complex64, device=x.device)
    N.check(N.lib().thz_split_pre(N.ptr(x), N.ptr(u), F, H, W, r0, c0, Hp, Wp, Pr, Pc, N.ptr(line_twiddles(Hp, x.device)),
                                  N.ptr(line_twiddles(Wp, x.device)), 1 if conj_tw else 0, float(scale),
                                  N.current_stream_ptr(x.device)), "thz_split_pre")
    return u


def split_post(v, y, region, Hp, Wp, Pr, Pc, conj_tw=False, scale=1.0):
    """v complex64 [F..., Pr Pc, Hp/Pr, Wp/Pc] -> y [F..., H, W] = the region (H, W, r0, c0) of the merged canvas."""
    H, W, r0, c0 = region
    F = y.numel() // max(H * W, 1)
    N.check(N.lib().thz_split_post(N.ptr(v), N.ptr(y), F, H, W, r0, c0, Hp, Wp, Pr, Pc, N.ptr(line_twiddles(Hp, y.device)),
                                   N.ptr(line_twiddles(Wp, y.device)), 1 if conj_tw else 0, float(scale),
                                   N.current_stream_ptr(y.device)), "thz_split_post")
    return y


# ------------------------------------------------------------------------------- decimated transfer functions
def split_tf_vectors(rowvec, colvec, scal, Pr, Pc):
    """Full-grid separable vectors (asm_host.tf_vectors, FFT-bin order) -> what the Mr x Mc sub-problems consume (tf_mode 0):
    rowtau [C Pr Pc, Mr, 2] (chunked where the static column kernels want it), colk2 [C Pr Pc, Mc], scal [C Pr Pc, 2] and the
    chunked flag; sub-channel index = (c Pr + a) Pc + b.  None if the keep mask does not fold into row thresholds."""
    tau = AH.tf_row_thresholds(rowvec, colvec, scal)
    if tau is None:
        return None
    C, Hp, Wp = rowvec.shape[0], rowvec.shape[1], colvec.shape[1]
    Mr, Mc = Hp // Pr, Wp // Pc
    pr, pc = N.slot_to_bin(Mr), N.slot_to_bin(Mc)
    full = torch.stack([rowvec[:, :, 0], tau], dim=2)                                   # [C, Hp, 2], bin Pr k + a
    rt = full.reshape(C, Mr, Pr, 2).permute(0, 2, 1, 3)[:, :, pr]                       # [C, Pr, Mr(slot), 2]
    rt = rt[:, :, None].expand(C, Pr, Pc, Mr, 2).reshape(C * Pr * Pc, Mr, 2).contiguous()
    ck = colvec[:, :, 0].reshape(C, Mc, Pc).permute(0, 2, 1)[:, :, pc]                  # [C, Pc, Mc(slot)]
    ck = ck[:, None].expand(C, Pr, Pc, Mc).reshape(C * Pr * Pc, Mc).contiguous()
    sc = scal[:, None].expand(C, Pr * Pc, 2).reshape(C * Pr * Pc, 2).contiguous()
    chunked = AH.row_vectors_chunked(Mr)
    rt_k = AH.chunk_row_vectors(rt, N.plan_radices(Mr)[-1]) if chunked else rt
    return rt_k, ck, sc, chunked, rt


def split_tables_from_natural(Hn, Pr, Pc):
    """Natural-order transfer function [C, Hp, Wp] (host or device) -> tf_mode 1 tables of the sub-problems
    [C Pr Pc, Mc, Mr] (table[c'][slot_c][slot_r], include/thzdoe.h)."""
    C, Hp, Wp = Hn.shape
    Mr, Mc = Hp // Pr, Wp // Pc
    pr, pc = N.slot_to_bin(Mr).to(Hn.device), N.slot_to_bin(Mc).to(Hn.device)
    sub = Hn.reshape(C, Mr, Pr, Mc, Pc).permute(0, 2, 4, 1, 3)                          # [C, Pr, Pc, Mr, Mc]
    sub = sub[:, :, :, pr][:, :, :, :, pc]
    return sub.transpose(-1, -2).reshape(C * Pr * Pc, Mc, Mr).contiguous()


def _abs_bin_full(n_full, P, a, M):
    """int32 [M]: |centred frequency index| on the FULL length-n_full grid of the bin every slot of sub-problem a holds."""
    q = N.slot_to_bin(M) * P + a
    return torch.where(q < n_full - n_full // 2, q, n_full - q).to(torch.int32)


def split_tf_table_device(Hp, Wp, spacing, wavelengths, z, bandlimit, bandlimit_type, device, Pr, Pc):
    """tf_mode 1 tables of the sub-problems [C Pr Pc, Mc, Mr], expanded ON THE DEVICE from the reference's own angles on the
    unique quarter of the full grid (asm_host.tf_angle_quarter; thz_tf_table_from_angles once per sub-problem)."""
    rowvec, colvec, scal = AH.tf_vectors(Hp, Wp, spacing, wavelengths, z, bandlimit, bandlimit_type)
    sv = split_tf_vectors(rowvec, colvec, scal, Pr, Pc)
    if sv is None:
        Hc = AH.tf_centred_reference_order(Hp, Wp, spacing, wavelengths, z, bandlimit, bandlimit_type)
        return split_tables_from_natural(torch.fft.ifftshift(Hc, dim=(-2, -1)), Pr, Pc).to(device)
    _, ck, _, _, rt = sv
    C = rowvec.shape[0]
    Mr, Mc = Hp // Pr, Wp // Pc
    angq = AH.tf_angle_quarter(Hp, Wp, spacing, wavelengths, z).to(device)              # [C, Hp/2+1, Wp/2+1]
    Hu, Wu = angq.shape[1], angq.shape[2]
    d_rt, d_ck = rt.to(device), ck.to(device)
    rabs = [_abs_bin_full(Hp, Pr, a, Mr).to(device) for a in range(Pr)]
    cabs = [_abs_bin_full(Wp, Pc, b, Mc).to(device) for b in range(Pc)]
    table = torch.empty(C * Pr * Pc, Mc, Mr, dtype=torch.complex64, device=device)
    stream = N.current_stream_ptr(device)
    for c in range(C):
        for a in range(Pr):
            for b in range(Pc):
                ch = (c * Pr + a) * Pc + b
                N.check(N.lib().thz_tf_table_from_angles(N.ptr(angq[c]), 1, Hu, Wu, N.ptr(d_rt[ch]), N.ptr(d_ck[ch]), N.ptr(rabs[a]),
                                                         N.ptr(cabs[b]), Mr, Mc, N.ptr(table[ch]), stream), "thz_tf_table_from_angles")
    return table


# ------------------------------------------------------------------------------- the plan
class SplitAsmPlan:
    """Same contract as functional.AsmPlan (geometry attributes, run(x, y, conj)) for a canvas with an edge above max_line():
    split -> ONE un-padded fused propagation of C Pr Pc channels -> merge.  DOE / pointwise-element fusion is not offered on
    this path (the callers apply those with their own kernels first)."""

    def __init__(self, B, C, H, W, pad_h, pad_w, Hp, Wp, unpad, device, Pr, Pc, vectors=None, table=None):
        self.B, self.C, self.H, self.W = B, C, H, W
        self.pad_h, self.pad_w, self.Hp, self.Wp = pad_h, pad_w, Hp, Wp
        self.unpad = unpad
        if unpad:
            self.outH, self.outW, self.out_r0, self.out_c0 = H, W, pad_h, pad_w
        else:
            self.outH, self.outW, self.out_r0, self.out_c0 = Hp, Wp, 0, 0
        self.device = device
        self.Pr, self.Pc = Pr, Pc
        Mr, Mc = Hp // Pr, Wp // Pc
        if vectors is not None:
            rt, ck, sc, chunked = vectors[:4]
            self.sub = Fn.AsmPlan(B, C * Pr * Pc, Mr, Mc, 0, 0, Mr, Mc, False, device, rt, ck, sc, None, 0, row_chunked=chunked)
        else:
            self.sub = Fn.AsmPlan(B, C * Pr * Pc, Mr, Mc, 0, 0, Mr, Mc, False, device, None, None, None, table, 1)
        self.tf_mode = self.sub.tf_mode

    def _group(self, B):
        per_b = self.C * self.Hp * self.Wp * 8 * 2
        return max(1, min(B, SCRATCH_BYTES // max(per_b, 1)))

    def run(self, x, y, conj, doe_mode=0, hmap=None, coef=None, xsaved=None, gh=None, hmap_bstride=0, elem=None):
        if doe_mode != 0 or (elem is not None and (elem[0] is not None or elem[1] is not None)):
            raise NotImplementedError("DOE / element fusion is not available on canvases above %d points per edge" % max_line())
        region_in = (self.H, self.W, self.pad_h, self.pad_w)
        region_out = (self.outH, self.outW, self.out_r0, self.out_c0)
        if conj:
            region_in, region_out = region_out, region_in
        B = x.shape[0]
        g = self._group(B)
        for b0 in range(0, B, g):
            xb, yb = x[b0:b0 + g], y[b0:b0 + g]
            u = split_pre(xb, region_in, self.Hp, self.Wp, self.Pr, self.Pc)
            u = u.reshape(xb.shape[0], self.C * self.Pr * self.Pc, self.Hp // self.Pr, self.Wp // self.Pc)
            v = torch.empty_like(u)
            self.sub.run(u, v, conj=conj)
            del u
            split_post(v, yb, region_out, self.Hp, self.Wp, self.Pr, self.Pc, scale=1.0 / (self.Pr * self.Pc))
        return y


def split_or_none(Hp, Wp):
    """(Pr, Pc) for a canvas that needs the split, None for one that does not; NotImplementedError if an edge cannot be served."""
    if not needs_split(Hp, Wp):
        return None
    Pr, Pc = split_factor(Hp), split_factor(Wp)
    if Pr is None or Pc is None:
        raise NotImplementedError("canvas %d x %d: an edge above %d points must be 2 or 4 times a length the radix plans cover "
                                  "(prime factors <= 7, at most %d points)" % (Hp, Wp, max_line(), max_line()))
    return Pr, Pc


def table_plan(B, C, H, W, pad_h, pad_w, Hp, Wp, unpad, device, Hn):
    """A cached-table plan for a natural-order transfer function Hn [C, Hp, Wp] (device): functional.AsmPlan, or the split
    plan when an edge is above max_line().  Used by the convolution-type callers (RSC_prop, chirp-z)."""
    sp = split_or_none(Hp, Wp)
    if sp is None:
        pr, pc = N.slot_to_bin(Hp).to(Hn.device), N.slot_to_bin(Wp).to(Hn.device)
        table = Hn[:, pr][:, :, pc].transpose(1, 2).contiguous()
        return Fn.AsmPlan(B, C, H, W, pad_h, pad_w, Hp, Wp, unpad, device, None, None, None, table, 1)
    return SplitAsmPlan(B, C, H, W, pad_h, pad_w, Hp, Wp, unpad, device, sp[0], sp[1], table=split_tables_from_natural(Hn, *sp))


# ------------------------------------------------------------------------------- stand-alone 2-D FFT
def fft2_split(x, inverse=False, ortho=False):
    """Natural-order 2-D DFT of fields whose edge is above max_line(): split, batched thz_fft2_c2c of the Pr Pc sub-arrays,
    interleave (X[Pr k + a, Pc l + b] = sub_ab[k, l])."""
    H, W = x.shape[-2], x.shape[-1]
    sp = split_or_none(H, W)
    Pr, Pc = sp
    Mr, Mc = H // Pr, W // Pc
    P2 = Pr * Pc
    scale = (1.0 / np.sqrt(P2)) if ortho else ((1.0 / P2) if inverse else 1.0)
    u = split_pre(x.reshape(-1, H, W), (H, W, 0, 0), H, W, Pr, Pc, conj_tw=inverse, scale=scale)
    X = Fn.fft2_c2c(u.reshape(-1, Mr, Mc), inverse=inverse, ortho=ortho)
    del u
    out = torch.empty_like(x)
    out.view(-1, Mr, Pr, Mc, Pc).copy_(X.view(-1, Pr, Pc, Mr, Mc).permute(0, 3, 1, 4, 2))
    return out
