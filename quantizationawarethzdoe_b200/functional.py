"""torch.autograd.Functions over the C ABI: explicit forward kernels and explicit adjoint kernels.

Nothing in here computes on the CPU or through torch.fft: each Function validates that its inputs
live on a CUDA device, allocates outputs / workspaces with torch (the caching allocator owns all
device memory), and calls libthzdoe.so on torch's current stream.
"""
import ctypes

import torch

from . import _native as N
from . import asm_host as AH

BASE_PLANE_THICKNESS = 2 * 1e-3   # Components/QuantizedDOE.py:23

# tuning knobs (bench.py / tests may override): fields per kernel group, K2 tile width, rows per CTA
import os as _os

TUNE = {"bc_chunk": int(_os.environ.get("THZ_BC_CHUNK", "0")), "k2_cols": 0, "lines": 0,   # 0 = library defaults
        "czt_impl": "auto"}
CZT_IMPL = {"auto": 0, "tc": 1, "simt": 2}      # thz_toeplitz_gemm_desc.impl

_ws_cache = {}      # (device, stream) -> [tensor, handed_out_during_capture]
_ws_retired = []    # superseded buffers that a CUDA graph may still reference: kept alive for the life of the process


def _asm_call(desc, device):
    """The one place the fused ASM pipeline is invoked (tests/ replace it with the CPU replay of the kernels)."""
    N.check(N.lib().thz_asm_propagate(ctypes.byref(desc), N.current_stream_ptr(device)), "thz_asm_propagate")


def _workspace(numel, device):
    """Cached complex64 scratch, one per (device, stream): two streams never share a buffer, so concurrent calls cannot
    race on it.  Grown by replacement, but a buffer that was ever handed out while a CUDA graph was being captured is
    retired, not freed -- the graph has its address baked in, and freeing it would let replays scribble over whatever
    the caching allocator put there next."""
    stream = (N.current_stream_ptr(device).value or 0) if device.type == "cuda" else 0
    key = (str(device), stream)
    ent = _ws_cache.get(key)
    capturing = torch.cuda.is_current_stream_capturing() if device.type == "cuda" else False
    if ent is None or ent[0].numel() < numel:
        if ent is not None and ent[1]:
            _ws_retired.append(ent[0])
        ent = _ws_cache[key] = [torch.empty(numel, dtype=torch.complex64, device=device), False]
    if capturing:
        ent[1] = True
    return ent[0]


def _c64(t, name):
    N.require_cuda(t, name)
    if t.dtype != torch.complex64:
        if t.dtype in (torch.float32, torch.float16, torch.bfloat16):
            t = t.to(torch.complex64)
        else:
            raise TypeError("%s must be complex64 (got %s): the sm_100a path computes in fp32" % (name, t.dtype))
    return t.contiguous()


class AsmPlan:
    """Everything static about one ASM_prop call: geometry + device copies of the transfer-function
    vectors / table and DOE coefficients.  Built once per (shape, spacing, wavelengths, z) by ASM_prop."""

    def __init__(self, B, C, H, W, pad_h, pad_w, Hp, Wp, unpad, device, rowvec, colvec, scal, table, tf_mode, row_chunked=False):
        self.B, self.C, self.H, self.W = B, C, H, W
        self.pad_h, self.pad_w, self.Hp, self.Wp = pad_h, pad_w, Hp, Wp
        self.unpad = unpad
        if unpad:
            self.outH, self.outW, self.out_r0, self.out_c0 = H, W, pad_h, pad_w
        else:
            self.outH, self.outW, self.out_r0, self.out_c0 = Hp, Wp, 0, 0
        self.device = device
        self.rowvec = rowvec.to(device) if rowvec is not None else None
        self.colvec = colvec.to(device) if colvec is not None else None
        self.scal = scal.to(device) if scal is not None else None
        self.table = table.to(device) if table is not None else None
        self.tf_mode = tf_mode
        self.row_chunked = 1 if (row_chunked and tf_mode == 0) else 0   # layout of rowvec, see thz_asm_desc.tf_row_chunked
        self.tw_h = N.twiddles(Hp, device)
        self.tw_w = N.twiddles(Wp, device)
        self._descs = {}

    def run(self, x, y, conj, doe_mode=0, hmap=None, coef=None, xsaved=None, gh=None, hmap_bstride=0, elem=None, gh_mode=0,
            levels=None):
        """One thz_asm_propagate call.  Forward: x [B,C,H,W] -> y [B,C,outH,outW].
        Adjoint (conj=1): x = grad [B,C,outH,outW] -> y [B,C,H,W] (regions swapped)."""
        B, C = self.B, self.C
        if not conj:
            inH, inW, in_r0, in_c0 = self.H, self.W, self.pad_h, self.pad_w
            outH, outW, out_r0, out_c0 = self.outH, self.outW, self.out_r0, self.out_c0
        else:
            inH, inW, in_r0, in_c0 = self.outH, self.outW, self.out_r0, self.out_c0
            outH, outW, out_r0, out_c0 = self.H, self.W, self.pad_h, self.pad_w
        # the descriptor of a (batch, direction, DOE mode) triple is built once; per call only the pointers change
        key = (x.shape[0], bool(conj), doe_mode, TUNE["bc_chunk"], TUNE["k2_cols"], TUNE["lines"])
        ent = self._descs.get(key)
        if ent is None:
            elems = AH.workspace_elems(x.shape[0], C, inH, outH, self.Wp, TUNE["bc_chunk"], Hp=self.Hp)
            d = AH.build_desc(None, None, x.shape[0], C, inH, inW, self.Hp, self.Wp, in_r0, in_c0, outH, outW, out_r0, out_c0,
                              self.tf_mode, 1 if conj else 0, self.rowvec, self.colvec, self.scal, self.table,
                              doe_mode, BASE_PLANE_THICKNESS, None, None, None, None, self.tw_h, self.tw_w, None,
                              bc_chunk=TUNE["bc_chunk"], tune_k2_cols=TUNE["k2_cols"], tune_lines=TUNE["lines"],
                              tf_row_chunked=self.row_chunked)
            ent = self._descs[key] = (d, elems)
        d, elems = ent
        ws = _workspace(elems, x.device)
        d.x, d.y, d.ws, d.ws_bytes = N.ptr(x), N.ptr(y), N.ptr(ws), ws.numel() * 8
        d.doe_hmap, d.doe_coef, d.doe_xsaved, d.doe_gh = N.ptr(hmap), N.ptr(coef), N.ptr(xsaved), N.ptr(gh)
        d.doe_hmap_bstride = int(hmap_bstride)
        d.doe_gh_mode = int(gh_mode)           # 1: gh is an NVLS multicast address, partial sums are added (parallel.FusedGradReduce)
        # quantised DOE: (int32 level map [H,W], complex64 per-level transmissions [C,L]) -- the static row kernels look the
        # transmission up instead of evaluating it per pixel and wavelength (thz_asm_desc.doe_level_*); same results
        lidx, lphase = levels if levels is not None else (None, None)
        d.doe_level_idx, d.doe_level_phase = N.ptr(lidx), N.ptr(lphase)
        d.doe_levels = int(lphase.shape[-1]) if lphase is not None else 0
        # pointwise elements in front of the propagation (aperture mask, lens kernel): on load in a forward pass, conjugated in
        # the epilogue of an adjoint pass (thz_asm_desc.elem_*)
        mask, mul = elem if elem is not None else (None, None)
        d.elem_mode = 0 if (mask is None and mul is None) else (2 if conj else 1)
        d.elem_mask, d.elem_mul = N.ptr(mask), N.ptr(mul)
        _asm_call(d, x.device)
        return y


class AsmPropagateFn(torch.autograd.Function):
    """y = ASM(x).  Backward = the adjoint pipeline with conj(H) (explicit kernel, not autograd replay)."""

    @staticmethod
    def forward(ctx, x, plan, mask=None, mul=None):
        """mask (float32 [H,W]) / mul (complex64 [C,H,W]): fixed pointwise elements in front of the propagation, fused."""
        x = _c64(x, "field.data")
        y = torch.empty(x.shape[0], plan.C, plan.outH, plan.outW, dtype=torch.complex64, device=x.device)
        plan.run(x, y, conj=0, elem=(mask, mul))
        ctx.plan, ctx.elem = plan, (mask, mul)
        return y

    @staticmethod
    def backward(ctx, g):
        plan = ctx.plan
        g = _c64(g, "grad_output")
        gx = torch.empty(g.shape[0], plan.C, plan.H, plan.W, dtype=torch.complex64, device=g.device)
        plan.run(g, gx, conj=1, elem=ctx.elem)
        return gx, None, None, None


class DoeAsmFn(torch.autograd.Function):
    """y = ASM(x * p(h)) in one fused pipeline; backward returns grad wrt x (if needed) and wrt h."""

    @staticmethod
    def forward(ctx, x, hmap, plan, coef, mask=None, mul=None, reducer=None, levels=None):
        """reducer (parallel.FusedGradReduce or None): grad_height is summed over the data-parallel ranks inside the adjoint.
        levels (level map int32 [H,W], per-level transmissions complex64 [C,L]) or None: the height map is quantised, hmap =
        lut[level map] -- the row kernels then consume the quantiser's level indices directly (level_phase_table)."""
        x = _c64(x, "field.data")
        N.require_cuda(hmap, "height_map")
        hmap = hmap.to(torch.float32).contiguous()
        y = torch.empty(x.shape[0], plan.C, plan.outH, plan.outW, dtype=torch.complex64, device=x.device)
        plan.run(x, y, conj=0, doe_mode=1, hmap=hmap, coef=coef, elem=(mask, mul), levels=levels)
        ctx.plan, ctx.coef, ctx.elem, ctx.reducer, ctx.levels = plan, coef, (mask, mul), reducer, levels
        ctx.save_for_backward(x, hmap)
        return y

    @staticmethod
    def backward(ctx, g):
        plan, coef = ctx.plan, ctx.coef
        x, hmap = ctx.saved_tensors
        g = _c64(g, "grad_output")
        need_x, need_h = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gx = torch.empty_like(x) if need_x else None
        if need_h and ctx.reducer is not None:
            red = ctx.reducer
            if (red.H, red.W) != (plan.H, plan.W):
                raise ValueError("FusedGradReduce was built for a %d x %d height map, the DOE is %d x %d" % (red.H, red.W, plan.H, plan.W))
            plan.run(g, gx, conj=1, doe_mode=2, hmap=hmap, coef=coef, xsaved=x, gh=red.target(), elem=ctx.elem, gh_mode=1,
                     levels=ctx.levels)
            gh = red.finish().clone()       # the replica is recycled two passes later; autograd may keep what it is handed
        elif need_h:
            gh = torch.empty(plan.H, plan.W, dtype=torch.float32, device=g.device)
            plan.run(g, gx, conj=1, doe_mode=2, hmap=hmap, coef=coef, xsaved=x, gh=gh, elem=ctx.elem, levels=ctx.levels)
        else:
            gh = None
            # grad wrt x only: adjoint ASM (conj(m) of the pointwise elements in its epilogue), then the conj(p) multiply
            gtmp = torch.empty_like(x)
            plan.run(g, gtmp, conj=1, elem=ctx.elem)
            N.check(N.lib().thz_doe_modulate_bwd(N.ptr(gtmp), None, N.ptr(hmap), N.ptr(coef), BASE_PLANE_THICKNESS,
                                                 N.ptr(gx), None, x.shape[0], x.shape[1], x.shape[2], x.shape[3],
                                                 N.current_stream_ptr(g.device)), "thz_doe_modulate_bwd")
        return gx, gh, None, None, None, None, None, None


def level_phase_table(lut, coef):
    """Transmission of every level for every wavelength, complex64 [C, L] = p_c(lut[l]): the modulation kernel itself on a unit
    field over the LUT, hence bit-identical to what the fused prologue evaluates per pixel."""
    lut = lut.detach().to(torch.float32).reshape(1, -1).contiguous()
    C, L = coef.shape[0], lut.shape[1]
    ones = torch.ones(1, C, 1, L, dtype=torch.complex64, device=lut.device)
    out = torch.empty_like(ones)
    N.check(N.lib().thz_doe_modulate_fwd(N.ptr(ones), N.ptr(out), N.ptr(lut), N.ptr(coef), BASE_PLANE_THICKNESS, 1, C, 1, L,
                                         N.current_stream_ptr(lut.device)), "thz_doe_modulate_fwd")
    return out.reshape(C, L)


def doe_asm_sweep(x, hmaps, prop, coef, spacing, wavelengths, mask=None, mul=None):
    """B candidate DOEs over ONE input field in a single fused pass (forward only; loss-landscape sweeps):
    x complex64 [1,C,H,W], hmaps float32 [Bc,H,W] -> prop(x * p(hmaps[b])) for every b, complex64 [Bc,C,outH,outW].
    The field is broadcast over the candidates; the height map is per entry (thz_asm_desc.doe_hmap_bstride)."""
    x = _c64(x, "field.data")
    N.require_cuda(hmaps, "height maps")
    hmaps = hmaps.to(torch.float32).contiguous()
    Bc, H, W = hmaps.shape
    if x.shape[0] != 1 or tuple(x.shape[-2:]) != (H, W):
        raise ValueError("doe_asm_sweep needs one input field [1,C,H,W] matching the height maps")
    C = x.shape[1]
    plan = prop._get_plan(Bc, C, H, W, spacing, wavelengths, x.device)
    xb = x.expand(Bc, C, H, W).contiguous()
    y = torch.empty(Bc, C, plan.outH, plan.outW, dtype=torch.complex64, device=x.device)
    plan.run(xb, y, conj=0, doe_mode=1, hmap=hmaps, coef=coef, hmap_bstride=H * W, elem=(mask, mul))
    return y


class DoeModulateFn(torch.autograd.Function):
    """y = x * p(h) (stand-alone; used when a DOE output is consumed by something other than ASM_prop)."""

    @staticmethod
    def forward(ctx, x, hmap, coef):
        x = _c64(x, "field.data")
        N.require_cuda(hmap, "height_map")
        hmap = hmap.to(torch.float32).contiguous()
        y = torch.empty_like(x)
        B, C, H, W = x.shape
        N.check(N.lib().thz_doe_modulate_fwd(N.ptr(x), N.ptr(y), N.ptr(hmap), N.ptr(coef), BASE_PLANE_THICKNESS,
                                             B, C, H, W, N.current_stream_ptr(x.device)), "thz_doe_modulate_fwd")
        ctx.coef = coef
        ctx.save_for_backward(x, hmap)
        return y

    @staticmethod
    def backward(ctx, g):
        x, hmap = ctx.saved_tensors
        g = _c64(g, "grad_output")
        B, C, H, W = x.shape
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gh = torch.empty(H, W, dtype=torch.float32, device=g.device) if ctx.needs_input_grad[1] else None
        N.check(N.lib().thz_doe_modulate_bwd(N.ptr(g), N.ptr(x), N.ptr(hmap), N.ptr(ctx.coef), BASE_PLANE_THICKNESS,
                                             N.ptr(gx), N.ptr(gh), B, C, H, W, N.current_stream_ptr(g.device)),
                "thz_doe_modulate_bwd")
        return gx, gh, None


# ----------------------------------------------------------------------------- height map / quantizers
def _f32(t, name):
    N.require_cuda(t, name)
    return t.to(torch.float32).contiguous()


class HeightFromWeightFn(torch.autograd.Function):
    """h = hmax * sigmoid(clamp(w, -c, c))  (Components/QuantizedDOE.py:277)."""

    @staticmethod
    def forward(ctx, w, hmax, clampv):
        w = _f32(w, "weight_height_map")
        h = torch.empty_like(w)
        N.check(N.lib().thz_height_fwd(N.ptr(w), hmax, clampv, N.ptr(h), w.numel(), N.current_stream_ptr(w.device)), "thz_height_fwd")
        ctx.hmax, ctx.clampv = hmax, clampv
        ctx.save_for_backward(w)
        return h

    @staticmethod
    def backward(ctx, g):
        (w,) = ctx.saved_tensors
        g = _f32(g, "grad")
        gw = torch.empty_like(w)
        N.check(N.lib().thz_height_bwd(N.ptr(g), N.ptr(w), ctx.hmax, ctx.clampv, N.ptr(gw), w.numel(),
                                       N.current_stream_ptr(w.device)), "thz_height_bwd")
        return gw, None, None


class SteQuantizeFn(torch.autograd.Function):
    """STEQuantizationFunction (Components/QuantizedDOE.py:1239-1253) on a height map; identity backward.
    Returns (q, idx)."""

    @staticmethod
    def forward(ctx, h, lut):
        h = _f32(h, "height_map")
        lut = _f32(lut, "lut")
        q = torch.empty_like(h)
        idx = torch.empty(h.shape, dtype=torch.int32, device=h.device)
        N.check(N.lib().thz_quant_ste_fwd(N.ptr(h), 0, 0.0, 0.0, N.ptr(lut), lut.numel(), N.ptr(q), N.ptr(idx), None,
                                          h.numel(), N.current_stream_ptr(h.device)), "thz_quant_ste_fwd")
        ctx.mark_non_differentiable(idx)
        return q, idx

    @staticmethod
    def backward(ctx, g, _gidx):
        return g.clone(), None


class SteFromWeightFn(torch.autograd.Function):
    """Fused sigmoid height construction + STE level selection (STEQuantizedDOELayer.preprocessed_height_map,
    Components/QuantizedDOE.py:1379-1388).  Returns (q, idx)."""

    @staticmethod
    def forward(ctx, w, lut, hmax, clampv):
        w = _f32(w, "weight_height_map")
        lut = _f32(lut, "lut")
        q = torch.empty_like(w)
        idx = torch.empty(w.shape, dtype=torch.int32, device=w.device)
        N.check(N.lib().thz_quant_ste_fwd(N.ptr(w), 1, hmax, clampv, N.ptr(lut), lut.numel(), N.ptr(q), N.ptr(idx), None,
                                          w.numel(), N.current_stream_ptr(w.device)), "thz_quant_ste_fwd")
        ctx.hmax, ctx.clampv = hmax, clampv
        ctx.save_for_backward(w)
        ctx.mark_non_differentiable(idx)
        return q, idx

    @staticmethod
    def backward(ctx, g, _gidx):
        (w,) = ctx.saved_tensors
        g = _f32(g, "grad")
        gw = torch.empty_like(w)
        N.check(N.lib().thz_height_bwd(N.ptr(g), N.ptr(w), ctx.hmax, ctx.clampv, N.ptr(gw), w.numel(),
                                       N.current_stream_ptr(w.device)), "thz_height_bwd")
        return gw, None, None, None


class NnQuantizeFn(torch.autograd.Function):
    """NearestNeighborSearch / PolyGrad / SigmoidGrad (Components/quantization.py:59-122). kind: 0/1/2."""

    @staticmethod
    def forward(ctx, x, lut, mid, s, kind):
        x = _f32(x, "thickness")
        lut = _f32(lut, "lut")
        mid = _f32(mid, "lut_midvals")
        q = torch.empty_like(x)
        idx = torch.empty(x.shape, dtype=torch.int32, device=x.device)
        N.check(N.lib().thz_quant_nn_fwd(N.ptr(x), N.ptr(lut), lut.numel(), N.ptr(mid), mid.numel(), N.ptr(q), N.ptr(idx),
                                         x.numel(), N.current_stream_ptr(x.device)), "thz_quant_nn_fwd")
        ctx.s, ctx.kind = float(s), int(kind)
        ctx.save_for_backward(x, idx, lut)
        ctx.mark_non_differentiable(idx)
        return q, idx

    @staticmethod
    def backward(ctx, g, _gidx):
        x, idx, lut = ctx.saved_tensors
        g = _f32(g, "grad")
        gx = torch.empty_like(x)
        N.check(N.lib().thz_quant_nn_bwd(N.ptr(g), N.ptr(x), N.ptr(idx), N.ptr(lut), lut.numel(), ctx.s, ctx.kind, N.ptr(gx),
                                         x.numel(), N.current_stream_ptr(x.device)), "thz_quant_nn_bwd")
        return gx, None, None, None, None


class PsqFn(torch.autograd.Function):
    """PSQuantizedDOELayer.preprocessed_height_map (Components/QuantizedDOE.py:1193-1207)."""

    @staticmethod
    def forward(ctx, w, hmax, levels, tau):
        w = _f32(w, "weight_height_map")
        out = torch.empty_like(w)
        d = torch.empty_like(w)
        N.check(N.lib().thz_quant_psq_fwd(N.ptr(w), hmax, levels, tau, N.ptr(out), N.ptr(d), w.numel(),
                                          N.current_stream_ptr(w.device)), "thz_quant_psq_fwd")
        ctx.save_for_backward(d)
        return out

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        return g * d, None, None, None


class GumbelV3Fn(torch.autograd.Function):
    """SoftGumbelQuantizedDOELayerv3 level selection (Components/QuantizedDOE.py:794-860), noise supplied.
    Returns (height_map, idx)."""

    @staticmethod
    def forward(ctx, w, lut, noise, hmax, kfac, c_s, tau, tau_max, beta, phase_input=False):
        w = _f32(w, "weight_init_phase")
        lut = _f32(lut, "lut")
        noise = _f32(noise, "gumbel noise")
        L = lut.numel()
        assert noise.numel() == L * w.numel(), "noise must be [1,L,H,W]"
        out = torch.empty_like(w)
        idx = torch.empty(w.shape, dtype=torch.int32, device=w.device)
        d = torch.empty_like(w)
        s = float(tau_max / tau)
        N.check(N.lib().thz_quant_gumbel_v3_fwd(N.ptr(w), N.ptr(lut), L, N.ptr(noise), hmax, kfac, c_s, tau, tau_max, s,
                                                float(beta), float(1 - beta), 1 if phase_input else 0, N.ptr(out), N.ptr(idx), N.ptr(d),
                                                w.numel(),
                                                N.current_stream_ptr(w.device)), "thz_quant_gumbel_v3_fwd")
        ctx.save_for_backward(d)
        ctx.mark_non_differentiable(idx)
        return out, idx

    @staticmethod
    def backward(ctx, g, _gidx):
        (d,) = ctx.saved_tensors
        return (g * d,) + (None,) * 9


class GumbelNaiveFn(torch.autograd.Function):
    """NaiveGumbelQuantizedDOELayer level selection (Components/QuantizedDOE.py:1022-1031). logits [H,W,L]."""

    @staticmethod
    def forward(ctx, logits, lut, noise, tau):
        logits = _f32(logits, "logits")
        lut = _f32(lut, "lut")
        noise = _f32(noise, "gumbel noise")
        L = lut.numel()
        n = logits.numel() // L
        q = torch.empty(logits.shape[:-1], dtype=torch.float32, device=logits.device)
        idx = torch.empty(logits.shape[:-1], dtype=torch.int32, device=logits.device)
        dq = torch.empty_like(logits)
        N.check(N.lib().thz_quant_gumbel_naive_fwd(N.ptr(logits), N.ptr(noise), N.ptr(lut), L, float(tau), N.ptr(q), N.ptr(idx),
                                                   N.ptr(dq), n, N.current_stream_ptr(logits.device)),
                "thz_quant_gumbel_naive_fwd")
        ctx.save_for_backward(dq)
        ctx.mark_non_differentiable(idx)
        return q, idx

    @staticmethod
    def backward(ctx, g, _gidx):
        (dq,) = ctx.saved_tensors
        return g.unsqueeze(-1) * dq, None, None, None


class SoftmaxQuantizeFn(torch.autograd.Function):
    """SoftmaxBasedQuantization.forward (Components/quantization.py:128-161) on one thickness map; noise=None is the plain
    softmax branch.  Returns (q, idx).  Backward: autograd through diff / max|diff| incl. the path through the max."""

    @staticmethod
    def forward(ctx, thickness, lut, noise, c, tau, s, hard):
        t = _f32(thickness, "thickness")
        lut = _f32(lut, "lut")
        L, n = lut.numel(), t.numel()
        if noise is not None:
            noise = _f32(noise, "gumbel noise")
            assert noise.numel() == L * n, "noise must be [1,L,H,W]"
        q = torch.empty_like(t)
        idx = torch.empty(t.shape, dtype=torch.int32, device=t.device)
        need = ctx.needs_input_grad[0]
        A, Bm, E = (torch.empty_like(t), torch.empty_like(t), torch.empty_like(t)) if need else (None, None, None)
        stats = torch.empty(4, dtype=torch.float32, device=t.device)
        N.check(N.lib().thz_quant_softmax_fwd(N.ptr(t), N.ptr(lut), L, N.ptr(noise), float(c), float(tau), float(s), 1 if hard else 0,
                                              N.ptr(q), N.ptr(idx), N.ptr(A), N.ptr(Bm), N.ptr(E), N.ptr(stats), n,
                                              N.current_stream_ptr(t.device)), "thz_quant_softmax_fwd")
        if need:
            ctx.save_for_backward(A, Bm, E, stats)
        ctx.mark_non_differentiable(idx)
        return q, idx

    @staticmethod
    def backward(ctx, g, _gidx):
        A, Bm, E, stats = ctx.saved_tensors
        g = _f32(g, "grad")
        gt = torch.empty_like(A)
        N.check(N.lib().thz_quant_softmax_bwd(N.ptr(g), N.ptr(A), N.ptr(Bm), N.ptr(E), N.ptr(stats), N.ptr(gt), A.numel(),
                                              N.current_stream_ptr(g.device)), "thz_quant_softmax_bwd")
        return (gt,) + (None,) * 6


SCORE_FUNCS = {"sigmoid": 0, "log": 1, "poly": 2, "sine": 3, "chirp": 4}


def score_thickness(thickness, lut, s=5., func="sigmoid"):
    """score_thickness (Components/quantization.py:36-55): thickness [N,1,H,W], lut [L] or [1,L,1,1] -> scores [N,L,H,W].
    Forward only: the reference's own backward through this function raises (in-place normalisation, :41)."""
    t = _f32(thickness.detach(), "thickness")
    lut = _f32(torch.as_tensor(lut).reshape(-1).to(t.device), "lut")
    if func not in SCORE_FUNCS:
        raise ValueError("func must be one of %s" % sorted(SCORE_FUNCS))
    if t.dim() != 4 or t.shape[1] != 1:
        raise ValueError("thickness must be [N,1,H,W]")
    Nb, _, H, W = t.shape
    scores = torch.empty(Nb, lut.numel(), H, W, dtype=torch.float32, device=t.device)
    stats = torch.empty(4, dtype=torch.float32, device=t.device)
    N.check(N.lib().thz_score_thickness(N.ptr(t), N.ptr(lut), lut.numel(), float(s), SCORE_FUNCS[func], N.ptr(scores), N.ptr(stats),
                                        Nb, H * W, N.current_stream_ptr(t.device)), "thz_score_thickness")
    return scores


class FieldMulFn(torch.autograd.Function):
    """y = x * m for a fixed pointwise element m (complex [C,H,W] per wavelength, or a real mask [H,W]); backward x conj(m)."""

    @staticmethod
    def forward(ctx, x, m):
        x = _c64(x, "field.data")
        N.require_cuda(m, "element")
        real = not m.is_complex()
        per_channel = m.dim() == 3
        if (real and m.dtype != torch.float32) or (not real and m.dtype != torch.complex64) or not m.is_contiguous():
            raise ValueError("pointwise element must be a contiguous float32 [H,W] mask or complex64 [C,H,W] kernel")
        B, C, H, W = x.shape
        if tuple(m.shape[-2:]) != (H, W) or (per_channel and m.shape[0] != C):
            raise ValueError("pointwise element does not match the field")
        ctx.m, ctx.real, ctx.per_channel = m, real, per_channel
        return FieldMulFn._run(x, m, real, per_channel, 0)

    @staticmethod
    def _run(x, m, real, per_channel, conj):
        B, C, H, W = x.shape
        y = torch.empty_like(x)
        N.check(N.lib().thz_field_mul(N.ptr(x), N.ptr(m), N.ptr(y), B * C, C, H * W, 1 if per_channel else 0, 1 if real else 0, conj,
                                      N.current_stream_ptr(x.device)), "thz_field_mul")
        return y

    @staticmethod
    def backward(ctx, g):
        return FieldMulFn._run(_c64(g, "grad_output"), ctx.m, ctx.real, ctx.per_channel, 1), None


_bluestein_fft2 = {}


def fft2_c2c(x, inverse=False, ortho=False):
    """Stand-alone natural-order batched 2-D FFT over the last two dims (thz_fft2_c2c); sizes with a prime factor > 7 go
    through the chirp-z path (bluestein.BluesteinFft2), edges above 16384 points through one outer split (longline.fft2_split)."""
    x = _c64(x, "input")
    H, W = x.shape[-2], x.shape[-1]
    from . import bluestein as BL
    from . import longline as LL
    if LL.needs_split(H, W) and LL.split_factor(H) is not None and LL.split_factor(W) is not None:
        return LL.fft2_split(x, inverse, ortho)       # an edge above 16384 points: one outer decimation step (longline.py)
    if not (BL.length_supported(H) and BL.length_supported(W)):
        key = (H, W, bool(inverse), bool(ortho), str(x.device))
        plan = _bluestein_fft2.get(key)
        if plan is None:
            plan = _bluestein_fft2[key] = BL.BluesteinFft2(H, W, bool(inverse), bool(ortho), x.device)
        return plan(x)
    batch = x.numel() // (H * W)
    y = torch.empty_like(x)
    ws = _workspace(batch * H * W, x.device)
    N.check(N.lib().thz_fft2_c2c(N.ptr(x), N.ptr(y), batch, H, W, 1 if inverse else 0, 1 if ortho else 0,
                                 N.ptr(N.twiddles(H, x.device)), N.ptr(N.twiddles(W, x.device)), N.ptr(ws),
                                 ws.numel() * 8, N.current_stream_ptr(x.device)), "thz_fft2_c2c")
    return y


# ----------------------------------------------------------------------------- chirp-z propagation
def _toeplitz_gemm(batch, M, N_, K, g, L, off, sm, sk, conj_g, B, sb, pro, conj_pro, C, sc, epi, conj_epi, device):
    scratch = _workspace(batch * K * N_, device) if pro is not None else None
    d = N.ToeplitzGemmDesc()
    d.batch, d.M, d.N, d.K = batch, M, N_, K
    d.g, d.L, d.off, d.sm, d.sk = N.ptr(g), L, off, sm, sk
    d.conj_g, d.conj_pro, d.conj_epi, d.impl = conj_g, conj_pro, conj_epi, CZT_IMPL[TUNE.get('czt_impl', 'auto')]
    d.B, (d.sb_b, d.sb_k, d.sb_n) = N.ptr(B), sb
    d.pro = N.ptr(pro)
    d.C, (d.sc_b, d.sc_m, d.sc_n) = N.ptr(C), sc
    d.epi = N.ptr(epi)
    d.scratch = N.ptr(scratch)
    N.check(N.lib().thz_toeplitz_gemm(ctypes.byref(d), N.current_stream_ptr(device)), "thz_toeplitz_gemm")


class CztDevicePlan:
    """Device copies of a czt_host.CztPlan."""

    def __init__(self, hp, device):
        self.P, self.Q = hp.P.to(device), hp.Q.to(device)
        self.gy, self.gx = hp.gy.to(device), hp.gx.to(device)
        self.Ly, self.Lx = hp.Ly, hp.Lx
        self.C, self.H, self.W, self.M1, self.M2 = hp.C, hp.H, hp.W, hp.M1, hp.M2


class CztFn(torch.autograd.Function):
    """out = Q * (Ty . (P * x) . Tx^T) per wavelength, two Toeplitz GEMMs; backward = the two adjoint GEMMs."""

    @staticmethod
    def forward(ctx, x, p):
        x = _c64(x, "field.data")
        Bn, C, H, W = x.shape
        M1, M2 = p.M1, p.M2
        out = torch.empty(Bn, C, M1, M2, dtype=torch.complex64, device=x.device)
        u1 = torch.empty(C, M1, W, dtype=torch.complex64, device=x.device)
        for b in range(Bn):
            _toeplitz_gemm(C, M1, W, H, p.gy, p.Ly, H, 1, -1, 0, x[b], (H * W, W, 1), p.P, 0, u1, (M1 * W, W, 1), None, 0, x.device)
            _toeplitz_gemm(C, M2, M1, W, p.gx, p.Lx, W, 1, -1, 0, u1, (M1 * W, 1, W), None, 0, out[b], (M1 * M2, 1, M2), p.Q, 0,
                           x.device)
        ctx.p = p
        ctx.in_shape = x.shape
        return out

    @staticmethod
    def backward(ctx, g):
        p = ctx.p
        g = _c64(g, "grad_output")
        Bn, C, H, W = ctx.in_shape
        M1, M2 = p.M1, p.M2
        gx = torch.empty(Bn, C, H, W, dtype=torch.complex64, device=g.device)
        v = torch.empty(C, M1, W, dtype=torch.complex64, device=g.device)
        for b in range(Bn):
            # V[c,k1,w] = sum_k2 conj(Tx[k2,w]) conj(Q) G        (m = w, k = k2, n = k1)
            _toeplitz_gemm(C, W, M1, M2, p.gx, p.Lx, W, -1, 1, 1, g[b], (M1 * M2, 1, M2), p.Q, 1, v, (M1 * W, 1, W), None, 0, g.device)
            # gx[c,h,w] = conj(P) sum_k1 conj(Ty[k1,h]) V         (m = h, k = k1, n = w)
            _toeplitz_gemm(C, H, W, M1, p.gy, p.Ly, H, -1, 1, 1, v, (M1 * W, W, 1), None, 0, gx[b], (H * W, W, 1), p.P, 1, g.device)
        return gx, None
