"""ctypes binding of the sm_100a C-ABI library (include/thzdoe.h).

PyTorch is plumbing here: it owns device memory and streams; every hot-path computation happens in
libthzdoe.so.  There is NO CPU or torch fallback: if the library is missing or CUDA is unavailable,
`lib()` raises.
"""
import ctypes
import os
import threading

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("THZ_LIB") or os.path.join(_HERE, "csrc", "libthzdoe.so")      # THZ_LIB: A/B builds (tools/build_variant.sh)

THZ_OK = 0
THZ_E_UNSUPPORTED = -3
_ERR_EXC = {-1: ValueError, -2: ValueError, -3: NotImplementedError, -4: RuntimeError, -5: RuntimeError, -6: RuntimeError}


class ThzError(RuntimeError):
    pass


class AsmDesc(ctypes.Structure):
    """Mirror of `thz_asm_desc` (include/thzdoe.h)."""
    _fields_ = [
        ("B", ctypes.c_int32), ("C", ctypes.c_int32),
        ("inH", ctypes.c_int32), ("inW", ctypes.c_int32),
        ("Hp", ctypes.c_int32), ("Wp", ctypes.c_int32),
        ("in_r0", ctypes.c_int32), ("in_c0", ctypes.c_int32),
        ("outH", ctypes.c_int32), ("outW", ctypes.c_int32),
        ("out_r0", ctypes.c_int32), ("out_c0", ctypes.c_int32),
        ("x", ctypes.c_void_p), ("y", ctypes.c_void_p),
        ("tf_mode", ctypes.c_int32), ("tf_conj", ctypes.c_int32),
        ("tf_rowvec", ctypes.c_void_p), ("tf_colvec", ctypes.c_void_p),
        ("tf_scal", ctypes.c_void_p), ("tf_table", ctypes.c_void_p),
        ("doe_mode", ctypes.c_int32), ("doe_base", ctypes.c_float),
        ("doe_hmap", ctypes.c_void_p), ("doe_coef", ctypes.c_void_p),
        ("doe_xsaved", ctypes.c_void_p), ("doe_gh", ctypes.c_void_p),
        ("tw_h", ctypes.c_void_p), ("tw_w", ctypes.c_void_p),
        ("ws", ctypes.c_void_p), ("ws_bytes", ctypes.c_uint64),
        ("bc_chunk", ctypes.c_int32), ("tune_k2_cols", ctypes.c_int32),
        ("tune_lines", ctypes.c_int32), ("stages", ctypes.c_int32),
        ("slab_parts", ctypes.c_int32), ("slab_row0", ctypes.c_int32),
        ("slab_rows", ctypes.c_int32), ("slab_blocked", ctypes.c_int32),
        ("slab_ptrs", ctypes.c_void_p * 8),
        ("tf_row_chunked", ctypes.c_int32), ("reserved2", ctypes.c_int32),
        ("doe_hmap_bstride", ctypes.c_int64),
        ("elem_mode", ctypes.c_int32), ("doe_gh_mode", ctypes.c_int32),
        ("elem_mask", ctypes.c_void_p), ("elem_mul", ctypes.c_void_p),
        ("doe_level_idx", ctypes.c_void_p), ("doe_level_phase", ctypes.c_void_p),
        ("doe_levels", ctypes.c_int32), ("reserved4", ctypes.c_int32),
    ]


class ToeplitzGemmDesc(ctypes.Structure):
    """Mirror of `thz_toeplitz_gemm_desc` (include/thzdoe.h)."""
    _fields_ = [
        ("batch", ctypes.c_int32), ("M", ctypes.c_int32), ("N", ctypes.c_int32), ("K", ctypes.c_int32),
        ("g", ctypes.c_void_p),
        ("L", ctypes.c_int32), ("off", ctypes.c_int32), ("sm", ctypes.c_int32), ("sk", ctypes.c_int32),
        ("conj_g", ctypes.c_int32), ("conj_pro", ctypes.c_int32), ("conj_epi", ctypes.c_int32), ("impl", ctypes.c_int32),
        ("B", ctypes.c_void_p),
        ("sb_b", ctypes.c_int64), ("sb_k", ctypes.c_int64), ("sb_n", ctypes.c_int64),
        ("pro", ctypes.c_void_p),
        ("C", ctypes.c_void_p),
        ("sc_b", ctypes.c_int64), ("sc_m", ctypes.c_int64), ("sc_n", ctypes.c_int64),
        ("epi", ctypes.c_void_p),
        ("scratch", ctypes.c_void_p),
    ]


_lib = None
_lib_lock = threading.Lock()


def _declare(l):
    vp, i32, u64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_uint64
    l.thz_version.restype = ctypes.c_int
    l.thz_last_error.restype = ctypes.c_char_p
    l.thz_fft_plan_info.argtypes = [i32, ctypes.POINTER(i32), ctypes.POINTER(i32)]
    l.thz_fft_slot_to_bin.argtypes = [i32, ctypes.POINTER(i32)]
    l.thz_fft_twiddles.argtypes = [i32, ctypes.POINTER(ctypes.c_float)]
    l.thz_fft_is_static.argtypes = [i32]
    l.thz_asm_workspace_bytes.argtypes = [ctypes.POINTER(AsmDesc)]
    l.thz_asm_workspace_bytes.restype = u64
    l.thz_asm_propagate.argtypes = [ctypes.POINTER(AsmDesc), vp]
    l.thz_fft2_c2c.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, u64, vp]
    f32 = ctypes.c_float
    l.thz_doe_modulate_fwd.argtypes = [vp, vp, vp, vp, f32, i32, i32, i32, i32, vp]
    l.thz_doe_modulate_bwd.argtypes = [vp, vp, vp, vp, f32, vp, vp, i32, i32, i32, i32, vp]
    l.thz_height_fwd.argtypes = [vp, f32, f32, vp, u64, vp]
    l.thz_height_bwd.argtypes = [vp, vp, f32, f32, vp, u64, vp]
    l.thz_quant_ste_fwd.argtypes = [vp, i32, f32, f32, vp, i32, vp, vp, vp, u64, vp]
    l.thz_quant_nn_fwd.argtypes = [vp, vp, i32, vp, i32, vp, vp, u64, vp]
    l.thz_quant_nn_bwd.argtypes = [vp, vp, vp, vp, i32, f32, i32, vp, u64, vp]
    l.thz_quant_psq_fwd.argtypes = [vp, f32, i32, f32, vp, vp, u64, vp]
    l.thz_quant_gumbel_v3_fwd.argtypes = [vp, vp, i32, vp, f32, f32, f32, f32, f32, f32, f32, f32, i32, vp, vp, vp, u64, vp]
    l.thz_quant_gumbel_naive_fwd.argtypes = [vp, vp, vp, i32, f32, vp, vp, vp, u64, vp]
    l.thz_toeplitz_gemm.argtypes = [ctypes.POINTER(ToeplitzGemmDesc), vp]
    l.thz_quant_softmax_fwd.argtypes = [vp, vp, i32, vp, f32, f32, f32, i32, vp, vp, vp, vp, vp, vp, u64, vp]
    l.thz_quant_softmax_bwd.argtypes = [vp, vp, vp, vp, vp, vp, u64, vp]
    l.thz_score_thickness.argtypes = [vp, vp, i32, f32, i32, vp, vp, i32, u64, vp]
    l.thz_tf_row_thresholds.argtypes = [i32, i32, i32, vp, vp, vp, vp]
    l.thz_field_mul.argtypes = [vp, vp, vp, i32, i32, u64, i32, i32, i32, vp]
    l.thz_normmse_loss.argtypes = [vp, vp, i32, u64, vp, vp, vp, vp]
    l.thz_tf_table_from_angles.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, i32, i32, vp, vp]
    l.thz_split_pre.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, i32, f32, vp]
    l.thz_split_post.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, i32, f32, vp]
    l.thz_normmse_loss_each.argtypes = [vp, vp, i32, i32, u64, vp, vp, vp]
    l.thz_adam_step.argtypes = [vp, vp, vp, vp, vp, u64, f32, f32, f32, f32, f32, i32, i32, vp]
    l.thz_launch_count.restype = u64
    l.thz_launch_count_class.restype = u64
    l.thz_launch_count_class.argtypes = [i32]
    l.thz_profile_enable.argtypes = [i32]
    l.thz_profile_read.argtypes = [i32, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(i32)]
    for name in EXPORTS:
        getattr(l, name)   # AttributeError here means the .so is stale: rebuild
    return l


# every symbol include/thzdoe.h declares (tests check the library exports all of them)
EXPORTS = [
    "thz_version", "thz_last_error", "thz_fft_plan_info", "thz_fft_slot_to_bin", "thz_fft_twiddles",
    "thz_asm_workspace_bytes", "thz_asm_propagate", "thz_fft2_c2c",
    "thz_doe_modulate_fwd", "thz_doe_modulate_bwd", "thz_height_fwd", "thz_height_bwd",
    "thz_quant_ste_fwd", "thz_quant_nn_fwd", "thz_quant_nn_bwd", "thz_quant_psq_fwd",
    "thz_quant_gumbel_v3_fwd", "thz_quant_gumbel_naive_fwd",
    "thz_launch_count", "thz_launch_count_class", "thz_profile_enable", "thz_profile_read", "thz_toeplitz_gemm", "thz_tf_row_thresholds",
    "thz_normmse_loss", "thz_adam_step", "thz_fft_is_static", "thz_field_mul",
    "thz_quant_softmax_fwd", "thz_quant_softmax_bwd", "thz_score_thickness", "thz_normmse_loss_each", "thz_tf_table_from_angles",
    "thz_split_pre", "thz_split_post",
]


def load_library(path=LIB_PATH):
    """dlopen the library and declare prototypes (no CUDA call is made)."""
    if not os.path.isfile(path):
        raise ThzError(
            "native library %s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or python -m quantizationawarethzdoe_b200.build). There is no CPU fallback." % path)
    return _declare(ctypes.CDLL(path))


def lib():
    """The process-wide library handle; raises if it cannot be loaded."""
    global _lib
    if _lib is None:
        with _lib_lock:
            if _lib is None:
                _lib = load_library()
    return _lib


def check(rc, what="thzdoe"):
    if rc == THZ_OK:
        return
    msg = lib().thz_last_error().decode("utf-8", "replace")
    raise _ERR_EXC.get(rc, ThzError)("%s failed (%d): %s" % (what, rc, msg))


def require_cuda(t, name="tensor"):
    if not t.is_cuda:
        raise ThzError("%s must live on a CUDA device (got %s); this package has no CPU path" % (name, t.device))


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def current_stream_ptr(device):
    """torch's current CUDA stream on `device` as a cudaStream_t (the raw-handle query: this runs before every launch)."""
    if _raw_stream is not None and device.index is not None:
        return ctypes.c_void_p(_raw_stream(device.index))
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


# ----------------------------------------------------------------------------- host-side tables
_tw_host = {}
_tw_dev = {}
_perm_host = {}


def twiddles_host(n):
    """tw[m] = exp(-2 pi i m / n), float64 math rounded to complex64 (host tensor, cached)."""
    t = _tw_host.get(n)
    if t is None:
        m = np.arange(n, dtype=np.float64)
        w = np.exp(-2j * np.pi * m / n).astype(np.complex64)
        t = torch.from_numpy(w)
        _tw_host[n] = t
    return t


def twiddles(n, device):
    key = (n, str(device))
    t = _tw_dev.get(key)
    if t is None:
        t = twiddles_host(n).to(device)
        _tw_dev[key] = t
    return t


def slot_to_bin(n, fn=None):
    """Digit-reversal permutation of the length-n plan (host int64 tensor, cached)."""
    p = _perm_host.get(n)
    if p is None:
        buf = (ctypes.c_int32 * n)()
        f = fn if fn is not None else lib().thz_fft_slot_to_bin
        rc = f(n, buf)
        if rc != 0:
            raise NotImplementedError("FFT length %d has a prime factor > 7, unsupported by the sm_100a plans" % n)
        p = torch.tensor(list(buf), dtype=torch.int64)
        _perm_host[n] = p
    return p


def plan_radices(n):
    """Radices of the length-n plan, first stage first (thz_fft_plan_info)."""
    rad, ns = (ctypes.c_int32 * 16)(), ctypes.c_int32(0)
    check(lib().thz_fft_plan_info(int(n), rad, ctypes.byref(ns)), "thz_fft_plan_info")
    return [int(rad[i]) for i in range(ns.value)]


def ptr(t):
    """Device address of a tensor (None -> NULL); an int is taken as a raw address (multicast mappings have no tensor)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(int(t)) if isinstance(t, int) else ctypes.c_void_p(t.data_ptr())
