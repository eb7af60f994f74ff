"""Shared test helpers: golden fixtures, error metric, CPU replay library."""
import ctypes
import os
import subprocess

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
EMUL_SRC = os.path.join(ROOT, "tests", "emul", "emul.cpp")
EMUL_SO = os.path.join(ROOT, "tests", "emul", "libthz_emul.so")


def rel_l2(a, b):
    a = torch.as_tensor(a)
    b = torch.as_tensor(b)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def record(test, **vals):
    """Append measured errors to gpurun_out/parity_errors.jsonl (copied to profiles/ per round) and echo them: the
    tolerances in the tests are bars, these are the distances actually measured on the GPU."""
    import json
    line = json.dumps(dict(test=test, **{k: (float(v) if isinstance(v, (int, float)) else v) for k, v in vals.items()}))
    print("PARITY", line)
    try:
        d = os.path.join(ROOT, "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_errors.jsonl"), "a") as f:
            f.write(line + "\n")
    except OSError:
        pass


def golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    out = {}
    for k in z.files:
        v = z[k]
        if v.dtype.kind in "US":
            out[k] = str(v)
        elif v.ndim == 0:
            out[k] = v.item()
        else:
            out[k] = torch.from_numpy(v)
    return out


def golden_names(prefix):
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith(prefix) and f.endswith(".npz"))


def asm_case_kwargs(g):
    """Constructor kwargs for ASM_prop / oracle from an asm_* fixture."""
    ps = g["padding_scale"]
    ps = None if ps.numel() == 1 and float(ps[0]) < 0 else [float(ps[0]), float(ps[1])]
    return dict(padding_scale=ps, bandlimit_type=g["bandlimit_type"], do_padding=bool(g["do_padding"]),
                do_unpad_after_pad=bool(g["do_unpad"]))


def emul_lib():
    """Build (if stale) and load the CPU replay of the CUDA kernel bodies (tests/emul/emul.cpp)."""
    from quantizationawarethzdoe_b200 import _native as N
    csrc = os.path.join(ROOT, "quantizationawarethzdoe_b200", "csrc")
    deps = [EMUL_SRC, os.path.join(ROOT, "include", "thzdoe.h")] + [
        os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
    # THZ_EMUL_ASAN=1: AddressSanitizer build of the replay (shared memory and workspaces are heap buffers there, so every
    # out-of-bounds access of a kernel body is caught).  Needs LD_PRELOAD=$(gcc -print-file-name=libasan.so) and
    # ASAN_OPTIONS=detect_leaks=0 on the pytest command line; compute-sanitizer is not available on the GPU pool.
    asan = os.environ.get("THZ_EMUL_ASAN") == "1"
    so = EMUL_SO.replace(".so", "_asan.so") if asan else EMUL_SO
    stale = not os.path.isfile(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps)
    if stale:
        extra = ["-fsanitize=address", "-fno-omit-frame-pointer", "-O1", "-g"] if asan else ["-O2"]
        subprocess.check_call(["g++", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared"] + extra +
                              ["-I/usr/local/cuda/include", "-o", so, EMUL_SRC])
    E = ctypes.CDLL(so)
    E.thz_emul_asm_propagate.argtypes = [ctypes.POINTER(N.AsmDesc), ctypes.c_int]
    E.thz_emul_slot_to_bin.argtypes = [ctypes.c_int32, ctypes.POINTER(ctypes.c_int32)]
    E.thz_emul_plan_info.argtypes = [ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]
    vp, i32 = ctypes.c_void_p, ctypes.c_int32
    E.thz_emul_fft2_c2c.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp]
    f32, u64 = ctypes.c_float, ctypes.c_uint64
    E.thz_emul_softmaxq.argtypes = [vp, vp, i32, vp, f32, f32, f32, i32, vp, vp, vp, vp, vp, vp, u64]
    E.thz_emul_softmaxq_bwd.argtypes = [vp, vp, vp, vp, vp, vp, u64]
    E.thz_emul_score_thickness.argtypes = [vp, vp, i32, f32, i32, vp, i32, u64]
    E.thz_emul_split_pre.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, i32, f32]
    E.thz_emul_split_post.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, i32, f32]
    return E
