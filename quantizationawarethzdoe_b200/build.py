"""Build libthzdoe.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m quantizationawarethzdoe_b200.build [--force]
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libthzdoe.so")
STAMP = os.path.join(CSRC, ".build_stamp")
SOURCES = ["thz_api.cu", "thz_asm.cu", "thz_asm_p2_k1.cu", "thz_asm_p2_k2.cu", "thz_asm_p2_k3.cu", "thz_doe.cu", "thz_split.cu", "thz_czt.cu", "thz_czt_tc.cu", "thz_train.cu"]
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "--threads", "4",
]


def _digest():
    h = hashlib.sha256()
    names = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".inc")))
    names.append(os.path.join("..", "..", "include", "thzdoe.h"))
    for n in names:
        with open(os.path.join(CSRC, n), "rb") as f:
            h.update(n.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS + _extra_flags()).encode())
    return h.hexdigest()


def _extra_flags():
    """Experiment switches (e.g. THZ_NVCC_EXTRA="-DTHZ_NO_F32X2"); empty in normal builds."""
    return os.environ.get("THZ_NVCC_EXTRA", "").split()


def build(force=False, verbose=False):
    """Compile every .cu into one shared library; skipped when sources are unchanged."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.isfile(os.path.join(CSRC, s))]
    dig = _digest()
    if not force and os.path.isfile(OUT) and os.path.isfile(STAMP) and open(STAMP).read().strip() == dig:
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for s in srcs:
        o = s[:-3] + ".o"
        objs.append(o)
        cmd = [nvcc] + NVCC_FLAGS + _extra_flags() + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s" % " ".join(cmd))
    cmd = [nvcc, "-shared", "-o", OUT] + objs + ["-lcudart"]
    subprocess.check_call(cmd)
    with open(STAMP, "w") as f:
        f.write(dig)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
