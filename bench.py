#!/usr/bin/env python3
"""Headline benchmark: ASM + quantized-DOE forward+backward Msamples/s at a 4096^2 padded grid x lambda batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One *step* = one pass of the hot path over one batch of synthetic fields:
  4-level STE level selection -> fused DOE phase modulation + band-limited ASM forward
  -> loss gradient (L = 1/2 |y|^2, i.e. g = y) -> ASM adjoint -> grad wrt field, grad wrt height map
  -> sigmoid/STE chain to the DOE weights [-> NCCL all-reduce of the weight gradient when N > 1].
Workload (BASELINE.json metric shape): x = (1, 16, 2048, 2048) complex64 per GPU, 2x pad -> 4096^2,
lambda_c = 1 mm (1 + 0.01 c), dx = 0.5 mm, z = 100 mm, eps = 2.66, tan d = 0.003, hmax = 1 mm.
Sample = one padded grid point of one (batch, lambda) field: 16 * 4096^2 samples per step per GPU.

`value`    device-resident throughput (inputs already in HBM), CUDA events, max over ranks.
`e2e`      the same step driven from HOST buffers through the module API: pinned H2D copy of the
           fields and D2H read of the weight gradient inside the timed region.
`roofline` algorithmic bytes (42 B / sample) / step time vs the measured HBM copy peak,
           plus a per-kernel breakdown from CUDA events recorded on the launching stream.
`cpu_baseline` the oracle port of the reference's CPU path on a bounded sample (one field), same box.
--impl reference: that CPU path alone, as the reference arm.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

mm = 1e-3
N_FIELD = int(os.environ.get("THZ_BENCH_N", "2048"))        # unpadded edge; padded = 2N = 4096
C_LAMBDA = int(os.environ.get("THZ_BENCH_C", "16"))
LEVELS = 4
Z = 0.1
SPACING = 0.5 * mm
MATERIAL = [2.66, 0.003]
HMAX = 1 * mm
BYTES_PER_SAMPLE = 42.0   # fwd 20 + bwd 22 bytes per padded sample at 2x pad
METRIC = "ASM+DOE fwd+bwd Msamples/s at 4096^2 pad x lambda batch"


def wavelengths(C):
    return [1 * mm * (1 + 0.01 * c) for c in range(C)]


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, "/tmp/thz_clocks_%d.csv" % os.getpid()

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, smax, reasons = [], [], set()
        for line in open(self.path):
            p = [t.strip() for t in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                smax.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        hi = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(hi) if hi else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_step(x, w, lams):
    """The reference's CPU path for this step (oracle port: same ops as the reference modules, torch CPU)."""
    from oracle import asm_oracle as AO, doe_oracle as DO
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    h = DO.ste_quantize(DO.sigmoid_height(wr[0, 0], HMAX), DO.linear_lut(HMAX, LEVELS))
    y = AO.asm_forward(DO.modulate(xr, h, lams, MATERIAL[0], MATERIAL[1]), lams, SPACING, Z)
    gx, gw = torch.autograd.grad(y, (xr, wr), y.detach())
    return gx, gw


def time_cpu(n_field, c_sample, steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    lams = wavelengths(c_sample)
    x = torch.randn(1, c_sample, n_field, n_field, dtype=torch.complex64)
    torch.manual_seed(1)
    w = torch.randn(1, 1, n_field, n_field)
    for _ in range(warmup):
        cpu_reference_step(x, w, lams)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_reference_step(x, w, lams)
        ts.append(time.perf_counter() - t0)
    t = sum(ts) / len(ts)
    samples = c_sample * (2 * n_field) ** 2
    return samples / t / 1e6, t


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, min(args.warmup, 2))
    steps = min(steps, 5)
    val, t = time_cpu(N_FIELD, 1, steps, warmup)
    cores = os.cpu_count() or 1
    sample = "1 of %d wavelength fields of the workload (%d^2 -> %d^2 pad), oracle port of the reference CPU path, %d torch threads" % (
        C_LAMBDA, N_FIELD, 2 * N_FIELD, cores)
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "STE 4-level DOE + band-limited ASM fwd+bwd, %d^2 field -> %d^2 pad, bounded sample of 1 field per step" % (N_FIELD, 2 * N_FIELD)},
        "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, local_rank, world):
    import torch.distributed as dist
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField, STEQuantizedDOELayer, _native as N, functional as Fn
    import ctypes

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = N.lib()
    for k, env in (("bc_chunk", "THZ_BC_CHUNK"), ("k2_cols", "THZ_K2_COLS"), ("lines", "THZ_LINES")):
        if os.environ.get(env):
            Fn.TUNE[k] = int(os.environ[env])
    lams = wavelengths(C_LAMBDA)
    B, C, n = 1, C_LAMBDA, N_FIELD
    Np = 2 * n
    torch.manual_seed(rank)
    x_host = torch.randn(B, C, n, n, dtype=torch.complex64).pin_memory()
    torch.manual_seed(1)          # DOE weights are replicated across ranks (data-parallel)
    doe = STEQuantizedDOELayer(dict(doe_size=[n, n], doe_dxy=SPACING, doe_level=LEVELS, height_constraint_max=HMAX,
                                    tolerance=None, material=MATERIAL), {}, device=dev)
    asm = ASM_prop(z_distance=Z, device=dev, kernel_mode=os.environ.get("THZ_KERNEL_MODE", "inregister"))
    asm.check_Zc = False
    x_dev = x_host.to(dev).requires_grad_(True)
    lam_t = torch.tensor(lams, dtype=torch.float32, device=dev)        # built once: the modules key their plans on these objects
    sp_t = torch.tensor([SPACING, SPACING], dtype=torch.float32, device=dev)
    gw_host = torch.empty(1, 1, n, n, dtype=torch.float32).pin_memory()

    def step(x):
        """One hot-path pass; returns the gradient wrt the DOE weights (all-reduced when world > 1)."""
        field = ElectricField(x, wavelengths=lam_t, spacing=sp_t, device=dev)
        y = asm(doe(field)).data
        gx, gw = torch.autograd.grad(y, (x, doe.weight_height_map), y.detach())
        if world > 1:
            dist.all_reduce(gw)
        return gw

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    warmup = max(3, args.warmup)
    for _ in range(warmup):
        step(x_dev)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = lib.thz_launch_count()
    ms_total = timed(lambda: step(x_dev), args.steps)
    launches = (lib.thz_launch_count() - l0) // max(1, args.steps)
    clocks = sampler.stop() if sampler else None
    ms_step = ms_total / args.steps
    samples_per_step = world * B * C * Np * Np
    value = samples_per_step / (ms_step * 1e-3) / 1e6

    # ---- e2e: host buffers in, host gradient out, through the same module API.  Every step copies ITS input
    # fields from pinned host memory and reads its weight gradient back; the copy of step i+1 runs on a second
    # stream while step i computes (two device buffers), which is how a user would feed a stream of fields.
    ncopy = int(os.environ.get("THZ_E2E_COPY_STREAMS", "1"))       # the H2D copy may be split over several streams
    copy_streams = [torch.cuda.Stream(device=dev) for _ in range(ncopy)]
    bufs = [torch.empty_like(x_dev).detach() for _ in range(2)]
    ready = [[torch.cuda.Event() for _ in range(ncopy)] for _ in range(2)]      # H2D of the buffer finished
    freed = [torch.cuda.Event() for _ in range(2)]      # compute that read the buffer finished
    parts = [slice(k * C // ncopy, (k + 1) * C // ncopy) for k in range(ncopy)]

    def h2d(b, wait):
        for k, cs in enumerate(copy_streams):
            with torch.cuda.stream(cs):
                if wait:
                    cs.wait_event(freed[b])
                bufs[b][:, parts[k]].copy_(x_host[:, parts[k]], non_blocking=True)
                ready[b][k].record(cs)

    def e2e_run(nsteps):
        cur = torch.cuda.current_stream(dev)
        h2d(0, False)
        for i in range(nsteps):
            b = i & 1
            if i + 1 < nsteps:
                h2d(b ^ 1, i >= 1)
            for ev in ready[b]:
                cur.wait_event(ev)
            gw = step(bufs[b].detach().requires_grad_(True))     # fresh leaf over the same storage
            freed[b].record(cur)
            gw_host.copy_(gw, non_blocking=True)
        cur.synchronize()

    e2e_run(2)
    e2e_steps = max(2, min(args.steps, 10))
    ms_e2e = timed(lambda: e2e_run(e2e_steps), 1) / e2e_steps
    e2e_val = samples_per_step / (ms_e2e * 1e-3) / 1e6

    # ---- per-kernel attribution with CUDA events on the launching stream (separate pass, slight overhead)
    kernels = None
    psteps = min(args.steps, 5)
    if rank == 0:
        lib.thz_profile_enable(1)
    for _ in range(psteps):          # every rank runs these steps (they contain the all-reduce); rank 0 records events
        step(x_dev)
    barrier()
    if rank == 0:
        ms_sum = (ctypes.c_float * 8)()
        cnt = (ctypes.c_int32 * 8)()
        lib.thz_profile_read(8, ms_sum, cnt)
        lib.thz_profile_enable(0)
        names = ["row_fft_fwd", "column_fft_H_ifft", "row_ifft_epilogue", "fft2_col", "doe_modulate", "quantizer", "czt"]
        fields_per_step = B * C
        # algorithmic bytes per field per launch class (fwd + bwd launches pooled), complex64, 2x pad
        alg = {"row_fft_fwd": 8 * (n * n + n * Np), "column_fft_H_ifft": 16 * n * Np, "row_ifft_epilogue": 8 * (n * Np + 1.5 * n * n)}
        kernels = {}
        for i, nm in enumerate(names):
            if cnt[i] == 0:
                continue
            per_step_ms = ms_sum[i] / psteps
            ent = {"ms_per_step": per_step_ms, "launches_per_step": cnt[i] / psteps}
            if nm in alg:
                gbs = alg[nm] * fields_per_step * 2 / (per_step_ms * 1e-3) / 1e9     # x2: forward and adjoint pass
                ent["achieved_gbs"] = gbs
            kernels[nm] = ent

    if rank != 0:
        return
    peak, peak_src = measured_peaks()
    step_achieved = BYTES_PER_SAMPLE * (B * C * Np * Np) / (ms_step * 1e-3) / 1e9      # per GPU, whole step
    dom = max(kernels.items(), key=lambda kv: kv[1]["ms_per_step"])[0] if kernels else None
    for ent in (kernels or {}).values():
        if "achieved_gbs" in ent:
            ent["frac"] = ent["achieved_gbs"] / peak
    # DRAM traffic of the dominant kernel per launch, from the committed ncu --set full capture (scaled to this launch size)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_k2_dram_traffic.json")) as f:
            t = json.load(f)
        traffic = (t["dram_bytes_read"] + t["dram_bytes_write"]) * (B * C) / t["fields_per_launch"]
    except Exception:
        pass
    dk = (kernels or {}).get(dom, {})
    dom_alg_bytes = 16.0 * n * Np * B * C                  # column pass: read + write of the live rows, per launch
    cpu_val, cpu_t = (None, None)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_val, cpu_t = time_cpu(N_FIELD, 1, 3, 1)
        cores = os.cpu_count() or 1
        cpu = {"value": cpu_val, "unit": "Msamples/s", "cores": cores, "kind": "port",
               "sample": "1 of %d wavelength fields (%d^2 -> %d^2 pad), 1 warm-up + mean of 3, %d torch threads, %.2f s/step" % (
                   C, n, Np, cores, cpu_t)}
    out = {
        "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "metric shape: STE %d-level DOE + band-limited ASM fwd+bwd, x=(1,%d,%d,%d) c64 per GPU, 2x pad -> %d^2, "
                               "z=100 mm, dx=0.5 mm, lambda=1 mm(1+0.01c)" % (LEVELS, C, n, n, Np),
                   "fields_per_gpu": B * C, "samples_per_step": samples_per_step, "kernel_mode": asm.kernel_mode,
                   "l2": "inputs larger than L2 (512 MiB fields + 2 x 1 GiB intermediate spectra per step)", "tune": dict(Fn.TUNE),
                   "parallelism": "dp%d over wavelengths, NCCL all-reduce of grad(weights)" % world if world > 1 else "single GPU"},
        "e2e": {"value": e2e_val, "unit": "Msamples/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": x_host.numel() * 8,
                "d2h_bytes_per_step": gw_host.numel() * 4},
        "gpu_launches": int(launches),
        "clocks": clocks,
        # dominant kernel (the column pass K2): algorithmic bytes per launch / its mean launch duration (CUDA events on the
        # launching stream); `step` = the whole hot path at SURVEY 8d's 42 B per padded sample.
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": dk.get("achieved_gbs"), "peak": peak, "unit": "GB/s",
                     "frac": dk.get("frac"), "traffic": traffic, "algorithmic_bytes_per_launch": dom_alg_bytes,
                     "launch_ms": (dk.get("ms_per_step") / dk.get("launches_per_step")) if dk else None,
                     "peak_source": peak_src,
                     "step": {"achieved": step_achieved, "frac": step_achieved / peak, "bytes_per_sample": BYTES_PER_SAMPLE,
                              "scope": "whole step: 6 FFT-pipeline launches + 2 level-selection launches, per GPU"},
                     "kernels": kernels},
        "fields_per_s": value * 1e6 / (Np * Np),
    }
    if cpu:
        out["cpu_baseline"] = cpu
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, local_rank, world)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
