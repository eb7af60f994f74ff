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
# FP32 lane-operations (FMA = 1) per padded sample of one fwd + bwd step at Np = 4096 = 16^3:
#   a radix-16 butterfly is 144 adds + 24 multiplies, its 15 twiddle multiplies 60, the twiddle power tree of a first /
#   last stage 56 -> per point 14.25 (twiddled stage) / 10.5 (last stage) + 3.5 (power tree, stages 0 of each transform);
#   per direction the row passes transform N Np points each and the column pass 2 Np^2: 3 Np^2 point-passes x 3 stages;
#   H generation ~25 per column-pass point, DOE phase ~30 per live input point.
FP32_LANE_OPS_PER_POINT_PASS = (14.25 + 14.25 + 10.5) + 3.5
FP32_LANE_OPS_K2_PER_POINT = 2 * FP32_LANE_OPS_PER_POINT_PASS + 25.0          # column FFT + H + column iFFT, per padded point
FP32_LANE_OPS_PER_SAMPLE = 2 * (3 * FP32_LANE_OPS_PER_POINT_PASS + 25.0) + 2 * 0.25 * 30.0
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


# ------------------------------------------------------------------------------------------- secondary configs
def _bf16_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["bf16_tflops"]), "measured bf16 burst (MEASURED_PEAKS.json)"
    except Exception:
        return 1674.1, "fallback (B200_PROFILING.md)"


def _event_ms(fn, reps, warm, dev):
    st = torch.cuda.current_stream(dev)
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        fn()
    e1.record(st)
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / reps


FP32_LANES_PER_SM = 128


def _fp32_peak(dev, sm_mhz):
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    return sms * FP32_LANES_PER_SM * 2 * (sm_mhz or 1965.0) * 1e6 / 1e12      # TFLOP/s, FMA = 2


def secondary_single(dev, lib, sm_mhz, hbm_peak):
    """The other BASELINE.json configurations that fit one GPU (SURVEY 8d shapes), device-event timings, warm:
    C2 (1000 -> 2000 pad, 8-level STE DOE, fwd + adjoint; eager and as a replayed CUDA graph), C3 (CZT 2048^2 -> 1024^2,
    16 wavelengths, forward: ms, 3xTF32 tensor TFLOP/s, fraction of 1/2 x the measured bf16 peak, which kernel ran),
    C4 (3-layer DONN, 200 -> 400 pad, batch 1024, fwd + bwd: samples/s vs the compulsory-I/O HBM ceiling and the FP32 ceiling)."""
    from quantizationawarethzdoe_b200 import ASM_prop, CZT_prop, ElectricField, STEQuantizedDOELayer, functional as Fn
    out = {}
    lam1, sp = torch.tensor([1 * mm], device=dev), torch.tensor([SPACING, SPACING], device=dev)
    # ---- C2
    n = 1000
    torch.manual_seed(0)
    doe = STEQuantizedDOELayer(dict(doe_size=[n, n], doe_dxy=SPACING, doe_level=8, height_constraint_max=HMAX, tolerance=None,
                                    material=MATERIAL), {}, device=dev)
    asm = ASM_prop(z_distance=Z, device=dev)
    asm.check_Zc = False
    x = torch.randn(1, 1, n, n, dtype=torch.complex64, device=dev)

    def c2_step():
        y = asm(doe(ElectricField(x, wavelengths=lam1, spacing=sp, device=dev))).data
        torch.autograd.grad(y, doe.weight_height_map, y.detach())

    side = torch.cuda.Stream(device=dev)          # AccumulateGrad nodes bind to the stream of the first forward: not the legacy one
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        ms = _event_ms(c2_step, 30, 5, dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            c2_step()
        ms_graph = _event_ms(graph.replay, 100, 5, dev)
    torch.cuda.current_stream(dev).wait_stream(side)
    smp = (2 * n) ** 2
    out["c2_step_1000_to_2000_8level"] = {
        "eager_ms": ms, "cuda_graph_ms": ms_graph, "Msamples_per_s_graph": smp / ms_graph / 1e3,
        "hbm_frac_graph": BYTES_PER_SAMPLE * smp / (ms_graph * 1e-3) / 1e9 / hbm_peak, "kernel_mode": asm.resolved_kernel_mode}
    del graph
    # ---- C3
    H, M, C = 2048, 1024, 16
    xc = torch.randn(1, C, H, H, dtype=torch.complex64, device=dev)
    lams = torch.tensor(wavelengths(C), dtype=torch.float32, device=dev)
    prop = CZT_prop(z_distance=0.5, device=dev)
    fc = ElectricField(xc, wavelengths=lams, spacing=sp, device=dev)
    tc0, simt0 = lib.thz_launch_count_class(8), lib.thz_launch_count_class(6)
    ms = _event_ms(lambda: prop(fc, M, M, 0.1 * mm, 0.1 * mm), 5, 2, dev)
    tc_l, simt_l = lib.thz_launch_count_class(8) - tc0, lib.thz_launch_count_class(6) - simt0
    flops = 8.0 * (M * H * H + M * H * M) * C
    bf16, bf16_src = _bf16_peak()
    tflops = 3 * flops / (ms * 1e-3) / 1e12
    out["c3_czt_2048_to_1024_16lambda"] = {
        "forward_ms": ms, "tensor_tflops_3xtf32": tflops, "complex_gemm_tflops": flops / (ms * 1e-3) / 1e12,
        "peak_tf32_tflops": 0.5 * bf16, "peak_source": "1/2 x " + bf16_src, "frac": tflops / (0.5 * bf16),
        "kernel": "tcgen05 3xTF32 (thz_k_toeplitz_gemm_tc)" if tc_l and not simt_l else "CUDA-core fp32 (thz_k_toeplitz_gemm)",
        "tcgen05_launches": int(tc_l), "cuda_core_gemm_launches": int(simt_l)}
    del xc, fc, prop
    # ---- C5 on ONE GPU: the reference point of the slab-decomposed runs at N > 1 (same grid, forward + adjoint)
    n5 = int(os.environ.get("THZ_BENCH_SLAB_N", "8192"))
    x5 = torch.randn(1, 1, n5, n5, dtype=torch.complex64, device=dev).requires_grad_(True)
    asm5 = ASM_prop(z_distance=Z, device=dev, kernel_mode="inregister")
    asm5.check_Zc = False
    f5 = ElectricField(x5, wavelengths=lam1, spacing=sp, device=dev)

    def c5_step():
        y = asm5(f5).data
        torch.autograd.grad(y, x5, y.detach())

    ms = _event_ms(c5_step, 5, 2, dev)
    smp = (2 * n5) ** 2
    out["c5_single_gpu_%d_padded" % (2 * n5)] = {"fwd_bwd_ms": ms, "Msamples_per_s": smp / ms / 1e3,
                                                 "hbm_frac": 40.0 * smp / (ms * 1e-3) / 1e9 / hbm_peak}
    del x5, f5, asm5
    torch.cuda.empty_cache()
    # ---- beyond the shared-memory line: one 16384^2 field on a 32768^2 canvas (longline.py: 2 x 2 split around the fused pipeline)
    if os.environ.get("THZ_BENCH_LONG", "1") == "1" and torch.cuda.mem_get_info(dev)[0] > (64 << 30):
        n6 = 16384
        x6 = torch.randn(1, 1, n6, n6, dtype=torch.complex64, device=dev).requires_grad_(True)
        asm6 = ASM_prop(z_distance=Z, device=dev)
        asm6.check_Zc = False
        f6 = ElectricField(x6, wavelengths=lam1, spacing=sp, device=dev)
        g6 = torch.randn(1, 1, n6, n6, dtype=torch.complex64, device=dev)

        def c6_step():
            y = asm6(f6).data
            torch.autograd.grad(y, x6, g6)

        ms = _event_ms(c6_step, 3, 1, dev)
        smp = (2 * n6) ** 2
        out["long_canvas_%d_padded" % (2 * n6)] = {"fwd_bwd_ms": ms, "Msamples_per_s": smp / ms / 1e3,
                                                   "kernel_mode": asm6.resolved_kernel_mode,
                                                   "peak_gpu_gib": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 1)}
        del x6, f6, g6, asm6
        torch.cuda.empty_cache()
    # ---- C4
    n, B, layers = 200, 1024, 3
    does = [STEQuantizedDOELayer(dict(doe_size=[n, n], doe_dxy=SPACING, doe_level=4, height_constraint_max=HMAX, tolerance=None,
                                      material=MATERIAL), {}, device=dev) for _ in range(layers)]
    asms = [ASM_prop(z_distance=0.05, device=dev) for _ in range(layers)]
    for a in asms:
        a.check_Zc = False
    xd = torch.randn(B, 1, n, n, dtype=torch.complex64, device=dev)

    def donn_step():
        f = ElectricField(xd, wavelengths=lam1, spacing=sp, device=dev)
        for d, a in zip(does, asms):
            f = a(d(f))
        y = f.data
        torch.autograd.grad(y, [d.weight_height_map for d in does], y.detach())

    ms = _event_ms(donn_step, 10, 3, dev)
    out["c4_donn_3layer_200_batch1024"] = _donn_report(ms, B, n, layers, hbm_peak, _fp32_peak(dev, sm_mhz))
    return out


def _donn_report(ms, B, n, layers, hbm_peak, fp32_peak_tf):
    Np = 2 * n
    import math
    io_bytes = 40.0 * n * n * layers                                   # SURVEY 8d small-grid variant: compulsory I/O per sample
    fft_flop = 5.0 * Np * math.log2(Np)                                # one length-Np complex FFT
    flop = layers * 2 * (2 * n + 2 * Np) * fft_flop                    # per sample: (n rows + Np cols x 2 + n rows) per pass, fwd + bwd
    sps = B / (ms * 1e-3)
    return {"ms_per_step": ms, "samples_per_s": sps, "hbm_ceiling_samples_per_s": hbm_peak * 1e9 / io_bytes,
            "frac_hbm_ceiling": sps * io_bytes / (hbm_peak * 1e9), "fft_gflop_per_sample": flop / 1e9,
            "achieved_fp32_tflops": sps * flop / 1e12, "fp32_peak_tflops": fp32_peak_tf,
            "fp32_ceiling_samples_per_s": fp32_peak_tf * 1e12 / flop, "frac_fp32_ceiling": sps * flop / 1e12 / fp32_peak_tf}


def secondary_multi(dev, rank, world, lib, sm_mhz, hbm_peak):
    """N > 1 (one rank per GPU): (a) ONE 16384^2 padded grid (config 5) through the slab-decomposed FFT, forward + adjoint,
    strong scaling -- total ms (max over ranks), per-stage ms of rank 0, which transport ran; (b) the 3-layer DONN (config 4)
    data-parallel over the batch with one NCCL all-reduce of the three weight gradients per step (weak scaling)."""
    import torch.distributed as dist
    from quantizationawarethzdoe_b200 import ASM_prop, ElectricField, STEQuantizedDOELayer, parallel as P
    out = {}
    lam1, sp = torch.tensor([1 * mm], device=dev), torch.tensor([SPACING, SPACING], device=dev)

    def max_ms(v):
        t = torch.tensor([v], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---- C5 slab
    n = int(os.environ.get("THZ_BENCH_SLAB_N", "8192"))
    lo, hi = P.shard_range(n, rank, world)
    torch.manual_seed(100 + rank)
    xl = torch.randn(1, 1, hi - lo, n, dtype=torch.complex64, device=dev).requires_grad_(True)
    slab = P.SlabAsm(z_distance=Z, transport=os.environ.get("THZ_SLAB_TRANSPORT", "auto"))
    fl = ElectricField(xl, wavelengths=lam1, spacing=sp, device=dev)

    def slab_fwd_bwd():
        y = slab(fl).data
        torch.autograd.grad(y, xl, y.detach())

    for _ in range(3):
        slab_fwd_bwd()
    torch.cuda.synchronize(dev)
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        slab_fwd_bwd()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = max_ms(e0.elapsed_time(e1) / reps)
    P.SLAB_TIMINGS = []
    slab_fwd_bwd()
    torch.cuda.synchronize(dev)
    marks, P.SLAB_TIMINGS = P.SLAB_TIMINGS, None
    stages = []
    for i in range(1, len(marks)):
        if marks[i][0] != "start":
            stages.append([marks[i][0], marks[i - 1][1].elapsed_time(marks[i][1])])
    smp = (2 * n) ** 2
    out["c5_slab_%d_padded" % (2 * n)] = {
        "fwd_bwd_ms": ms, "scaling": "strong", "n_gpus": world, "transport": "peer memory (kernels scatter / gather over NVLink)"
        if slab._slabs is not None else "nccl all_to_all", "Msamples_per_s": smp / ms / 1e3,
        "hbm_frac_aggregate": 40.0 * smp / (ms * 1e-3) / 1e9 / (hbm_peak * world),      # 20 B fwd + 20 B adjoint per sample, no DOE
        "rank0_stage_ms": stages}
    del slab, xl, fl
    torch.cuda.empty_cache()
    # ---- C3 sharded over wavelengths: the 16 lambda of config 3 dealt over the ranks (strong scaling, no collective forward)
    from quantizationawarethzdoe_b200 import CZT_prop
    H, M, C = 2048, 1024, 16
    if C % world == 0:
        clo, chi = P.shard_range(C, rank, world)
        lam_all = wavelengths(C)
        torch.manual_seed(300)
        xc = torch.randn(1, chi - clo, H, H, dtype=torch.complex64, device=dev)
        prop = CZT_prop(z_distance=0.5, device=dev)
        fc = ElectricField(xc, wavelengths=torch.tensor(lam_all[clo:chi], dtype=torch.float32, device=dev), spacing=sp, device=dev)
        tc0, simt0 = lib.thz_launch_count_class(8), lib.thz_launch_count_class(6)
        for _ in range(2):
            prop(fc, M, M, 0.1 * mm, 0.1 * mm)
        torch.cuda.synchronize(dev)
        dist.barrier()
        e0.record()
        for _ in range(5):
            prop(fc, M, M, 0.1 * mm, 0.1 * mm)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = max_ms(e0.elapsed_time(e1) / 5)
        flops = 8.0 * (M * H * H + M * H * M) * C
        bf16, bf16_src = _bf16_peak()
        tfl = 3 * flops / (ms * 1e-3) / 1e12
        out["c3_czt_16lambda_sharded"] = {
            "forward_ms": ms, "scaling": "strong", "n_gpus": world, "wavelengths_per_gpu": chi - clo, "tensor_tflops_3xtf32_aggregate": tfl,
            "frac_of_aggregate_tf32_peak": tfl / (0.5 * bf16 * world), "peak_source": "1/2 x " + bf16_src,
            "kernel": "tcgen05 3xTF32" if lib.thz_launch_count_class(8) > tc0 and lib.thz_launch_count_class(6) == simt0 else "CUDA-core fp32"}
        del xc, fc, prop
        torch.cuda.empty_cache()
    # ---- C4 data parallel
    n, B, layers = 200, 1024, 3
    torch.manual_seed(0)
    does = [STEQuantizedDOELayer(dict(doe_size=[n, n], doe_dxy=SPACING, doe_level=4, height_constraint_max=HMAX, tolerance=None,
                                      material=MATERIAL), {}, device=dev) for _ in range(layers)]
    asms = [ASM_prop(z_distance=0.05, device=dev) for _ in range(layers)]
    for a in asms:
        a.check_Zc = False
    torch.manual_seed(200 + rank)
    xd = torch.randn(B, 1, n, n, dtype=torch.complex64, device=dev)
    params = [d.weight_height_map for d in does]

    def donn_step():
        f = ElectricField(xd, wavelengths=lam1, spacing=sp, device=dev)
        for d, a in zip(does, asms):
            f = a(d(f))
        y = f.data
        gws = torch.autograd.grad(y, params, y.detach())
        flat = torch.cat([g.reshape(-1) for g in gws])
        dist.all_reduce(flat)
        return flat

    for _ in range(3):
        donn_step()
    torch.cuda.synchronize(dev)
    dist.barrier()
    e0.record()
    for _ in range(10):
        donn_step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = max_ms(e0.elapsed_time(e1) / 10)
    rep = _donn_report(ms, B * world, n, layers, hbm_peak * world, _fp32_peak(dev, sm_mhz) * world)
    rep.update({"scaling": "weak", "n_gpus": world, "batch_per_gpu": B})
    out["c4_donn_dp_batch%d_per_gpu" % B] = rep
    return out


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
    asm = ASM_prop(z_distance=Z, device=dev, kernel_mode=os.environ.get("THZ_KERNEL_MODE", "auto"))    # what ASM_prop ships as its default
    asm.check_Zc = False
    x_dev = x_host.to(dev).requires_grad_(True)
    lam_t = torch.tensor(lams, dtype=torch.float32, device=dev)        # built once: the modules key their plans on these objects
    sp_t = torch.tensor([SPACING, SPACING], dtype=torch.float32, device=dev)
    gw_host = torch.empty(1, 1, n, n, dtype=torch.float32).pin_memory()

    # N > 1: the sum of the weight gradient over the ranks forms inside the adjoint's last kernel (multimem.red through the
    # NVSwitch, parallel.FusedGradReduce) where NVLS multicast exists; otherwise one NCCL all-reduce after the backward pass
    fused_reduce = False
    if world > 1 and os.environ.get("THZ_BENCH_FUSED_REDUCE", "1") == "1":
        from quantizationawarethzdoe_b200 import parallel as P
        ok = torch.tensor([1.0 if P.FusedGradReduce.available(dev) else 0.0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) > 0:
            P.fuse_gradient_allreduce(doe)
            fused_reduce = True

    def step(x):
        """One hot-path pass; returns the gradient wrt the DOE weights (summed over the ranks when world > 1)."""
        field = ElectricField(x, wavelengths=lam_t, spacing=sp_t, device=dev)
        y = asm(doe(field)).data
        gx, gw = torch.autograd.grad(y, (x, doe.weight_height_map), y.detach())
        if world > 1 and not fused_reduce:
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

    import glob
    nodes = len(glob.glob("/sys/devices/system/node/node[0-9]*"))
    numa_note = ("%d NUMA node(s) visible to this process, %d CPUs; pinned buffers are first-touched by the rank that owns them" % (
        nodes, os.cpu_count() or 0)) + ("" if nodes > 1 else
                                        " -- a single-node (virtualised) host: there is no second node to bind a rank's buffers to, so at "
                                        "N > 2 the ranks' H2D copies share one host memory / PCIe root path")
    e2e_run(2)
    e2e_steps = max(2, min(args.steps, 10))
    ms_e2e = timed(lambda: e2e_run(e2e_steps), 1) / e2e_steps
    e2e_val = samples_per_step / (ms_e2e * 1e-3) / 1e6
    numa_note += "; aggregate H2D %.1f GB/s over %d rank(s)" % (world * x_host.numel() * 8 / (ms_e2e * 1e-3) / 1e9, world)

    # ---- per-kernel attribution with CUDA events on the launching stream (separate pass, slight overhead)
    kernels = None
    psteps = min(args.steps, 5)
    if rank == 0:
        lib.thz_profile_enable(1)
    for _ in range(psteps):          # every rank runs these steps (they contain the all-reduce); rank 0 records events
        step(x_dev)
    barrier()
    if rank == 0:
        ms_sum = (ctypes.c_float * 10)()
        cnt = (ctypes.c_int32 * 10)()
        lib.thz_profile_read(10, ms_sum, cnt)
        lib.thz_profile_enable(0)
        names = ["row_fft_fwd", "column_fft_H_ifft", "row_ifft_epilogue", "fft2_col", "doe_modulate", "quantizer", "czt_cuda_core",
                 "train", "czt_tcgen05", "reserved"]
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

    peak, peak_src = measured_peaks()
    sm_mhz = (clocks or {}).get("sm_mhz") if rank == 0 else None
    secondary = None
    if not args.no_secondary:
        try:
            secondary = secondary_single(dev, lib, sm_mhz, peak) if world == 1 else secondary_multi(dev, rank, world, lib, sm_mhz, peak)
        except Exception as e:          # a secondary line must never cost the headline
            secondary = {"error": "%s: %s" % (type(e).__name__, e)}
    if rank != 0:
        return
    step_achieved = BYTES_PER_SAMPLE * (B * C * Np * Np) / (ms_step * 1e-3) / 1e9      # per GPU, whole step
    dom = max(kernels.items(), key=lambda kv: kv[1]["ms_per_step"])[0] if kernels else None
    for ent in (kernels or {}).values():
        if "achieved_gbs" in ent:
            ent["frac"] = ent["achieved_gbs"] / peak
    # DRAM traffic of the dominant kernel per launch: NOT measured in this run -- taken from the committed ncu --set full
    # capture named in traffic_source and scaled to this launch size
    traffic, traffic_src = None, None
    for fn in ("r02_k2_dram_traffic.json", "r01_k2_dram_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", fn)) as f:
                t = json.load(f)
            traffic = (t["dram_bytes_read"] + t["dram_bytes_write"]) * (B * C) / t["fields_per_launch"]
            traffic_src = "ncu --set full capture profiles/%s (dram__bytes_read.sum + dram__bytes_write.sum), rescaled to %d fields; not measured live" % (fn, B * C)
            break
        except Exception:
            continue
    dk = (kernels or {}).get(dom, {})
    dom_alg_bytes = 16.0 * n * Np * B * C                  # column pass: read + write of the live rows, per launch
    # FP32 co-limit (SURVEY 7 / DESIGN 3.2): lane-operations the radix-16 pipeline needs per padded sample of one fwd + bwd
    # step, counted from the butterfly structure, against 148 SMs x 128 FP32 lanes x the SM clock sampled during the run.
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    lane_rate = sms * FP32_LANES_PER_SM * (sm_mhz or 1965.0) * 1e6
    fp32_min_ms = FP32_LANE_OPS_PER_SAMPLE * (B * C * Np * Np) / lane_rate * 1e3
    hbm_min_ms = BYTES_PER_SAMPLE * (B * C * Np * Np) / (peak * 1e9) * 1e3
    k2_lane_ops = FP32_LANE_OPS_K2_PER_POINT * (Np * Np * B * C)
    k2_launch_ms = (dk.get("ms_per_step") / dk.get("launches_per_step")) if dk else None
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
                   "kernel_mode_resolved": asm.resolved_kernel_mode, "inregister_estimate": asm.inregister_estimate,
                   "l2": "inputs larger than L2 (512 MiB fields + 2 x 1 GiB intermediate spectra per step)", "tune": dict(Fn.TUNE),
                   "parallelism": ("dp%d over wavelengths, grad(weights) summed %s" % (
                       world, "inside the adjoint's last kernel (multimem.red over NVSwitch multicast + one symmetric-memory barrier)"
                       if fused_reduce else "by one NCCL all-reduce per step")) if world > 1 else "single GPU"},
        "e2e": {"value": e2e_val, "unit": "Msamples/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": x_host.numel() * 8,
                "d2h_bytes_per_step": gw_host.numel() * 4, "numa": numa_note},
        "gpu_launches": int(launches),
        "clocks": clocks,
        # dominant kernel (the column pass K2): algorithmic bytes per launch / its mean launch duration (CUDA events on the
        # launching stream).  `bound` names the limit that actually binds: the kernel moves its algorithmic bytes once
        # (traffic ~ 0.98 x algorithmic) at a fraction of the HBM peak because the FP32 pipe saturates first; both are given.
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": dk.get("achieved_gbs"), "peak": peak, "unit": "GB/s",
                     "frac": dk.get("frac"), "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": dom_alg_bytes, "launch_ms": k2_launch_ms,
                     "peak_source": peak_src,
                     "binding_limit": "fp32" if fp32_min_ms > hbm_min_ms else "hbm",
                     "fp32_issue": {"note": "analytic lane-operation count of the radix-16 pipeline (butterflies, twiddles, power trees, "
                                            "H generation), FMA = 1 lane-op; peak = SMs x 128 lanes x sampled SM clock",
                                    "lane_ops_per_sample": FP32_LANE_OPS_PER_SAMPLE, "peak_lane_ops_per_s": lane_rate,
                                    "step_min_ms": fp32_min_ms, "step_frac": fp32_min_ms / ms_step,
                                    "kernel_lane_ops_per_launch": k2_lane_ops,
                                    "kernel_min_ms": k2_lane_ops / lane_rate * 1e3,
                                    "kernel_frac": (k2_lane_ops / lane_rate * 1e3 / k2_launch_ms) if k2_launch_ms else None},
                     "step": {"achieved": step_achieved, "frac": step_achieved / peak, "bytes_per_sample": BYTES_PER_SAMPLE,
                              "hbm_min_ms": hbm_min_ms,
                              "scope": "whole step: 6 FFT-pipeline launches + 2 level-selection launches, per GPU"},
                     "kernels": kernels},
        "fields_per_s": value * 1e6 / (Np * Np),
        "secondary": secondary,
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
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary configurations (C2/C3/C4; slab + DONN DP when N > 1)")
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
