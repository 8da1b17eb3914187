#!/usr/bin/env python
"""bench.py -- NLML+gradient evaluations per second of the GPz hot path on B200.

Metric (BASELINE.json): "NLML+grad evals/sec at (n,d,m) per covariance mode; 1/2/4/8 GPU vs CPU ref".
Headline workload (north_star target): synthetic n=1e6, d=10, m=1000, method VC, heteroscedastic, k=1.
A "step" = one full objective+gradient evaluation GPz(theta) (GPz/GPz.m:1-263) at a fresh theta
(theta0 + small perturbations, as a line search visits).  With N GPUs the n rows are sharded
(strong scaling: total n fixed) and every evaluation does two NCCL allreduces.

  value  : evals/s with theta already in HBM and the result left in HBM (gpz_eval_dev), CUDA events
           on the library's stream, max over ranks.
  e2e    : evals/s through the host-buffer C-ABI call a MATLAB/MEX caller makes (gpz_eval): theta H2D
           and (f, grad, stats) D2H inside the timed region, every step.
  --impl reference : the reference algorithm's CPU path.  The reference is MATLAB (not installable
           here), so this arm times the NumPy restatement oracle/gpz_oracle.py ("port") with all
           host threads on a bounded row sample of the same workload and scales linearly in n.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n, d, m, method, extras).  "gpus": the GPU count BASELINE.json states the config on -- those workloads are WEAK
    # scaling (rows per GPU fixed at n / gpus, every rank generates its own rows), the others shard a fixed n (strong).
    "target": (1_000_000, 10, 1000, "VC", {}),                         # north-star headline; = configs[3] per GPU shape
    "cfg3": (1_000_000, 10, 500, "VD", {}),                            # configs[2]: 1-GPU roofline capture
    "cfg4": (10_000_000, 10, 1000, "VC", {"gpus": 8}),                 # configs[3]: n = 1e7 sharded over 8 GPUs
    "cfg5": (1_000_000, 32, 2000, "GC", {"gpus": 4, "psi": True}),     # configs[4]: GC + input noise over 4 GPUs
    "photoz": (60_000, 5, 100, "VC", {}),
    "small": (20_000, 10, 256, "VC", {}),
    "m1000_small_n": (20_000, 10, 1000, "VC", {}),                      # the m x m solve dominates: what 8 GPUs see of it
}
METRIC = "NLML+grad evals/sec at (n,d,m) per covariance mode"


def workload_desc(name, n, d, m, method):
    ex = WORKLOADS[name][4]
    return (f"synthetic n={n} d={d} m={m} {method}{' + input noise Psi (d x d x n)' if ex.get('psi') else ''} heteroscedastic k=1 ({name}"
            + (f": {WORKLOADS[name][0] // ex['gpus']} rows per GPU, BASELINE states it on {ex['gpus']} GPUs" if ex.get("gpus") else "") + ")")


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def fp64_peak_tflops(torch, dev, reps=10, N=8192):
    """Same protocol as MEASURED_PEAKS.json's GEMM figure, in fp64: cuBLAS DGEMM N^3, best of reps."""
    a = torch.randn(N, N, dtype=torch.float64, device=dev)
    b = torch.randn(N, N, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    torch.cuda.synchronize(dev)
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize(dev)
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0 * N ** 3 / (best * 1e-3) / 1e12


def make_problem(name, seed=0):
    from gpz_b200 import synth
    n, d, m, method, _ = WORKLOADS[name]
    X, Y = synth.make_data(n, d, seed=seed)
    theta0 = synth.make_theta0(X, Y, method, m, het=True, seed=seed + 1)
    return n, d, m, method, X, Y, theta0


def make_shard(name, rank, world, seed=0):
    """This rank's rows.  Strong workloads: the contiguous block [n rank / world, n (rank+1) / world) of the one seeded data
    set.  Weak workloads ("gpus" set): n / gpus rows generated by this rank alone (seed + 1000 + rank: no 800 MB host array
    per process), theta0 from a fixed 100 000-row sample so that it is the same on every rank."""
    from gpz_b200 import synth
    n, d, m, method, ex = WORKLOADS[name]
    if not ex.get("gpus"):
        X, Y = synth.make_data(n, d, seed=seed)
        theta0 = synth.make_theta0(X, Y, method, m, het=True, seed=seed + 1)
        lo, hi = (n * rank) // world, (n * (rank + 1)) // world
        return dict(n_total=n, rows=hi - lo, X=X[lo:hi], Y=Y[lo:hi], Psi=None, theta0=theta0, scaling="strong")
    rows = n // ex["gpus"]
    Xs, Ys = synth.make_data(100_000, d, seed=seed)
    theta0 = synth.make_theta0(Xs, Ys, method, m, het=True, seed=seed + 1)
    X, Y = synth.make_data(rows, d, seed=seed + 1000 + rank)
    Psi = synth.make_psi(rows, d, method, seed=seed + 2000 + rank) if ex.get("psi") else None
    return dict(n_total=rows * world, rows=rows, X=X, Y=Y, Psi=Psi, theta0=theta0, scaling="weak")


def thetas_for(theta0, count, seed=100):
    rng = np.random.default_rng(seed)
    return [theta0 + 0.01 * rng.standard_normal(theta0.size) for _ in range(count)]


# ----------------------------------------------------------------------------------------------
def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU legs are meant to use every host core."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass


def cpu_oracle_time(name, n_sample, reps, seed=0):
    """Seconds per evaluation of the NumPy restatement on an n_sample-row sample of the workload."""
    from gpz_b200 import synth
    from oracle import gpz_oracle as O
    use_all_host_threads()
    n, d, m, method, ex = WORKLOADS[name]
    ns = min(n, n_sample)
    X, Y = synth.make_data(ns, d, seed=seed)
    theta0 = synth.make_theta0(X, Y, method, m, het=True, seed=seed + 1)
    model = O.Model(d=d, k=1, m=m, method=method, heteroscedastic=True)
    X, Y = np.array(X), np.array(Y)
    Psi = np.array(synth.make_psi(ns, d, method, seed=seed + 2000)) if ex.get("psi") else None
    ths = thetas_for(theta0, reps)
    ts = []
    for th in ths:
        t0 = time.perf_counter()
        O.GPz(th, model, X, Y, Psi)
        ts.append(time.perf_counter() - t0)
    # the m x m SVD pseudo-inverse does not grow with n: time it alone so only the n-proportional part is scaled
    A = np.random.default_rng(0).standard_normal((m, m + 8))
    S = A @ A.T + np.eye(m)
    t0 = time.perf_counter()
    O.inv_logdet(S)
    t_svd = time.perf_counter() - t0
    return ns, ts, t_svd


def scale_cpu_time(sec_sample, t_svd, n, ns):
    """t(n) = a*n + c with c = the n-independent SVD: scale only the n-proportional part."""
    t_svd = min(t_svd, 0.5 * sec_sample)
    return (sec_sample - t_svd) * (n / ns) + t_svd


def cpu_sample_rows(name, budget):
    """Rows of the bounded CPU sample: ~budget 'flop-like units' per evaluation (dense paths: 4 GEMMs of 2 n m^2 plus the
    per-basis loops; C modes + Psi: the reference's scalar loop over n x m with d x d solves, ~25 000 (i, j) pairs per second)."""
    n, d, m, method, ex = WORKLOADS[name]
    if ex.get("psi") and method[1] == "C":
        return max(16, min(n, int(budget / 4.0e10 * 1.0e5 / m)))
    return max(2000, min(n, int(budget / (m * m + 40.0 * m * d * d))))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    name = args.workload
    n, d, m, method, ex = WORKLOADS[name]
    if ex.get("gpus"):                                   # weak workloads: the job the GPU arm runs at this GPU count
        n = n // ex["gpus"] * max(1, args.gpus)
    ns = args.cpu_sample or cpu_sample_rows(name, 4.0e10)                                   # ~4-6 s per step
    total = args.warmup + args.steps
    ns, ts, t_svd = cpu_oracle_time(name, ns, total)
    timed = ts[args.warmup:] if len(ts) > args.warmup else ts
    sec_sample = float(np.mean(timed))
    sec_full = scale_cpu_time(sec_sample, t_svd, n, ns)
    value = 1.0 / sec_full
    cores = os.cpu_count() or 1
    # linearity of t(n) = a n + c on this box: one more evaluation at half the sample
    _, th, _ = cpu_oracle_time(name, max(1, ns // 2), 1)
    half_pred = (sec_sample - min(t_svd, 0.5 * sec_sample)) * 0.5 + min(t_svd, 0.5 * sec_sample)
    lin = {"rows": [max(1, ns // 2), ns], "seconds": [th[0], sec_sample], "residual_at_half": (th[0] - half_pred) / th[0]}
    try:
        lin["tracked_run_at_5e4_1e5_rows"] = json.load(open(os.path.join(ROOT, "profiles", "r02_cpu_linearity.json")))
    except Exception:
        pass
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_full * 1e3, "higher_is_better": True,
        "scaling": "weak" if ex.get("gpus") else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "linearity": lin,
        "config": {"workload": workload_desc(name, n, d, m, method),
                   "note": "reference is MATLAB (no MATLAB/Octave here): NumPy restatement of GPz.m/getPHI.m/inv_logdet.m, "
                           "same operation sequence, OpenBLAS threads = all host cores"},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port",
                         "sample": f"{ns} of {n} rows per step, same d/m/mode; t(n)=a*n+c with c = the m x m SVD ({t_svd:.2f} s) "
                                   f"timed alone, a*n scaled x{n / ns:.1f}; measured {sec_sample:.3f} s/eval on the sample"},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    from gpz_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    name = args.workload
    _, d, m, method, wex = WORKLOADS[name]
    sh = make_shard(name, rank, world)                               # this rank's rows (strong: a block of the one data set)
    n, theta0, scaling = sh["n_total"], sh["theta0"], sh["scaling"]
    Xs, Ys, Psis = sh["X"], sh["Y"], sh["Psi"]
    lo, hi = 0, sh["rows"]
    t0 = time.perf_counter()
    ctx = L.Context(L.make_model(d, 1, m, method, True), Xs, Ys, Psis, device=local)
    if world > 1:
        uid = [L.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])
    upload_s = time.perf_counter() - t0
    p = theta0.size
    W, K = args.warmup, args.steps
    ths = thetas_for(theta0, W + K)
    d_th = [torch.from_numpy(t).to(dev) for t in ths]
    d_out = torch.empty(p + 5, dtype=torch.float64, device=dev)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident arm ------------------------------------------------------------------
    for i in range(W):
        ctx.eval_dev(d_th[i].data_ptr(), d_out.data_ptr())
    ctx.sync()
    barrier()
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gram_ms, tgemm_ms, phases = [], [], []
    e0.record(stream)
    for i in range(K):
        ctx.eval_dev(d_th[W + i].data_ptr(), d_out.data_ptr())
    e1.record(stream)
    ctx.sync()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    tm = ctx.last_timing()            # CUDA events recorded inside the last timed evaluation
    kt = ctx.kernel_timing()          # ... and around single kernels of it
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / K
    f_last = float(d_out[0].item())

    # ---- end-to-end arm: host theta in, host (f, g, stats) out, every step -----------------------
    for i in range(min(W, 2)):
        ctx.eval(ths[i])
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        f_e2e, g_e2e, st = ctx.eval(ths[W + i])
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_step = float(t.item()) / K
    assert abs(f_e2e - f_last) <= 1e-12 * abs(f_last), (f_e2e, f_last)   # same theta -> same answer on both arms

    # ---- a few iterations of the training loop (gpz_train: minFunc L-BFGS + line search on the device) -- reported
    # beside the metric, outside every timed region; all ranks take part (the objective all-reduces)
    train = None
    if args.train_iters > 0:
        try:
            barrier()
            _, _, _, ti = ctx.train(theta0, theta0, -np.inf, max_iter=args.train_iters, training_only=1)
            train = {"iterations": ti["iterations"], "fun_evals": ti["fun_evals"], "ms_total": ti["ms_total"],
                     "ms_in_evals": ti["ms_eval"], "optimizer_ms_per_iteration": (ti["ms_total"] - ti["ms_eval"]) / max(1, ti["iterations"]),
                     "nlogML_start": float(ctx.eval(theta0)[0]), "nlogML_end": ti["f"], "exit": ti["message"]}
        except Exception as e:                                           # never lose the bench line to the extra leg
            train = {"error": str(e)[:200]}
        barrier()

    # ---- second run with a 10 % validation mask: the evaluation also builds PHI on the validation rows and
    # returns validRMSE / validLL (GPz.m:239-259); reported beside the metric, not part of it
    valid_run = None
    m_bases = ctx.model.m
    if args.valid_frac > 0:
        try:
            ctx.close()
            ctx = None
            va_mask = (np.arange(lo, hi) % max(2, int(round(1.0 / args.valid_frac)))) == 0
            ctx2 = L.Context(L.make_model(d, 1, m, method, True), Xs, Ys, Psis, None, ~va_mask, va_mask, device=local)
            if world > 1:
                uid = [L.comm_unique_id() if rank == 0 else None]
                dist.broadcast_object_list(uid, src=0)
                ctx2.comm_init(rank, world, uid[0])
            st2 = torch.cuda.ExternalStream(ctx2.stream(), device=dev)
            for i in range(2):
                ctx2.eval_dev(d_th[i].data_ptr(), d_out.data_ptr())
            ctx2.sync()
            barrier()
            kv = min(K, 5)
            v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            v0.record(st2)
            for i in range(kv):
                ctx2.eval_dev(d_th[W + i].data_ptr(), d_out.data_ptr())
            v1.record(st2)
            ctx2.sync()
            barrier()
            tv = torch.tensor([v0.elapsed_time(v1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tv, op=dist.ReduceOp.MAX)
            stats_v = d_out[-4:].tolist()
            valid_run = {"valid_fraction": float(va_mask.mean()), "steps": kv, "ms_per_step": float(tv.item()) / kv,
                         "value": 1e3 * kv / float(tv.item()), "unit": "evals/s", "validRMSE": stats_v[2], "validLL": stats_v[3]}
            ctx2.close()
        except Exception as e:
            valid_run = {"error": str(e)[:200]}
        barrier()

    # ---- roofline of the dominant kernel -------------------------------------------------------------
    peak = fp64_peak_tflops(torch, dev) if rank == 0 else None
    line = None
    if rank == 0:
        n_loc = hi - lo
        flops_tgemm = 2.0 * n_loc * m * m                  # algorithmic fp64 flops of ONE of the two n x m x m GEMMs (F_gemm/2)
        prof = {}
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "summary.json")))
        except Exception:
            pass
        measured = {}
        try:
            measured = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        fp64_equiv = {"t_gemm_tflops": flops_tgemm / (tm["tgemm_kernel"] * 1e-3) / 1e12,
                      "gram_tflops": flops_tgemm / (tm["gram_kernel"] * 1e-3) / 1e12,
                      "eval_tflops": (4.0 * n * m * m / world) / (ms_step * 1e-3) / 1e12,
                      "dgemm_peak_tflops": peak,
                      "note": "algorithmic fp64 flops (2nm^2 per GEMM, 4nm^2 per eval) per second; the fp64 DMMA pipe ceiling is "
                              "the cuBLAS DGEMM 8192^3 figure measured in this run -- values above it come from the error-free "
                              "int8-slice (Ozaki) GEMMs on the tcgen05 tensor cores"}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "profiles", "peaks.json")))
        except Exception:
            pass
        if tm["int8_slices"] > 0:
            i8 = peaks.get("int8_tops_sustained")
            src = ("profiles/peaks.json int8_tops_sustained: cuBLASLt IGEMM 8192^3 (torch._int_mm) back to back for 4 s on this pool's B200 "
                   "(tools/measure_peaks.py; the kernel is timed inside a long step, so the sustained figure; burst: %s)" % peaks.get("int8_tops"))
            if not i8:
                bf16 = measured.get("bf16_tflops_sustained") or 1400.0
                i8, src = 2.0 * bf16, "2 x sustained bf16 (profiles/peaks.json absent: kind::i8 issues at twice the bf16 rate)"
            ach = tm["i8_gemms_ops"] / (tm["i8_gemms_ms"] * 1e-3) / 1e12
            roof = {"bound": "tensor",
                    "kernel": "ozmma_kernel (hand-written tcgen05.mma kind::i8 cta_group::2 N=256 over two digit levels, TMA, TMEM double "
                              "buffer; %d base-256 digits, levels folded in fp64 registers, fused nu/H epilogue): T = PHI*iSigma, one launch "
                              "over all rows" % tm["int8_slices"],
                    "achieved": ach, "peak": i8, "unit": "TFLOP/s", "frac": ach / i8,
                    "ops": "int8 multiply-adds x2 executed by that launch (2 n MP^2 per digit pair, s(s+1)/2 pairs)",
                    "traffic": prof.get("ozmma_tgemm_dram_bytes_per_launch") if n_loc == 1000000 and name == "target" else None,
                    "peak_source": src,
                    "kernel_ms": tm["i8_gemms_ms"], "executed_ops": tm["i8_gemms_ops"],
                    "int8_slices": tm["int8_slices"], "int8_gram": tm["int8_gram"]}
        else:
            ach = flops_tgemm / (tm["tgemm_kernel"] * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": "tgemm_kernel (T = PHI*iSigma, DMMA.8x8x4 fp64)",
                    "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": prof.get("tgemm_dram_bytes_per_launch"),
                    "peak_source": "measured in this run: cuBLAS DGEMM 8192^3 fp64, best of 10, CUDA events "
                                   "(MEASURED_PEAKS.json has no fp64 figure; tcgen05 has no fp64 kind, DMMA is the fp64 tensor path)",
                    "algorithmic_flops_per_launch": flops_tgemm, "kernel_ms": tm["tgemm_kernel"]}
        roof["fp64_equivalent"] = fp64_equiv
        # the PHI build against the HBM roofline (north star: ">= 60 %"; SURVEY 8d: B_PHI = 8 (n m + n d [+ Psi] + m d + g_dim) bytes),
        # on the PHI kernel's OWN CUDA-event time (gpz_kernel_timing[0]); the kernel is fp64-pipe bound at d = 10
        hbm = measured.get("hbm_gbs") or 6650.0
        g_dim = int(p) - m * d - 3 * m - 1
        b_phi = 8.0 * (n_loc * m + n_loc * d + (n_loc * d * d if wex.get("psi") else 0) + m * d + g_dim)
        t_phi = kt["phi_build"] if kt["phi_build"] > 0 else tm["phi"]
        roof["phi_path"] = {"algorithmic_bytes": b_phi, "ms": t_phi, "achieved_gbs": b_phi / (t_phi * 1e-3) / 1e9,
                            "peak_gbs": hbm, "frac": b_phi / (t_phi * 1e-3) / 1e9 / hbm,
                            "timed": "the PHI kernel (+ the ordered sum of its row-dot partials), CUDA events" if kt["phi_build"] > 0 else "the phi phase",
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if measured.get("hbm_gbs") else "fallback 6.65 TB/s"}
        roof["kernel_ms"] = {k_: round(float(v), 3) for k_, v in kt.items() if v >= 0}
        roof["phase_ms"] = {k_: round(float(v), 3) for k_, v in tm.items() if k_ not in ("i8_gemms_ops", "int8_slices", "int8_gram")}
        cpu = None
        if world == 1 and not args.no_cpu:
            ns = args.cpu_sample or cpu_sample_rows(name, 1.2e11)                                  # ~10-20 s of CPU work
            ns, ts, t_svd = cpu_oracle_time(name, ns, 2)
            sec = scale_cpu_time(min(ts), t_svd, n, ns)
            cpu = {"value": 1.0 / sec, "unit": "evals/s", "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": f"{ns} of {n} rows, same d/m/mode, best of 2; t(n)=a*n+c with c = the m x m SVD ({t_svd:.2f} s), "
                             f"a*n scaled x{n / ns:.1f}; NumPy restatement of the MATLAB reference, {min(ts):.2f} s on the sample"}
        line = {
            "metric": METRIC, "value": 1e3 / ms_step, "unit": "evals/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_desc(name, n, d, m, method), "rows_per_gpu": hi - lo,
                       "parallelism": f"rows sharded over {world} GPU(s), 2 NCCL allreduces per eval" if world > 1 else "1 GPU",
                       "n_total": int(n),
                       "cache": "inputs larger than L2: PHI/H working set %.1f GB per GPU" % (16.0 * (hi - lo) * m_bases / 1e9),
                       "dataset_upload_s": round(upload_s, 3), "theta_len": int(p)},
            "e2e": {"value": 1.0 / e2e_step, "unit": "evals/s", "h2d_bytes_per_step": 8 * int(p), "d2h_bytes_per_step": 8 * (int(p) + 5),
                    "ms_per_step": e2e_step * 1e3},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "train": train,
            "with_validation": valid_run,
            "check": {"nlogML_last": f_last, "trainRMSE": float(st["trainRMSE"])},
        }
        print(json.dumps(line), flush=True)
    if ctx is not None:
        ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_single_process(args):
    """The same metric through gpz_create_multi / gpz_multi_eval: ONE process, one caller thread, N GPUs (the way one MATLAB
    interpreter would drive them) -- no torchrun.  Host theta in, host (f, g, stats) out every step, so the
    whole line is an end-to-end number; the per-rank device times of the last evaluation are reported beside it."""
    import torch

    from gpz_b200 import _lib as L
    N = args.gpus
    if torch.cuda.device_count() < N:
        raise SystemExit(f"--single-process --gpus {N}: only {torch.cuda.device_count()} device(s) visible")
    name = args.workload
    n, d, m, method, wex = WORKLOADS[name]
    if wex.get("gpus"):
        raise SystemExit("--single-process takes the strong-scaling workloads (one host data set split inside the library)")
    n, d, m, method, X, Y, theta0 = make_problem(name)
    t0 = time.perf_counter()
    mc = L.MultiContext(L.make_model(d, 1, m, method, True), X, Y, ngpus=N)
    upload_s = time.perf_counter() - t0
    W, K = args.warmup, args.steps
    ths = thetas_for(theta0, W + K)
    for i in range(W):
        mc.eval(ths[i])
    sampler = ClockSampler(0)
    sampler.start()
    t0 = time.perf_counter()
    for i in range(K):
        f, g, st = mc.eval(ths[W + i])
    sec = (time.perf_counter() - t0) / K
    clocks = sampler.stop()
    ranks = [mc.rank_timing(r) for r in range(N)]
    p = int(theta0.size)
    line = {"metric": METRIC, "value": 1.0 / sec, "unit": "evals/s", "n_gpus": N, "steps": K, "warmup": W, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "mode": "single-process: gpz_create_multi, one caller thread, one worker thread + NCCL rank per device",
            "config": {"workload": workload_desc(name, n, d, m, method), "n_total": int(n), "dataset_upload_s": round(upload_s, 3),
                       "theta_len": p, "timing": "host wall clock around gpz_multi_eval (blocking): theta H2D and f/g/stats D2H included"},
            "e2e": {"value": 1.0 / sec, "unit": "evals/s", "h2d_bytes_per_step": 8 * p * N, "d2h_bytes_per_step": 8 * (p + 5) * N,
                    "ms_per_step": sec * 1e3},
            "clocks": clocks, "rank_device_ms": [round(float(r["total"]), 3) for r in ranks],
            "check": {"nlogML_last": f, "trainRMSE": float(st["trainRMSE"])}}
    print(json.dumps(line), flush=True)
    mc.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="target", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=0, help="rows of the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--valid-frac", type=float, default=0.1, help="validation share of the extra 'with_validation' run (0 = skip)")
    ap.add_argument("--train-iters", type=int, default=5, help="iterations of the device-resident training loop reported in 'train' (0 = skip)")
    ap.add_argument("--single-process", action="store_true", help="drive --gpus N devices from this one process (gpz_create_multi), no torchrun")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if args.single_process:
        return run_single_process(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
