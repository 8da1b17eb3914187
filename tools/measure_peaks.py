"""Measures the roofline denominators MEASURED_PEAKS.json does not hold -- dense int8 (cuBLASLt IGEMM through torch._int_mm)
and fp64 (cuBLAS DGEMM) -- with the same protocol the driver used for bf16 (8192^3, best of 10 = burst; back to back for a
few seconds = sustained), plus the bf16 and HBM-copy figures again as a cross-check.  Run on the GPU box:
    python tools/measure_peaks.py gpurun_out/peaks.json
The result is copied to profiles/peaks.json (tracked); bench.py divides by these."""
import json
import subprocess
import sys
import time

import torch


def clocks():
    try:
        out = subprocess.run(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.sw_power_cap",
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip()
        return out
    except Exception:
        return ""


def time_op(fn, reps):
    fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def sustained(fn, seconds):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    count, t0, mid = 0, time.time(), ""
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(8):
            fn()
        count += 8
        torch.cuda.synchronize()
        if not mid and time.time() - t0 > seconds / 2:
            mid = clocks()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / count, mid


def main():
    out = {"gpu_name": torch.cuda.get_device_name(0), "torch": torch.__version__, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    N = 8192
    ops = 2.0 * N ** 3
    a8 = torch.randint(-128, 128, (N, N), dtype=torch.int8, device="cuda")
    b8 = torch.randint(-128, 128, (N, N), dtype=torch.int8, device="cuda")
    ms = time_op(lambda: torch._int_mm(a8, b8), 10)
    out["int8_tops"] = ops / (ms * 1e-3) / 1e12
    ms, ck = sustained(lambda: torch._int_mm(a8, b8), 4.0)
    out["int8_tops_sustained"] = ops / (ms * 1e-3) / 1e12
    out["int8_clocks_mid"] = ck
    del a8, b8
    a = torch.randn(N, N, dtype=torch.bfloat16, device="cuda")
    b = torch.randn(N, N, dtype=torch.bfloat16, device="cuda")
    ms = time_op(lambda: torch.matmul(a, b), 10)
    out["bf16_tflops"] = ops / (ms * 1e-3) / 1e12
    ms, ck = sustained(lambda: torch.matmul(a, b), 4.0)
    out["bf16_tflops_sustained"] = ops / (ms * 1e-3) / 1e12
    out["bf16_clocks_mid"] = ck
    del a, b
    a = torch.randn(N, N, dtype=torch.float64, device="cuda")
    b = torch.randn(N, N, dtype=torch.float64, device="cuda")
    ms = time_op(lambda: torch.matmul(a, b), 10)
    out["fp64_tflops"] = ops / (ms * 1e-3) / 1e12
    ms, ck = sustained(lambda: torch.matmul(a, b), 4.0)
    out["fp64_tflops_sustained"] = ops / (ms * 1e-3) / 1e12
    out["fp64_clocks_mid"] = ck
    del a, b
    x = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda")
    y = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda")
    ms = time_op(lambda: y.copy_(x), 10)
    out["hbm_gbs"] = 2.0 * x.numel() * 2 / (ms * 1e-3) / 1e9
    out["how"] = ("torch._int_mm int8 8192^3 (cuBLASLt IGEMM, int32 accumulate), torch.matmul bf16 / fp64 8192^3: best of 10 with CUDA events "
                  "(burst) and back to back for 4 s (sustained, clocks sampled mid-way: sm MHz, max MHz, W, power-cap flag); "
                  "y.copy_(x) over 1 Gi bf16 elements (read + write bytes)")
    path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/peaks.json"
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
