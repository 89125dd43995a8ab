"""First-contact GPU script: FP64 peak calibration + timing of the fused kernel at the BASELINE shapes."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fermi_jl_b200 as fb

out = {}
eng = fb.Engine(0)
out["dmma_tflops"] = eng.fp64_peak(0, 300.0)
out["dfma_tflops"] = eng.fp64_peak(1, 300.0)
a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b = torch.randn_like(a)
for _ in range(2): torch.matmul(a, b)
torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5): torch.matmul(a, b)
e.record(); torch.cuda.synchronize()
out["cublas_dgemm_tflops"] = 5 * 2 * 8192**3 / (s.elapsed_time(e) * 1e-3) / 1e12
del a, b
print(json.dumps(out), flush=True)
for name, (o, v) in {"c2": (5, 53), "c3": (15, 93), "c4": (24, 114), "mid": (10, 160)}.items():
    x = fb.synth.make_inputs(o, v, naux=32)
    t0 = time.time(); eng.upload_conv(o, v, x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv); tu = time.time() - t0
    best = None
    for rep in range(3):
        e_, st = eng.compute(0, -1)
        best = st if best is None or st["kernel_ms"] < best["kernel_ms"] else best
    out[name] = {"o": o, "v": v, "E": e_, "kernel_ms": best["kernel_ms"], "tflops": best["flops"] / best["kernel_ms"] / 1e9,
                 "upload_s": tu, "items": best["n_items"]}
    print(name, json.dumps(out[name]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gpu_first.json", "w"), indent=1)
