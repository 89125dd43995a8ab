"""C5 proxy: v=400 with a reduced occupied space (same per-item work as o=40/v=400 up to K = v+o)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fermi_jl_b200 as fb
o, v = int(sys.argv[1]) if len(sys.argv) > 1 else 6, 400
eng = fb.Engine(0)
x = fb.synth.make_inputs(o, v, naux=32)
t0 = time.time(); eng.upload_conv(o, v, x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv); tu = time.time() - t0
res = {}
for fl in (0, 3):
    eng.set_debug_flags(fl)
    best = min((eng.compute(0, -1)[1] for _ in range(2)), key=lambda s: s["kernel_ms"])
    res[f"flags{fl}"] = {"kernel_ms": best["kernel_ms"], "tflops": best["flops"] / best["kernel_ms"] / 1e9}
print(json.dumps({"o": o, "v": v, "upload_s": tu, **res}))
