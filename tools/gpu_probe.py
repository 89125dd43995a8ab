"""GPU probe: DMMA issue sweep (warps/SM x ILP) and the phase profile of the fused kernel."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fermi_jl_b200 as fb

eng = fb.Engine(0)
out = {"sweep": {}, "shapes": {}}
if "--nosweep" not in sys.argv:
    for w in (4, 8, 16, 32):
        for ilp in (1, 2, 4, 8, 16):
            out["sweep"][f"w{w}_ilp{ilp}"] = round(eng.dmma_sweep(ilp, w), 2)
    print(json.dumps(out["sweep"]), flush=True)
shapes = {"c3": (15, 93), "c4": (24, 114), "mid": (10, 160)}
for name, (o, v) in shapes.items():
    x = fb.synth.make_inputs(o, v, naux=32)
    eng.upload_conv(o, v, x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
    best = None
    eng.set_profiling(False)
    for rep in range(3):
        e_, st = eng.compute(0, -1)
        if best is None or st["kernel_ms"] < best["kernel_ms"]:
            best = st
    dbg = {}
    for fl in (1, 2, 3):
        eng.set_debug_flags(fl)
        bt = min(eng.compute(0, -1)[1]["kernel_ms"] for _ in range(2))
        dbg[f"flags{fl}_ms"] = round(bt, 3)
    eng.set_debug_flags(0)
    eng.set_profiling(True)
    e_, stp = eng.compute(0, -1)
    prof = eng.last_profile()
    tot = prof["total"]
    out["shapes"][name] = {"o": o, "v": v, "E": e_, "kernel_ms": best["kernel_ms"],
                           "tflops": best["flops"] / best["kernel_ms"] / 1e9, "prof_kernel_ms": stp["kernel_ms"], "dbg": dbg,
                           "phase_frac": {k: round(val / tot, 4) for k, val in prof.items()}}
    print(name, json.dumps(out["shapes"][name]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gpu_probe.json", "w"), indent=1)
