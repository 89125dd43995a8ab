"""Phase profile of the fused kernel (instrumented variant): python tools/gpu_phase.py [o v] ...  -> gpurun_out/gpu_phase.json
Cycles are summed over CTAs; printed per CTA-item (mean cycles per work item) for warp 0 and warp 12."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fermi_jl_b200 as fb

args = [int(a) for a in sys.argv[1:] if a.lstrip("-").isdigit()]
shapes = list(zip(args[0::2], args[1::2])) or [(24, 114)]
eng = fb.Engine(0)
variant = 1 if "--v1" in sys.argv else 2
eng.set_kernel_variant(variant)
out = {"variant": variant}
for o, v in shapes:
    x = fb.synth.make_inputs(o, v, naux=32)
    eng.upload_conv(o, v, x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
    eng.set_profiling(False)
    best = min((eng.compute(0, -1)[1] for _ in range(3)), key=lambda s: s["kernel_ms"])
    dbg = {}
    for fl in (1, 2, 3):
        eng.set_debug_flags(fl)
        dbg[f"flags{fl}_ms"] = round(min(eng.compute(0, -1)[1]["kernel_ms"] for _ in range(2)), 3)
    eng.set_debug_flags(0)
    eng.set_profiling(True)
    e_, stp = eng.compute(0, -1)
    prof = eng.last_profile()
    n_items = best["n_items"]
    rec = {"o": o, "v": v, "E": e_, "kernel_ms": best["kernel_ms"], "tflops": best["flops"] / best["kernel_ms"] / 1e9,
           "prof_kernel_ms": stp["kernel_ms"], "dbg": dbg, "n_items": n_items,
           "cycles_per_item": {k: round(val / n_items, 1) for k, val in prof.items()},
           "frac": {k: round(val / prof["total"], 4) for k, val in prof.items() if not k.startswith("g3_")},
           "g3_frac": {k: round(val / prof["g3_total"], 4) for k, val in prof.items() if k.startswith("g3_")}}
    out[f"o{o}v{v}"] = rec
    print(json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/gpu_phase_v{variant}.json", "w"), indent=1)
