"""A/B timing of kernel variants x debug flags:  python tools/gpu_ab.py o v [o v ...] [--flags 0,8,1,9] [--variants 1,2]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fermi_jl_b200 as fb

def opt(name, default):
    for i, a in enumerate(sys.argv):
        if a == name:
            return [int(x) for x in sys.argv[i + 1].split(",")]
    return default

pos = []
skip = False
for a in sys.argv[1:]:
    if skip: skip = False; continue
    if a.startswith("--"): skip = True; continue
    pos.append(int(a))
shapes = list(zip(pos[0::2], pos[1::2])) or [(24, 114)]
flags = opt("--flags", [0, 8])
variants = opt("--variants", [1, 2])
eng = fb.Engine(0)
out = []
for o, v in shapes:
    x = fb.synth.make_inputs(o, v, naux=32)
    eng.upload_conv(o, v, x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
    for var in variants:
        eng.set_kernel_variant(var)
        for fl in flags:
            eng.set_debug_flags(fl)
            rs = [eng.compute(0, -1) for _ in range(3)]
            best = min(rs, key=lambda r: r[1]["kernel_ms"])
            rec = {"o": o, "v": v, "variant": var, "flags": fl, "ms": round(best[1]["kernel_ms"], 3),
                   "tflops": round(best[1]["flops"] / best[1]["kernel_ms"] / 1e9, 2), "E": best[0]}
            out.append(rec)
            print(json.dumps(rec), flush=True)
    eng.set_debug_flags(0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gpu_ab.json", "w"), indent=1)
