"""How much do edge tiles cost?  Same o, neighbouring v with and without a remainder tile."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fermi_jl_b200 as fb
eng = fb.Engine(0)
o = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for v in [int(a) for a in sys.argv[2:]] or [112, 114, 116, 120, 124, 128]:
    x = fb.synth.make_inputs(o, v, naux=24)
    eng.upload_conv(o, v, x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
    res = {}
    for fl in (0, 3):
        eng.set_debug_flags(fl)
        st = min((eng.compute(0, -1)[1] for _ in range(2)), key=lambda s: s["kernel_ms"])
        res[fl] = (round(st["kernel_ms"], 2), round(st["flops"] / st["kernel_ms"] / 1e9, 2))
    eng.set_debug_flags(0)
    print(json.dumps({"o": o, "v": v, "full_ms_tf": res[0], "kloops_only_ms_tf": res[3]}), flush=True)
