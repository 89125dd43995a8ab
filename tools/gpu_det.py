import sys; sys.path.insert(0, "/root/repo")
import fermi_jl_b200 as fb, json
eng = fb.Engine(0)
for o, v in [(24, 114), (10, 160)]:
    x = fb.synth.make_inputs(o, v, naux=32)
    eng.upload_conv(o, v, x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
    for det in (0, 1):
        eng.set_deterministic(det)
        rs = [eng.compute(0, -1) for _ in range(5)]
        print(json.dumps({"o": o, "v": v, "deterministic": det, "kernel_ms": min(r[1]["kernel_ms"] for r in rs), "distinct_E": len(set(r[0] for r in rs))}), flush=True)
