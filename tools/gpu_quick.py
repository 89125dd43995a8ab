"""Kernel-only timing at a few shapes (best of n), with the parity check against stored energies of the previous build when given:
python tools/gpu_quick.py [o v]...   -> one JSON line per shape"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fermi_jl_b200 as fb
args = [int(a) for a in sys.argv[1:] if a.isdigit()]
shapes = list(zip(args[0::2], args[1::2])) or [(24, 114), (15, 93), (10, 160), (5, 53)]
eng = fb.Engine(0)
peak = eng.fp64_peak(0, 200.0)
out = []
for o, v in shapes:
    x = fb.synth.make_inputs(o, v, naux=32)
    eng.upload_conv(o, v, x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
    runs = [eng.compute(0, -1) for _ in range(5)]
    best = min(runs, key=lambda r: r[1]["kernel_ms"])
    tf = best[1]["flops"] / best[1]["kernel_ms"] / 1e9
    rec = {"o": o, "v": v, "E": best[0], "kernel_ms": round(best[1]["kernel_ms"], 3), "tflops": round(tf, 2), "frac": round(tf / peak, 4)}
    if o * v <= 24 * 114:
        import oracle
        npair = o * (o + 1) // 2
        tb, te = fb.host.pair_range_triplets(o, max(0, npair - 3), npair)
        eng.set_triplet_window(tb, te)
        e_w, _ = eng.compute(0, -1)
        eng.set_triplet_window(0, -1)
        rec["dE_window"] = e_w - oracle.pt_gemm(x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv, t_begin=tb, t_end=te)
    out.append(rec)
    print(json.dumps(rec), flush=True)
sweep = {}
if "--sweep" in sys.argv:   # DMMA issue study: can w warps per SM with `ilp` independent accumulators each saturate the FP64 tensor pipe?
    for ilp in (4, 8, 16):
        for w in (4, 8, 16):
            sweep[f"ilp{ilp}_warps{w}"] = round(eng.dmma_sweep(ilp, w), 2)
    print(json.dumps(sweep), flush=True)
if "--noskew" in sys.argv:
    eng.set_debug_flags(256)
    for o, v in shapes[:2]:
        x = fb.synth.make_inputs(o, v, naux=32)
        eng.upload_conv(o, v, x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
        best = min((eng.compute(0, -1) for _ in range(5)), key=lambda r: r[1]["kernel_ms"])
        print(json.dumps({"noskew": True, "o": o, "v": v, "kernel_ms": round(best[1]["kernel_ms"], 3)}), flush=True)
    eng.set_debug_flags(0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"peak": peak, "shapes": out, "dmma_sweep": sweep}, open("gpurun_out/gpu_quick.json", "w"), indent=1)
