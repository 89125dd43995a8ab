"""Item order A/B (block-major vs triplet-major) + cost-weighted shard balance:  python tools/gpu_order.py o v [o v ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fermi_jl_b200 as fb

pos = [int(a) for a in sys.argv[1:] if not a.startswith("--")]
shapes = list(zip(pos[0::2], pos[1::2])) or [(24, 114)]
reps = 1 if "--once" in sys.argv else 3
eng = fb.Engine(0)
out = []
for o, v in shapes:
    x = fb.synth.make_inputs(o, v, naux=32)
    eng.upload_conv(o, v, x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
    for order in (1, 0):
        eng.set_item_order(order)
        best = min((eng.compute(0, -1) for _ in range(reps)), key=lambda r: r[1]["kernel_ms"])
        rec = {"o": o, "v": v, "order": order, "ms": round(best[1]["kernel_ms"], 3),
               "tflops": round(best[1]["flops"] / best[1]["kernel_ms"] / 1e9, 2), "E": best[0]}
        if "--shards" in sys.argv:
            rec["shard_ms_8"] = [round(eng.compute(*eng.shard_items(r, 8))[1]["kernel_ms"], 3) for r in range(8)]
        out.append(rec)
        print(json.dumps(rec), flush=True)
    eng.set_item_order(1)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gpu_order.json", "w"), indent=1)
