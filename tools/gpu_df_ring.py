"""DF route: all o slabs resident vs the slab ring (fpt_set_df_ring): time, E(T), device memory.
python tools/gpu_df_ring.py [o v naux]...  -> gpurun_out/gpu_df_ring.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fermi_jl_b200 as fb

args = [int(a) for a in sys.argv[1:] if a.isdigit()]
shapes = list(zip(args[0::3], args[1::3], args[2::3])) or [(15, 93, 420), (24, 160, 600)]
out = []
for o, v, naux in shapes:
    x = fb.synth.make_inputs(o, v, naux=naux, conventional=False)
    df = (o, v, naux, x.T1, x.T2, x.BOO, x.BOV, x.BVV, x.fo, x.fv)
    rec = {"o": o, "v": v, "naux": naux, "modes": []}
    for ob in (-1, 4, 2, 1):
        eng = fb.Engine(0)
        eng.set_df_ring(ob)
        eng.triples_df(*df)
        best = min((eng.triples_df(*df) for _ in range(3)), key=lambda r: r[1]["total_ms"])
        rec["modes"].append({"ring_block": ob, "E": best[0], "total_ms": best[1]["total_ms"], "gpu_ms": best[1]["kernel_ms"], "launches": best[1]["n_launches"],
                             "device_MB": eng.device_bytes() / 1e6})
        eng.close()
    out.append(rec)
    print(json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gpu_df_ring.json", "w"), indent=1)
