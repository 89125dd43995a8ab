"""Staging sweep at the C4 shape on one GPU: pinned slot size x store flavour x host threads -> end-to-end time and timeline."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fermi_jl_b200 as fb
o, v = 24, 114
x = fb.synth.make_inputs(o, v, naux=64)
page = [np.asfortranarray(getattr(x, k)) for k in ("T1", "T2", "OVVV", "OOOV", "OVOV", "fo", "fv")]
out = []
for kb in (128, 256, 512, 1024, 2048, 4096):
    for nt in ("0", "1"):
        os.environ["FERMI_PT_B200_NT"] = nt
        os.environ["FERMI_PT_B200_PIECE_KB"] = str(kb)
        eng = fb.Engine(0)
        for th in (4, 8, 16):
            eng.set_host_threads(th)
            best = None
            for rep in range(4):
                t0 = time.perf_counter()
                e, st = eng.triples_conv(o, v, *page)
                ms = (time.perf_counter() - t0) * 1e3
                tl = eng.last_timeline()
                if best is None or ms < best["total_ms"]:
                    best = {"piece_kb": kb, "nt": nt, "threads": th, "total_ms": round(ms, 2), "stage_ms": round(tl["host_stage_ms"], 2), "h2d_done_ms": round(tl["h2d_done_ms"], 2)}
            out.append(best)
            print(json.dumps(best), flush=True)
        eng.close()
json.dump(out, open("gpurun_out/gpu_stage_sweep.json", "w"), indent=1)
