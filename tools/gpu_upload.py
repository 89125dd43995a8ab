"""Upload-path and GEMM timing on one GPU: end-to-end fpt_triples_conv at the C4 shape from pageable and from pinned host
arrays for several staging-thread counts (with the library's own timeline of the call), the DF route at C3, and the K3/K5 GEMM
at the shapes those routes use.  python tools/gpu_upload.py [o v]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fermi_jl_b200 as fb

o, v = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (24, 114)
x = fb.synth.make_inputs(o, v, naux=64)
names = ("T1", "T2", "OVVV", "OOOV", "OVOV", "fo", "fv")
page = [np.asfortranarray(getattr(x, k)) for k in names]
pinned = [torch.from_numpy(np.ascontiguousarray(a.ravel(order="F"))).pin_memory() for a in page]
res = {"o": o, "v": v, "h2d_MB": sum(a.size * 8 for a in page) / 1e6, "cpus": len(os.sched_getaffinity(0)), "runs": []}
eng = None


def run(label, arrs, threads):
    eng.set_host_threads(threads)
    best = None
    for rep in range(4):
        t0 = time.perf_counter()
        e, st = eng.triples_conv(o, v, *arrs)
        ms = (time.perf_counter() - t0) * 1e3
        if best is None or ms < best["total_ms"]:
            best = {"label": label, "threads": threads, "total_ms": ms, "kernel_ms": st["kernel_ms"], "E": e, **eng.last_timeline()}
    res["runs"].append(best)
    print(json.dumps(best), flush=True)


for nt in ("0", "1"):       # bounce copies with ordinary / non-temporal stores (FERMI_PT_B200_NT, read when the handle is created)
    os.environ["FERMI_PT_B200_NT"] = nt
    if eng is not None:
        eng.close()
    eng = fb.Engine(0)
    for th in (1, 2, 4, 8, 16):
        run("pageable nt=" + nt, page, th)
run("pinned", pinned, 16)
# DF route at the benzene shape
xo = fb.synth.make_inputs(15, 93, naux=420)
for rep in range(3):
    t0 = time.perf_counter()
    e, st = eng.triples_df(15, 93, 420, xo.T1, xo.T2, xo.BOO, xo.BOV, xo.BVV, xo.fo, xo.fv)
    df = {"label": "df_c3", "total_ms": (time.perf_counter() - t0) * 1e3, "kernel_ms": st["kernel_ms"], "E": e, **eng.last_timeline()}
t0 = time.perf_counter()
e, st = eng.triples_conv(15, 93, xo.T1, xo.T2, xo.OVVV, xo.OOOV, xo.OVOV, xo.fo, xo.fv)
df["conv_total_ms"] = (time.perf_counter() - t0) * 1e3
df["E_conv"] = e
res["df"] = df
print(json.dumps(df), flush=True)
# the GEMM alone: DF assembly of Pt at C3 / C4-like / C5-like shapes, quarter transforms at nbf = 144
peak = eng.fp64_peak(0, 200.0)
res["fp64_peak"] = peak
res["gemm"] = []
for (M, N, K, what) in [(15 * 93, 93 * 93, 420, "DF Pt C3"), (24 * 114, 114 * 114, 512, "DF Pt C4-like"), (4 * 400, 400 * 400, 1024, "DF Pt C5-like, 4 of 40 p"),
                        (144 ** 3, 24, 144, "quarter 1 nbf=144"), (144 * 144 * 24, 114, 144, "quarter 2v"), (144 * 24 * 114, 114, 144, "quarter 3vv"),
                        (24 * 114 * 114, 114, 144, "quarter 4 OVVV"), (8192, 8192, 1024, "square")]:
    tf = eng.gemm_bench(M, N, K, 5)
    res["gemm"].append({"M": M, "N": N, "K": K, "what": what, "tflops": tf, "frac_of_dmma_peak": tf / peak})
    print(json.dumps(res["gemm"][-1]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gpu_upload.json", "w"), indent=1)
