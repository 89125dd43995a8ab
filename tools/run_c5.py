"""Full synthetic o=40/v=400 (T) on one GPU (BASELINE config 5).  The 20.5 GB OVVV block is generated on the device
(torch is plumbing here: cuBLAS only builds the *synthetic inputs*), handed to the library as device pointers, and the
result is checked against the CPU oracle on the triplets of the last (i,j) pair."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fermi_jl_b200 as fb

o, v, naux = (int(sys.argv[1]), int(sys.argv[2]), 64) if len(sys.argv) > 2 else (40, 400, 64)
check = "--nocheck" not in sys.argv
x = fb.synth.make_inputs(o, v, naux=naux, conventional=False)
dev = torch.device("cuda", 0)
f64 = torch.float64
BOV = torch.from_numpy(np.ascontiguousarray(x.BOV)).to(dev)   # (Q,i,a)
BVV = torch.from_numpy(np.ascontiguousarray(x.BVV)).to(dev)   # (Q,a,b)
BOO = torch.from_numpy(np.ascontiguousarray(x.BOO)).to(dev)
# column-major (i fastest) == C-order arrays with reversed axes
OVVV = torch.empty((v, v, v, o), dtype=f64, device=dev)        # [c][b][a][i]
for c0 in range(0, v, 16):
    OVVV[c0:c0 + 16] = torch.einsum("Qia,Qbc->cbai", BOV, BVV[:, :, c0:c0 + 16])
OOOV = torch.einsum("Qij,Qka->akji", BOO, BOV).contiguous()     # [a][k][j][i]
OVOV = torch.einsum("Qia,Qjb->bjai", BOV, BOV).contiguous()     # [b][j][a][i]
T1 = torch.from_numpy(np.ascontiguousarray(x.T1.ravel(order="F"))).to(dev)
T2 = torch.from_numpy(np.ascontiguousarray(x.T2.ravel(order="F"))).to(dev)
fo = torch.from_numpy(x.fo.copy()).to(dev); fv = torch.from_numpy(x.fv.copy()).to(dev)
torch.cuda.synchronize()
eng = fb.Engine(0)
t0 = time.time()
eng.upload_conv(o, v, T1, T2, OVVV, OOOV, OVOV, fo, fv)
t_up = time.time() - t0
n_items = eng.num_items()
ntrip = o * (o + 1) * (o + 2) // 6 - o
flops = 12.0 * v ** 3 * (v + o) * ntrip
t0 = time.time()
e, st = eng.compute(0, -1)
wall = time.time() - t0
res = {"o": o, "v": v, "E_T": e, "kernel_ms": st["kernel_ms"], "wall_s": wall, "prep_s": t_up, "n_items": n_items,
       "triplets": ntrip, "tflops": flops / st["kernel_ms"] / 1e9, "triplets_per_s": ntrip / (st["kernel_ms"] * 1e-3),
       "fp64_dmma_peak_tflops": eng.fp64_peak(0, 300.0)}
res["frac_of_peak"] = res["tflops"] / res["fp64_dmma_peak_tflops"]
print(json.dumps(res), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/c5_result.json", "w"), indent=1)
if check:
    import oracle
    npair = o * (o + 1) // 2
    tb, te = fb.host.pair_range_triplets(o, npair - 1, npair)
    eng.set_triplet_window(tb, te)
    e_part, _ = eng.compute(0, -1)
    eng.set_triplet_window(0, -1)
    try:
        h = {"OVVV": OVVV.cpu().numpy().reshape(-1).reshape((o, v, v, v), order="F"),
             "OOOV": OOOV.cpu().numpy().reshape(-1).reshape((o, o, o, v), order="F"),
             "OVOV": OVOV.cpu().numpy().reshape(-1).reshape((o, v, o, v), order="F")}
        del OVVV
        t0 = time.time()
        ref = oracle.pt_gemm(x.T1, x.T2, h["OVVV"], h["OOOV"], h["OVOV"], x.fo, x.fv, t_begin=tb, t_end=te)
        res.update({"check_triplets": [tb, te], "E_part_gpu": e_part, "E_part_oracle": ref, "dE": e_part - ref,
                    "oracle_s": time.time() - t0, "oracle_threads": oracle.num_threads()})
    except MemoryError as ex:
        res["check_error"] = repr(ex)
    print(json.dumps(res), flush=True)
    json.dump(res, open("gpurun_out/c5_result.json", "w"), indent=1)
