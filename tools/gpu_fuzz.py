"""Randomised parity sweep on the GPU: random (o, v) shapes, routes (conventional / DF / AO / sparse AO), item orders, kernel
variants, triplet windows and shard counts, each compared with the CPU oracle.  python tools/gpu_fuzz.py [n_cases] [seed]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fermi_jl_b200 as fb
import oracle
from oracle import pt_numpy as PN

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed0)
eng = fb.Engine(0)
try:                                   # the experimental epilogue-warp kernel is only in -DFPT_WITH_VARIANT2 builds
    eng.set_kernel_variant(2)
    variants = [1, 1, 2]
except fb.FermiException:
    variants = [1]
eng.set_kernel_variant(1)
worst, fails = 0.0, []
t0 = time.time()
for case in range(n_cases):
    route = rng.choice(["conv", "conv", "df", "ao", "sparse"])
    o = int(rng.integers(1, 9))
    v = int(rng.integers(1, 150)) if route in ("conv", "df") else int(rng.integers(2, 30))
    seed = int(rng.integers(1 << 30))
    order = int(rng.integers(0, 2))
    variant = int(rng.choice(variants))
    eng.set_kernel_variant(variant)
    eng.set_item_order(order)
    desc = {"case": case, "route": route, "o": o, "v": v, "order": order, "variant": variant}
    if route in ("conv", "df"):
        x = fb.synth.make_inputs(o, v, naux=int(rng.integers(1, 40)), seed=seed)
        args = (x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
        if route == "df":
            e, _ = eng.triples_df(o, v, x.naux, x.T1, x.T2, x.BOO, x.BOV, x.BVV, x.fo, x.fv)
            ref = oracle.pt_gemm(*args)
        else:
            eng.upload_conv(o, v, *args)
            nfull = o * (o + 1) * (o + 2) // 6
            tb = int(rng.integers(0, nfull)); te = int(rng.integers(tb, nfull + 1))
            if rng.random() < 0.5:
                tb, te = 0, nfull
            eng.set_triplet_window(tb, te)
            world = int(rng.integers(1, 6))
            e = sum(eng.compute(*eng.shard_items(r, world))[0] for r in range(world))
            ref = oracle.pt_gemm(*args, t_begin=tb, t_end=te)
            desc.update({"window": [tb, te], "world": world})
    else:
        ndocc = o + int(rng.integers(0, 3)); dv = int(rng.integers(0, 3))
        nbf = ndocc + v + dv
        dc = ndocc - o
        AO, C, T1, T2, fo, fv = fb.synth.make_ao_inputs(nbf, ndocc, dc, dv, seed=seed)
        Co, Cv = np.asfortranarray(C[:, dc:ndocc]), np.asfortranarray(C[:, ndocc:nbf - dv])
        OVVV, OOOV, OVOV = PN.mo_blocks_from_ao(AO, C, ndocc, dc, dv)
        ref = oracle.pt_gemm(T1, T2, OVVV, OOOV, OVOV, fo, fv)
        desc["nbf"] = nbf
        if route == "ao":
            e, _ = eng.triples_ao(nbf, o, v, T1, T2, AO, Co, Cv, fo, fv)
        else:
            idx, vals = PN.sparse_from_dense(AO)
            e, _ = eng.triples_ao_sparse(nbf, o, v, T1, T2, idx, vals, Co, Cv, fo, fv)
    d = abs(e - ref)
    worst = max(worst, d)
    if not d < 1e-9:
        fails.append(dict(desc, e=e, ref=ref))
        print("FAIL", desc, e, ref, flush=True)
eng.set_kernel_variant(1); eng.set_item_order(1)
res = {"cases": n_cases, "seed": seed0, "worst_abs_dE": worst, "fails": fails, "seconds": time.time() - t0, "library": fb.load_library().fpt_version().decode()}
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
# one file per (cases, seed): a later, smaller run (the 40-case pytest) cannot overwrite the record of a bigger one
json.dump(res, open(f"gpurun_out/gpu_fuzz_{n_cases}_{seed0}.json", "w"), indent=1)
sys.exit(1 if fails else 0)
