"""Locate wrong items of a variant build: full-run E(T) with the production and the variant library, accumulated per block
(atomic per-block sums are not available, so the per-block energies are taken from separate launches over each block's item
range, and the full-run total is printed next to their sum).  python tools/gpu_blockdiff.py variant.so o v"""
import json, os, subprocess, sys
lib, o, v = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
code = r'''
import sys, json, os
sys.path.insert(0, os.getcwd())
import fermi_jl_b200 as fb
o, v = int(sys.argv[1]), int(sys.argv[2])
x = fb.synth.make_inputs(o, v, naux=16)
eng = fb.Engine(0)
eng.upload_conv(o, v, x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
full = [eng.compute(0, -1)[0] for _ in range(3)]
nb, nt = fb.host.num_blocks(v), fb.host.num_triplets(o)
per_block = [eng.compute(b * nt, (b + 1) * nt)[0] for b in range(nb)]
# big chunks of blocks (more concurrency, more L2 pressure)
chunks = [eng.compute(b * nt, min(nb, b + 10) * nt)[0] for b in range(0, nb, 10)]
print(json.dumps({"full": full, "per_block": per_block, "chunks": chunks}))
'''
res = {}
for name, L in (("prod", None), ("var", lib)):
    env = dict(os.environ)
    if L:
        env["FERMI_PT_B200_LIB"] = os.path.abspath(L)
    r = subprocess.run([sys.executable, "-c", code, str(o), str(v)], env=env, capture_output=True, text=True)
    if r.returncode:
        print(r.stderr[-800:]); sys.exit(1)
    res[name] = json.loads(r.stdout.strip().splitlines()[-1])
p, q = res["prod"], res["var"]
print("full prod", p["full"], "sum blocks", sum(p["per_block"]))
print("full var ", q["full"], "sum blocks", sum(q["per_block"]))
bad = [(b, a - c) for b, (a, c) in enumerate(zip(q["per_block"], p["per_block"])) if abs(a - c) > 1e-12]
print("blocks differing when run alone:", bad[:40], len(bad))
badc = [(b, a - c) for b, (a, c) in enumerate(zip(q["chunks"], p["chunks"])) if abs(a - c) > 1e-12]
print("10-block chunks differing:", badc)
