"""Generates tests/golden/pt_golden.json: E(T) of small seeded inputs computed three independent ways on the CPU
(the line-by-line numpy transcription of ijk.jl, the explicit-GEMM form of ijk2.jl, and -- for the smallest shapes --
the spin-orbital brute force).  Run from the repo root:  python tests/golden/make_golden.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import fermi_jl_b200 as fb  # noqa: E402
from oracle import pt_numpy as P  # noqa: E402

CASES = [(1, 1, 3, 11), (1, 4, 3, 12), (2, 3, 5, 13), (3, 4, 6, 14), (3, 5, 7, 15), (2, 7, 5, 16), (4, 9, 8, 17),
         (5, 19, 16, 18), (3, 21, 9, 19), (6, 17, 12, 20)]
out = []
for o, v, naux, seed in CASES:
    x = fb.synth.make_inputs(o, v, naux=naux, seed=seed)
    a = (x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
    rec = {"o": o, "v": v, "naux": naux, "seed": seed, "E_ijk": P.pt_ijk(*a), "E_ijk2": P.pt_ijk2(*a)}
    if o <= 3 and v <= 5:
        rec["E_spinorbital"] = P.pt_spinorbital_bruteforce(*a)
    out.append(rec)
    print(rec)
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "pt_golden.json"), "w") as fh:
    json.dump(out, fh, indent=1)
