"""Small multi-item case for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fermi_jl_b200 as fb, oracle
o, v = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3, 21)
x = fb.synth.make_inputs(o, v, naux=6, seed=4)
eng = fb.Engine(0)
e, st = eng.triples_conv(o, v, x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
ref = oracle.pt_gemm(x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
print("E", e, "ref", ref, "dE", e - ref, "items", st["n_items"])
assert abs(e - ref) < 1e-9
