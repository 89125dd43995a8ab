"""Small cases of the round-2 additions for compute-sanitizer (memcheck / synccheck): DF slab ring, symmetric-half and split uploads,
Float32 entry, DF-CCSD ladder, MP2, deterministic deal."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fermi_jl_b200 as fb, oracle
from oracle import cc_numpy as C
o, v, naux = 9, 41, 24
x = fb.synth.make_inputs(o, v, naux=naux, seed=4)
a = (x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
ref = oracle.pt_gemm(*a)
eng = fb.Engine(0)
e, _ = eng.triples_conv(o, v, *a); assert abs(e - ref) < 1e-9           # split call (o >= 8); arrays below 4 MB go in full
big = fb.synth.make_inputs(6, 90, naux=16, seed=5)                          # OVVV 35 MB: symmetric halves
ab = (big.T1, big.T2, big.OVVV, big.OOOV, big.OVOV, big.fo, big.fv)
e, st = eng.triples_conv(6, 90, *ab); assert abs(e - oracle.pt_gemm(*ab)) < 1e-9 and st["h2d_bytes"] < 0.7 * sum(t.size * 8 for t in ab)
for ob in (1, 4):
    eng.set_df_ring(ob)
    e, _ = eng.triples_df(o, v, naux, x.T1, x.T2, x.BOO, x.BOV, x.BVV, x.fo, x.fv); assert abs(e - ref) < 1e-9
eng.set_df_ring(0)
eng.set_deterministic(True)
e, _ = eng.triples_conv(o, v, *a); assert abs(e - ref) < 1e-9
eng.set_deterministic(False)
f32 = [np.asfortranarray(t.astype(np.float32)) for t in a]
e, _ = eng.triples_conv_f32(o, v, *f32); assert abs(e - oracle.pt_gemm(*[np.asfortranarray(t.astype(np.float64)) for t in f32])) < 1e-9
new0 = np.asfortranarray(0.01 * np.random.default_rng(1).standard_normal((o, o, v, v)))
got = new0.copy(order="F"); eng.ccsd_ladder_df(o, v, naux, x.T1, x.T2, x.BVV, got)
assert np.abs(got - C.ladder_df(new0.copy(order="F"), x.T1, x.T2, x.BVV)).max() < 1e-12
e, _ = eng.mp2_df(o, v, naux, x.BOV, x.fo, x.fv); assert abs(e - C.mp2_df(x.BOV, x.fo, x.fv)) < 1e-10 * abs(e)
print("extras ok")
