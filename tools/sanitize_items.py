"""Single work items of every kind for compute-sanitizer (racecheck / memcheck / synccheck) on the final build: o = 3, v = 48 (three
full tiles).  Triplets by position in the reference's list: 5 = (2,1,0) i > j > k, 7 = (2,2,0) i = j, 1 = (1,0,0) j = k; tile
triples by block number: 5 = (2,1,0) six W slots, 2 = (1,1,0) three slots, 0 = (0,0,0) one slot.
python tools/sanitize_items.py [case ...]   (default: all)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fermi_jl_b200 as fb, oracle
CASES = {"six_slot_ijk": (5, 5, 6), "six_slot_i_eq_j": (7, 5, 6), "six_slot_j_eq_k": (1, 5, 6), "three_slot_ijk": (5, 2, 3),
         "three_slot_i_eq_j": (7, 2, 3), "one_slot_ijk": (5, 0, 1), "three_items_in_a_row": (5, 4, 7)}
o, v = 3, 48
x = fb.synth.make_inputs(o, v, naux=6, seed=4)
a = (x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
eng = fb.Engine(0)
eng.upload_conv(o, v, *a)
for name in (sys.argv[1:] or list(CASES)):
    t, b0, b1 = CASES[name]
    eng.set_triplet_window(t, t + 1)
    e, st = eng.compute(b0, b1)
    print(name, "E", e, "items", st["n_items"], flush=True)
eng.set_triplet_window(0, -1)
e, _ = eng.compute(0, -1)
ref = oracle.pt_gemm(*a)
print("full", e, "ref", ref, "dE", e - ref)
assert abs(e - ref) < 1e-9
