"""One (T) evaluation at a given shape for compute-sanitizer runs on variant builds (FERMI_PT_B200_LIB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fermi_jl_b200 as fb
o, v = int(sys.argv[1]), int(sys.argv[2])
x = fb.synth.make_inputs(o, v, naux=16)
eng = fb.Engine(0)
eng.upload_conv(o, v, x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
n = eng.num_items()
ie = int(sys.argv[3]) if len(sys.argv) > 3 else -1
print("E", eng.compute(0, ie)[0], "items", n)
