"""Run the fused kernel once (or n times) at a given shape -- target for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fermi_jl_b200 as fb
o, v = int(sys.argv[1]), int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1
x = fb.synth.make_inputs(o, v, naux=32)
eng = fb.Engine(0)
eng.upload_conv(o, v, x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
for _ in range(n):
    e, st = eng.compute(0, -1)
print(e, st["kernel_ms"], st["flops"] / st["kernel_ms"] / 1e9)
