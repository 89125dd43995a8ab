"""Compare E(T) of two builds of the library over a list of shapes: python tools/gpu_cmp.py libA.so libB.so o v [o v ...]"""
import json, os, subprocess, sys
libs = sys.argv[1:3]
shapes = sys.argv[3:]
code = r'''
import sys, json, os
sys.path.insert(0, os.getcwd())
import fermi_jl_b200 as fb
pos=[int(a) for a in sys.argv[1:]]
eng=fb.Engine(0)
out=[]
for o,v in zip(pos[0::2],pos[1::2]):
    x=fb.synth.make_inputs(o,v,naux=16)
    eng.upload_conv(o,v,x.T1,x.T2,x.OVVV,x.OOOV,x.OVOV,x.fo,x.fv)
    es=[eng.compute(0,-1)[0] for _ in range(3)]
    out.append((o,v,es))
print(json.dumps(out))
'''
res = {}
for L in libs:
    env = dict(os.environ, FERMI_PT_B200_LIB=os.path.abspath(L))
    r = subprocess.run([sys.executable, "-c", code] + shapes, env=env, capture_output=True, text=True)
    res[L] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else r.stderr[-500:]
a, b = res[libs[0]], res[libs[1]]
if isinstance(a, str) or isinstance(b, str):
    print(a, b)
else:
    for (o, v, ea), (_, _, eb) in zip(a, b):
        print(o, v, "A:", ea, "B:", eb, "maxdiff", max(abs(x - y) for x in ea for y in eb))
