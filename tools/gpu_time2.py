"""Interleaved timing of two builds: python tools/gpu_time2.py libA.so libB.so o v [o v ...] (best of 6 each, alternating processes)"""
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
    x=fb.synth.make_inputs(o,v,naux=32)
    eng.upload_conv(o,v,x.T1,x.T2,x.OVVV,x.OOOV,x.OVOV,x.fo,x.fv)
    ts=sorted(eng.compute(0,-1)[1]["kernel_ms"] for _ in range(6))
    out.append((o,v,ts[0],ts[2]))
print(json.dumps(out))
'''
best = {}
for rep in range(2):
    for L in libs:
        env = dict(os.environ, FERMI_PT_B200_LIB=os.path.abspath(L))
        r = subprocess.run([sys.executable, "-c", code] + shapes, env=env, capture_output=True, text=True)
        for o, v, t0, t2 in json.loads(r.stdout.strip().splitlines()[-1]):
            k = (L, o, v)
            best[k] = min(best.get(k, 1e30), t0)
for (L, o, v), t in sorted(best.items(), key=lambda kv: (kv[0][1], kv[0][2], kv[0][0])):
    print(o, v, os.path.basename(L), round(t, 3))
