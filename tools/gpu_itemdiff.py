"""Per-item E(T) of the production and a variant library: python tools/gpu_itemdiff.py variant.so o v"""
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
n = eng.num_items()
print(json.dumps([eng.compute(i, i + 1)[0] for i in range(n)]))
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
nt = o * (o + 1) * (o + 2) // 6 - o
trips = [(i, j, k) for i in range(o) for j in range(i + 1) for k in range(j + 1) if not (i == j == k)]
bad = {}
for it, (a, b) in enumerate(zip(res["var"], res["prod"])):
    if abs(a - b) > 1e-13:
        blk, u = divmod(it, nt)
        bad.setdefault(trips[u], []).append(blk)
def dec(b):
    A = 0
    while (A + 1) * (A + 2) * (A + 3) // 6 <= b: A += 1
    b -= A * (A + 1) * (A + 2) // 6
    B = 0
    while (B + 1) * (B + 2) // 2 <= b: B += 1
    return (A, B, b - B * (B + 1) // 2)
print("items", len(res["prod"]), "single-item launches differing:", sum(len(v) for v in bad.values()))
for t, blks in bad.items():
    print(t, len(blks), [dec(b) for b in blks[:30]])
