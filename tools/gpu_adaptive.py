"""Adaptive shard balance of a single-process multi-GPU handle: kernel time (max over GPUs) of repeated computes and one-call evaluations.
python tools/gpu_adaptive.py [ngpu] [o v]  -> gpurun_out/gpu_adaptive.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fermi_jl_b200 as fb
args = [int(a) for a in sys.argv[1:] if a.isdigit()]
n = args[0] if args else torch.cuda.device_count()
o, v = (args[1], args[2]) if len(args) >= 3 else (24, 114)
x = fb.synth.make_inputs(o, v, naux=32)
a = (x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
out = {"ngpu": n, "o": o, "v": v}
one = fb.Engine(0)
one.upload_conv(o, v, *a)
e1, st1 = min((one.compute(0, -1) for _ in range(3)), key=lambda r: r[1]["kernel_ms"])
out["one_gpu_kernel_ms"] = st1["kernel_ms"]
one.close()
for adaptive in (0, 1):
    eng = fb.Engine(list(range(n)))
    eng.set_adaptive_shards(adaptive)
    eng.upload_conv(o, v, *a)
    ks, es = [], []
    for _ in range(12):
        e, st = eng.compute(0, -1)
        ks.append(round(st["kernel_ms"], 3)); es.append(e)
    calls = []
    for _ in range(10):
        e, st = eng.triples_conv(o, v, *a)
        calls.append(round(st["kernel_ms"], 3)); es.append(e)
    out["adaptive" if adaptive else "static"] = {"compute_kernel_ms": ks, "one_call_kernel_ms": calls, "max_dE": max(abs(t - e1) for t in es),
                                                 "efficiency_last": st1["kernel_ms"] / (n * ks[-1])}
    eng.close()
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gpu_adaptive.json", "w"), indent=1)
