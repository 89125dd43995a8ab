"""Kernel time per tile triple (block-major work list: the items of one block are contiguous) -> where a shape loses time:
python tools/gpu_blocktime.py [o v]   -> gpurun_out/gpu_blocktime_o<o>v<v>.json  (one record per block: tiles, sizes, ms, flop)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fermi_jl_b200 as fb

args = [int(a) for a in sys.argv[1:] if a.isdigit()]
o, v = (args + [24, 114])[:2] if len(args) >= 2 else (24, 114)
eng = fb.Engine(0)
peak = eng.fp64_peak(0, 200.0)
x = fb.synth.make_inputs(o, v, naux=32)
eng.upload_conv(o, v, x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
n = eng.num_items()
vp = (v + 3) // 4 * 4
nt = (vp + 15) // 16
split = (vp & 15) == 4 and vp > 16
def tsize(t):
    if split:
        return 16 if t < nt - 2 else (12 if t == nt - 2 else 8)
    return min(16, vp - 16 * t)
nb = nt * (nt + 1) * (nt + 2) // 6
per = n // nb
full = min((eng.compute(0, -1)[1]["kernel_ms"] for _ in range(3)))
recs = []
b = 0
for A in range(nt):
    for B in range(A + 1):
        for C in range(B + 1):
            ms = min(eng.compute(b * per, (b + 1) * per)[1]["kernel_ms"] for _ in range(2))
            recs.append({"block": b, "tiles": [A, B, C], "sizes": [tsize(A), tsize(B), tsize(C)], "ms": ms})
            b += 1
out = {"o": o, "v": v, "peak": peak, "full_ms": full, "sum_ms": sum(r["ms"] for r in recs), "items_per_block": per, "blocks": recs}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/gpu_blocktime_o{o}v{v}.json", "w"))
print(json.dumps({k: out[k] for k in ("o", "v", "peak", "full_ms", "sum_ms", "items_per_block")}))
