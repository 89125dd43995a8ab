"""AO -> MO route timing at an (H2O)6/cc-pVDZ-like dimension: nbf=144, 6 frozen core of 30 occupied -> o=24, v=114.
The AO tensor (3.4 GB) is built on the device from a symmetric factor (torch is plumbing for the synthetic input only)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fermi_jl_b200 as fb

nbf, ndocc, dc = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (144, 30, 6)
o, v = ndocc - dc, nbf - ndocc
rng = np.random.default_rng(1)
naux = 32
ts, bs = fb.synth.default_scales(o, v, naux)
B = bs * rng.standard_normal((naux, nbf, nbf)); B = 0.5 * (B + B.transpose(0, 2, 1))
C, _ = np.linalg.qr(rng.standard_normal((nbf, nbf)))
T1 = np.asfortranarray(ts * rng.standard_normal((o, v)))
T2 = ts * rng.standard_normal((o, o, v, v)); T2 = np.asfortranarray(0.5 * (T2 + T2.transpose(1, 0, 3, 2)))
fo = -np.sort(rng.uniform(0.3, 2.0, o))[::-1].copy(); fv = np.sort(rng.uniform(0.1, 3.0, v))
dev = torch.device("cuda", 0)
Bd = torch.from_numpy(B).to(dev)
AO = torch.einsum("Qmn,Qrs->srnm", Bd, Bd).contiguous()      # C-order [s][r][n][m] == column-major [m,n,r,s]
Co = np.asfortranarray(C[:, dc:ndocc]); Cv = np.asfortranarray(C[:, ndocc:])
# MO blocks for the conventional route (reference values), via the factor
Bov = np.einsum("Qmn,mi,na->Qia", B, Co, Cv); Bvv = np.einsum("Qmn,ma,nb->Qab", B, Cv, Cv); Boo = np.einsum("Qmn,mi,nj->Qij", B, Co, Co)
F = np.asfortranarray
OVVV = F(np.einsum("Qia,Qbc->iabc", Bov, Bvv, optimize=True)); OOOV = F(np.einsum("Qij,Qka->ijka", Boo, Bov, optimize=True))
OVOV = F(np.einsum("Qia,Qjb->iajb", Bov, Bov, optimize=True))
eng = fb.Engine(0)
e_conv, st_conv = eng.triples_conv(o, v, T1, T2, OVVV, OOOV, OVOV, fo, fv)
res = {"nbf": nbf, "o": o, "v": v, "E_conv": e_conv, "conv_total_ms": st_conv["total_ms"], "conv_h2d_MB": st_conv["h2d_bytes"] / 1e6}
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e_ao, st = eng.triples_ao(nbf, o, v, T1, T2, AO, Co, Cv, fo, fv)     # AO tensor already on the device
    res.update({"E_ao_dev": e_ao, "ao_dev_total_ms": (time.perf_counter() - t0) * 1e3, "ao_dev_upload_ms": st["upload_ms"],
                "ao_dev_kernel_ms": st["kernel_ms"]})
AOh = AO.cpu().numpy().reshape(-1)      # host copy, column-major flat
del AO
for rep in range(2):
    t0 = time.perf_counter()
    e_aoh, st = eng.triples_ao(nbf, o, v, T1, T2, AOh, Co, Cv, fo, fv)   # AO tensor from (pageable) host memory
    res.update({"E_ao_host": e_aoh, "ao_host_total_ms": (time.perf_counter() - t0) * 1e3, "ao_host_h2d_GB": st["h2d_bytes"] / 1e9})
res["dE_ao_minus_conv"] = res["E_ao_dev"] - e_conv
res["quarter_flops"] = 2.0 * nbf * (nbf ** 3 * o + nbf ** 2 * o * (v + o) + nbf * o * (v * v + v * o + o * o) + o * (v ** 3 + v * o * v + o * o * v))
print(json.dumps(res), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gpu_ao.json", "w"), indent=1)
