"""DF route timing at the C3 (benzene/cc-pVDZ fc) shape: upload_df = H2D of B factors + on-device assembly (K3) + prep."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fermi_jl_b200 as fb
o, v, naux = 15, 93, 420
x = fb.synth.make_inputs(o, v, naux=naux)
eng = fb.Engine(0)
out = {}
for rep in range(3):
    t0 = time.time(); e_df, st = eng.triples_df(o, v, naux, x.T1, x.T2, x.BOO, x.BOV, x.BVV, x.fo, x.fv); t_df = time.time() - t0
    t0 = time.time(); e_cv, st2 = eng.triples_conv(o, v, x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv); t_cv = time.time() - t0
out = {"o": o, "v": v, "naux": naux, "E_df": e_df, "E_conv": e_cv, "dE": e_df - e_cv, "df_total_ms": t_df * 1e3, "df_upload_ms": st["upload_ms"],
       "df_kernel_ms": st["kernel_ms"], "conv_total_ms": t_cv * 1e3, "conv_upload_ms": st2["upload_ms"], "h2d_df": st["h2d_bytes"], "h2d_conv": st2["h2d_bytes"],
       "df_assembly_flops": 2.0 * naux * (o * v * v * v + o * o * o * v + o * v * o * v)}
print(json.dumps(out))
json.dump(out, open("gpurun_out/df_c3.json", "w"), indent=1)
