"""Synthetic o=40 / v=400 (T) (BASELINE config 5) on 1..8 GPUs.

    python tools/run_c5.py [--shape 40x400] [--steps 1] [--host] [--check 402] [--tag name]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/run_c5.py ...        (N > 1)

One process per GPU with rank handles (fpt_create_rank): torch.distributed only carries the NCCL id and the barriers.
Inputs:
  default   the 20.5 GB OVVV block is generated on every rank's own GPU (torch / cuBLAS is plumbing for the *synthetic input*
            only) and handed over as device pointers -- `value` is then the resident-operand number and there is no e2e;
  --host    rank 0 builds OVVV / OOOV / OVOV slab by slab into files under /dev/shm (or /tmp), every rank maps them: ordinary
            pageable host memory, as a Julia caller would hold it; `e2e` = fpt_triples_conv from those arrays (sharded H2D +
            all-gather + prep + kernel + all-reduce).
  --check n (needs --host, N = 1): E(T) of three windows of n/3 triplets -- the first, the middle and the last of the
            reference's i >= j >= k list -- against the CPU oracle.
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fermi_jl_b200 as fb

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="40x400", help="occupied x virtual (one option: torchrun's own parser claims short prefixes such as --v)")
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--host", action="store_true")
ap.add_argument("--check", type=int, default=0)
ap.add_argument("--tag", default=None)
args = ap.parse_args()
o, v = (int(t) for t in args.shape.split("x"))
naux = 64
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
    cpu_group = dist.new_group(backend="gloo")
    box = [fb.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    eng = fb.Engine(local, rank=rank, world=world, nccl_id=box[0])
else:
    eng = fb.Engine(local)


def cpu_barrier():
    if dist is not None:
        dist.barrier(group=cpu_group)


def max_over_ranks(x):
    if dist is None:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


x = fb.synth.make_inputs(o, v, naux=naux, conventional=False)
ntrip = o * (o + 1) * (o + 2) // 6 - o
flops = 12.0 * v ** 3 * (v + o) * ntrip
res = {"o": o, "v": v, "n_gpus": world, "triplets": ntrip, "flops": flops, "inputs": "host (pageable, memory-mapped)" if args.host else "device-generated"}
t_gen = time.time()
if args.host:
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.statvfs("/dev/shm").f_bavail * os.statvfs("/dev/shm").f_frsize > 30e9 else "/tmp"
    paths = {k: os.path.join(base, f"fpt_c5_{k}_{o}_{v}.bin") for k in ("OVVV", "OOOV", "OVOV")}
    shapes = {"OVVV": (o, v, v, v), "OOOV": (o, o, o, v), "OVOV": (o, v, o, v)}
    if rank == 0:
        BOV, BVV, BOO = x.BOV, x.BVV, x.BOO
        # column-major [i,a,b,c] == C-order [c][b][a][i]; one slab of 8 c's at a time:  out[(c,b),(a,i)] = sum_Q BVV[Q,b,c] BOV[Q,i,a]
        mm = np.memmap(paths["OVVV"], dtype=np.float64, mode="w+", shape=(v, v, v * o))
        R = np.ascontiguousarray(BOV.transpose(0, 2, 1).reshape(naux, v * o))           # [Q][(a,i)]
        for c0 in range(0, v, 8):
            L = np.ascontiguousarray(BVV[:, :, c0:c0 + 8].transpose(2, 1, 0).reshape(-1, naux))   # [(c,b)][Q]
            mm[c0:c0 + 8] = (L @ R).reshape(-1, v, v * o)
        mm.flush(); del mm
        np.asfortranarray(np.einsum("Qij,Qka->ijka", BOO, BOV, optimize=True)).ravel(order="F").tofile(paths["OOOV"])
        np.asfortranarray(np.einsum("Qia,Qjb->iajb", BOV, BOV, optimize=True)).ravel(order="F").tofile(paths["OVOV"])
    cpu_barrier()
    harr = {k: np.memmap(paths[k], dtype=np.float64, mode="r").reshape(shapes[k], order="F") for k in paths}
    inputs = (x.T1, x.T2, harr["OVVV"], harr["OOOV"], harr["OVOV"], x.fo, x.fv)
else:
    f64 = torch.float64
    BOV = torch.from_numpy(np.ascontiguousarray(x.BOV)).to(dev)   # (Q,i,a)
    BVV = torch.from_numpy(np.ascontiguousarray(x.BVV)).to(dev)   # (Q,a,b)
    BOO = torch.from_numpy(np.ascontiguousarray(x.BOO)).to(dev)
    OVVV = torch.empty((v, v, v, o), dtype=f64, device=dev)        # C-order [c][b][a][i] == column-major [i,a,b,c]
    for c0 in range(0, v, 16):
        OVVV[c0:c0 + 16] = torch.einsum("Qia,Qbc->cbai", BOV, BVV[:, :, c0:c0 + 16])
    OOOV = torch.einsum("Qij,Qka->akji", BOO, BOV).contiguous()
    OVOV = torch.einsum("Qia,Qjb->bjai", BOV, BOV).contiguous()
    T1 = torch.from_numpy(np.ascontiguousarray(x.T1.ravel(order="F"))).to(dev)
    T2 = torch.from_numpy(np.ascontiguousarray(x.T2.ravel(order="F"))).to(dev)
    fo = torch.from_numpy(x.fo.copy()).to(dev); fv = torch.from_numpy(x.fv.copy()).to(dev)
    torch.cuda.synchronize()
    inputs = (T1, T2, OVVV, OOOV, OVOV, fo, fv)
res["generate_s"] = time.time() - t_gen
cpu_barrier()

# ---- resident operands: upload once, time `steps` computes ----
t0 = time.time()
eng.upload_conv(o, v, *inputs)
res["upload_s"] = max_over_ranks(time.time() - t0)
res["work_items"] = eng.num_items()
kern, walls = [], []
for _ in range(args.steps):
    cpu_barrier()
    t0 = time.perf_counter()
    e, st = eng.compute(0, -1)
    walls.append(max_over_ranks(time.perf_counter() - t0))
    kern.append(max_over_ranks(st["kernel_ms"]))
wall = min(walls)
res.update({"E_T": e, "kernel_ms_max_over_ranks": min(kern), "step_s": wall, "value_tflops": flops / wall / 1e12,
            "triplets_per_s": ntrip / wall, "kernel_tflops_per_gpu": flops / world / (min(kern) * 1e-3) / 1e12})
if rank == 0:
    peak = (fb.Engine(local) if world > 1 else eng).fp64_peak(0, 300.0)
    res["fp64_dmma_peak_tflops"] = peak
    res["roofline_frac"] = res["kernel_tflops_per_gpu"] / peak
    print(json.dumps(res), flush=True)

# ---- end to end from the host arrays ----
if args.host:
    e2e = []
    for _ in range(max(1, args.steps)):
        cpu_barrier()
        t0 = time.perf_counter()
        e2, st2 = eng.triples_conv(o, v, *inputs)
        e2e.append(max_over_ranks(time.perf_counter() - t0))
        tl = eng.last_timeline()
    res["e2e"] = {"step_s": min(e2e), "value_tflops": flops / min(e2e) / 1e12, "E_T": e2, "h2d_bytes": st2["h2d_bytes"], "breakdown_ms": tl}

# ---- oracle check on three windows of the reference's triplet list ----
if args.check and args.host and world == 1:
    import oracle
    nfull = o * (o + 1) * (o + 2) // 6
    w = max(1, args.check // 3)
    wins = [(0, w), (nfull // 2 - w // 2, nfull // 2 - w // 2 + w), (nfull - w, nfull)]
    eng.upload_conv(o, v, *inputs)
    oracle.use_blas("openblas")
    checks = []
    for tb, te in wins:
        eng.set_triplet_window(tb, te)
        eg, _ = eng.compute(0, -1)
        t0 = time.time()
        ec = oracle.pt_gemm(x.T1, x.T2, harr["OVVV"], harr["OOOV"], harr["OVOV"], x.fo, x.fv, t_begin=tb, t_end=te)
        checks.append({"triplets": [tb, te], "E_gpu": eg, "E_oracle": ec, "dE": eg - ec, "oracle_s": time.time() - t0})
        print(json.dumps(checks[-1]), flush=True)
    oracle.use_blas("own")
    eng.set_triplet_window(0, -1)
    res["oracle_check"] = {"windows": checks, "worst_abs_dE": max(abs(c["dE"]) for c in checks), "oracle_threads": oracle.num_threads(),
                           "triplets_checked": sum(c["triplets"][1] - c["triplets"][0] for c in checks)}

if rank == 0:
    print(json.dumps(res), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    tag = args.tag or f"c5_o{o}_v{v}_n{world}"
    json.dump(res, open(f"gpurun_out/{tag}.json", "w"), indent=1)
    if args.host:
        for p in paths.values():
            try:
                os.remove(p)
            except OSError:
                pass
cpu_barrier()
if dist is not None:
    eng.close()
    dist.destroy_process_group()
