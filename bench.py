#!/usr/bin/env python
"""bench.py -- RCCSD(T) perturbative-triples throughput of the B200 path (and the CPU reference arm).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c3|c2|c1|c5|oXvY] [--impl reference]

One "step" = one full (T) correction over the workload (all non-zero-weight triplets i>=j>=k of the configuration).
  value  : sustained FP64 TFLOP/s (algorithmic flops 12 v^3 (v+o) per triplet) with the operands already
           resident in HBM -- the fused kernel (+ final reduction), timed with CUDA events / max over ranks.
  e2e    : the same metric through the public API (RCCSDpT(ccsd, moints, B200()) -> fpt_triples_conv) from HOST
           buffers: pinned-host -> device copies, layout prep, kernel and the 8-byte result read are all inside the
           timed region.  At N>1 rank 0 uploads, the raw arrays are broadcast once over NCCL, every rank preps and
           computes its contiguous shard of the static work list, and E(T) is one scalar all-reduce.
  roofline / cpu_baseline / clocks: see DESIGN.md "Measurement".
With --impl reference the CPU restatement of the reference algorithm (oracle/, OpenMP, all host cores) is timed on a
bounded triplet sample of the same workload (the reference itself is Julia and cannot run in this image).
"""
from __future__ import annotations

import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    "c1": ("c1_h2o_dz", 5, 19, 32),
    "c2": ("c2_h2o_tz", 5, 53, 48),
    "c3": ("c3_benzene_dz_df", 15, 93, 420),
    "c4": ("c4_h2o6_dz", 24, 114, 64),
    "c5": ("c5_synth_o40_v400", 40, 400, 64),
}
METRIC = "RCCSD(T) sustained FP64 throughput (algorithmic flops 12*v^3*(v+o) per triplet / wall time)"
UNIT = "TFLOP/s"


def parse_workload(name):
    if name in WORKLOADS:
        return WORKLOADS[name]
    m = re.fullmatch(r"o(\d+)v(\d+)", name)
    if not m:
        raise SystemExit(f"unknown workload {name}")
    return (name, int(m.group(1)), int(m.group(2)), 64)


def algorithmic_flops(o, v, ntrip):
    return 12.0 * v ** 3 * (v + o) * ntrip


def n_triplets(o):
    return o * (o + 1) * (o + 2) // 6 - o


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, f"/tmp/fpt_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.remove(self.path)
        except OSError:
            pass
        if sm:
            busy = [x for x in sm if x > 0.5 * max(sm)] or sm
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_reference_sample(x, o, v, budget_s=15.0):
    """Time the CPU restatement (oracle.pt_gemm, OpenMP over triplets, one single-threaded dgemm per contraction) on a bounded
    sample of trailing (i,j) pairs.  The GEMMs run on the faster of the oracle's own kernel and the OpenBLAS bundled with scipy
    (a short calibration on the last pair decides); the other one's calibration rate is reported beside it."""
    import oracle
    import fermi_jl_b200 as fb
    # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1 to every rank)
    threads = max(oracle.num_threads(), len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    npair = o * (o + 1) // 2
    args = (x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)

    def count(tb, te):   # zero-weight i=j=k triplets inside the range do no work
        ntr = 0
        i = j = k = 0
        for t in range(te):
            if t >= tb and not (i == j == k):
                ntr += 1
            k += 1
            if k > j:
                k = 0; j += 1
                if j > i:
                    j = 0; i += 1
        return ntr

    # calibration: the last max(threads, 8) triplets with each BLAS
    ntot = o * (o + 1) * (o + 2) // 6
    cb = max(0, ntot - max(threads, 8) - 1)
    rates = {}
    for which in ("own", "openblas"):
        if oracle.use_blas(which) != which:
            continue
        t0 = time.perf_counter()
        oracle.pt_gemm(*args, t_begin=cb, t_end=ntot, nthreads=threads)
        rates[which] = algorithmic_flops(o, v, count(cb, ntot)) / (time.perf_counter() - t0)
    best = max(rates, key=rates.get)
    oracle.use_blas(best)
    per_trip = algorithmic_flops(o, v, 1) / rates[best]
    want = max(1, int(budget_s / max(per_trip, 1e-9)))
    pr0 = npair
    while pr0 > 0:
        tb, te = fb.host.pair_range_triplets(o, pr0 - 1, npair)
        pr0 -= 1
        if te - tb >= want:
            break
    tb, te = fb.host.pair_range_triplets(o, pr0, npair)
    t0 = time.perf_counter()
    e = oracle.pt_gemm(*args, t_begin=tb, t_end=te, nthreads=threads)
    dt = time.perf_counter() - t0
    oracle.use_blas("own")
    ntr = count(tb, te)
    other = {k: {"value": r / 1e12, "unit": UNIT, "sample": f"calibration on the last {ntot - cb} triplets"} for k, r in rates.items() if k != best}
    return {"E": e, "seconds": dt, "triplets": ntr, "flops": algorithmic_flops(o, v, ntr), "threads": threads,
            "triplet_range": (tb, te), "blas": best, "other": other}


def workload_text(name, o, v, naux, route):
    integrals = "conventional integrals" if route == "conv" else f"density-fitted integrals, naux={naux}, B factors handed over, (ia|bd) etc. assembled on the GPU"
    return f"{name} (o={o}, v={v}, {integrals}, synthetic symmetric inputs, seed 20240517)"


def config_dict(name, o, v, naux, route, world):
    """The `config` object, identical in both arms (the driver compares them): what is computed, not how fast."""
    import fermi_jl_b200 as fb
    return {"workload": workload_text(name, o, v, naux, route),
            "o": o, "v": v, "triplets": n_triplets(o), "work_items": fb.host.num_items(o, v), "route": route,
            "parallelism": f"static contiguous, cost-weighted shards of the block-major (tile triple, triplet) work list over {world} GPU(s)",
            "e2e_inputs": ("pageable host arrays; " + ("the symmetry-unique halves of OVVV, T2, OVOV cross PCIe; " if route == "conv" else
                                                       "B factors and the a <= b half of T2 cross PCIe, the (ov|vv) slabs are assembled on the GPUs; ")
                           + "sharded H2D + ncclAllGather + scalar ncclAllReduce inside the library (no torch.distributed on the data path)"),
            "l2": "operands (P layout %.0f MB) exceed the 126 MB L2; no explicit flush" % (o * ((v + 3) // 4 * 4) ** 2 * ((v + o + 15) // 16 * 16) * 8 / 1e6)}


def run_reference(args, name, o, v, naux, route):
    import fermi_jl_b200 as fb
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    x = fb.synth.make_inputs(o, v, naux=naux)
    vals, secs = [], []
    s = None
    for it in range(args.warmup + args.steps):
        s = cpu_reference_sample(x, o, v, budget_s=args.cpu_budget)
        if it >= args.warmup:
            vals.append(s["flops"] / s["seconds"] / 1e12)
            secs.append(s["seconds"])
    val = statistics.mean(vals)
    sample = f"triplets [{s['triplet_range'][0]},{s['triplet_range'][1]}) of the i>=j>=k list ({s['triplets']} non-zero-weight) per step"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(secs),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(name, o, v, naux, route, args.gpus),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": s["threads"], "kind": "port", "blas": s["blas"], "other_blas": s["other"], "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference = CPU restatement of Fermi.jl ijk2.jl (oracle/pt_oracle.c: OpenMP over triplets, single-threaded dgemm per contraction "
                    "from the faster of its own kernel and scipy's OpenBLAS, as ijk.jl:45-46 arranges it); Julia is not available in this image"}
    _emit(line)


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL's version banner, torchrun) write to file descriptor 1; the contract is ONE JSON line on stdout.
    Everything but that line is sent to stderr: fd 1 is pointed at fd 2 until `_emit`."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def _emit(line: dict):
    text = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, text)
    else:
        os.write(_REAL_STDOUT, text)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--route", default=None, choices=["conv", "df"], help="conventional MO integrals or density-fitted B factors (default: df for c3, conv otherwise)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    name, o, v, naux = parse_workload(args.workload)
    route = args.route or ("df" if args.workload == "c3" else "conv")
    if args.impl == "reference":
        return run_reference(args, name, o, v, naux, route)

    import torch
    import fermi_jl_b200 as fb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if "NCCL_DEBUG" not in os.environ:
            os.environ["NCCL_DEBUG"] = "WARN"    # NCCL's default prints a version banner on stdout; stdout carries one JSON line
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")   # host-side barrier: an NCCL barrier parks a spinning kernel on every waiting GPU
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # The engine.  N = 1: a plain handle.  N > 1: one process per GPU, every process holds a *rank handle* (fpt_create_rank ->
    # ncclCommInitRank); torch.distributed is the control plane only (it carries the 128-byte NCCL id here, and barriers / the
    # max-over-ranks of the timings below); every byte of the data path -- sharded H2D, all-gather, all-reduce -- is the library's.
    if world == 1:
        eng = fb.Engine(local)
    else:
        box = [fb.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        eng = fb.Engine(local, rank=rank, world=world, nccl_id=box[0])

    # ---- synthetic inputs (same seed on every rank): ordinary pageable numpy arrays, column-major like the reference's ----
    x = fb.synth.make_inputs(o, v, naux=naux)
    names = ("T1", "T2", "OVVV", "OOOV", "OVOV", "fo", "fv") if route == "conv" else ("T1", "T2", "BOO", "BOV", "BVV", "fo", "fv")
    harr = {k: np.asfortranarray(getattr(x, k), dtype=np.float64) for k in names}
    h2d_bytes = sum(a.size * 8 for a in harr.values())

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(val):
        if dist is None:
            return val
        t = torch.tensor([val], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- leg 1: operands resident in HBM (collective at N > 1: every rank computes its shard, one scalar all-reduce) ----
    if route == "conv":
        eng.upload_conv(o, v, *[harr[k] for k in names])
    else:
        eng.upload_df(o, v, naux, *[harr[k] for k in names])
    n_items = eng.num_items()
    ntrip = n_triplets(o)
    flops = algorithmic_flops(o, v, ntrip)
    for _ in range(args.warmup):
        eng.compute(0, -1)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    kernel_ms, launches, e_gpu = 0.0, 0, 0.0
    for _ in range(args.steps):
        e_gpu, st = eng.compute(0, -1)     # synchronous: returns after E(T) is on the host
        kernel_ms += st["kernel_ms"]
        launches += 2                       # triples_kernel + reduce_partials on this rank's GPU
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    step_ms = max_over_ranks(wall_ms / args.steps)
    kern_ms = max_over_ranks(kernel_ms / args.steps)
    value = flops / (step_ms * 1e-3) / 1e12

    # ---- leg 2: end to end through the public API, from pageable host arrays (what a Julia ccall passes) ----
    ccsd = fb.RCCSD(0.0, 0.0, 0.0, harr["T1"], harr["T2"])
    if route == "conv":
        moints = fb.IntegralHelper({"OVVV": harr["OVVV"], "OOOV": harr["OOOV"], "OVOV": harr["OVOV"], "Fii": harr["fo"], "Faa": harr["fv"]})
    else:   # a DF helper holds only the B factors (DFERI.jl:15-69); the 4-index blocks are assembled on the GPU
        moints = fb.IntegralHelper({"BOO": harr["BOO"], "BOV": harr["BOV"], "BVV": harr["BVV"], "Fii": harr["fo"], "Faa": harr["fv"]}, eri_type="RIFIT")

    last_stats = {}

    def e2e_step(engine):
        res = fb.RCCSDpT(ccsd, moints, fb.B200(), engine=engine)
        last_stats.update(res.stats)
        return res.correction

    def time_e2e(engine, collective):
        for _ in range(3):
            e2e_step(engine)
        if collective:
            barrier()
        t0 = time.perf_counter()
        e = 0.0
        tl = []
        for _ in range(args.steps):
            e = e2e_step(engine)
            tl.append(engine.last_timeline())
        if collective:
            barrier()
        ms = (time.perf_counter() - t0) * 1e3 / args.steps
        br = {k: statistics.mean(t[k] for t in tl) for k in tl[0]}
        return e, ms, br

    e_e2e, e2e_ms, breakdown = time_e2e(eng, True)
    e2e_ms = max_over_ranks(e2e_ms)
    # bytes that crossed PCIe in one end-to-end step, as counted by the library (every rank pulls its own share: sum over ranks);
    # host arrays of 4 MB and more cross as their symmetry-unique halves, so this is below the size of the arrays (h2d_bytes)
    moved = float(last_stats.get("h2d_bytes", 0.0))
    if dist is not None:
        t = torch.tensor([moved], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        moved = float(t.item())
    e2e_value = flops / (e2e_ms * 1e-3) / 1e12

    # ---- leg 3 (N > 1): the same call on ONE process driving all N GPUs (fpt_create(ngpu = N), what the Julia glue uses);
    # the other ranks wait at the barrier ----
    e2e_handle = None
    if world > 1:
        barrier()
        dist.barrier(group=cpu_group)
        if rank == 0:
            try:
                engh = fb.Engine(list(range(world)))
                e_h, ms_h, br_h = time_e2e(engh, False)
                engh.close()
                e2e_handle = {"value": flops / (ms_h * 1e-3) / 1e12, "unit": UNIT, "ms_per_step": ms_h, "E_T": e_h, "breakdown_ms": br_h,
                              "what": f"fpt_triples_conv on a single-process handle over {world} GPUs (fpt_create(ngpu={world})), pageable host inputs"}
            except fb.FermiException as ex:
                e2e_handle = {"error": str(ex)}
        dist.barrier(group=cpu_group)    # the other ranks wait here on the CPU, their GPUs idle
        barrier()

    if rank == 0:
        # ---- roofline denominator: FP64 tensor-pipe peak measured live (MEASURED_PEAKS.json has no FP64 entry) ----
        peak = fb.Engine(local).fp64_peak(0, 300.0) if world > 1 else eng.fp64_peak(0, 300.0)
        achieved = flops / world / (kern_ms * 1e-3) / 1e12   # per GPU: this rank's share of the flops / its kernel time
        traffic = None
        tfile = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tfile):
            traffic = json.load(open(tfile)).get(name)
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": UNIT, "frac": achieved / peak, "traffic": traffic,
                "kernel": "fpt::triples_kernel (fused W build + permutation + V + denominators + energy)",
                "peak_source": "live DMMA.8x8x4 register-resident stream on this GPU (fpt_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
                "kernel_ms_per_launch": kern_ms}
        cpu = None
        if not args.no_cpu_baseline and world == 1:   # the CPU baseline is a rank-0, N = 1 measurement
            s = cpu_reference_sample(x, o, v, budget_s=args.cpu_budget)
            eng.set_triplet_window(*s["triplet_range"])   # the same triplets on the GPU
            e_s, _ = eng.compute(0, -1)
            eng.set_triplet_window(0, -1)
            cpu = {"value": s["flops"] / s["seconds"] / 1e12, "unit": UNIT, "cores": s["threads"], "kind": "port", "blas": s["blas"],
                   "sample": f"triplets [{s['triplet_range'][0]},{s['triplet_range'][1]}) of the i>=j>=k list "
                             f"({s['triplets']} non-zero-weight), {s['seconds']:.1f} s of oracle/pt_oracle.c (OpenMP, {s['blas']} dgemm)",
                   "other_blas": s.get("other"),
                   "dE_gpu_minus_cpu_on_sample_Eh": e_s - s["E"]}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": config_dict(name, o, v, naux, route, world),
                "triplets_per_s": ntrip / (step_ms * 1e-3), "E_T": e_gpu, "E_T_e2e": e_e2e,
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(moved), "host_array_bytes": h2d_bytes, "d2h_bytes_per_step": 8,
                        "breakdown_ms": breakdown},
                "e2e_handle": e2e_handle,
                "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "clocks": clocks}
        _emit(line)
    if dist is not None:
        dist.barrier()
        eng.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
