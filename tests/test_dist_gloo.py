"""world_size-2 gloo test (CPU) of the multi-GPU host logic: every rank evaluates its contiguous shard of the static work
list (with the CPU emulator standing in for the kernel), one scalar all-reduce gives E(T), which must equal the oracle."""
import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import fermi_jl_b200 as fb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o, v = 4, 13
    # rank 0 owns the inputs; they are broadcast once (the NCCL broadcast of the GPU run)
    x = fb.synth.make_inputs(o, v, naux=6, seed=21)
    names = ("T1", "T2", "OVVV", "OOOV", "OVOV", "fo", "fv")
    bufs = {}
    for k in names:
        a = np.ascontiguousarray(getattr(x, k).ravel(order="F"))
        t = torch.from_numpy(a.copy()) if rank == 0 else torch.zeros(a.size, dtype=torch.float64)
        dist.broadcast(t, src=0)
        bufs[k] = t.numpy()
    L = ctypes.CDLL(os.path.join(ROOT, "tests", "emul", "libfpt_emul.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    L.fpt_emulate.argtypes = [ctypes.c_int, ctypes.c_int] + [dp] * 7 + [ctypes.c_int] + [ctypes.c_longlong] * 4 + [
        dp, ctypes.POINTER(ctypes.c_longlong)]
    # this rank's part of the work list: the library's own cost-weighted split (same code as fpt_shard_items)
    L.fpt_emul_shard.argtypes = [ctypes.c_int] * 5 + [ctypes.POINTER(ctypes.c_longlong)] * 2 + [dp]
    sb, se, share = ctypes.c_longlong(), ctypes.c_longlong(), ctypes.c_double()
    assert L.fpt_emul_shard(o, v, 1, rank, world, ctypes.byref(sb), ctypes.byref(se), ctypes.byref(share)) == 0
    ib, ie = sb.value, se.value
    assert abs(share.value - 1.0 / world) < 0.1
    e, n = ctypes.c_double(), ctypes.c_longlong()
    rc = L.fpt_emulate(o, v, *[bufs[k].ctypes.data_as(dp) for k in names], 1, 0, -1, ib, ie, ctypes.byref(e), ctypes.byref(n))
    assert rc == 0
    t = torch.tensor([e.value], dtype=torch.float64)
    dist.all_reduce(t)
    if rank == 0:
        out.put(float(t.item()))
    dist.destroy_process_group()


def test_two_rank_sharded_energy(built):
    import torch.multiprocessing as mp
    import fermi_jl_b200 as fb
    import oracle
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    e = out.get(timeout=5)
    x = fb.synth.make_inputs(4, 13, naux=6, seed=21)
    ref = oracle.pt_gemm(x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
    assert abs(e - ref) < 1e-13, (e, ref)
