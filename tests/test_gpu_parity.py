"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on identical seeded
inputs.  Gate: |E(T)_gpu - E(T)_oracle| < 1e-9 Eh (north star), with |E(T)| ~ 1e-2..1e-1 Eh by construction."""
import numpy as np
import pytest

import fermi_jl_b200 as fb
import oracle

TOL = 1e-9  # Eh, FP64 (BASELINE.json north_star)
pytestmark = pytest.mark.gpu


def _args(x):
    return (x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)


def _expected_h2d(x, halves=True):
    """Bytes a conventional upload from pageable host arrays moves: arrays of 4 MB and more cross PCIe as their symmetry-unique half."""
    o, v = x.o, x.v
    tri = v * (v + 1) // 2
    n = (x.T1.size + x.OOOV.size + x.fo.size + x.fv.size) * 8
    for full, half in ((x.T2.size, o * o * tri), (x.OVOV.size, o * o * tri), (x.OVVV.size, o * v * tri)):
        n += 8 * (half if halves and full * 8 >= (4 << 20) else full)
    return n


@pytest.mark.parametrize("o", [1, 2, 3, 5])
@pytest.mark.parametrize("v", [1, 2, 7, 16, 19, 33, 53])
def test_shape_sweep_conv_and_df(engine, o, v):
    x = fb.synth.make_inputs(o, v, naux=9, seed=100 + 7 * o + v)
    ref = oracle.pt_gemm(*_args(x)) if o * v > 40 else oracle.pt_naive(*_args(x))
    e, st = engine.triples_conv(o, v, *_args(x))
    assert abs(e - ref) < TOL, (e, ref)
    e2, _ = engine.triples_df(o, v, x.naux, x.T1, x.T2, x.BOO, x.BOV, x.BVV, x.fo, x.fv)
    assert abs(e2 - ref) < TOL, (e2, ref)


def test_c1_h2o_dz_shape(engine):
    x = fb.synth.make_inputs(5, 19, naux=32)
    ref = oracle.pt_naive(*_args(x))
    e, _ = engine.triples_conv(5, 19, *_args(x))
    assert abs(e - ref) < TOL, (e, ref)


def test_c2_h2o_tz_shape(engine):
    x = fb.synth.make_inputs(5, 53, naux=48)
    ref = oracle.pt_gemm(*_args(x))
    e, _ = engine.triples_conv(5, 53, *_args(x))
    assert abs(e - ref) < TOL, (e, ref)


def test_c3_benzene_df_shape(engine):
    o, v, naux = 15, 93, 420
    x = fb.synth.make_inputs(o, v, naux=naux)
    ref = oracle.pt_gemm(*_args(x))
    e_df, _ = engine.triples_df(o, v, naux, x.T1, x.T2, x.BOO, x.BOV, x.BVV, x.fo, x.fv)
    e_cv, _ = engine.triples_conv(o, v, *_args(x))
    assert abs(e_df - ref) < TOL, (e_df, ref)
    assert abs(e_cv - ref) < TOL, (e_cv, ref)


def test_c4_h2o6_shape_partial_and_sharded(engine):
    """(H2O)6 shape: the oracle checks the triplets of i = 21, 22, 23 -- 829 of the 2600 positions of the reference's list, 32 % -- in one
    window; the full run is checked through size-independent properties: shards add up to the whole, and the DF route reproduces
    the conventional one."""
    o, v = 24, 114
    x = fb.synth.make_inputs(o, v, naux=64)
    engine.upload_conv(o, v, *_args(x))
    n = engine.num_items()
    npair = o * (o + 1) // 2
    tb, te = fb.host.pair_range_triplets(o, 21 * 22 // 2, npair)
    assert te - tb == 829 and te == 2600
    engine.set_triplet_window(tb, te)
    part, _ = engine.compute(0, -1)
    ref = oracle.pt_gemm(*_args(x), t_begin=tb, t_end=te)
    assert abs(part - ref) < TOL, (part, ref)
    engine.set_triplet_window(0, -1)
    assert engine.num_items() == n == fb.host.num_items(o, v)
    full, st = engine.compute(0, -1)
    ranges = [engine.shard_items(r, 8) for r in range(8)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n and all(ranges[r][1] == ranges[r + 1][0] for r in range(7))
    shards = [engine.compute(*rg)[0] for rg in ranges]
    assert abs(sum(shards) - full) < 1e-11, (sum(shards), full)
    engine.set_item_order(0)                      # triplet-major order: same items, same energy
    full0, _ = engine.compute(0, -1)
    shards0 = [engine.compute(*engine.shard_items(r, 3))[0] for r in range(3)]
    engine.set_item_order(1)
    assert abs(full0 - full) < 1e-11 and abs(sum(shards0) - full) < 1e-11
    e_df, _ = engine.triples_df(o, v, 64, x.T1, x.T2, x.BOO, x.BOV, x.BVV, x.fo, x.fv)
    assert abs(e_df - full) < TOL, (e_df, full)


def test_device_pointer_inputs(engine):
    import torch
    x = fb.synth.make_inputs(3, 21, naux=8, seed=5)
    ref = oracle.pt_gemm(*_args(x))
    dev = [torch.from_numpy(np.ascontiguousarray(a.ravel(order="F"))).cuda() for a in _args(x)]
    e, _ = engine.triples_conv(3, 21, *dev)
    assert abs(e - ref) < TOL


def test_interface_mirror(engine):
    x = fb.synth.make_inputs(3, 10, naux=8, seed=9)
    ccsd = fb.RCCSD(0.0, -0.2, -76.2, x.T1, x.T2)
    moints = fb.IntegralHelper({"OVVV": x.OVVV, "OOOV": x.OOOV, "OVOV": x.OVOV, "Fii": x.fo, "Faa": x.fv})
    res = fb.RCCSDpT(ccsd, moints)  # algorithm appended from Options["pt_alg"], PerturbativeTriples.jl:51-53
    ref = oracle.pt_naive(*_args(x))
    assert abs(res.correction - ref) < TOL
    assert abs(res.energy - (ref + ccsd.energy)) < TOL
    with pytest.raises(fb.FermiException):
        fb.RCCSDpT(ccsd, "not an integral helper", fb.B200())


def test_single_process_multi_gpu_handle(built):
    """ngpu > 1 handle (what the Julia glue uses): sharded H2D + ncclAllGather of the raw arrays, prep on every GPU, sharded
    compute + one scalar all-reduce must give the single-GPU answer -- conventional (arrays above and below the 1 MB sharding
    threshold, several OVVV chunks), DF (p-sliced assembly + broadcasts), device-resident inputs, and the AO route."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    eng = fb.Engine(list(range(min(n, 8))))
    for o, v, naux, seed in [(6, 37, 10, 11), (9, 70, 40, 12), (3, 130, 24, 13)]:
        x = fb.synth.make_inputs(o, v, naux=naux, seed=seed)
        ref = oracle.pt_gemm(*_args(x))
        e, st = eng.triples_conv(o, v, *_args(x))
        assert abs(e - ref) < TOL, (o, v, e, ref)
        e2, _ = eng.triples_df(o, v, x.naux, x.T1, x.T2, x.BOO, x.BOV, x.BVV, x.fo, x.fv)
        assert abs(e2 - ref) < TOL, (o, v, e2, ref)
        dev = [torch.from_numpy(np.ascontiguousarray(a.ravel(order="F"))).cuda(0) for a in _args(x)]
        e3, _ = eng.triples_conv(o, v, *dev)
        assert abs(e3 - ref) < TOL, (o, v, e3, ref)
        eng.set_df_ring(2)                       # slab ring on every GPU, each takes its shard of every block triple's launch
        e4, _ = eng.triples_df(o, v, x.naux, x.T1, x.T2, x.BOO, x.BOV, x.BVV, x.fo, x.fv)
        eng.set_df_ring(0)
        assert abs(e4 - ref) < TOL, (o, v, e4, ref)
        f32 = [np.asfortranarray(a.astype(np.float32)) for a in _args(x)]
        e5, _ = eng.triples_conv_f32(o, v, *f32)  # widened on GPU 0, then the device-input route (broadcast to the peers)
        assert abs(e5 - oracle.pt_gemm(*[np.asfortranarray(a.astype(np.float64)) for a in f32])) < TOL
        # the 8(f) routines: the ladder splits the range of a over the GPUs, MP2 runs on the first one
        from oracle import cc_numpy as C
        new0 = np.asfortranarray(0.01 * np.random.default_rng(seed).standard_normal((o, o, v, v)))
        want = C.ladder_df(new0.copy(order="F"), x.T1, x.T2, x.BVV)
        got = new0.copy(order="F")
        eng.ccsd_ladder_df(o, v, x.naux, x.T1, x.T2, x.BVV, got)
        assert np.abs(got - want).max() < 1e-12 * max(1.0, float(np.abs(want - new0).max()))
        e_mp2, _ = eng.mp2_df(o, v, x.naux, x.BOV, x.fo, x.fv)
        assert abs(e_mp2 - C.mp2_df(x.BOV, x.fo, x.fv)) < 1e-12 * max(1.0, abs(e_mp2))
    eng.close()


def _rank_worker(rank, world, idfile, out):
    import os, sys, time
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import numpy as np
    import fermi_jl_b200 as fb
    if rank == 0:
        with open(idfile + ".tmp", "wb") as fh:
            fh.write(fb.nccl_unique_id())
        os.replace(idfile + ".tmp", idfile)
    while not os.path.exists(idfile):
        time.sleep(0.01)
    eng = fb.Engine(rank, rank=rank, world=world, nccl_id=open(idfile, "rb").read())
    res = []
    for o, v, naux, seed in [(6, 37, 10, 11), (9, 70, 40, 12)]:
        x = fb.synth.make_inputs(o, v, naux=naux, seed=seed)
        a = (x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
        e, _ = eng.triples_conv(o, v, *a)                      # collective: sharded H2D, all-gather, shard, all-reduce
        e2, _ = eng.triples_df(o, v, x.naux, x.T1, x.T2, x.BOO, x.BOV, x.BVV, x.fo, x.fv)
        eng.upload_conv(o, v, *a)
        e3, _ = eng.compute(0, -1)
        eng.set_df_ring(2)
        e4, _ = eng.triples_df(o, v, x.naux, x.T1, x.T2, x.BOO, x.BOV, x.BVV, x.fo, x.fv)   # slab ring, collective
        eng.set_df_ring(0)
        # repeated calls of one shape: the adaptive balance moves the shard boundaries after every second call, identically on every rank
        # (a ring evaluation leaves nothing resident: upload again first)
        eng.upload_conv(o, v, *a)
        rep = [eng.compute(0, -1)[0] for _ in range(7)] + [eng.triples_conv(o, v, *a)[0] for _ in range(5)]
        res.append((e, e2, e3, e4) + tuple(rep))
    eng.close()
    out.put((rank, res))


def test_one_process_per_gpu_rank_handles(built):
    """fpt_create_rank (ncclCommInitRank): two processes, one GPU each; every rank must obtain the full E(T) of the oracle."""
    import os, tempfile
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    idfile = os.path.join(tempfile.mkdtemp(), "nccl_id")
    procs = [ctx.Process(target=_rank_worker, args=(r, 2, idfile, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for n, (o, v, naux, seed) in enumerate([(6, 37, 10, 11), (9, 70, 40, 12)]):
        x = fb.synth.make_inputs(o, v, naux=naux, seed=seed)
        ref = oracle.pt_gemm(*_args(x))
        for r in range(2):
            for e in got[r][n]:
                assert abs(e - ref) < TOL, (r, o, v, e, ref)


def test_pageable_pinned_and_async_calls(engine):
    """Inputs large enough to go through the pinned bounce ring (OVVV 23 MB > the 4 MB slots): pageable numpy arrays, pinned
    torch tensors and the asynchronous form must agree to the last bit of the item-order noise; the asynchronous call has
    consumed its inputs when it returns (they are overwritten before fpt_wait)."""
    import torch
    o, v = 7, 74
    x = fb.synth.make_inputs(o, v, naux=20, seed=31)
    ref = oracle.pt_gemm(*_args(x))
    e_pg, st = engine.triples_conv(o, v, *_args(x))
    assert abs(e_pg - ref) < TOL, (e_pg, ref)
    assert st["h2d_bytes"] == _expected_h2d(x)        # OVVV (22.7 MB) crosses PCIe as its b <= c half
    pinned = [torch.from_numpy(np.ascontiguousarray(a.ravel(order="F"))).pin_memory() for a in _args(x)]
    e_pin, _ = engine.triples_conv(o, v, *pinned)
    assert abs(e_pin - e_pg) < 1e-13
    for threads in (1, 3):
        engine.set_host_threads(threads)
        e_t, _ = engine.triples_conv(o, v, *_args(x))
        assert abs(e_t - e_pg) < 1e-13
    scratch = [np.array(a, order="F", copy=True) for a in _args(x)]
    engine.triples_conv_async(o, v, *scratch)
    for a in scratch:
        a[...] = np.nan                       # inputs were consumed
    with pytest.raises(fb.FermiException):
        engine.compute(0, -1)                 # one call in flight: collect it first
    e_as, st = engine.wait()
    assert abs(e_as - e_pg) < 1e-13 and st["kernel_ms"] > 0
    engine.triples_df_async(o, v, x.naux, x.T1, x.T2, x.BOO, x.BOV, x.BVV, x.fo, x.fv)
    e_df, _ = engine.wait()
    assert abs(e_df - ref) < TOL
    with pytest.raises(fb.FermiException):
        engine.wait()                         # nothing in flight
    tl = engine.last_timeline()
    assert 0 < tl["operands_ready_ms"] <= tl["kernel_begin_ms"] <= tl["kernel_end_ms"]


def test_current_device_is_restored_and_foreign_device_pointers_are_rejected(engine):
    import torch
    x = fb.synth.make_inputs(3, 21, naux=8, seed=5)
    if torch.cuda.device_count() >= 2:
        torch.cuda.set_device(1)
        e, _ = engine.triples_conv(3, 21, *_args(x))          # engine lives on GPU 0
        assert torch.cuda.current_device() == 1
        assert abs(e - oracle.pt_gemm(*_args(x))) < TOL
        other = [torch.from_numpy(np.ascontiguousarray(a.ravel(order="F"))).cuda(1) for a in _args(x)]
        with pytest.raises(fb.FermiException, match="lives on GPU 1"):
            engine.triples_conv(3, 21, *other)
        torch.cuda.set_device(0)
    # inputs produced on a side stream without any caller-side synchronisation: the library orders itself behind them
    side = torch.cuda.Stream()
    host = [torch.from_numpy(np.ascontiguousarray(a.ravel(order="F"))).pin_memory() for a in _args(x)]
    with torch.cuda.stream(side):
        torch.cuda._sleep(20_000_000)
        dev = [t.to("cuda:0", non_blocking=True) for t in host]
    e, _ = engine.triples_conv(3, 21, *dev)
    assert abs(e - oracle.pt_gemm(*_args(x))) < TOL


def test_sparse_list_with_out_of_range_index_is_an_error(engine):
    from oracle import pt_numpy as PN
    AO, C, T1, T2, fo, fv = fb.synth.make_ao_inputs(9, 3, 0, 0, seed=6)
    idx, vals = PN.sparse_from_dense(AO, threshold=0.0)
    Co, Cv = np.asfortranarray(C[:, :3]), np.asfortranarray(C[:, 3:])
    with pytest.raises(fb.FermiException, match="zero-based"):
        engine.triples_ao_sparse(9, 3, 6, T1, T2, (idx + 1).astype(np.int16), vals, Co, Cv, fo, fv)   # a one-based list
    e, _ = engine.triples_ao_sparse(9, 3, 6, T1, T2, idx.astype(np.int16), vals, Co, Cv, fo, fv)
    assert np.isfinite(e)


def test_gemm_kernel_shapes(engine):
    """K3/K5 GEMM on awkward shapes through the DF route: naux odd (8-byte copies) and even (16-byte copies), M, N not multiples of
    the 128 x 128 tile, o*v below one tile."""
    for o, v, naux in [(2, 5, 3), (4, 37, 64), (5, 50, 131), (3, 66, 258)]:
        x = fb.synth.make_inputs(o, v, naux=naux, seed=naux)
        ref = oracle.pt_gemm(*_args(x))
        e, _ = engine.triples_df(o, v, naux, x.T1, x.T2, x.BOO, x.BOV, x.BVV, x.fo, x.fv)
        assert abs(e - ref) < TOL, (o, v, naux, e, ref)


@pytest.mark.parametrize("o,v", [(7, 5), (6, 6), (12, 20), (9, 31), (4, 44), (3, 68)])
def test_more_shapes(engine, o, v):
    """occupied > virtual, square, many (i,j) pairs, 12+8 edge split (v=44 -> vp=44: 16,16,12; v=68 -> vp=68: 16x3,12,8)."""
    x = fb.synth.make_inputs(o, v, naux=11, seed=1000 + o * v)
    ref = oracle.pt_gemm(*_args(x))
    e, _ = engine.triples_conv(o, v, *_args(x))
    assert abs(e - ref) < TOL, (e, ref)


def test_handle_reuse_across_shapes_and_layouts(engine):
    """gradient_findif-style reuse: the same handle serves many calls (FiniteDifferences.jl:48-74 makes 6N of them);
    C-ordered / non-contiguous numpy inputs are accepted (the wrapper hands over column-major copies)."""
    for o, v, seed in [(3, 9, 1), (5, 23, 2), (2, 17, 3), (5, 23, 2)]:
        x = fb.synth.make_inputs(o, v, naux=7, seed=seed)
        ref = oracle.pt_gemm(*_args(x))
        c_ordered = [np.ascontiguousarray(a) for a in _args(x)]
        e, _ = engine.triples_conv(o, v, *c_ordered)
        assert abs(e - ref) < TOL
    big = np.zeros((4, 4, 8, 8))
    x = fb.synth.make_inputs(2, 4, naux=5, seed=9)
    big[::2, ::2, ::2, ::2] = x.T2
    view = big[::2, ::2, ::2, ::2]
    e, _ = engine.triples_conv(2, 4, x.T1, view, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
    assert abs(e - oracle.pt_naive(*_args(x))) < TOL


def test_random_item_ranges_add_up(engine):
    o, v = 6, 29
    x = fb.synth.make_inputs(o, v, naux=9, seed=77)
    engine.upload_conv(o, v, *_args(x))
    n = engine.num_items()
    full, _ = engine.compute(0, -1)
    rng = np.random.default_rng(5)
    cuts = sorted(set([0, n] + list(rng.integers(0, n, 7))))
    parts = [engine.compute(int(cuts[t]), int(cuts[t + 1]))[0] for t in range(len(cuts) - 1)]
    assert abs(sum(parts) - full) < 1e-13
    assert abs(full - oracle.pt_gemm(*_args(x))) < TOL
    assert engine.compute(5, 5)[0] == 0.0


def test_error_paths(engine):
    x = fb.synth.make_inputs(2, 3, naux=3, seed=1)
    with pytest.raises(fb.FermiException):
        engine.triples_conv(0, 3, *_args(x))
    with pytest.raises(fb.FermiException):
        engine.triples_df(2, 3, 0, x.T1, x.T2, x.BOO, x.BOV, x.BVV, x.fo, x.fv)
    eng2 = fb.Engine(0)
    with pytest.raises(fb.FermiException):
        eng2.compute(0, -1)          # nothing uploaded
    with pytest.raises(fb.FermiException):
        fb.Engine(99)                # no such device
    eng2.close()


def test_df_odd_aux_sizes(engine):
    for naux in (1, 5, 37):
        x = fb.synth.make_inputs(3, 13, naux=naux, seed=naux)
        ref = oracle.pt_gemm(*_args(x))
        e, _ = engine.triples_df(3, 13, naux, x.T1, x.T2, x.BOO, x.BOV, x.BVV, x.fo, x.fv)
        assert abs(e - ref) < TOL, (naux, e, ref)


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("o,v", [(3, 20), (5, 53), (4, 44), (2, 7), (6, 33)])
def test_kernel_variants_agree(engine, o, v, variant):
    """Both kernel variants (1: RMW by the DMMA warps between k-loops; 2: epilogue warps fed through TMEM parking,
    fpt_triples2.cuh) must give the oracle's E(T)."""
    x = fb.synth.make_inputs(o, v, naux=16, seed=11)
    ref = oracle.pt_gemm(*_args(x))
    engine.upload_conv(o, v, *_args(x))
    default = engine.kernel_variant
    if variant == 2:
        try:
            engine.set_kernel_variant(2)
        except fb.FermiException:
            pytest.skip("the experimental epilogue-warp kernel is only in builds with -DFPT_WITH_VARIANT2")
    try:
        engine.set_kernel_variant(variant)
        e, _ = engine.compute(0, -1)
        e_b, _ = engine.compute(0, -1)          # repeated launch on the same handle: TMEM is allocated and released per launch
    finally:
        engine.set_kernel_variant(default)
    assert abs(e - ref) < TOL and abs(e_b - e) < 1e-13, (e, e_b, ref)   # CTAs pull items dynamically: last-bit differences


@pytest.mark.parametrize("nbf,ndocc,drop_occ,drop_vir", [(9, 3, 0, 0), (14, 4, 1, 2), (30, 6, 1, 0), (47, 5, 0, 3)])
def test_ao_route_matches_oracle(engine, nbf, ndocc, drop_occ, drop_vir):
    """fpt_triples_ao (GPU quarter transforms) against the oracle fed with the einsum transcription of
    Chonky.jl:28-114, including frozen-core / dropped-virtual slices and AO dimensions that are not multiples of 4."""
    from oracle import pt_numpy as PN
    AO, C, T1, T2, fo, fv = fb.synth.make_ao_inputs(nbf, ndocc, drop_occ, drop_vir, seed=5)
    OVVV, OOOV, OVOV = PN.mo_blocks_from_ao(AO, C, ndocc, drop_occ, drop_vir)
    ref = oracle.pt_gemm(T1, T2, OVVV, OOOV, OVOV, fo, fv)
    o, v = T1.shape
    Co, Cv = np.asfortranarray(C[:, drop_occ:ndocc]), np.asfortranarray(C[:, ndocc:nbf - drop_vir])
    e, st = engine.triples_ao(nbf, o, v, T1, T2, AO, Co, Cv, fo, fv)
    assert abs(e - ref) < TOL, (e, ref)
    # through the interface mirror: an IntegralHelper without cached MO blocks but with the AO helper and orbitals
    ao = fb.IntegralHelper({"ERI": AO})
    moints = fb.IntegralHelper({"Fii": fo, "Faa": fv}, aoints=ao, C=C, ndocc=ndocc, drop_occ=drop_occ, drop_vir=drop_vir)
    res = fb.RCCSDpT(fb.RCCSD(0.0, 0.0, -1.0, T1, T2), moints, fb.B200())
    assert abs(res.correction - ref) < TOL
    with pytest.raises(fb.FermiException):
        engine.triples_ao(nbf, o, nbf, T1, T2, AO, Co, Cv, fo, fv)      # o + v > nbf


def test_full_size_invariances(engine):
    """Size-independent properties at the full (H2O)6 shape (o=24, v=114), where the oracle only covers a window:
    E(T) is invariant under a relabelling of the virtual orbitals and of the occupied orbitals (the tile a label falls in,
    the padding and the i>=j>=k / a>=b>=c orderings all change), and it is homogeneous of degree 2 in the amplitudes."""
    o, v = 24, 114
    x = fb.synth.make_inputs(o, v, naux=48, seed=9)
    e0, _ = engine.triples_conv(o, v, *_args(x))
    rng = np.random.default_rng(1)
    pv, po = rng.permutation(v), rng.permutation(o)
    F = np.asfortranarray
    ev, _ = engine.triples_conv(o, v, F(x.T1[:, pv]), F(x.T2[:, :, pv][:, :, :, pv]), F(x.OVVV[:, pv][:, :, pv][:, :, :, pv]),
                                F(x.OOOV[:, :, :, pv]), F(x.OVOV[:, pv][:, :, :, pv]), x.fo, x.fv[pv].copy())
    assert abs(ev - e0) < 1e-12, (ev, e0)
    eo, _ = engine.triples_conv(o, v, F(x.T1[po]), F(x.T2[po][:, po]), F(x.OVVV[po]), F(x.OOOV[po][:, po][:, :, po]),
                                F(x.OVOV[po][:, :, po]), x.fo[po].copy(), x.fv)
    assert abs(eo - e0) < 1e-12, (eo, e0)
    s = 1.75
    es, _ = engine.triples_conv(o, v, F(s * x.T1), F(s * x.T2), x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
    assert abs(es - s * s * e0) < 1e-12, (es, s * s * e0)


@pytest.mark.parametrize("nbf,ndocc,drop_occ,drop_vir,thr,itype", [(9, 3, 0, 0, 0.0, np.int16), (14, 4, 1, 2, 0.0, np.int32),
                                                                    (26, 6, 1, 0, 1e-4, np.int16)])
def test_sparse_ao_route_matches_oracle(engine, nbf, ndocc, drop_occ, drop_vir, thr, itype):
    """fpt_triples_ao_sparse (sparse AO list -> dense tensor on the GPU -> quarter transforms) against the oracle fed with MO
    blocks from the dense tensor *rebuilt from the same list* (so that a screening threshold changes both sides alike)."""
    from oracle import pt_numpy as PN
    AO, C, T1, T2, fo, fv = fb.synth.make_ao_inputs(nbf, ndocc, drop_occ, drop_vir, seed=6)
    idx, vals = PN.sparse_from_dense(AO, threshold=thr)
    AOs = np.zeros_like(AO)
    for (m, n, r, s), V in zip(idx, vals):
        for a, b, c, d in ((m, n, r, s), (n, m, r, s), (m, n, s, r), (n, m, s, r), (r, s, m, n), (s, r, m, n), (r, s, n, m), (s, r, n, m)):
            AOs[a, b, c, d] = V
    OVVV, OOOV, OVOV = PN.mo_blocks_from_ao(AOs, C, ndocc, drop_occ, drop_vir)
    ref = oracle.pt_gemm(T1, T2, OVVV, OOOV, OVOV, fo, fv)
    o, v = T1.shape
    Co, Cv = np.asfortranarray(C[:, drop_occ:ndocc]), np.asfortranarray(C[:, ndocc:nbf - drop_vir])
    e, st = engine.triples_ao_sparse(nbf, o, v, T1, T2, idx.astype(itype), vals, Co, Cv, fo, fv)
    assert abs(e - ref) < TOL, (e, ref)
    ao = fb.IntegralHelper({"ERI": fb.FermiSparse(idx, vals)}, eri_type="SparseERI")
    moints = fb.IntegralHelper({"Fii": fo, "Faa": fv}, aoints=ao, C=C, ndocc=ndocc, drop_occ=drop_occ, drop_vir=drop_vir)
    res = fb.RCCSDpT(fb.RCCSD(0.0, 0.0, -1.0, T1, T2), moints, fb.B200())
    assert abs(res.correction - ref) < TOL


def test_single_precision_inputs_are_widened(engine):
    """`@set precision single` hands Float32 arrays (IntegralHelper.jl:58-68); the path widens them and computes in FP64:
    E(T) must equal the oracle evaluated on the same (rounded) values."""
    o, v = 4, 23
    x = fb.synth.make_inputs(o, v, naux=12, seed=3)
    f32 = [np.asfortranarray(a.astype(np.float32)) for a in _args(x)]
    ref = oracle.pt_gemm(*[np.asfortranarray(a.astype(np.float64)) for a in f32])
    moints = fb.IntegralHelper({"OVVV": f32[2], "OOOV": f32[3], "OVOV": f32[4], "Fii": f32[5], "Faa": f32[6]})
    res = fb.RCCSDpT(fb.RCCSD(0.0, 0.0, -1.0, f32[0], f32[1]), moints, fb.B200())
    assert abs(res.correction - ref) < TOL
    assert abs(res.correction - oracle.pt_gemm(*_args(x))) < 1e-6 * abs(ref) * 100   # and close to the FP64 problem
    assert res.stats["h2d_bytes"] == sum(a.size * 4 for a in f32)                    # the arrays crossed PCIe in 4-byte form
    # the Float32 entry points directly, conventional and density-fitted, at a size that goes through the pinned ring
    o, v, naux = 9, 70, 40
    x = fb.synth.make_inputs(o, v, naux=naux, seed=5)
    r = lambda a: np.asfortranarray(a.astype(np.float32))
    w = lambda a: np.asfortranarray(a.astype(np.float64))
    c32 = [r(a) for a in _args(x)]
    e32, st = engine.triples_conv_f32(o, v, *c32)
    assert abs(e32 - oracle.pt_gemm(*[w(a) for a in c32])) < TOL
    assert st["h2d_bytes"] == sum(a.size * 4 for a in c32)
    d32 = [r(a) for a in (x.T1, x.T2, x.BOO, x.BOV, x.BVV, x.fo, x.fv)]
    e32d, _ = engine.triples_df_f32(o, v, naux, *d32)
    B = [w(a) for a in d32]
    ref_df = oracle.pt_gemm(B[0], B[1], np.asfortranarray(np.einsum("Qia,Qbc->iabc", B[3], B[4], optimize=True)),
                            np.asfortranarray(np.einsum("Qij,Qka->ijka", B[2], B[3], optimize=True)),
                            np.asfortranarray(np.einsum("Qia,Qjb->iajb", B[3], B[3], optimize=True)), B[5], B[6])
    assert abs(e32d - ref_df) < TOL
    with pytest.raises(fb.FermiException):
        engine.triples_conv_f32(0, v, *c32)


@pytest.mark.parametrize("o,v", [(3, 127), (4, 126), (6, 124)])
def test_nine_group_k_all_distinct_tile_triples(engine, o, v):
    """K = v + o in (128, 144] -> nine kappa-groups, vp = 128 -> eight full tiles: the shape on which a ring refill overtaking
    the consumers' shared-memory reads (missing cross-proxy fence, see producer_loop) showed up first."""
    x = fb.synth.make_inputs(o, v, naux=16)
    ref = oracle.pt_gemm(*_args(x))
    for _ in range(3):
        e, _ = engine.triples_conv(o, v, *_args(x))
        assert abs(e - ref) < TOL, (e, ref)


def test_randomised_sweep(engine):
    """tools/gpu_fuzz.py: random shapes x routes (conventional / DF / AO / sparse AO) x item orders x kernel variants x
    triplet windows x shard counts against the oracle (40 cases here; profiles/r02c_gpu_fuzz_250_7.json holds a 250-case run of tools/gpu_fuzz.py)."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gpu_fuzz.py"), "40", "11"], capture_output=True, text=True, cwd=root)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]


@pytest.mark.parametrize("o,v", [(8, 30), (9, 70), (13, 41), (24, 57)])
def test_split_call_matches_single_launch(engine, o, v):
    """fpt_triples_conv with o >= 8 runs as a split call (operands of the occupied indices p < pA first, kernel over the triplets with
    i < pA, the rest of OVVV staged behind it, second kernel): same E(T) as the oracle and as the unsplit call, with two, three and
    four phases, from pageable and from pinned host arrays (the latter take the cudaMemcpy2DAsync row gather)."""
    import os
    import torch
    x = fb.synth.make_inputs(o, v, naux=12, seed=40 + o)
    ref = oracle.pt_gemm(*_args(x))
    e_split, st = engine.triples_conv(o, v, *_args(x))
    assert st["n_launches"] >= 9          # two fused kernels + two reductions among them
    res = {}
    for nph in ("0", "3", "4"):
        os.environ["FERMI_PT_B200_SPLIT"] = nph
        try:
            res[nph] = engine.triples_conv(o, v, *_args(x))
        finally:
            del os.environ["FERMI_PT_B200_SPLIT"]
    e_whole, st0 = res["0"]
    assert abs(e_split - ref) < TOL and abs(e_whole - ref) < TOL, (e_split, e_whole, ref)
    assert abs(e_split - e_whole) < 1e-13 and abs(res["3"][0] - e_whole) < 1e-13 and abs(res["4"][0] - e_whole) < 1e-13
    assert res["3"][1]["n_launches"] >= st["n_launches"] > st0["n_launches"]      # (o = 8 has no room for a third multiple-of-4 slice)
    assert st["h2d_bytes"] == st0["h2d_bytes"] == res["4"][1]["h2d_bytes"] == _expected_h2d(x)
    pinned = [torch.from_numpy(np.ascontiguousarray(a.ravel(order="F"))).pin_memory() for a in _args(x)]
    e_pin, stp = engine.triples_conv(o, v, *pinned)
    assert abs(e_pin - e_split) < 1e-13
    assert stp["h2d_bytes"] == _expected_h2d(x, halves=False)      # pinned arrays are read by the DMA engines, in full


@pytest.mark.parametrize("o,v", [(15, 93), (9, 70), (3, 127)])
def test_symmetry_unique_halves(engine, o, v):
    """Pageable host arrays of 4 MB and more cross PCIe as the half their index symmetry leaves free (b <= c of OVVV, a <= b of T2 and
    OVOV), the mirror images are written on the GPU: same E(T) as the full upload and as the oracle, about half the bytes.  An array
    that is not symmetric is detected by the spot check and read in full."""
    x = fb.synth.make_inputs(o, v, naux=16, seed=7 + o)
    ref = oracle.pt_gemm(*_args(x))
    e_half, st = engine.triples_conv(o, v, *_args(x))
    engine.set_symmetric_inputs(False)
    try:
        e_full, st_full = engine.triples_conv(o, v, *_args(x))
    finally:
        engine.set_symmetric_inputs(True)
    assert abs(e_half - ref) < TOL and abs(e_full - ref) < TOL, (e_half, e_full, ref)
    assert abs(e_half - e_full) < 1e-13
    assert st["h2d_bytes"] == _expected_h2d(x) and st_full["h2d_bytes"] == _expected_h2d(x, halves=False)
    assert st["h2d_bytes"] < 0.62 * st_full["h2d_bytes"]
    # upload + compute (the staged form) takes the same route
    engine.upload_conv(o, v, *_args(x))
    e_staged, _ = engine.compute(0, -1)
    assert abs(e_staged - e_half) < 1e-13
    # break the (b, c) symmetry of OVVV: the spot check must notice, the array goes in full, and what is computed is what the full
    # upload computes from the same array
    y = fb.synth.make_inputs(o, v, naux=16, seed=7 + o)
    y.OVVV = np.asfortranarray(y.OVVV * (1.0 + 0.25 * np.arange(v)[None, None, None, :] / v))
    e_asym, st_asym = engine.triples_conv(o, v, *_args(y))
    engine.set_symmetric_inputs(False)
    try:
        e_asym_full, _ = engine.triples_conv(o, v, *_args(y))
    finally:
        engine.set_symmetric_inputs(True)
    assert abs(e_asym - e_asym_full) < 1e-13 and abs(e_asym - e_half) > 1e-6
    assert st_asym["h2d_bytes"] == _expected_h2d(y) + 8 * (y.OVVV.size - o * v * (v * (v + 1) // 2))


@pytest.mark.parametrize("o,v,naux,blocks", [(5, 19, 9, (1, 2, 4)), (9, 31, 37, (1, 4)), (13, 41, 64, (2, 5)), (15, 93, 420, (4,))])
def test_df_slab_ring(built, o, v, naux, blocks):
    """DF route with only 3 * block occupied slabs resident (re-assembled from the B factors on the fly, block triple by block triple,
    explicit triplet lists + slot maps): same E(T) as the materialised route and as the oracle; the handle's device memory shrinks
    accordingly.  Fresh handles, so that the buffer pool shows what each mode needs."""
    x = fb.synth.make_inputs(o, v, naux=naux, seed=3 * o + v)
    ref = oracle.pt_gemm(*_args(x))
    df = (o, v, naux, x.T1, x.T2, x.BOO, x.BOV, x.BVV, x.fo, x.fv)
    full = fb.Engine(0)
    full.set_df_ring(-1)
    e_full, _ = full.triples_df(*df)
    bytes_full = full.device_bytes()
    full.close()
    assert abs(e_full - ref) < TOL
    vp, Kp = (v + 3) // 4 * 4, (v + o + 15) // 16 * 16
    slab = vp * vp * Kp * 8
    for ob in blocks:
        ring = fb.Engine(0)
        ring.set_df_ring(ob)
        e_ring, st = ring.triples_df(*df)
        assert abs(e_ring - ref) < TOL, (ob, e_ring, ref)
        assert abs(e_ring - e_full) < 1e-12
        assert st["n_items"] == fb.host.num_items(o, v)
        if 3 * ob < o:
            # (+ the ring's triplet lists and slot maps)
            assert ring.device_bytes() <= bytes_full - (o - 3 * ob) * slab + 65536, (ob, ring.device_bytes(), bytes_full)
        ring.triples_df_async(*df)               # the asynchronous form takes the same route
        e_async, _ = ring.wait()
        assert abs(e_async - e_ring) < 1e-13
        with pytest.raises(fb.FermiException):
            ring.compute(0, -1)                  # nothing is left resident after a ring evaluation
        ring.close()


def test_deterministic_mode_is_bitwise_reproducible(engine):
    """fpt_set_deterministic: items are dealt statically to the CTAs, so repeated evaluations return the same bits (the default dynamic
    counter only promises ~1e-16 relative); the value itself agrees with the dynamic mode to the usual noise."""
    o, v = 6, 70
    x = fb.synth.make_inputs(o, v, naux=12, seed=77)
    engine.upload_conv(o, v, *_args(x))
    e_dyn, _ = engine.compute(0, -1)
    engine.set_deterministic(True)
    try:
        runs = [engine.compute(0, -1)[0] for _ in range(6)]
        one_call = [engine.triples_conv(o, v, *_args(x))[0] for _ in range(3)]
    finally:
        engine.set_deterministic(False)
    assert len(set(runs)) == 1 and len(set(one_call)) == 1, (runs, one_call)
    assert abs(runs[0] - e_dyn) < 1e-13 and abs(one_call[0] - e_dyn) < 1e-13
    assert abs(runs[0] - oracle.pt_gemm(*_args(x))) < TOL
