"""CPU tests of the kernel's index algebra: tests/emul/emul_main.cpp runs the same fpt_layout.h code the CUDA kernel
uses (item decode, block/slot construction, GEMM descriptors, swizzled slot addressing incl. the fast RMW form, the
column-wise energy stage) with plain loops instead of tensor cores, and must reproduce the oracle."""
import ctypes
import os

import numpy as np
import pytest

import fermi_jl_b200 as fb
import oracle

_dp = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def emul(built):
    L = ctypes.CDLL(os.path.join(os.path.dirname(__file__), "emul", "libfpt_emul.so"))
    L.fpt_emulate.argtypes = [ctypes.c_int, ctypes.c_int] + [_dp] * 7 + [ctypes.c_int] + [ctypes.c_longlong] * 4 + [
        _dp, ctypes.POINTER(ctypes.c_longlong)]

    def run(x, ib=0, ie=-1, order=1, tb=0, te=-1):
        arrs = [np.asfortranarray(a) for a in (x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)]
        e, n = ctypes.c_double(), ctypes.c_longlong()
        rc = L.fpt_emulate(x.o, x.v, *[a.ctypes.data_as(_dp) for a in arrs], order, tb, te, ib, ie, ctypes.byref(e),
                           ctypes.byref(n))
        assert rc == 0, f"emulator self-check failed with code {rc}"
        return e.value, n.value

    return run


@pytest.mark.parametrize("o,v", [(1, 4), (2, 3), (3, 7), (3, 16), (2, 19), (3, 20), (2, 33), (3, 28), (2, 40), (4, 17)])
def test_emulator_matches_oracle(emul, o, v):
    x = fb.synth.make_inputs(o, v, naux=8, seed=3)
    e, n = emul(x)
    ref = oracle.pt_gemm(x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
    assert abs(e - ref) < 1e-13, (e, ref)
    assert n == fb.host.num_items(o, v)


@pytest.mark.parametrize("order", [0, 1])
def test_item_shards_add_up_and_match_triplet_ranges(emul, order):
    o, v = 4, 21
    x = fb.synth.make_inputs(o, v, naux=8, seed=5)
    full, n = emul(x, order=order)
    parts = [emul(x, *fb.host.shard_items(n, r, 3), order=order)[0] for r in range(3)]
    assert abs(sum(parts) - full) < 1e-15
    npair = o * (o + 1) // 2
    tb, te = fb.host.pair_range_triplets(o, npair - 3, npair)
    part, npart = emul(x, order=order, tb=tb, te=te)
    ref = oracle.pt_gemm(x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv, t_begin=tb, t_end=te)
    assert abs(part - ref) < 1e-14
    assert 0 < npart < n
    # a window that starts and ends inside pairs, cutting through zero-weight diagonal entries
    part2, _ = emul(x, order=order, tb=3, te=13)
    ref2 = oracle.pt_gemm(x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv, t_begin=3, t_end=13)
    assert abs(part2 - ref2) < 1e-14


@pytest.mark.parametrize("o,v,world", [(4, 21, 3), (6, 50, 8), (24, 114, 8), (5, 19, 2), (40, 400, 8)])   # the last one: C5, 33.5 M items
@pytest.mark.parametrize("order", [0, 1])
def test_cost_weighted_shards_partition_the_work_list(built, o, v, world, order):
    """fpt_shard_items' split (shard_items / block_cost in fpt_layout.h, run here on the CPU): the parts are contiguous,
    cover [0, n_items) exactly once, and carry equal shares of the estimated cost."""
    L = ctypes.CDLL(os.path.join(os.path.dirname(__file__), "emul", "libfpt_emul.so"))
    L.fpt_emul_shard.argtypes = [ctypes.c_int] * 5 + [ctypes.POINTER(ctypes.c_longlong)] * 2 + [_dp]
    n = fb.host.num_items(o, v)
    prev_end, shares = 0, []
    for r in range(world):
        b, e, sh = ctypes.c_longlong(), ctypes.c_longlong(), ctypes.c_double()
        assert L.fpt_emul_shard(o, v, order, r, world, ctypes.byref(b), ctypes.byref(e), ctypes.byref(sh)) == 0
        assert b.value == prev_end and e.value >= b.value
        prev_end = e.value
        shares.append(sh.value)
    assert prev_end == n
    assert abs(sum(shares) - 1.0) < 1e-12
    if n >= 50 * world:
        assert max(shares) < 1.05 / world, shares


def test_adaptive_rebalance_converges(built):
    """rebalance_fractions (what a multi-GPU handle applies after every second call of one shape): for a cost density the model does not
    know -- here piecewise, with a cheap middle and an expensive tail -- repeated updates from the 'measured' shard times drive the
    imbalance from 20 % to below 0.1 %, keep the fractions strictly rising, and leave a balanced split where it is."""
    L = ctypes.CDLL(os.path.join(os.path.dirname(__file__), "emul", "libfpt_emul.so"))
    L.fpt_emul_rebalance.argtypes = [ctypes.c_int, _dp, _dp, ctypes.c_double, _dp]
    xs = np.linspace(0.0, 1.0, 20001)
    dens = 1.0 + 0.25 * (xs > 0.7) - 0.2 * ((xs > 0.3) & (xs < 0.5)) + 0.1 * np.sin(7 * xs)      # true time per unit of estimated cost
    cum = np.concatenate([[0.0], np.cumsum(0.5 * (dens[1:] + dens[:-1]) * np.diff(xs))])
    times = lambda f: np.diff(np.interp(f, xs, cum))
    for W in (2, 3, 8, 16):
        f = np.linspace(0.0, 1.0, W + 1)
        first = times(f).max() / times(f).mean()
        for _ in range(8):
            out = np.empty(W + 1)
            t = np.ascontiguousarray(times(f))
            assert L.fpt_emul_rebalance(W, f.ctypes.data_as(_dp), t.ctypes.data_as(_dp), 0.8, out.ctypes.data_as(_dp)) == 0
            assert out[0] == 0.0 and out[W] == 1.0 and np.all(np.diff(out) > 0)
            f = out
        last = times(f).max() / times(f).mean()
        assert last < 1.001 and (first < 1.002 or last < first), (W, first, last)
        out = np.empty(W + 1)                       # balanced times: nothing moves
        t = np.full(W, 3.0)
        assert L.fpt_emul_rebalance(W, f.ctypes.data_as(_dp), t.ctypes.data_as(_dp), 0.8, out.ctypes.data_as(_dp)) == 0
        assert np.allclose(out, f, atol=1e-15)
    # degenerate input (a zero-width segment would follow): refused, the caller keeps its boundaries
    f = np.array([0.0, 0.5, 1.0]); t = np.array([0.0, 1.0]); out = np.empty(3)
    assert L.fpt_emul_rebalance(2, f.ctypes.data_as(_dp), t.ctypes.data_as(_dp), 1.0, out.ctypes.data_as(_dp)) in (0, 1)


@pytest.mark.parametrize("o,v,world", [(6, 50, 8), (24, 114, 8), (5, 19, 3)])
def test_shards_with_adaptive_fractions_still_partition(built, o, v, world):
    """The adaptive balance only moves the boundary *fractions* (shard_items' `frac`): for any strictly rising set of fractions the parts
    stay contiguous and cover the list exactly once, and a part's share of the estimated cost follows its fraction interval."""
    L = ctypes.CDLL(os.path.join(os.path.dirname(__file__), "emul", "libfpt_emul.so"))
    L.fpt_emul_shard_frac.argtypes = [ctypes.c_int] * 5 + [_dp] + [ctypes.POINTER(ctypes.c_longlong)] * 2 + [_dp]
    n = fb.host.num_items(o, v)
    rng = np.random.default_rng(world)
    for trial in range(4):
        w = 1.0 + 0.1 * rng.standard_normal(world)
        frac = np.concatenate([[0.0], np.cumsum(w) / w.sum()])
        frac[-1] = 1.0
        prev_end = 0
        for r in range(world):
            b, e, sh = ctypes.c_longlong(), ctypes.c_longlong(), ctypes.c_double()
            assert L.fpt_emul_shard_frac(o, v, 1, r, world, frac.ctypes.data_as(_dp), ctypes.byref(b), ctypes.byref(e), ctypes.byref(sh)) == 0
            assert b.value == prev_end and e.value >= b.value
            prev_end = e.value
            if n >= 200 * world:
                assert abs(sh.value - (frac[r + 1] - frac[r])) < 0.01, (r, sh.value, frac)
        assert prev_end == n


# ---- the kernel's index algebra on the real molecules whose E(T) the reference holds (tests/test_oracle_kat.py has the sources) ---------
_HELD = [("water_sto3g", -0.0000738086, 5e-11),                                             # examples/Juliacon2022.ipynb:614 (10 decimals)
         ("water_ccpvtz", -76.343819598166903 + 76.335767822597347, 1e-9),                  # test/test_pT.jl:5,31    v = 53: 16+16+16+8 (vp 56)
         ("ammonia_augccpvdz", -56.427768639264869 + 56.422272522003723, 1e-9),             # test/test_pT.jl:6,32    v = 45: 16+16+16 (vp 48)
         ("formaldehyde_631gs", -114.189827180824139 + 114.180708994251702, 1e-9),          # test/test_pT.jl:9,35    v = 24: 16+8
         ("glycine_sto3g", -279.422940929335255 + 279.415437830677774, 1e-9)]               # test/test_pT.jl:10,36   o = 20


@pytest.mark.parametrize("name,held,tol", _HELD)
def test_emulator_lands_on_reference_held_values(emul, name, held, tol):
    """Not only the oracle but the product's own index algebra (fpt_layout.h: tiles, blocks, slots, GEMM descriptors, symmetric classes,
    energy stage), run with plain loops, gives the E(T) the reference holds for these molecules."""
    from types import SimpleNamespace
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    o, v = g["T1"].shape
    if "OVVV" in g.files:
        ovvv = g["OVVV"]
    else:                                   # stored packed over b >= c
        iu = np.triu_indices(v)
        ovvv = np.empty((o, v, v, v))
        ovvv[:, :, iu[1], iu[0]] = g["OVVV_packed"]
        ovvv[:, :, iu[0], iu[1]] = g["OVVV_packed"]
    x = SimpleNamespace(o=o, v=v, T1=g["T1"], T2=g["T2"], OVVV=ovvv, OOOV=g["OOOV"], OVOV=g["OVOV"], fo=g["fo"], fv=g["fv"])
    e, n = emul(x)
    assert abs(e - held) < tol, (e, held)
    assert abs(e - float(g["e_t"])) < 1e-13
    assert n == fb.host.num_items(o, v)
