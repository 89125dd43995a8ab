"""CPU tests of the oracle itself: the C restatement (naive loops and GEMM form) against the numpy transcriptions,
the spin-orbital brute force and the committed golden vectors."""
import json
import os

import numpy as np
import pytest

import fermi_jl_b200 as fb
import oracle
from oracle import pt_numpy as P

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pt_golden.json")))


def _args(x):
    return (x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)


@pytest.mark.parametrize("rec", GOLD, ids=lambda r: f"o{r['o']}v{r['v']}")
def test_golden_vectors(rec):
    x = fb.synth.make_inputs(rec["o"], rec["v"], naux=rec["naux"], seed=rec["seed"])
    for f in (oracle.pt_naive, oracle.pt_gemm):
        e = f(*_args(x))
        assert abs(e - rec["E_ijk"]) < 1e-13, (f.__name__, e, rec["E_ijk"])
    assert abs(rec["E_ijk"] - rec["E_ijk2"]) < 1e-13
    if "E_spinorbital" in rec:
        assert abs(rec["E_ijk"] - rec["E_spinorbital"]) < 1e-13


def test_numpy_transcriptions_agree_with_bruteforce():
    x = fb.synth.make_inputs(2, 4, naux=5, seed=3)
    a = _args(x)
    e = P.pt_ijk(*a)
    assert abs(P.pt_ijk2(*a) - e) < 1e-14
    assert abs(P.pt_ijk_loops(*a) - e) < 1e-14
    assert abs(P.pt_spinorbital_bruteforce(*a) - e) < 1e-14


def test_unsymmetric_inputs_break_ijk_vs_ijk2():
    """the two reference algorithms agree only for inputs with the physical symmetries -- the oracle follows
    ijk.jl index for index, so it must reproduce that disagreement (guards against an accidental symmetrisation)."""
    rng = np.random.default_rng(0)
    o, v = 2, 3
    T1, T2 = rng.standard_normal((o, v)), rng.standard_normal((o, o, v, v))
    OVVV, OOOV, OVOV = rng.standard_normal((o, v, v, v)), rng.standard_normal((o, o, o, v)), rng.standard_normal((o, v, o, v))
    fo, fv = -np.sort(rng.uniform(0.5, 2, o))[::-1], np.sort(rng.uniform(0.5, 2, v))
    e1 = P.pt_ijk(T1, T2, OVVV, OOOV, OVOV, fo, fv)
    assert abs(oracle.pt_naive(T1, T2, OVVV, OOOV, OVOV, fo, fv) - e1) < 1e-10 * max(1.0, abs(e1))
    assert abs(P.pt_ijk2(T1, T2, OVVV, OOOV, OVOV, fo, fv) - e1) > 1e-6


def test_triplet_ranges_add_up():
    x = fb.synth.make_inputs(5, 11, naux=6, seed=8)
    a = _args(x)
    full = oracle.pt_gemm(*a)
    ntrip = 5 * 6 * 7 // 6
    cuts = [0, 4, 9, 20, ntrip]
    parts = [oracle.pt_gemm(*a, t_begin=cuts[i], t_end=cuts[i + 1]) for i in range(len(cuts) - 1)]
    assert abs(sum(parts) - full) < 1e-15
    assert abs(P.pt_ijk(*a) - full) < 1e-14
    assert abs(oracle.pt_gemm(*a, nthreads=1) - full) < 1e-15


def test_degenerate_shapes():
    for o, v in [(1, 1), (1, 5), (2, 1), (3, 2)]:
        x = fb.synth.make_inputs(o, v, naux=3, seed=1)
        a = _args(x)
        e = oracle.pt_naive(*a)
        assert np.isfinite(e)
        assert abs(oracle.pt_gemm(*a) - e) < 1e-15
        assert abs(P.pt_ijk(*a) - e) < 1e-15
    # o = 1: the only triplet is i=j=k, weight 0 (ijk.jl:133)
    x = fb.synth.make_inputs(1, 6, naux=3, seed=2)
    assert oracle.pt_naive(*_args(x)) == 0.0


def test_synth_symmetries_and_dump_roundtrip(tmp_path):
    x = fb.synth.make_inputs(3, 5, naux=4, seed=4)
    assert np.allclose(x.T2, x.T2.transpose(1, 0, 3, 2))
    assert np.allclose(x.OVVV, x.OVVV.transpose(0, 1, 3, 2))
    assert np.allclose(x.OOOV, x.OOOV.transpose(1, 0, 2, 3))
    assert np.allclose(x.OVOV, x.OVOV.transpose(2, 3, 0, 1))
    assert x.T2.flags.f_contiguous and x.OVVV.flags.f_contiguous
    fb.synth.dump(x, str(tmp_path / "d"))
    y = fb.synth.load(str(tmp_path / "d"))
    for k in ("T1", "T2", "OVVV", "OOOV", "OVOV", "fo", "fv", "BOO", "BOV", "BVV"):
        assert np.array_equal(getattr(x, k), getattr(y, k))


def test_ao_to_mo_transcription_against_factorised_form():
    """oracle.pt_numpy.mo_blocks_from_ao (Chonky.jl:28-114) vs. transforming the DF-like factor first."""
    import fermi_jl_b200 as fb
    from oracle import pt_numpy as PN
    nbf, ndocc, dc, dv = 12, 4, 1, 2
    rng = np.random.default_rng(3)
    B = rng.standard_normal((7, nbf, nbf)); B = 0.5 * (B + B.transpose(0, 2, 1))
    AO = np.einsum("Qmn,Qrs->mnrs", B, B)
    C, _ = np.linalg.qr(rng.standard_normal((nbf, nbf)))
    Co, Cv = C[:, dc:ndocc], C[:, ndocc:nbf - dv]
    Bov = np.einsum("Qmn,mi,na->Qia", B, Co, Cv); Bvv = np.einsum("Qmn,ma,nb->Qab", B, Cv, Cv); Boo = np.einsum("Qmn,mi,nj->Qij", B, Co, Co)
    OVVV, OOOV, OVOV = PN.mo_blocks_from_ao(AO, C, ndocc, dc, dv)
    assert np.abs(OVVV - np.einsum("Qia,Qbc->iabc", Bov, Bvv)).max() < 1e-12
    assert np.abs(OOOV - np.einsum("Qij,Qka->ijka", Boo, Bov)).max() < 1e-12
    assert np.abs(OVOV - np.einsum("Qia,Qjb->iajb", Bov, Bov)).max() < 1e-12
    assert OVVV.shape == (3, 6, 6, 6) and OVVV.flags.f_contiguous


def test_oracle_invariances():
    """Properties the full-size GPU tests rely on: E(T) is invariant under relabelling virtual / occupied orbitals and
    homogeneous of degree 2 in (T1, T2)."""
    import fermi_jl_b200 as fb
    o, v = 4, 9
    x = fb.synth.make_inputs(o, v, naux=8, seed=9)
    F = np.asfortranarray
    e0 = oracle.pt_gemm(x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
    rng = np.random.default_rng(1)
    pv, po = rng.permutation(v), rng.permutation(o)
    ev = oracle.pt_gemm(F(x.T1[:, pv]), F(x.T2[:, :, pv][:, :, :, pv]), F(x.OVVV[:, pv][:, :, pv][:, :, :, pv]),
                        F(x.OOOV[:, :, :, pv]), F(x.OVOV[:, pv][:, :, :, pv]), x.fo, x.fv[pv].copy())
    eo = oracle.pt_gemm(F(x.T1[po]), F(x.T2[po][:, po]), F(x.OVVV[po]), F(x.OOOV[po][:, po][:, :, po]),
                        F(x.OVOV[po][:, :, po]), x.fo[po].copy(), x.fv)
    es = oracle.pt_gemm(F(1.75 * x.T1), F(1.75 * x.T2), x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
    assert abs(ev - e0) < 1e-15 and abs(eo - e0) < 1e-15 and abs(es - 1.75 ** 2 * e0) < 1e-15


def test_sparse_ao_list_transcription():
    """oracle.pt_numpy.ovvv_from_sparse (the case-by-case scatter of Sparse.jl:316-393) equals the dense contraction of the
    tensor the list was taken from -- i.e. expanding every unique integral to its 8 images (what the GPU does) is what the
    reference's gamma cases amount to."""
    from oracle import pt_numpy as PN
    import fermi_jl_b200 as fb
    nbf, ndocc, dc, dv = 8, 3, 1, 1
    AO, C, *_ = fb.synth.make_ao_inputs(nbf, ndocc, dc, dv, seed=2)
    idx, vals = PN.sparse_from_dense(AO)
    assert len(vals) == (nbf * (nbf + 1) // 2) * (nbf * (nbf + 1) // 2 + 1) // 2
    OVVV_sparse = PN.ovvv_from_sparse(idx, vals, C, ndocc, dc, dv)
    OVVV, _, _ = PN.mo_blocks_from_ao(AO, C, ndocc, dc, dv)
    assert np.abs(OVVV_sparse - OVVV).max() < 1e-13


def test_oracle_blas_backends_agree():
    """pt_gemm on its own register-blocked kernel and on the single-threaded OpenBLAS bundled with scipy (the arrangement of
    ijk.jl:45-46) give the same E(T); the timed CPU baseline uses whichever is faster."""
    import fermi_jl_b200 as fb
    x = fb.synth.make_inputs(4, 19, naux=8, seed=3)
    a = (x.T1, x.T2, x.OVVV, x.OOOV, x.OVOV, x.fo, x.fv)
    assert oracle.use_blas("own") == "own"
    e_own = oracle.pt_gemm(*a)
    if oracle.use_blas("openblas") != "openblas":
        pytest.skip("no OpenBLAS in this environment")
    try:
        e_ob = oracle.pt_gemm(*a)
    finally:
        oracle.use_blas("own")
    assert abs(e_own - e_ob) < 1e-13
    assert abs(e_own - oracle.pt_naive(*a)) < 1e-13
