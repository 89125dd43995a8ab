"""Host-side staging logic on the CPU (no GPU): the strided views the uploads build (fpt_stage.h: OVVV chunks with an occupied slice and
the b <= c prefix, the a <= b halves of T2 and OVOV), the packed-stream copy the staging threads run (arbitrary byte ranges, pieces
that cut rows), and the packed layouts the device kernels index (prep_pt_particle_tri, expand_t2_tri, expand_ovov_tri in
fpt_aux_kernels.cuh) -- both sides of that contract are checked against numpy gathers."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def stage():
    so = os.path.join(HERE, "emul", "libfpt_stage.so")
    src = os.path.join(HERE, "emul", "stage_main.cpp")
    hdr = os.path.join(HERE, "..", "fermi.jl_b200", "csrc", "fpt_stage.h")
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(so):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", cuda_inc, "-o", so, src])
    L = ctypes.CDLL(so)
    L.fpt_test_view_copy.restype = ctypes.c_longlong
    L.fpt_test_view_copy.argtypes = [ctypes.c_int] * 8 + [_dp] + [ctypes.c_longlong] * 3 + [ctypes.c_int, _dp]

    def run(kind, A, o, v, p0=0, np_=0, c0=0, cn=0, flag=0, off=0, nb=None, piece=1 << 12, nt=1):
        src_p = A.ravel(order="F").ctypes.data_as(_dp) if A is not None else None
        total = L.fpt_test_view_copy(kind, o, v, p0, np_, c0, cn, flag, src_p, 0, 0, piece, nt, None)
        if nb is None:
            nb = total - off
        out = np.empty(nb // 8)
        flat = np.ascontiguousarray(A.ravel(order="F"))
        got = L.fpt_test_view_copy(kind, o, v, p0, np_, c0, cn, flag, flat.ctypes.data_as(_dp), off, nb, piece, nt, out.ctypes.data_as(_dp))
        assert got == total
        return out, total

    return run


def _tri(c):
    return c * (c + 1) // 2


@pytest.mark.parametrize("o,v,p0,npp,c0,cn,half", [(3, 5, 0, 3, 0, 5, 1), (4, 7, 1, 2, 2, 3, 1), (6, 9, 4, 2, 0, 9, 0), (5, 6, 0, 5, 3, 2, 1), (8, 11, 4, 4, 5, 6, 1)])
def test_ovvv_chunk_view_and_prep_indexing(stage, o, v, p0, npp, c0, cn, half):
    rng = np.random.default_rng(o * 100 + v)
    A = np.asfortranarray(rng.standard_normal((o, v, v, v)))
    packed, total = stage(0, A, o, v, p0, npp, c0, cn, half)
    # what the view promises: for c in the chunk, rows (a, b) with b <= c (half) or all b, the occupied slice, packed
    want = []
    for c in range(c0, c0 + cn):
        nb = c + 1 if half else v
        want.append(A[p0:p0 + npp, :, :nb, c].ravel(order="F"))
    want = np.concatenate(want)
    assert total == want.size * 8 and np.array_equal(packed, want)
    if half:   # prep_pt_particle_tri: src[np v (tri(c) - tri(c0) + b) + pl + np y] = OVVV[p0 + pl, y, b, c]
        for c in range(c0, c0 + cn):
            for b in range(c + 1):
                blk = packed[npp * v * (_tri(c) - _tri(c0) + b): npp * v * (_tri(c) - _tri(c0) + b + 1)].reshape((npp, v), order="F")
                assert np.array_equal(blk, A[p0:p0 + npp, :, b, c])


@pytest.mark.parametrize("o,v", [(1, 1), (2, 3), (3, 7), (5, 4)])
def test_t2_and_ovov_half_views_and_expand_indexing(stage, o, v):
    rng = np.random.default_rng(7 * o + v)
    T2 = rng.standard_normal((o, o, v, v))
    T2 = np.asfortranarray(0.5 * (T2 + T2.transpose(1, 0, 3, 2)))
    tri, total = stage(1, T2, o, v)
    assert total == 8 * o * o * _tri(v)
    o2 = o * o
    full = np.empty_like(T2)
    for i in range(o):          # expand_t2_tri: a <= b: src[o^2 (tri(b) + a) + i + o j], else src[o^2 (tri(a) + b) + j + o i]
        for j in range(o):
            for a in range(v):
                for b in range(v):
                    full[i, j, a, b] = tri[o2 * (_tri(b) + a) + i + o * j] if a <= b else tri[o2 * (_tri(a) + b) + j + o * i]
    assert np.array_equal(full, T2)
    B = rng.standard_normal((4, o, v))
    OVOV = np.asfortranarray(np.einsum("Qia,Qjb->iajb", B, B))
    OVOV = np.asfortranarray(0.5 * (OVOV + OVOV.transpose(2, 3, 0, 1)))       # exactly symmetric
    tri2, total2 = stage(2, OVOV, o, v)
    assert total2 == 8 * o2 * _tri(v)
    full2 = np.empty_like(OVOV)
    for i in range(o):          # expand_ovov_tri: a <= b: src[o^2 tri(b) + j o (b+1) + i + o a], else src[o^2 tri(a) + i o (a+1) + j + o b]
        for a in range(v):
            for j in range(o):
                for b in range(v):
                    full2[i, a, j, b] = (tri2[o2 * _tri(b) + j * o * (b + 1) + i + o * a] if a <= b
                                         else tri2[o2 * _tri(a) + i * o * (a + 1) + j + o * b])
    assert np.array_equal(full2, OVOV)


@pytest.mark.parametrize("nt", [0, 1])
def test_packed_ranges_cut_rows_and_pieces(stage, nt):
    """Any byte range of the packed stream (a GPU's share of a sharded upload), copied in pieces that cut rows and slabs, equals the same
    range of the whole stream -- with and without non-temporal stores."""
    o, v = 6, 13
    rng = np.random.default_rng(3)
    A = np.asfortranarray(rng.standard_normal((o, v, v, v)))
    whole, total = stage(0, A, o, v, 2, 3, 1, 9, 1, nt=nt)
    for off, nb, piece in [(0, total, 64), (8, 1000, 24), (4096, total - 4096, 4096), (total - 40, 40, 512), (1032, 7 * 512, 1 << 20)]:
        part, _ = stage(0, A, o, v, 2, 3, 1, 9, 1, off=off, nb=nb, piece=piece, nt=nt)
        assert np.array_equal(part, whole[off // 8:(off + nb) // 8]), (off, nb, piece)
    T2 = np.asfortranarray(rng.standard_normal((o, o, v, v)))
    w2, t2 = stage(1, T2, o, v, nt=nt)
    for W in (2, 3, 8):          # the sharding rule of `distribute`: parts of ceil(n / W) doubles rounded up to 512
        n = t2 // 8
        part = ((n + W - 1) // W + 511) & ~511
        got = np.concatenate([stage(1, T2, o, v, off=8 * min(n, g * part), nb=8 * (min(n, (g + 1) * part) - min(n, g * part)), piece=2048, nt=nt)[0]
                              for g in range(W)])
        assert np.array_equal(got, w2)
