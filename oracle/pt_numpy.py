"""CPU ORACLE (test infrastructure, NOT product code) -- numpy restatement of Fermi.jl's RCCSD(T).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product path (fermi.jl_b200) never does.

Follows, line for line, the reference's default (T) algorithm
    src/Methods/CoupledCluster/PerturbativeTriples/ijk.jl:20-150
(`RCCSDpT(ccsd, moints, ::ijk)`), and cross-checks it against the explicit-GEMM form of
    src/Methods/CoupledCluster/PerturbativeTriples/ijk2.jl:21-190
and against an independent spin-orbital brute-force (T) (`pt_spinorbital_bruteforce`).

PARITY PIN: the reference itself (Julia) cannot run in the build container, and its only direct (T)
tests need SCF+CCSD inputs.  The pin is therefore (a) this transcription == the ijk2 transcription ==
spin-orbital brute force, and (b) the known answer printed by the reference for water/STO-3G
(examples/Juliacon2022.ipynb:613-615, E(T) = -0.0000738086, CCSD(T) = -75.0187834019), reproduced by
oracle/mini_ccsd.py feeding this function (see tests/test_oracle_kat.py).

All arrays are in the reference's index order: T1[i,a], T2[i,j,a,b], OVVV[i,a,b,c]=(ia|bc),
OOOV[i,j,k,a]=(ij|ka), OVOV[i,a,j,b]=(ia|jb), fo[i], fv[a].  numpy arrays of any memory order are
accepted (indexing is by axis, not by layout).
"""
from __future__ import annotations

import itertools

import numpy as np


def _energy_from_WV(W, V, Dijk, fv, dij, djk):
    """ijk.jl:120-136 -- scalar a>=b>=c loop, vectorised over an index mask (same arithmetic)."""
    v = W.shape[0]
    a, b, c = np.meshgrid(np.arange(v), np.arange(v), np.arange(v), indexing="ij")
    m = (a >= b) & (b >= c)
    a, b, c = a[m], b[m], c[m]
    Wabc, Wacb, Wbac, Wbca, Wcab, Wcba = W[a, b, c], W[a, c, b], W[b, a, c], W[b, c, a], W[c, a, b], W[c, b, a]
    Vabc, Vacb, Vbac, Vbca, Vcab, Vcba = V[a, b, c], V[a, c, b], V[b, a, c], V[b, c, a], V[c, a, b], V[c, b, a]
    X = Wabc * Vabc + Wacb * Vacb + Wbac * Vbac + Wbca * Vbca + Wcab * Vcab + Wcba * Vcba  # :129
    Y = Vabc + Vbca + Vcab  # :130
    Z = Vacb + Vbac + Vcba  # :131
    Ef = (Y - 2 * Z) * (Wabc + Wbca + Wcab) + (Z - 2 * Y) * (Wacb + Wbac + Wcba) + 3 * X  # :132
    Dd = Dijk - fv[a] - fv[b] - fv[c]  # :122,124,127
    dab = (a == b).astype(float)
    dbc = (b == c).astype(float)
    return float(np.sum(Ef * (2 - dij - djk) / (Dd * (1 + dab + dbc))))  # :133


def triplet_WV(T1, T2, OVVV, OOOV, OVOV, i, j, k):
    """W and V for one (i,j,k) -- ijk.jl:108-117 with the permutes of :24-32 undone."""
    es = np.einsum
    W = es("abd,cd->abc", OVVV[i], T2[k, j]) - es("lb,lca->abc", OOOV[:, i, j, :], T2[k])  # :109
    W += es("bad,cd->abc", OVVV[j], T2[k, i]) - es("la,lcb->abc", OOOV[:, j, i, :], T2[k])  # :110
    W += es("cad,bd->abc", OVVV[k], T2[j, i]) - es("la,lbc->abc", OOOV[:, k, i, :], T2[j])  # :111
    W += es("cbd,da->abc", OVVV[k], T2[j, i]) - es("lb,lac->abc", OOOV[:, k, j, :], T2[i])  # :112
    W += es("acd,db->abc", OVVV[i], T2[k, j]) - es("lc,lba->abc", OOOV[:, i, k, :], T2[j])  # :113
    W += es("bcd,da->abc", OVVV[j], T2[k, i]) - es("lc,lab->abc", OOOV[:, j, k, :], T2[i])  # :114
    V = (
        W
        + es("a,bc->abc", T1[i], OVOV[j, :, k, :])
        + es("ac,b->abc", OVOV[i, :, k, :], T1[j])
        + es("ab,c->abc", OVOV[i, :, j, :], T1[k])
    )  # :116
    return W, V


def pt_ijk(T1, T2, OVVV, OOOV, OVOV, fo, fv, triplets=None):
    """E(T) following ijk.jl:20-150.  `triplets` (iterable of (i,j,k), 0-based, i>=j>=k) restricts the
    sum to a subset (used for bounded CPU-baseline samples and for sharding tests)."""
    o, v = T1.shape
    if triplets is None:
        triplets = ((i, j, k) for i in range(o) for j in range(i + 1) for k in range(j + 1))  # :49,63,83
    Et = 0.0
    for i, j, k in triplets:
        W, V = triplet_WV(T1, T2, OVVV, OOOV, OVOV, i, j, k)
        Dijk = fo[i] + fo[j] + fo[k]
        Et += _energy_from_WV(W, V, Dijk, fv, float(i == j), float(j == k))
    return Et  # :145


def pt_ijk2(T1, T2, OVVV, OOOV, OVOV, fo, fv):
    """E(T) following the explicit-GEMM formulation ijk2.jl:26-33,110-176."""
    o, v = T1.shape

    def X(p, q, r):
        # X(p;q,r)[x,y,z] = sum_d OVVV[p,y,x,d] t2[r,q,z,d] - sum_l t2[p,l,y,x] OOOV[l,q,r,z]
        return np.einsum("yxd,zd->xyz", OVVV[p], T2[r, q]) - np.einsum("lyx,lz->xyz", T2[p], OOOV[:, q, r, :])

    Et = 0.0
    for i in range(o):
        for j in range(i + 1):
            for k in range(j + 1):
                W = X(j, i, k).copy()  # abc  :110-114
                W += X(k, i, j).transpose(0, 2, 1)  # acb  :116-123
                W += X(i, j, k).transpose(1, 0, 2)  # bac  :125-131
                W += X(k, j, i).transpose(2, 0, 1)  # bca  :133-139  (W[a,b,c] += X[b,c,a])
                W += X(i, k, j).transpose(1, 2, 0)  # cab  :141-147  (W[a,b,c] += X[c,a,b])
                W += X(j, k, i).transpose(2, 1, 0)  # cba  :149-153
                V = (
                    W
                    + np.einsum("a,bc->abc", T1[i], OVOV[j, :, k, :])
                    + np.einsum("ac,b->abc", OVOV[i, :, k, :], T1[j])
                    + np.einsum("ab,c->abc", OVOV[i, :, j, :], T1[k])
                )  # :155-157
                Et += _energy_from_WV(W, V, fo[i] + fo[j] + fo[k], fv, float(i == j), float(j == k))
    return Et


def pt_ijk_loops(T1, T2, OVVV, OOOV, OVOV, fo, fv):
    """Pure-Python scalar loops over (a,b,c) for the energy stage (tiny shapes only): the literal
    ijk.jl:120-136 nest, used to validate the vectorised `_energy_from_WV`."""
    o, v = T1.shape
    Et = 0.0
    for i in range(o):
        for j in range(i + 1):
            for k in range(j + 1):
                W, V = triplet_WV(T1, T2, OVVV, OOOV, OVOV, i, j, k)
                Dijk = fo[i] + fo[j] + fo[k]
                dij, djk = float(i == j), float(j == k)
                for a in range(v):
                    for b in range(a + 1):
                        for c in range(b + 1):
                            Dd = Dijk - fv[a] - fv[b] - fv[c]
                            X = (W[a, b, c] * V[a, b, c] + W[a, c, b] * V[a, c, b] + W[b, a, c] * V[b, a, c]
                                 + W[b, c, a] * V[b, c, a] + W[c, a, b] * V[c, a, b] + W[c, b, a] * V[c, b, a])
                            Y = V[a, b, c] + V[b, c, a] + V[c, a, b]
                            Z = V[a, c, b] + V[b, a, c] + V[c, b, a]
                            Ef = ((Y - 2 * Z) * (W[a, b, c] + W[b, c, a] + W[c, a, b])
                                  + (Z - 2 * Y) * (W[a, c, b] + W[b, a, c] + W[c, b, a]) + 3 * X)
                            Et += Ef * (2 - dij - djk) / (Dd * (1 + float(a == b) + float(b == c)))
    return Et


def pt_spinorbital_bruteforce(T1, T2, OVVV, OOOV, OVOV, fo, fv):
    """Independent check: textbook spin-orbital (T) with antisymmetrised integrals,
       E = (1/36) sum C (C + S) / D,  built from the closed-shell arrays.  O((2o)^3 (2v)^3 (2v+2o))
       -- only for o<=3, v<=5.  Shares no contraction code with pt_ijk."""
    o, v = T1.shape
    no, nv = 2 * o, 2 * v
    so = lambda P: (P // 2, P % 2)  # spatial index, spin
    # spin-orbital amplitudes
    t1 = np.zeros((no, nv))
    t2 = np.zeros((no, no, nv, nv))
    for I in range(no):
        i, si = so(I)
        for A in range(nv):
            a, sa = so(A)
            if si == sa:
                t1[I, A] = T1[i, a]
    for I, J, A, B in itertools.product(range(no), range(no), range(nv), range(nv)):
        (i, si), (j, sj), (a, sa), (b, sb) = so(I), so(J), so(A), so(B)
        val = 0.0
        if si == sa and sj == sb:
            val += T2[i, j, a, b]
        if si == sb and sj == sa:
            val -= T2[j, i, a, b]
        t2[I, J, A, B] = val
    # antisymmetrised physicist integrals <pq||rs> = (pr|qs) d(sp,sr) d(sq,ss) - (ps|qr) d(sp,ss) d(sq,sr)
    # needed blocks: <ei||bc> -> <vo||vv> via (ia|bc) ; <ma||jk> -> <ov||oo> via (ij|ka) ; <jk||bc> via (ia|jb)
    def vovv(E, I, B, C):  # <ei||bc> = (eb|ic) - (ec|ib)
        (e, se), (i, si), (b, sb), (c, sc) = so(E), so(I), so(B), so(C)
        r = 0.0
        if se == sb and si == sc:
            r += OVVV[i, c, e, b]  # (ic|eb)
        if se == sc and si == sb:
            r -= OVVV[i, b, e, c]  # (ib|ec)
        return r

    def ovoo(M, A, J, K):  # <ma||jk> = (mj|ak) - (mk|aj)
        (m, sm), (a, sa), (j, sj), (k, sk) = so(M), so(A), so(J), so(K)
        r = 0.0
        if sm == sj and sa == sk:
            r += OOOV[m, j, k, a]  # (mj|ka)
        if sm == sk and sa == sj:
            r -= OOOV[m, k, j, a]  # (mk|ja)
        return r

    def oovv(J, K, B, C):  # <jk||bc> = (jb|kc) - (jc|kb)
        (j, sj), (k, sk), (b, sb), (c, sc) = so(J), so(K), so(B), so(C)
        r = 0.0
        if sj == sb and sk == sc:
            r += OVOV[j, b, k, c]
        if sj == sc and sk == sb:
            r -= OVOV[j, c, k, b]
        return r

    VOVV = np.zeros((nv, no, nv, nv))
    for E, I, B, C in itertools.product(range(nv), range(no), range(nv), range(nv)):
        VOVV[E, I, B, C] = vovv(E, I, B, C)
    OVOO = np.zeros((no, nv, no, no))
    for M, A, J, K in itertools.product(range(no), range(nv), range(no), range(no)):
        OVOO[M, A, J, K] = ovoo(M, A, J, K)
    OOVV = np.zeros((no, no, nv, nv))
    for J, K, B, C in itertools.product(range(no), range(no), range(nv), range(nv)):
        OOVV[J, K, B, C] = oovv(J, K, B, C)
    eo = np.repeat(np.asarray(fo), 2)
    ev = np.repeat(np.asarray(fv), 2)

    # connected:   base_c[i,j,k,a,b,c] = sum_e t_jk^ae <ei||bc> - sum_m t_im^bc <ma||jk>
    base_c = np.einsum("jkae,eibc->ijkabc", t2, VOVV) - np.einsum("imbc,majk->ijkabc", t2, OVOO)
    base_d = np.einsum("ia,jkbc->ijkabc", t1, OOVV)

    def P_perm(x):
        # P(i/jk) f = f(ijk) - f(jik) - f(kji) ; P(a/bc) likewise on the last three axes
        y = x - x.transpose(1, 0, 2, 3, 4, 5) - x.transpose(2, 1, 0, 3, 4, 5)
        return y - y.transpose(0, 1, 2, 4, 3, 5) - y.transpose(0, 1, 2, 5, 4, 3)

    C = P_perm(base_c)
    S = P_perm(base_d)
    D = (eo[:, None, None, None, None, None] + eo[None, :, None, None, None, None] + eo[None, None, :, None, None, None]
         - ev[None, None, None, :, None, None] - ev[None, None, None, None, :, None] - ev[None, None, None, None, None, :])
    return float(np.sum(C * (C + S) / D) / 36.0)


def mo_blocks_from_ao(AOERI, C, ndocc, drop_occ=0, drop_vir=0):
    """The reference's dense AO -> MO contractions for the three blocks the (T) path reads -- transcription of
    src/Core/Integrals/ROIntegrals/Chonky.jl: compute_OOOV! (:28-48), compute_OVOV! (:72-92), compute_OVVV! (:94-114).
    o = (1+core):ndocc, v = (ndocc+1):(nbf-inac) there (1-based); returns (OVVV, OOOV, OVOV), Fortran-ordered."""
    nmo = C.shape[1]
    Co = C[:, drop_occ:ndocc]
    Cv = C[:, ndocc:nmo - drop_vir]
    F = np.asfortranarray
    OOOV = np.einsum("mnrs,mi,nj,rk,sa->ijka", AOERI, Co, Co, Co, Cv, optimize=True)   # :45
    OVOV = np.einsum("mnrs,mi,na,rj,sb->iajb", AOERI, Co, Cv, Co, Cv, optimize=True)   # :89
    OVVV = np.einsum("mnrs,mi,na,rb,sc->iabc", AOERI, Co, Cv, Cv, Cv, optimize=True)   # :111
    return F(OVVV), F(OOOV), F(OVOV)


def df_factors_from_ao(Bmn, C, ndocc, drop_occ=0, drop_vir=0):
    """MO-basis DF factors from the AO-basis ones, Bmn[Q, mu, nu] -- transcription of
    src/Core/Integrals/ROIntegrals/DFERI.jl: compute_BOO! (:15-29), compute_BOV! (:31-51), compute_BVV! (:53-69).
    Returns (BOO, BOV, BVV) with the auxiliary index first, Fortran-ordered (what fpt_triples_df is handed)."""
    nmo = C.shape[1]
    Co = C[:, drop_occ:ndocc]
    Cv = C[:, ndocc:nmo - drop_vir]
    F = np.asfortranarray
    BOO = np.einsum("Qmn,mi,nj->Qij", Bmn, Co, Co, optimize=True)   # :26
    BOV = np.einsum("Qmn,mi,na->Qia", Bmn, Co, Cv, optimize=True)   # :48
    BVV = np.einsum("Qmn,ma,nb->Qab", Bmn, Cv, Cv, optimize=True)   # :66
    return F(BOO), F(BOV), F(BVV)


def mo_blocks_from_df(BOO, BOV, BVV):
    """The three blocks the (T) path reads, assembled from DF factors -- transcription of DFERI.jl: compute_OOOV! (:88-112),
    compute_OVOV! (:139-154), compute_OVVV! (:156-180).  Returns (OVVV, OOOV, OVOV), Fortran-ordered."""
    F = np.asfortranarray
    OOOV = np.einsum("Qij,Qka->ijka", BOO, BOV, optimize=True)      # :109
    OVOV = np.einsum("Qia,Qjb->iajb", BOV, BOV, optimize=True)      # :151
    OVVV = np.einsum("Qia,Qbc->iabc", BOV, BVV, optimize=True)      # :177
    return F(OVVV), F(OOOV), F(OVOV)


def _index2(i, j):
    """Backend/Arrays.jl:34-40."""
    return (j * (j + 1)) // 2 + i if i < j else (i * (i + 1)) // 2 + j


def ovvv_from_sparse(indexes, data, C, ndocc, drop_occ=0, drop_vir=0):
    """Transcription of compute_OVVV! for the sparse AO list (Sparse.jl:316-393): every symmetry-unique integral is scattered,
    times Co[.,i], to the images the reference enumerates case by case (its gamma flags), giving the partially contracted
    i-nu-rho-sigma array, which is then contracted with Cv three times.  Pure-Python loops: small cases only."""
    nbf, nmo = C.shape
    Co = C[:, drop_occ:ndocc]
    Cv = C[:, ndocc:nmo - drop_vir]
    no = Co.shape[1]
    X = np.zeros((no, nbf, nbf, nbf))
    for (m, n, r, s), V in zip(indexes, data):
        g_mn = m != n
        g_rs = r != s
        g_ab = _index2(m, n) != _index2(r, s)
        Vm, Vn, Vr, Vs = V * Co[m], V * Co[n], V * Co[r], V * Co[s]
        if g_ab and g_mn and g_rs:
            X[:, n, r, s] += Vm; X[:, n, s, r] += Vm; X[:, m, r, s] += Vn; X[:, m, s, r] += Vn
            X[:, s, m, n] += Vr; X[:, s, n, m] += Vr; X[:, r, m, n] += Vs; X[:, r, n, m] += Vs
        elif g_ab and g_mn:
            X[:, n, r, s] += Vm; X[:, m, r, s] += Vn; X[:, s, m, n] += Vr; X[:, s, n, m] += Vr
        elif g_ab and g_rs:
            X[:, n, r, s] += Vm; X[:, n, s, r] += Vm; X[:, s, m, n] += Vr; X[:, r, m, n] += Vs
        elif g_mn and g_rs:
            X[:, n, r, s] += Vm; X[:, n, s, r] += Vm; X[:, m, r, s] += Vn; X[:, m, s, r] += Vn
        elif g_ab:
            X[:, n, r, s] += Vm; X[:, s, m, n] += Vr
        else:
            X[:, n, r, s] += Vm
    return np.asfortranarray(np.einsum("inrs,na,rb,sc->iabc", X, Cv, Cv, Cv, optimize=True))   # :389-391


def sparse_from_dense(AOERI, threshold=0.0):
    """Symmetry-unique entries (mu >= nu, rho >= sigma, (mu nu) >= (rho sigma)) of a dense AO tensor with |value| > threshold,
    as (indexes (nint,4) int16 zero-based, data) -- the shape of the list GaussianBasis.sparseERI_2e4c hands the reference
    (AtomicIntegrals.jl:48-52)."""
    nbf = AOERI.shape[0]
    idx, vals = [], []
    for m in range(nbf):
        for n in range(m + 1):
            for r in range(nbf):
                for s in range(r + 1):
                    if _index2(m, n) < _index2(r, s):
                        continue
                    V = AOERI[m, n, r, s]
                    if abs(V) > threshold:
                        idx.append((m, n, r, s)); vals.append(V)
    return np.array(idx, dtype=np.int16).reshape(-1, 4), np.array(vals)
