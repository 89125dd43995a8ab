"""CPU ORACLE SUPPORT (test infrastructure, NOT product code) -- Gaussian integrals for shells of any angular momentum.

oracle/mini_ccsd.py pins the (T) oracle on water / STO-3G and 6-31G with pure-Python integrals over s and p functions.  Config C2
of BASELINE.json is water / cc-pVTZ (o = 5, v = 53: d and f shells, real solid harmonics), for which the reference's own test
suite holds Psi4's CCSD and CCSD(T) totals (test/test_pT.jl:4-54).  This module provides what that needs: McMurchie-Davidson
one- and two-electron integrals over contracted Cartesian shells, compiled with numba, and the transformation to real solid
harmonics.  Basis functions are scaled to unit self-overlap at the end; energies and MO-basis quantities do not depend on the
scaling (nor on the order or sign convention of the harmonics), so none of the reference's conventions has to be matched.
"""
from __future__ import annotations

import math

import numpy as np
from numba import njit, prange

LMAX = 3


def cart_components(l):
    return [(lx, ly, l - lx - ly) for lx in range(l, -1, -1) for ly in range(l - lx, -1, -1)]


def sph_matrix(l):
    """(ncart, 2l+1): real solid harmonics of degree l as combinations of the monomials x^lx y^ly z^lz (not normalised)."""
    comps = cart_components(l)
    idx = {c: n for n, c in enumerate(comps)}

    def vec(terms):
        v = np.zeros(len(comps))
        for coef, (a, b, c) in terms:
            v[idx[(a, b, c)]] += coef
        return v

    if l == 0:
        rows = [[(1, (0, 0, 0))]]
    elif l == 1:
        rows = [[(1, (1, 0, 0))], [(1, (0, 1, 0))], [(1, (0, 0, 1))]]
    elif l == 2:
        rows = [[(1, (1, 1, 0))], [(1, (0, 1, 1))], [(2, (0, 0, 2)), (-1, (2, 0, 0)), (-1, (0, 2, 0))], [(1, (1, 0, 1))],
                [(1, (2, 0, 0)), (-1, (0, 2, 0))]]
    elif l == 3:
        rows = [[(3, (2, 1, 0)), (-1, (0, 3, 0))], [(1, (1, 1, 1))], [(4, (0, 1, 2)), (-1, (2, 1, 0)), (-1, (0, 3, 0))],
                [(2, (0, 0, 3)), (-3, (2, 0, 1)), (-3, (0, 2, 1))], [(4, (1, 0, 2)), (-1, (3, 0, 0)), (-1, (1, 2, 0))],
                [(1, (2, 0, 1)), (-1, (0, 2, 1))], [(1, (3, 0, 0)), (-3, (1, 2, 0))]]
    else:
        raise ValueError("l > 3 not tabulated")
    return np.array([vec(r) for r in rows]).T


@njit(cache=True)
def boys(nmax, x, out):
    """out[n] = F_n(x), n = 0..nmax"""
    if x < 1e-13:
        for n in range(nmax + 1):
            out[n] = 1.0 / (2 * n + 1)
        return
    ex = math.exp(-x)
    if x < 35.0:
        term = 1.0 / (2 * nmax + 1)
        s = term
        k = 1
        while True:
            term *= 2.0 * x / (2 * nmax + 2 * k + 1)
            s += term
            if term < 1e-17 * s:
                break
            k += 1
        out[nmax] = ex * s
        for n in range(nmax, 0, -1):
            out[n - 1] = (2.0 * x * out[n] + ex) / (2 * n - 1)
    else:
        out[0] = 0.5 * math.sqrt(math.pi / x) * math.erf(math.sqrt(x))
        for n in range(nmax):
            out[n + 1] = ((2 * n + 1) * out[n] - ex) / (2.0 * x)


@njit(cache=True)
def e_table(la, lb, a, b, Q, E):
    """E[i, j, t] (Hermite expansion coefficients of one Cartesian direction), i <= la, j <= lb"""
    p = a + b
    q = a * b / p
    E[:, :, :] = 0.0
    E[0, 0, 0] = math.exp(-q * Q * Q)
    for i in range(la):
        for t in range(i + 2):
            v = -(q * Q / a) * E[i, 0, t] if t <= i else 0.0
            if t > 0:
                v += E[i, 0, t - 1] / (2 * p)
            if t + 1 <= i:
                v += (t + 1) * E[i, 0, t + 1]
            E[i + 1, 0, t] = v
    for i in range(la + 1):
        for j in range(lb):
            for t in range(i + j + 2):
                v = (q * Q / b) * E[i, j, t] if t <= i + j else 0.0
                if t > 0:
                    v += E[i, j, t - 1] / (2 * p)
                if t + 1 <= i + j:
                    v += (t + 1) * E[i, j, t + 1]
                E[i, j + 1, t] = v


@njit(cache=True)
def r_table(L, alpha, PQ, R, F):
    """R[n, t, u, v] Hermite Coulomb integrals; R[0] is what is used"""
    r2 = PQ[0] * PQ[0] + PQ[1] * PQ[1] + PQ[2] * PQ[2]
    boys(L, alpha * r2, F)
    R[:L + 1, :L + 1, :L + 1, :L + 1] = 0.0
    f = 1.0
    for n in range(L + 1):
        R[n, 0, 0, 0] = f * F[n]
        f *= -2.0 * alpha
    for tot in range(1, L + 1):
        for t in range(tot + 1):
            for u in range(tot - t + 1):
                v = tot - t - u
                for n in range(L - tot + 1):
                    if t > 0:
                        val = PQ[0] * R[n + 1, t - 1, u, v]
                        if t > 1:
                            val += (t - 1) * R[n + 1, t - 2, u, v]
                    elif u > 0:
                        val = PQ[1] * R[n + 1, t, u - 1, v]
                        if u > 1:
                            val += (u - 1) * R[n + 1, t, u - 2, v]
                    else:
                        val = PQ[2] * R[n + 1, t, u, v - 1]
                        if v > 1:
                            val += (v - 1) * R[n + 1, t, u, v - 2]
                    R[n, t, u, v] = val


@njit(cache=True)
def _one_electron(ns, cen, sl, npr, ex, cf, off, comps, coff, ncart, atoms_z, atoms_r):
    S = np.zeros((ncart, ncart))
    T = np.zeros((ncart, ncart))
    V = np.zeros((ncart, ncart))
    L2 = 2 * LMAX + 3
    Ex = np.zeros((LMAX + 1, LMAX + 3, L2)); Ey = np.zeros((LMAX + 1, LMAX + 3, L2)); Ez = np.zeros((LMAX + 1, LMAX + 3, L2))
    R = np.zeros((L2, L2, L2, L2)); F = np.zeros(L2)
    PC = np.zeros(3)
    for sa in range(ns):
        for sb in range(ns):
            la, lb = sl[sa], sl[sb]
            A, B = cen[sa], cen[sb]
            for ia in range(npr[sa]):
                for ib in range(npr[sb]):
                    a, b = ex[sa, ia], ex[sb, ib]
                    c = cf[sa, ia] * cf[sb, ib]
                    p = a + b
                    e_table(la, lb + 2, a, b, A[0] - B[0], Ex)
                    e_table(la, lb + 2, a, b, A[1] - B[1], Ey)
                    e_table(la, lb + 2, a, b, A[2] - B[2], Ez)
                    pref = (math.pi / p) ** 1.5
                    Px = (a * A[0] + b * B[0]) / p; Py = (a * A[1] + b * B[1]) / p; Pz = (a * A[2] + b * B[2]) / p
                    for ka in range((la + 1) * (la + 2) // 2):
                        l1 = comps[coff[la] + ka, 0]; m1 = comps[coff[la] + ka, 1]; n1 = comps[coff[la] + ka, 2]
                        for kb in range((lb + 1) * (lb + 2) // 2):
                            l2 = comps[coff[lb] + kb, 0]; m2 = comps[coff[lb] + kb, 1]; n2 = comps[coff[lb] + kb, 2]
                            sx, sy, sz = Ex[l1, l2, 0], Ey[m1, m2, 0], Ez[n1, n2, 0]
                            S[off[sa] + ka, off[sb] + kb] += c * pref * sx * sy * sz
                            # kinetic: -1/2 d2/dx2 on the ket, per direction
                            tx = -2 * b * b * Ex[l1, l2 + 2, 0] + b * (2 * l2 + 1) * sx
                            if l2 >= 2:
                                tx -= 0.5 * l2 * (l2 - 1) * Ex[l1, l2 - 2, 0]
                            ty = -2 * b * b * Ey[m1, m2 + 2, 0] + b * (2 * m2 + 1) * sy
                            if m2 >= 2:
                                ty -= 0.5 * m2 * (m2 - 1) * Ey[m1, m2 - 2, 0]
                            tz = -2 * b * b * Ez[n1, n2 + 2, 0] + b * (2 * n2 + 1) * sz
                            if n2 >= 2:
                                tz -= 0.5 * n2 * (n2 - 1) * Ez[n1, n2 - 2, 0]
                            T[off[sa] + ka, off[sb] + kb] += c * pref * (tx * sy * sz + sx * ty * sz + sx * sy * tz)
                    for at in range(atoms_z.shape[0]):
                        PC[0] = Px - atoms_r[at, 0]; PC[1] = Py - atoms_r[at, 1]; PC[2] = Pz - atoms_r[at, 2]
                        L = la + lb
                        r_table(L, p, PC, R, F)
                        for ka in range((la + 1) * (la + 2) // 2):
                            l1 = comps[coff[la] + ka, 0]; m1 = comps[coff[la] + ka, 1]; n1 = comps[coff[la] + ka, 2]
                            for kb in range((lb + 1) * (lb + 2) // 2):
                                l2 = comps[coff[lb] + kb, 0]; m2 = comps[coff[lb] + kb, 1]; n2 = comps[coff[lb] + kb, 2]
                                val = 0.0
                                for t in range(l1 + l2 + 1):
                                    for u in range(m1 + m2 + 1):
                                        for v in range(n1 + n2 + 1):
                                            val += Ex[l1, l2, t] * Ey[m1, m2, u] * Ez[n1, n2, v] * R[0, t, u, v]
                                V[off[sa] + ka, off[sb] + kb] -= atoms_z[at] * c * 2.0 * math.pi / p * val
    return S, T, V


@njit(cache=True, parallel=True)
def _eri(ns, cen, sl, npr, ex, cf, off, comps, coff, ncart):
    G = np.zeros((ncart, ncart, ncart, ncart))
    npair = ns * (ns + 1) // 2
    L4 = 4 * LMAX + 1
    LP = 2 * LMAX + 1
    for pab in prange(npair):
        sa = 0
        while (sa + 1) * (sa + 2) // 2 <= pab:
            sa += 1
        sb = pab - sa * (sa + 1) // 2
        la, lb = sl[sa], sl[sb]
        na, nb = (la + 1) * (la + 2) // 2, (lb + 1) * (lb + 2) // 2
        A, B = cen[sa], cen[sb]
        Eabx = np.zeros((LMAX + 1, LMAX + 1, LP)); Eaby = np.zeros((LMAX + 1, LMAX + 1, LP)); Eabz = np.zeros((LMAX + 1, LMAX + 1, LP))
        Ecdx = np.zeros((LMAX + 1, LMAX + 1, LP)); Ecdy = np.zeros((LMAX + 1, LMAX + 1, LP)); Ecdz = np.zeros((LMAX + 1, LMAX + 1, LP))
        R = np.zeros((L4, L4, L4, L4)); F = np.zeros(L4)
        H = np.zeros((LP, LP, LP))
        PQ = np.zeros(3)
        for pcd in range(pab + 1):
            sc = 0
            while (sc + 1) * (sc + 2) // 2 <= pcd:
                sc += 1
            sd = pcd - sc * (sc + 1) // 2
            lc, ld = sl[sc], sl[sd]
            nc, nd = (lc + 1) * (lc + 2) // 2, (ld + 1) * (ld + 2) // 2
            C, D = cen[sc], cen[sd]
            blk = np.zeros((na, nb, nc, nd))
            lab, lcd = la + lb, lc + ld
            for ia in range(npr[sa]):
                for ib in range(npr[sb]):
                    a, b = ex[sa, ia], ex[sb, ib]
                    p = a + b
                    cab = cf[sa, ia] * cf[sb, ib]
                    e_table(la, lb, a, b, A[0] - B[0], Eabx)
                    e_table(la, lb, a, b, A[1] - B[1], Eaby)
                    e_table(la, lb, a, b, A[2] - B[2], Eabz)
                    Px = (a * A[0] + b * B[0]) / p; Py = (a * A[1] + b * B[1]) / p; Pz = (a * A[2] + b * B[2]) / p
                    for ic in range(npr[sc]):
                        for id_ in range(npr[sd]):
                            c, d = ex[sc, ic], ex[sd, id_]
                            q = c + d
                            ccd = cf[sc, ic] * cf[sd, id_]
                            e_table(lc, ld, c, d, C[0] - D[0], Ecdx)
                            e_table(lc, ld, c, d, C[1] - D[1], Ecdy)
                            e_table(lc, ld, c, d, C[2] - D[2], Ecdz)
                            PQ[0] = Px - (c * C[0] + d * D[0]) / q
                            PQ[1] = Py - (c * C[1] + d * D[1]) / q
                            PQ[2] = Pz - (c * C[2] + d * D[2]) / q
                            alpha = p * q / (p + q)
                            r_table(lab + lcd, alpha, PQ, R, F)
                            pref = cab * ccd * 2.0 * math.pi ** 2.5 / (p * q * math.sqrt(p + q))
                            for kc in range(nc):
                                l3 = comps[coff[lc] + kc, 0]; m3 = comps[coff[lc] + kc, 1]; n3 = comps[coff[lc] + kc, 2]
                                for kd in range(nd):
                                    l4 = comps[coff[ld] + kd, 0]; m4 = comps[coff[ld] + kd, 1]; n4 = comps[coff[ld] + kd, 2]
                                    # H[t,u,v] = sum_{tau,nu,phi} (-1)^(tau+nu+phi) Ecd[tau,nu,phi] R[t+tau, u+nu, v+phi]
                                    for t in range(lab + 1):
                                        for u in range(lab + 1 - t):
                                            for v in range(lab + 1 - t - u):
                                                s = 0.0
                                                for tau in range(l3 + l4 + 1):
                                                    ex_ = Ecdx[l3, l4, tau]
                                                    for nu in range(m3 + m4 + 1):
                                                        ey_ = Ecdy[m3, m4, nu]
                                                        for phi in range(n3 + n4 + 1):
                                                            sign = -1.0 if (tau + nu + phi) & 1 else 1.0
                                                            s += sign * ex_ * ey_ * Ecdz[n3, n4, phi] * R[0, t + tau, u + nu, v + phi]
                                                H[t, u, v] = s
                                    for ka in range(na):
                                        l1 = comps[coff[la] + ka, 0]; m1 = comps[coff[la] + ka, 1]; n1 = comps[coff[la] + ka, 2]
                                        for kb in range(nb):
                                            l2 = comps[coff[lb] + kb, 0]; m2 = comps[coff[lb] + kb, 1]; n2 = comps[coff[lb] + kb, 2]
                                            val = 0.0
                                            for t in range(l1 + l2 + 1):
                                                for u in range(m1 + m2 + 1):
                                                    for v in range(n1 + n2 + 1):
                                                        val += Eabx[l1, l2, t] * Eaby[m1, m2, u] * Eabz[n1, n2, v] * H[t, u, v]
                                            blk[ka, kb, kc, kd] += pref * val
            for ka in range(na):
                for kb in range(nb):
                    for kc in range(nc):
                        for kd in range(nd):
                            x = blk[ka, kb, kc, kd]
                            i, j, k, l = off[sa] + ka, off[sb] + kb, off[sc] + kc, off[sd] + kd
                            G[i, j, k, l] = x; G[j, i, k, l] = x; G[i, j, l, k] = x; G[j, i, l, k] = x
                            G[k, l, i, j] = x; G[l, k, i, j] = x; G[k, l, j, i] = x; G[l, k, j, i] = x
    return G


def integrals(shells, atoms):
    """shells: list of (centre xyz in bohr, l, exponents, coefficients for normalised primitives); atoms: list of (Z, xyz in bohr).
    Returns S, T, V, ERI over real solid harmonics scaled to unit self-overlap, and the nuclear repulsion energy."""
    ns = len(shells)
    maxp = max(len(s[2]) for s in shells)
    cen = np.array([s[0] for s in shells], float)
    sl = np.array([s[1] for s in shells], np.int64)
    npr = np.array([len(s[2]) for s in shells], np.int64)
    ex = np.zeros((ns, maxp)); cf = np.zeros((ns, maxp))
    for n, (_, l, es, cs) in enumerate(shells):
        for k, (a, c) in enumerate(zip(es, cs)):
            ex[n, k] = a
            cf[n, k] = c * (2 * a / math.pi) ** 0.75 * (4 * a) ** (l / 2)      # radial norm of the primitive, common to the shell
    comps, coff = [], []
    for l in range(LMAX + 1):
        coff.append(len(comps))
        comps += cart_components(l)
    comps = np.array(comps, np.int64); coff = np.array(coff, np.int64)
    off = np.zeros(ns, np.int64)
    ncart = 0
    for n in range(ns):
        off[n] = ncart
        ncart += (sl[n] + 1) * (sl[n] + 2) // 2
    az = np.array([z for z, _ in atoms], float); ar = np.array([r for _, r in atoms], float)
    S, T, V = _one_electron(ns, cen, sl, npr, ex, cf, off, comps, coff, ncart, az, ar)
    G = _eri(ns, cen, sl, npr, ex, cf, off, comps, coff, ncart)
    # cartesian -> real solid harmonics, then unit self-overlap
    nsph = int(sum(2 * l + 1 for l in sl))
    U = np.zeros((ncart, nsph))
    c0 = 0
    for n in range(ns):
        M = sph_matrix(int(sl[n]))
        U[off[n]:off[n] + M.shape[0], c0:c0 + M.shape[1]] = M
        c0 += M.shape[1]
    U = U / np.sqrt(np.diag(U.T @ S @ U))[None, :]
    S = U.T @ S @ U
    T = U.T @ T @ U
    V = U.T @ V @ U
    G = np.einsum("pqrs,pi->iqrs", G, U, optimize=True)
    G = np.einsum("iqrs,qj->ijrs", G, U, optimize=True)
    G = np.einsum("ijrs,rk->ijks", G, U, optimize=True)
    G = np.einsum("ijks,sl->ijkl", G, U, optimize=True)
    enuc = sum(atoms[a][0] * atoms[b][0] / np.linalg.norm(np.array(atoms[a][1]) - np.array(atoms[b][1]))
               for a in range(len(atoms)) for b in range(a))
    return S, T, V, G, float(enuc)
