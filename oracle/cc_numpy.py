"""CPU ORACLE (test infrastructure, NOT product code) -- numpy restatements of the two reference routines SURVEY.md 8(f) ranks beside
the (T) path.  Imported only by tests/ and tools/ as the checker.  Parity pin: oracle/mini_ccsd.py runs the reference's CCSD equations
(which contain this ladder term) to the energies the reference printed, and the MP2 guess energy below is the first iteration of
that run; no reference test isolates either routine.

  ladder_df / ladder_conv   cc_update_T2_v4_term!   src/Methods/CoupledCluster/RCCSD/RCCSDHelper.jl:204-220 (DF), :193-202 (VVVV)
  mp2_df / mp2_conv         RMP2_energy             src/Methods/MollerPlesset/RMP2/RMP2a.jl:91-143 (DF), :146-169 (conventional)
"""
from __future__ import annotations

import numpy as np


def ladder_df(newT2, T1, T2, Bvv):
    """RCCSDHelper.jl:204-220, loop for loop: per a, cdb[c,d,b] = Ba[Q,c] Bvv[Q,d,b]; newT2a[i,j,b] = tau[i,j,c,d] cdb[c,d,b];
    newT2[:,:,a,:] += newT2a.  Updates newT2 in place and returns it."""
    tau = T2 + np.einsum("ia,jb->ijab", T1, T1)                      # :208
    o, v = T1.shape
    for a in range(v):                                                # :214
        Ba = Bvv[:, :, a]                                             # :215
        cdb = np.einsum("Qc,Qdb->cdb", Ba, Bvv, optimize=True)        # :216
        newT2a = np.einsum("ijcd,cdb->ijb", tau, cdb, optimize=True)  # :217
        newT2[:, :, a, :] += newT2a                                   # :218
    return newT2


def ladder_conv(newT2, T1, T2, Vvvvv):
    """RCCSDHelper.jl:193-202: newT2[i,j,a,b] += tau[i,j,c,d] Vvvvv[c,a,d,b] with Vvvvv[c,a,d,b] = (ca|db)."""
    tau = T2 + np.einsum("ia,jb->ijab", T1, T1)
    newT2 += np.einsum("ijcd,cadb->ijab", tau, Vvvvv, optimize=True)
    return newT2


def mp2_df(Bov, fo, fv):
    """RMP2a.jl:91-143: per pair i <= j, Bab = Bi^T Bj (Bvo = permutedims(BOV, (1,3,2))), E += fac * sum_ab Bab[a,b] (2 Bab[a,b] -
    Bab[b,a]) / (fo[i] + fo[j] - fv[a] - fv[b]), fac = 2 for i != j."""
    o, v = len(fo), len(fv)
    Bvo = np.transpose(Bov, (0, 2, 1))                                # :92
    e_tot = 0.0
    for i in range(o):                                                # :113
        Bi = Bvo[:, :, i]
        for j in range(i, o):                                         # :120
            Bab = Bi.T @ Bvo[:, :, j]                                 # :123
            D = fo[i] + fo[j] - fv[:, None] - fv[None, :]             # :125-130
            e = float(np.sum(Bab * (2.0 * Bab - Bab.T) / D))          # :131
            e_tot += (2.0 if i != j else 1.0) * e                     # :134-135
    return e_tot


def mp2_conv(ovov, fo, fv):
    """RMP2a.jl:146-169: sum over b, a, j, i of ovov[i,a,j,b] (2 ovov[i,a,j,b] - ovov[i,b,j,a]) / (fo[i] + fo[j] - fv[a] - fv[b])."""
    D = fo[:, None, None, None] - fv[None, :, None, None] + fo[None, None, :, None] - fv[None, None, None, :]
    return float(np.sum(ovov * (2.0 * ovov - np.transpose(ovov, (0, 3, 2, 1))) / D))
