"""Deterministic synthetic RCCSD(T) inputs in the reference's array layouts.

The reference's `ijk` and `ijk2` algorithms agree only when the inputs carry the physical index
symmetries: T2[i,j,a,b] = T2[j,i,b,a], (ia|bc) = (ia|cb), (ij|ka) = (ji|ka),
(ia|jb) = (jb|ia).  All of them hold here by construction because the conventional integral blocks
are contracted from one symmetric factor B[Q,p,q] -- which also gives the DF form
(BOO/BOV/BVV, Q fastest, `DFERI.jl:15-69`) of exactly the same problem.

Arrays are returned Fortran-ordered (Julia column-major): the first index is the fastest.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass

import numpy as np

# The five BASELINE.json configs as (o, v, naux or None)
CONFIGS = {
    "c1_h2o_dz": (5, 19, None),
    "c2_h2o_tz": (5, 53, None),
    "c3_benzene_dz_df": (15, 93, 420),
    "c4_h2o6_dz": (24, 114, None),
    "c5_synth_o40_v400": (40, 400, None),
}


@dataclass
class PTInputs:
    """What `RCCSDpT(ccsd, moints, alg)` reads (ijk.jl:24-37): ccsd.T1/T2 and moints[...]."""

    o: int
    v: int
    naux: int
    T1: np.ndarray  # (o,v)
    T2: np.ndarray  # (o,o,v,v)
    fo: np.ndarray  # (o,)   moints["Fii"]
    fv: np.ndarray  # (v,)   moints["Faa"]
    BOO: np.ndarray  # (naux,o,o)
    BOV: np.ndarray  # (naux,o,v)
    BVV: np.ndarray  # (naux,v,v)
    OVVV: np.ndarray | None = None  # (o,v,v,v)
    OOOV: np.ndarray | None = None  # (o,o,o,v)
    OVOV: np.ndarray | None = None  # (o,v,o,v)
    seed: int = 0

    def conventional(self):
        """Materialise the 4-index blocks from B exactly as DFERI.jl:88-180 does."""
        if self.OVVV is None:
            F = lambda x: np.asfortranarray(x)
            self.OVVV = F(np.einsum("Qia,Qbc->iabc", self.BOV, self.BVV, optimize=True))
            self.OOOV = F(np.einsum("Qij,Qka->ijka", self.BOO, self.BOV, optimize=True))
            self.OVOV = F(np.einsum("Qia,Qjb->iajb", self.BOV, self.BOV, optimize=True))
        return self


def default_scales(o: int, v: int, naux: int, target_e: float = 0.05):
    """Amplitude / factor scales that put |E(T)| near `target_e` Eh (1e-2..1e-1), so the
    1e-9 Eh parity gate is a ~1e-8 relative test at every shape.  Heuristic fit:
    |E(T)| ~ 1e-7 * N_trip * v^3 * (v+o) * naux * (t/0.05)^2 * (b/0.1)^4."""
    b_scale = (0.02 / np.sqrt(naux)) ** 0.5  # integrals ~ sqrt(naux) b^2 ~ 0.02
    ntrip = o * (o + 1) * (o + 2) / 6.0
    e_unit = 1e-7 * ntrip * v ** 3 * (v + o) * naux * (b_scale / 0.1) ** 4
    t_scale = 0.05 * np.sqrt(target_e / e_unit)
    return float(t_scale), float(b_scale)


def make_inputs(o: int, v: int, naux: int = 64, seed: int = 20240517, conventional: bool = True,
                t_scale: float | None = None, b_scale: float | None = None) -> PTInputs:
    rng = np.random.default_rng(seed)
    n = o + v
    ts, bs = default_scales(o, v, naux)
    t_scale = ts if t_scale is None else t_scale
    b_scale = bs if b_scale is None else b_scale
    B = b_scale * rng.standard_normal((naux, n, n))
    B = 0.5 * (B + B.transpose(0, 2, 1))
    F = np.asfortranarray
    BOO = F(B[:, :o, :o])
    BOV = F(B[:, :o, o:])
    BVV = F(B[:, o:, o:])
    T1 = F(t_scale * rng.standard_normal((o, v)))
    T2 = t_scale * rng.standard_normal((o, o, v, v))
    T2 = F(0.5 * (T2 + T2.transpose(1, 0, 3, 2)))
    fo = -np.sort(rng.uniform(0.3, 2.0, o))[::-1].copy()  # ascending orbital energies, all < -0.3
    fv = np.sort(rng.uniform(0.1, 3.0, v))
    inp = PTInputs(o, v, naux, T1, T2, fo, fv, BOO, BOV, BVV, seed=seed)
    if conventional:
        inp.conventional()
    return inp


# ---------------------------------------------------------------------------------------------
# Flat dump container: raw little-endian f64, column-major, plus a JSON header.  A Julia-side
# exporter of real `ccsd.T1/T2` + `moints[...]` writes the same files.
# ---------------------------------------------------------------------------------------------
_FIELDS = ("T1", "T2", "fo", "fv", "BOO", "BOV", "BVV", "OVVV", "OOOV", "OVOV")


def dump(inp: PTInputs, path: str) -> None:
    os.makedirs(path, exist_ok=True)
    hdr = {"o": inp.o, "v": inp.v, "naux": inp.naux, "seed": inp.seed, "arrays": {}}
    for f in _FIELDS:
        a = getattr(inp, f)
        if a is None:
            continue
        hdr["arrays"][f] = list(a.shape)
        np.asfortranarray(a, dtype="<f8").ravel(order="F").tofile(os.path.join(path, f + ".f64"))
    with open(os.path.join(path, "header.json"), "w") as fh:
        json.dump(hdr, fh)


def load(path: str) -> PTInputs:
    with open(os.path.join(path, "header.json")) as fh:
        hdr = json.load(fh)
    arrs = {}
    for f, shape in hdr["arrays"].items():
        arrs[f] = np.fromfile(os.path.join(path, f + ".f64"), dtype="<f8").reshape(shape, order="F")
    z = lambda *s: np.zeros(s, order="F")
    o, v, naux = hdr["o"], hdr["v"], hdr.get("naux", 0)
    return PTInputs(o, v, naux, arrs["T1"], arrs["T2"], arrs["fo"], arrs["fv"],
                    arrs.get("BOO", z(0, o, o)), arrs.get("BOV", z(0, o, v)), arrs.get("BVV", z(0, v, v)),
                    arrs.get("OVVV"), arrs.get("OOOV"), arrs.get("OVOV"), seed=hdr.get("seed", 0))


def make_ao_inputs(nbf: int, ndocc: int, drop_occ: int = 0, drop_vir: int = 0, naux: int = 24, seed: int = 20240517):
    """Synthetic inputs of the AO route: an 8-fold symmetric AO ERI tensor (contracted from one symmetric factor), an
    orthogonal MO coefficient matrix, and amplitudes / orbital energies for the active space
    o = ndocc - drop_occ, v = nbf - ndocc - drop_vir.  Returns (AOERI, C, T1, T2, fo, fv), Fortran-ordered."""
    rng = np.random.default_rng(seed)
    o, v = ndocc - drop_occ, nbf - ndocc - drop_vir
    ts, bs = default_scales(o, v, naux)
    B = bs * rng.standard_normal((naux, nbf, nbf))
    B = 0.5 * (B + B.transpose(0, 2, 1))
    F = np.asfortranarray
    AOERI = F(np.einsum("Qmn,Qrs->mnrs", B, B, optimize=True))
    C, _ = np.linalg.qr(rng.standard_normal((nbf, nbf)))
    T1 = F(ts * rng.standard_normal((o, v)))
    T2 = ts * rng.standard_normal((o, o, v, v))
    T2 = F(0.5 * (T2 + T2.transpose(1, 0, 3, 2)))
    fo = -np.sort(rng.uniform(0.3, 2.0, o))[::-1].copy()
    fv = np.sort(rng.uniform(0.1, 3.0, v))
    return AOERI, F(C), T1, T2, fo, fv
