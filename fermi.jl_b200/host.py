"""Host-side mirror of the reference's RCCSD(T) interface, in Python because Julia is not available in this image
(the Julia glue a Fermi.jl maintainer would add is in julia/FermiB200.jl; same C ABI underneath).

Mirrors, name for name:
  RpTAlgorithm / get_rpt_alg / RCCSDpT(x...)     src/Methods/CoupledCluster/PerturbativeTriples/PerturbativeTriples.jl:1-63
  RCCSDpT(ccsd, moints, alg)                     .../ijk.jl:20-150   <- the drop-in boundary
  RCCSD struct (fields)                          src/Methods/CoupledCluster/RCCSD/RCCSD.jl:61-69
  IntegralHelper (cache + lazy getindex)         src/Core/Integrals/IntegralHelper.jl:50-56,111-118
  Options / output / FermiException              src/Core/Options.jl:40-199, src/Core/Output.jl:35-60
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np

from ._lib import Engine, FermiException

# ---- Options (only the keys the (T) path reads; Options.jl:40-104) -------------------------------------
_DEFAULT = {"pt_alg": 4, "printstyle": "none", "output": "fermi.out", "return_ints": False, "df": False}
Options = dict(_DEFAULT)


def output(fmt: str, *args, ending: str = "\n") -> None:
    """Output.jl:35-60."""
    style = Options["printstyle"]
    text = fmt.format(*args) + ending
    if style == "none":
        return
    if style in ("file", "both"):
        with open(Options["output"], "a") as fh:
            fh.write(text)
    if style in ("repl", "both"):
        print(text, end="")
    if style not in ("repl", "file", "both"):
        raise FermiException(f"printing style not recognized: {style}. Accepted `printstyle` values: repl, file, both, and none")


# ---- algorithm singletons (PerturbativeTriples.jl:1-11,41-49) ------------------------------------------------
class RpTAlgorithm:
    pass


class B200(RpTAlgorithm):
    """The new singleton: `struct B200 <: RpTAlgorithm end`; registered as pt_alg = 4."""


class ijk(RpTAlgorithm):
    """Reference CPU algorithms are not part of this package (no CPU fallback)."""


def get_rpt_alg():
    implemented = {4: B200()}
    N = Options["pt_alg"]
    try:
        return implemented[N]
    except KeyError:
        raise FermiException(f"implementation number {N} not available for RCCSD(T).")


# ---- inputs ------------------------------------------------------------------------------------------------
@dataclass
class FermiSparse:
    """Backend/Arrays.jl:9-12: `indexes::Vector{NTuple{N,Ti}}` (zero-based, here an (nint, N) integer array) and `data`."""
    indexes: np.ndarray
    data: np.ndarray


@dataclass
class RCCSD:
    """RCCSD.jl:61-69."""
    guessenergy: float
    correlation: float
    energy: float
    T1: np.ndarray
    T2: np.ndarray
    e_conv: float = 0.0
    t_conv: float = 0.0


class IntegralHelper:
    """The slice of IntegralHelper the (T) path touches: a string-keyed cache of MO arrays plus `eri_type`.
    For a DF helper (eri_type in {"RIFIT","JKFIT"}, ERITypes.jl:19) the 4-index keys are *not* materialised
    (the reference would do so lazily, DFERI.jl:88-180); the B200 path consumes BOO/BOV/BVV directly."""

    def __init__(self, cache: dict | None = None, eri_type: str = "Chonky", aoints: "IntegralHelper | None" = None,
                 C=None, ndocc: int | None = None, drop_occ: int = 0, drop_vir: int = 0):
        """`aoints` (an AO-basis helper holding "ERI"), `C` (MO coefficients, nbf x nmo), `ndocc`, `drop_occ`, `drop_vir`:
        what the reference's dense helper uses to build MO blocks on demand (Chonky.jl:28-114: I.orbitals.C,
        I.molecule.Nα, Options drop_occ / drop_vir).  With them and no cached "OVVV" the B200 path transforms on the GPU."""
        self.cache = dict(cache or {})
        self.eri_type = eri_type
        self.aoints = aoints
        self.C = C
        self.ndocc = ndocc
        self.drop_occ = drop_occ
        self.drop_vir = drop_vir

    @property
    def has_ao_route(self) -> bool:
        return (self.eri_type == "Chonky" and self.aoints is not None and "ERI" in self.aoints and self.C is not None
                and self.ndocc is not None)

    @property
    def ao_is_sparse(self) -> bool:
        """AO helper of type SparseERI: aoints["ERI"] is a FermiSparse (Arrays.jl:9-12) -- here any object with
        `.indexes` ((nint,4) zero-based) and `.data` ((nint,))."""
        return self.has_ao_route and hasattr(self.aoints["ERI"], "indexes")

    def orbital_blocks(self):
        """(Co, Cv) = C[:, (1+drop_occ):ndocc], C[:, (ndocc+1):(nbf-drop_vir)] (Chonky.jl:38-41, 1-based there)."""
        C = np.asarray(self.C)
        nmo = C.shape[1]
        return (np.asfortranarray(C[:, self.drop_occ:self.ndocc]), np.asfortranarray(C[:, self.ndocc:nmo - self.drop_vir]))

    @property
    def is_df(self) -> bool:
        return self.eri_type in ("RIFIT", "JKFIT", "AbstractDFERI")

    def __getitem__(self, key: str):
        if key in self.cache:
            return self.cache[key]
        raise FermiException(f"Invalid key for IntegralHelper: {key} (this mirror cannot compute integrals; fill the cache)")

    def __setitem__(self, key: str, val) -> None:
        self.cache[key] = val

    def __contains__(self, key: str) -> bool:
        return key in self.cache


_ENGINES: dict = {}


def _engine(device=None) -> Engine:
    key = tuple(device) if isinstance(device, (list, tuple)) else device
    if key not in _ENGINES:
        _ENGINES[key] = Engine(device)
    return _ENGINES[key]


class RCCSDpT:
    """Result struct `RCCSDpT{T}(CCSD, energy, correction)` (PerturbativeTriples.jl:21-25) whose constructor is the
    reference's varargs entry `RCCSDpT(x...)` (:51-63): with no algorithm among the arguments the one selected by
    Options["pt_alg"] is appended; `RCCSDpT(ccsd, moints, B200())` is the boundary method (ijk.jl:20)."""

    def __init__(self, *x, device=None, engine=None):
        """device: GPU ordinal or list of ordinals (one handle per distinct value is kept and reused); engine: an existing
        `Engine` (e.g. a rank handle of a one-process-per-GPU run, for which this constructor is a collective call)."""
        algs = [a for a in x if isinstance(a, RpTAlgorithm)]
        if not algs:
            x = x + (get_rpt_alg(),)
        if not (len(x) == 3 and isinstance(x[0], RCCSD) and isinstance(x[1], IntegralHelper) and isinstance(x[2], B200)):
            args = ", ".join(type(a).__name__ for a in x[:-1])
            raise FermiException(f"invalid arguments for RCCSD(T) method: ({args})")
        ccsd, moints, _ = x
        output("\n   • Perturbative Triples Started\n")
        output("   - Contraction Engine: B200 DMMA (libfermi_pt_b200)")
        T1, T2 = ccsd.T1, ccsd.T2
        o, v = T1.shape
        if tuple(T2.shape) != (o, o, v, v):
            raise FermiException(f"invalid T2 shape {tuple(T2.shape)} for T1 shape {(o, v)}")
        fo, fv = moints["Fii"], moints["Faa"]
        eng = engine if engine is not None else _engine(device)
        output("Computing energy contribution from occupied orbitals:")
        t0 = time.perf_counter()
        # `@set precision single` (IntegralHelper{Float32}, IntegralHelper.jl:58-68): Float32 arrays go down as they are
        single = lambda *arrs: all(isinstance(a, np.ndarray) and a.dtype == np.float32 for a in arrs)
        if moints.is_df and "OVVV" not in moints:
            BOO, BOV, BVV = moints["BOO"], moints["BOV"], moints["BVV"]
            naux = BOV.shape[0]
            if single(T1, T2, BOO, BOV, BVV, fo, fv):
                Et, st = eng.triples_df_f32(o, v, naux, T1, T2, BOO, BOV, BVV, fo, fv)
            else:
                Et, st = eng.triples_df(o, v, naux, T1, T2, BOO, BOV, BVV, fo, fv)
        elif "OVVV" not in moints and moints.has_ao_route:
            # the reference would now run compute_OVVV!/OOOV!/OVOV! on the CPU (Chonky.jl:28-114); the AO tensor goes to the GPU
            Co, Cv = moints.orbital_blocks()
            if Co.shape[1] != o or Cv.shape[1] != v:
                raise FermiException(f"orbital blocks ({Co.shape[1]} occupied, {Cv.shape[1]} virtual) do not match T1 {(o, v)}")
            if moints.ao_is_sparse:    # the reference's default AO container (Sparse.jl:78-151,236-393 on the CPU)
                eri = moints.aoints["ERI"]
                Et, st = eng.triples_ao_sparse(Co.shape[0], o, v, T1, T2, eri.indexes, eri.data, Co, Cv, fo, fv)
            else:
                Et, st = eng.triples_ao(Co.shape[0], o, v, T1, T2, moints.aoints["ERI"], Co, Cv, fo, fv)
        elif single(T1, T2, moints["OVVV"], moints["OOOV"], moints["OVOV"], fo, fv):
            Et, st = eng.triples_conv_f32(o, v, T1, T2, moints["OVVV"], moints["OOOV"], moints["OVOV"], fo, fv)
        else:
            Et, st = eng.triples_conv(o, v, T1, T2, moints["OVVV"], moints["OOOV"], moints["OVOV"], fo, fv)
        t = time.perf_counter() - t0
        output("Finished in {:5.5f} s", t)
        output("Final (T) contribution: {:15.10f}", Et)
        output("CCSD(T) energy:         {:15.10f}", Et + ccsd.energy)
        self.CCSD = ccsd
        self.energy = Et + ccsd.energy
        self.correction = Et
        self.stats = st


# ---- beside the (T) path: the DF-CCSD particle-particle ladder and the MP2 energy ----------------------------------
class RCCSDa:
    """Mirror of the reference's CCSD algorithm singleton (RCCSD.jl); only the ladder term is offloaded."""


def cc_update_T2_v4_term(newT2, T1, T2, moints: IntegralHelper, alg=None, device=None, engine=None):
    """`cc_update_T2_v4_term!(newT2, T1, T2, moints::IntegralHelper{T,<:AbstractDFERI}, alg::RCCSDa)` (RCCSDHelper.jl:204-220) on the
    GPU: newT2 is updated in place with the particle-particle ladder contraction; returns the call's stats."""
    if not moints.is_df:
        raise FermiException("cc_update_T2_v4_term: the B200 path offloads the density-fitted ladder (moints must hold BVV)")
    o, v = T1.shape
    Bvv = moints["BVV"]
    eng = engine if engine is not None else _engine(device)
    return eng.ccsd_ladder_df(o, v, Bvv.shape[0], T1, T2, Bvv, newT2)


def RMP2_energy(ints: IntegralHelper, alg=None, device=None, engine=None) -> float:
    """`RMP2_energy(ints, alg)` for RHF orbitals: density-fitted (RMP2a.jl:91-143) when the helper is a DF helper, conventional
    (RMP2a.jl:146-169, from ints["OVOV"]) otherwise.  Returns the MP2 correlation energy."""
    fo, fv = ints["Fii"], ints["Faa"]
    o, v = len(fo), len(fv)
    eng = engine if engine is not None else _engine(device)
    if ints.is_df:
        output(" Computing DF-MP2 Energy!")
        Bov = ints["BOV"]
        e, _ = eng.mp2_df(o, v, Bov.shape[0], Bov, fo, fv)
    else:
        output(" Computing MP2 Energy... ", ending="")
        e, _ = eng.mp2_conv(o, v, ints["OVOV"], fo, fv)
    return e


# ---- static work list helpers (mirror of fpt_layout.h) ---------------------------------------------------------
def num_blocks(v: int) -> int:
    """Tile triples A >= B >= C of the virtual range (tiles of 16 over roundup4(v))."""
    vp = (v + 3) // 4 * 4
    nt = (vp + 15) // 16
    return nt * (nt + 1) * (nt + 2) // 6


def num_triplets(o: int) -> int:
    """Non-zero-weight triplets i >= j >= k (ijk.jl:133 gives i = j = k weight 0)."""
    return o * (o + 1) * (o + 2) // 6 - o


def num_items(o: int, v: int) -> int:
    """Items of the full work list: (non-zero-weight triplet) x (tile triple)."""
    return num_blocks(v) * num_triplets(o)


def pair_range_triplets(o: int, pr0: int, pr1: int):
    """Positions [tb, te) of pairs [pr0, pr1) (pair index = i(i+1)/2 + j) in the reference's flattened i>=j>=k triplet
    list (k fastest, ijk.jl:49,63,83) -- the window `Engine.set_triplet_window` and the oracle's t_begin/t_end take."""

    def first_triplet(pr):
        i = 0
        while (i + 1) * (i + 2) // 2 <= pr:
            i += 1
        j = pr - i * (i + 1) // 2
        return i * (i + 1) * (i + 2) // 6 + j * (j + 1) // 2

    return first_triplet(pr0), first_triplet(pr1)


def shard_items(n_items: int, rank: int, world: int):
    """Equal-count contiguous split of an item range (host-side tests); GPUs use the cost-weighted `Engine.shard_items`."""
    return n_items * rank // world, n_items * (rank + 1) // world
