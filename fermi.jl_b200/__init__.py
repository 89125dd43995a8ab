"""fermi.jl_b200 -- B200-native drop-in for Fermi.jl's RCCSD(T) perturbative-triples path.

    csrc/      hand-written sm_100a CUDA kernels + the C ABI (libfermi_pt_b200.so)
    host.py    host-side mirror of the reference interface (RCCSDpT, RCCSD, IntegralHelper, B200)
    _lib.py    ctypes binding of the C ABI
    synth.py   synthetic inputs in the reference's array layouts
    julia/     the Julia glue a Fermi.jl maintainer would add (same C ABI)
"""
from . import build as _build_mod
from . import synth
from ._lib import Engine, FermiException, Stats, load_library, library_path, nccl_unique_id, EXPORTS
from .host import B200, FermiSparse, IntegralHelper, Options, RCCSD, RCCSDa, RCCSDpT, RMP2_energy, RpTAlgorithm, cc_update_T2_v4_term, get_rpt_alg, output

build_library = _build_mod.build

__all__ = ["Engine", "FermiException", "FermiSparse", "Stats", "load_library", "library_path", "nccl_unique_id", "EXPORTS", "B200", "IntegralHelper",
           "Options", "RCCSD", "RCCSDa", "RCCSDpT", "RMP2_energy", "cc_update_T2_v4_term", "RpTAlgorithm", "get_rpt_alg", "output", "synth", "build_library"]
