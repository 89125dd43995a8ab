"""CPU tests of the boundary: the shared library builds, loads and exports every symbol include/*.h declares; without a
GPU the product path fails loudly (no CPU fallback); the host mirror reproduces the reference's argument checking."""
import ctypes
import os
import re

import numpy as np
import pytest

import fermi_jl_b200 as fb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "fermi_pt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(fpt_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(built):
    L = ctypes.CDLL(fb.library_path())
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/fermi_pt_b200.h but not exported"
    assert set(fb.EXPORTS) <= set(names)


def test_version_and_error_strings(built):
    L = fb.load_library()
    assert b"sm_100a" in L.fpt_version()
    assert isinstance(L.fpt_last_error(), bytes)


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(fb.FermiException) as ei:
        fb.Engine(0)
    assert "no CPU fallback" in str(ei.value)
    x = fb.synth.make_inputs(2, 3, naux=3, seed=1)
    ccsd = fb.RCCSD(0.0, 0.0, 0.0, x.T1, x.T2)
    moints = fb.IntegralHelper({"OVVV": x.OVVV, "OOOV": x.OOOV, "OVOV": x.OVOV, "Fii": x.fo, "Faa": x.fv})
    with pytest.raises(fb.FermiException):
        fb.RCCSDpT(ccsd, moints, fb.B200())


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fermi.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "pt_oracle" not in src, f


def test_argument_checking_mirrors_reference():
    x = fb.synth.make_inputs(2, 3, naux=3, seed=1)
    ccsd = fb.RCCSD(0.0, 0.0, 0.0, x.T1, x.T2)
    # PerturbativeTriples.jl:51-63: anything but (RCCSD, IntegralHelper, alg) is rejected with the argument types listed
    with pytest.raises(fb.FermiException, match="invalid arguments for RCCSD\\(T\\) method"):
        fb.RCCSDpT(ccsd, "nope", fb.B200())
    # PerturbativeTriples.jl:3-11: unknown pt_alg
    old = fb.Options["pt_alg"]
    fb.Options["pt_alg"] = 7
    try:
        with pytest.raises(fb.FermiException, match="implementation number 7 not available"):
            fb.get_rpt_alg()
    finally:
        fb.Options["pt_alg"] = old
    assert isinstance(fb.get_rpt_alg(), fb.B200)
    # IntegralHelper: missing key
    with pytest.raises(fb.FermiException):
        fb.IntegralHelper({})["OVVV"]


def test_work_layout_matches_device_header():
    # host helpers mirror fpt_layout.h (tiles of 16 + remainder; items = tile triples x non-zero-weight triplets)
    assert fb.host.num_blocks(19) == 4           # vp=20 -> 2 tiles -> 4 blocks
    assert fb.host.num_triplets(5) == 5 * 6 * 7 // 6 - 5
    assert fb.host.num_items(5, 19) == 4 * 30
    assert fb.host.shard_items(10, 0, 3) == (0, 3) and fb.host.shard_items(10, 2, 3) == (6, 10)
    # pairs (2,0),(2,1),(2,2) of o=3 sit at positions [4, 10) of the reference's i>=j>=k list
    assert fb.host.pair_range_triplets(3, 3, 6) == (4, 10)
