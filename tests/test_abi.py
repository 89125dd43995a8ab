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


# ---- the Julia glue cannot be executed here (no Julia in the image): check its ccalls against the header mechanically -----------------
def _split_top(s):
    """split at top-level commas (brackets, braces and parentheses nest)"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _balanced(s, i):
    """s[i] == '(' -> index just past the matching ')'"""
    depth = 0
    for k in range(i, len(s)):
        depth += s[k] == "("
        depth -= s[k] == ")"
        if depth == 0:
            return k + 1
    raise AssertionError("unbalanced")


def _header_prototypes():
    import re
    h = re.sub(r"/\*.*?\*/", " ", open(os.path.join(ROOT, "include", "fermi_pt_b200.h")).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int|double|const char\s*\*)\s+(fpt_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", h, flags=re.S):
        params = [] if m.group(3).strip() in ("", "void") else [" ".join(p.split()) for p in _split_top(m.group(3))]
        protos[m.group(2)] = (" ".join(m.group(1).split()), [re.sub(r"\s*\b[A-Za-z_][A-Za-z0-9_]*$", "", p).strip() for p in params])
    return h, protos


_C_TO_JULIA = {"int": {"Cint"}, "long long": {"Clonglong"}, "double": {"Cdouble"},
               "const double*": {"Ptr{Cdouble}"}, "double*": {"Ptr{Cdouble}", "Ref{Cdouble}"},
               "const float*": {"Ptr{Cfloat}"}, "const int*": {"Ptr{Cint}"}, "const void*": {"Ptr{Cvoid}"},
               "fpt_handle*": {"Ptr{Cvoid}"}, "fpt_handle**": {"Ref{Ptr{Cvoid}}"}, "fpt_stats*": {"Ref{FptStats}", "Ptr{Cvoid}"},
               "const char*": {"Cstring"}}


def test_julia_glue_ccalls_match_the_header():
    import re
    jl = open(os.path.join(ROOT, "fermi.jl_b200", "julia", "FermiB200.jl")).read()
    jl = "\n".join(line.split("#")[0] if "ccall" not in line.split("#")[0] and line.lstrip().startswith("#") else line for line in jl.splitlines())
    _, protos = _header_prototypes()
    seen = 0
    for m in re.finditer(r"ccall\(\(:(fpt_[a-z0-9_]+), LIB\),\s*([A-Za-z]+),\s*\(", jl):
        name, ret = m.group(1), m.group(2)
        assert name in protos, f"{name} is not declared in include/fermi_pt_b200.h"
        cret, cparams = protos[name]
        assert ret in _C_TO_JULIA[cret], (name, ret, cret)
        t0 = m.end() - 1
        t1 = _balanced(jl, t0)
        jtypes = _split_top(jl[t0 + 1:t1 - 1])
        call_end = _balanced(jl, m.start() + len("ccall"))
        args = _split_top(jl[t1:call_end - 1].lstrip().lstrip(","))
        assert len(jtypes) == len(cparams), (name, jtypes, cparams)
        assert len(args) == len(jtypes), (name, len(args), len(jtypes), args)
        for jt, ct in zip(jtypes, cparams):
            key = ct.replace(" *", "*")
            assert key in _C_TO_JULIA, (name, ct)
            assert jt in _C_TO_JULIA[key], f"{name}: Julia passes {jt} where the header declares {ct}"
        seen += 1
    assert seen >= 12


def test_julia_stats_struct_matches_the_header():
    import re
    h, _ = _header_prototypes()
    body = re.search(r"typedef struct fpt_stats \{(.*?)\} fpt_stats;", h, flags=re.S).group(1)
    cfields = [(" ".join(f.split()[:-1]), f.split()[-1]) for f in (x.strip() for x in body.split(";")) if f]
    jl = open(os.path.join(ROOT, "fermi.jl_b200", "julia", "FermiB200.jl")).read()
    jbody = re.search(r"struct FptStats\n(.*?)\nend", jl, flags=re.S).group(1)
    jfields = [tuple(f.strip().split("::")) for f in re.split(r"[;\n]", jbody) if f.strip()]
    assert [n for _, n in cfields] == [n for n, _ in jfields]
    for (ct, n), (_, jt) in zip(cfields, jfields):
        assert jt in _C_TO_JULIA[ct], (n, ct, jt)
