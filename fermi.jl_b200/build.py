"""In-tree build of libfermi_pt_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfermi_pt_b200.so")
SOURCES = ["fpt_api.cu"]
HEADERS = ["fpt_api_handle.inl", "fpt_api_staging.inl", "fpt_api_upload.inl", "fpt_api_compute.inl", "fpt_api_ring.inl", "fpt_api_extras.inl", "fpt_api_ao.inl", "fpt_api_diag.inl",
           "fpt_triples.cuh", "fpt_triples2.cuh", "fpt_aux_kernels.cuh", "fpt_gemm.cuh", "fpt_internal.h", "fpt_stage.h", "fpt_ptx.cuh", "fpt_layout.h",
           os.path.join("..", "..", "include", "fermi_pt_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    nv = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nv):
        raise RuntimeError("nvcc not found: cannot build libfermi_pt_b200.so")
    return nv


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


LAUNCH_REGS = 96   # fpt_triples.cuh: the setmaxnreg budget (512 x 112 + 128 x 24 <= 640 x 96) assumes this launch allocation


def check_register_budget() -> None:
    """The fused kernel redistributes its launch-time register allocation with setmaxnreg; a different allocation would make
    the consumers' setmaxnreg.inc spin forever.  ptxas settles on 96 by itself (launch bound 640 threads, 1 CTA/SM), but
    nothing in the source pins it, so the built library is checked."""
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        return
    out = subprocess.run([cuobjdump, "-res-usage", LIB], capture_output=True, text=True).stdout.splitlines()
    seen = 0
    for n, line in enumerate(out):
        if "triples_kernelILb" in line and n + 1 < len(out):
            regs = int(out[n + 1].split("REG:")[1].split()[0])
            seen += 1
            if regs != LAUNCH_REGS:
                os.remove(LIB)
                raise RuntimeError(f"triples_kernel was compiled with {regs} registers per thread, the setmaxnreg budget needs {LAUNCH_REGS}")
    if seen == 0:
        raise RuntimeError("triples_kernel not found in the built library")


def build(force: bool = False, verbose: bool = False, defines=()) -> str:
    if force or needs_build():
        cmd = ([_nvcc()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB]
               + [os.path.join(CSRC, s) for s in SOURCES])
        subprocess.check_call(cmd)
        check_register_budget()
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
