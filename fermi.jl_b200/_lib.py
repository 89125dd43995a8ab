"""ctypes binding of libfermi_pt_b200.so -- the same C ABI (include/fermi_pt_b200.h) the Julia `ccall` glue uses.
Fails loudly when the library or a GPU is missing: there is no CPU fallback on the product path."""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build as _build

_dp = ctypes.POINTER(ctypes.c_double)
_LIB = None
DEFAULT_KERNEL_VARIANT = 1   # mirrors fpt_handle::kernel_variant in csrc/fpt_api.cu


class FermiException(Exception):
    """Mirror of Fermi.Options.FermiException (Options.jl:196-199)."""


class Stats(ctypes.Structure):
    _fields_ = [("upload_ms", ctypes.c_double), ("kernel_ms", ctypes.c_double), ("total_ms", ctypes.c_double),
                ("flops", ctypes.c_double), ("h2d_bytes", ctypes.c_double), ("n_items", ctypes.c_longlong),
                ("n_triplets", ctypes.c_longlong), ("n_launches", ctypes.c_int), ("n_sm", ctypes.c_int)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = ["fpt_create", "fpt_destroy", "fpt_triples_conv", "fpt_triples_df", "fpt_upload_conv", "fpt_upload_df", "fpt_triples_ao", "fpt_upload_ao", "fpt_triples_ao_sparse", "fpt_upload_ao_sparse",
           "fpt_nccl_unique_id", "fpt_create_rank", "fpt_set_host_threads", "fpt_set_symmetric_inputs", "fpt_set_df_ring", "fpt_device_bytes", "fpt_set_deterministic", "fpt_set_adaptive_shards", "fpt_triples_conv_async", "fpt_triples_df_async", "fpt_wait", "fpt_ccsd_ladder_df", "fpt_mp2_df", "fpt_mp2_conv", "fpt_triples_conv_f32", "fpt_triples_df_f32", "fpt_gemm_bench", "fpt_last_timeline",
           "fpt_num_items", "fpt_compute", "fpt_set_triplet_window", "fpt_set_item_order", "fpt_shard_items", "fpt_fp64_peak", "fpt_set_profiling", "fpt_set_kernel_variant", "fpt_set_debug_flags", "fpt_last_profile", "fpt_dmma_sweep", "fpt_last_error", "fpt_version"]


def library_path() -> str:
    return _build.LIB


def load_library():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("FERMI_PT_B200_LIB", _build.LIB)   # diagnostics: a variant build of the same ABI
    if not os.path.exists(path):
        raise FermiException(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback)")
    L = ctypes.CDLL(path)
    vp = ctypes.c_void_p
    L.fpt_create.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(vp)]
    L.fpt_destroy.argtypes = [vp]
    L.fpt_triples_conv.argtypes = [vp, ctypes.c_int, ctypes.c_int] + [vp] * 7 + [_dp, ctypes.POINTER(Stats)]
    L.fpt_triples_df.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [vp] * 7 + [_dp, ctypes.POINTER(Stats)]
    L.fpt_upload_conv.argtypes = [vp, ctypes.c_int, ctypes.c_int] + [vp] * 7
    L.fpt_upload_df.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [vp] * 7
    L.fpt_num_items.argtypes = [vp, ctypes.POINTER(ctypes.c_longlong)]
    L.fpt_compute.argtypes = [vp, ctypes.c_longlong, ctypes.c_longlong, _dp, ctypes.POINTER(Stats)]
    L.fpt_fp64_peak.argtypes = [vp, ctypes.c_int, ctypes.c_double, _dp]
    L.fpt_set_profiling.argtypes = [vp, ctypes.c_int]
    L.fpt_set_debug_flags.argtypes = [vp, ctypes.c_int]
    L.fpt_last_profile.argtypes = [vp, _dp]
    L.fpt_dmma_sweep.argtypes = [vp, ctypes.c_int, ctypes.c_int, _dp]
    L.fpt_last_error.restype = ctypes.c_char_p
    L.fpt_version.restype = ctypes.c_char_p
    L.fpt_set_kernel_variant.argtypes = [vp, ctypes.c_int]
    L.fpt_set_triplet_window.argtypes = [vp, ctypes.c_longlong, ctypes.c_longlong]
    L.fpt_set_item_order.argtypes = [vp, ctypes.c_int]
    L.fpt_shard_items.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong)]
    L.fpt_triples_ao.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [vp] * 7 + [_dp, ctypes.POINTER(Stats)]
    L.fpt_upload_ao.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [vp] * 7
    L.fpt_triples_ao_sparse.argtypes = ([vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_longlong, vp, ctypes.c_int]
                                        + [vp] * 5 + [_dp, ctypes.POINTER(Stats)])
    L.fpt_upload_ao_sparse.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_longlong, vp, ctypes.c_int] + [vp] * 5
    L.fpt_nccl_unique_id.argtypes = [vp]
    L.fpt_create_rank.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, ctypes.POINTER(vp)]
    L.fpt_set_host_threads.argtypes = [vp, ctypes.c_int]
    L.fpt_set_symmetric_inputs.argtypes = [vp, ctypes.c_int]
    L.fpt_set_df_ring.argtypes = [vp, ctypes.c_int]
    L.fpt_set_deterministic.argtypes = [vp, ctypes.c_int]
    L.fpt_set_adaptive_shards.argtypes = [vp, ctypes.c_int]
    L.fpt_device_bytes.argtypes = [vp, _dp]
    L.fpt_triples_conv_async.argtypes = [vp, ctypes.c_int, ctypes.c_int] + [vp] * 7
    L.fpt_triples_df_async.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [vp] * 7
    L.fpt_wait.argtypes = [vp, _dp, ctypes.POINTER(Stats)]
    L.fpt_ccsd_ladder_df.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, ctypes.POINTER(Stats)]
    L.fpt_mp2_df.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, _dp, ctypes.POINTER(Stats)]
    L.fpt_mp2_conv.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, _dp, ctypes.POINTER(Stats)]
    L.fpt_triples_conv_f32.argtypes = [vp, ctypes.c_int, ctypes.c_int] + [vp] * 7 + [_dp, ctypes.POINTER(Stats)]
    L.fpt_triples_df_f32.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [vp] * 7 + [_dp, ctypes.POINTER(Stats)]
    L.fpt_gemm_bench.argtypes = [vp, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int, _dp]
    L.fpt_last_timeline.argtypes = [vp, _dp]
    for f in EXPORTS:
        if f not in ("fpt_last_error", "fpt_version"):
            getattr(L, f).restype = ctypes.c_int
    _LIB = L
    return L


def _ptr(a):
    """Pointer for an input array: numpy (host, made Fortran-contiguous f64) or a torch tensor (host or CUDA).
    Returns (address, keepalive)."""
    if hasattr(a, "data_ptr"):  # torch tensor: must already be laid out column-major by the caller (see host.as_colmajor)
        return ctypes.c_void_p(a.data_ptr()), a
    arr = np.asfortranarray(a, dtype=np.float64)
    return ctypes.c_void_p(arr.ctypes.data), arr


def nccl_unique_id() -> bytes:
    """128-byte NCCL id for `Engine(device, rank=..., world=..., nccl_id=...)`: rank 0 creates it, the launcher's control plane
    (torch.distributed, MPI, a file ...) hands it to the other ranks."""
    L = load_library()
    buf = ctypes.create_string_buffer(128)
    if L.fpt_nccl_unique_id(buf) != 0:
        raise FermiException(L.fpt_last_error().decode())
    return buf.raw


class Engine:
    """Owns one fpt_handle.  Thin, 1:1 over the C ABI."""

    def __init__(self, device=None, rank=None, world=None, nccl_id=None):
        """device: None (current device), an int, or a list of ints (single-process multi-GPU handle).
        rank / world / nccl_id: one process per GPU (fpt_create_rank); uploads and computes are then collective calls."""
        self._L = load_library()
        self._h = ctypes.c_void_p()
        if rank is not None:
            idbuf = ctypes.create_string_buffer(bytes(nccl_id), 128) if nccl_id is not None else None
            self._check(self._L.fpt_create_rank(-1 if device is None else int(device), int(rank), int(world), idbuf, ctypes.byref(self._h)))
            return
        if isinstance(device, (list, tuple)):
            devs = (ctypes.c_int * len(device))(*device)
            n = len(device)
        else:
            devs = (ctypes.c_int * 1)(device) if device is not None else None
            n = 1
        self._check(self._L.fpt_create(n, devs, ctypes.byref(self._h)))

    def _check(self, rc):
        if rc != 0:
            raise FermiException(self._L.fpt_last_error().decode())

    def close(self):
        if self._h:
            self._L.fpt_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def triples_conv(self, o, v, T1, T2, OVVV, OOOV, OVOV, fo, fv):
        ps = [_ptr(a) for a in (T1, T2, OVVV, OOOV, OVOV, fo, fv)]
        e, st = ctypes.c_double(), Stats()
        self._check(self._L.fpt_triples_conv(self._h, o, v, *[p for p, _ in ps], ctypes.byref(e), ctypes.byref(st)))
        return e.value, st.asdict()

    def triples_conv_async(self, o, v, T1, T2, OVVV, OOOV, OVOV, fo, fv):
        """Returns once the arrays have been consumed; `wait()` collects (E(T), stats)."""
        ps = [_ptr(a) for a in (T1, T2, OVVV, OOOV, OVOV, fo, fv)]
        self._check(self._L.fpt_triples_conv_async(self._h, o, v, *[p for p, _ in ps]))

    def triples_df_async(self, o, v, naux, T1, T2, BOO, BOV, BVV, fo, fv):
        ps = [_ptr(a) for a in (T1, T2, BOO, BOV, BVV, fo, fv)]
        self._check(self._L.fpt_triples_df_async(self._h, o, v, naux, *[p for p, _ in ps]))

    def wait(self):
        e, st = ctypes.c_double(), Stats()
        self._check(self._L.fpt_wait(self._h, ctypes.byref(e), ctypes.byref(st)))
        return e.value, st.asdict()

    def set_host_threads(self, n: int):
        self._check(self._L.fpt_set_host_threads(self._h, n))

    def set_adaptive_shards(self, on: bool):
        """Multi-GPU handles: refine the shard boundaries from measured kernel times over repeated calls (default on)."""
        self._check(self._L.fpt_set_adaptive_shards(self._h, int(bool(on))))

    def set_deterministic(self, on: bool):
        """Static item -> CTA deal: E(T) is bitwise reproducible from run to run (default: dynamic work counter)."""
        self._check(self._L.fpt_set_deterministic(self._h, int(bool(on))))

    def set_df_ring(self, block: int):
        """-1: never; 0: automatic; n >= 1: the DF route keeps only 3 n occupied slabs resident and re-assembles them on the fly."""
        self._check(self._L.fpt_set_df_ring(self._h, int(block)))

    def device_bytes(self) -> float:
        b = ctypes.c_double()
        self._check(self._L.fpt_device_bytes(self._h, ctypes.byref(b)))
        return b.value

    def set_symmetric_inputs(self, on: bool):
        """on (default): pageable host inputs cross PCIe as their symmetry-unique halves; off: always in full."""
        self._check(self._L.fpt_set_symmetric_inputs(self._h, int(bool(on))))

    def gemm_bench(self, M: int, N: int, K: int, reps: int = 5) -> float:
        t = ctypes.c_double()
        self._check(self._L.fpt_gemm_bench(self._h, M, N, K, reps, ctypes.byref(t)))
        return t.value

    def last_timeline(self) -> dict:
        buf = (ctypes.c_double * 8)()
        self._check(self._L.fpt_last_timeline(self._h, buf))
        names = ["host_stage_ms", "h2d_done_ms", "operands_ready_ms", "kernel_begin_ms", "kernel_end_ms", "result_sent_ms"]
        return dict(zip(names, list(buf)[:6]))

    def triples_df(self, o, v, naux, T1, T2, BOO, BOV, BVV, fo, fv):
        ps = [_ptr(a) for a in (T1, T2, BOO, BOV, BVV, fo, fv)]
        e, st = ctypes.c_double(), Stats()
        self._check(self._L.fpt_triples_df(self._h, o, v, naux, *[p for p, _ in ps], ctypes.byref(e), ctypes.byref(st)))
        return e.value, st.asdict()

    @staticmethod
    def _ptr32(a):
        arr = np.asfortranarray(a, dtype=np.float32)
        return ctypes.c_void_p(arr.ctypes.data), arr

    def triples_conv_f32(self, o, v, T1, T2, OVVV, OOOV, OVOV, fo, fv):
        """Float32 arrays as they are (half the PCIe bytes), widened on the GPU, FP64 arithmetic."""
        ps = [self._ptr32(a) for a in (T1, T2, OVVV, OOOV, OVOV, fo, fv)]
        e, st = ctypes.c_double(), Stats()
        self._check(self._L.fpt_triples_conv_f32(self._h, o, v, *[p for p, _ in ps], ctypes.byref(e), ctypes.byref(st)))
        return e.value, st.asdict()

    def triples_df_f32(self, o, v, naux, T1, T2, BOO, BOV, BVV, fo, fv):
        ps = [self._ptr32(a) for a in (T1, T2, BOO, BOV, BVV, fo, fv)]
        e, st = ctypes.c_double(), Stats()
        self._check(self._L.fpt_triples_df_f32(self._h, o, v, naux, *[p for p, _ in ps], ctypes.byref(e), ctypes.byref(st)))
        return e.value, st.asdict()

    def ccsd_ladder_df(self, o, v, naux, T1, T2, BVV, newT2):
        """newT2[i,j,a,b] += sum_cd (T2 + T1 T1)[i,j,c,d] sum_Q BVV[Q,c,a] BVV[Q,d,b]  (RCCSDHelper.jl:204-220), in place: newT2 must be a
        Fortran-ordered float64 numpy array."""
        if not (isinstance(newT2, np.ndarray) and newT2.dtype == np.float64 and newT2.flags.f_contiguous and newT2.flags.writeable):
            raise FermiException("ccsd_ladder_df: newT2 must be a writeable Fortran-ordered float64 array (it is updated in place)")
        if tuple(newT2.shape) != (o, o, v, v):
            raise FermiException(f"ccsd_ladder_df: newT2 has shape {tuple(newT2.shape)}, expected {(o, o, v, v)}")
        ps = [_ptr(a) for a in (T1, T2, BVV)]
        st = Stats()
        self._check(self._L.fpt_ccsd_ladder_df(self._h, o, v, naux, *[p for p, _ in ps], ctypes.c_void_p(newT2.ctypes.data), ctypes.byref(st)))
        return st.asdict()

    def mp2_df(self, o, v, naux, BOV, fo, fv):
        ps = [_ptr(a) for a in (BOV, fo, fv)]
        e, st = ctypes.c_double(), Stats()
        self._check(self._L.fpt_mp2_df(self._h, o, v, naux, *[p for p, _ in ps], ctypes.byref(e), ctypes.byref(st)))
        return e.value, st.asdict()

    def mp2_conv(self, o, v, OVOV, fo, fv):
        ps = [_ptr(a) for a in (OVOV, fo, fv)]
        e, st = ctypes.c_double(), Stats()
        self._check(self._L.fpt_mp2_conv(self._h, o, v, *[p for p, _ in ps], ctypes.byref(e), ctypes.byref(st)))
        return e.value, st.asdict()

    def triples_ao(self, nbf, o, v, T1, T2, AOERI, Co, Cv, fo, fv):
        """(T) from the AO-basis ERI tensor and the occupied / virtual MO coefficient blocks (GPU AO -> MO transform)."""
        ps = [_ptr(a) for a in (T1, T2, AOERI, Co, Cv, fo, fv)]
        e, st = ctypes.c_double(), Stats()
        self._check(self._L.fpt_triples_ao(self._h, nbf, o, v, *[p for p, _ in ps], ctypes.byref(e), ctypes.byref(st)))
        return e.value, st.asdict()

    def triples_ao_sparse(self, nbf, o, v, T1, T2, indexes, data, Co, Cv, fo, fv):
        """(T) from the sparse AO list: `indexes` (nint, 4) zero-based int16/int32, `data` (nint,) (FermiSparse fields)."""
        idx = np.ascontiguousarray(indexes)
        if idx.dtype not in (np.int16, np.int32):
            idx = idx.astype(np.int32)
        if idx.ndim != 2 or idx.shape[1] != 4 or idx.shape[0] != len(data):
            raise FermiException(f"invalid sparse ERI list: indexes {idx.shape}, data {np.shape(data)}")
        vals = np.ascontiguousarray(data, dtype=np.float64)
        ps = [_ptr(a) for a in (T1, T2)]
        qs = [_ptr(a) for a in (Co, Cv, fo, fv)]
        e, st = ctypes.c_double(), Stats()
        self._check(self._L.fpt_triples_ao_sparse(self._h, nbf, o, v, ps[0][0], ps[1][0], len(vals), ctypes.c_void_p(idx.ctypes.data),
                                                  idx.dtype.itemsize, ctypes.c_void_p(vals.ctypes.data), *[p for p, _ in qs],
                                                  ctypes.byref(e), ctypes.byref(st)))
        return e.value, st.asdict()

    def upload_ao(self, nbf, o, v, T1, T2, AOERI, Co, Cv, fo, fv):
        ps = [_ptr(a) for a in (T1, T2, AOERI, Co, Cv, fo, fv)]
        self._check(self._L.fpt_upload_ao(self._h, nbf, o, v, *[p for p, _ in ps]))

    def upload_conv(self, o, v, T1, T2, OVVV, OOOV, OVOV, fo, fv):
        ps = [_ptr(a) for a in (T1, T2, OVVV, OOOV, OVOV, fo, fv)]
        self._check(self._L.fpt_upload_conv(self._h, o, v, *[p for p, _ in ps]))

    def upload_df(self, o, v, naux, T1, T2, BOO, BOV, BVV, fo, fv):
        ps = [_ptr(a) for a in (T1, T2, BOO, BOV, BVV, fo, fv)]
        self._check(self._L.fpt_upload_df(self._h, o, v, naux, *[p for p, _ in ps]))

    def num_items(self) -> int:
        n = ctypes.c_longlong()
        self._check(self._L.fpt_num_items(self._h, ctypes.byref(n)))
        return n.value

    def compute(self, item_begin: int = 0, item_end: int = -1):
        e, st = ctypes.c_double(), Stats()
        self._check(self._L.fpt_compute(self._h, item_begin, item_end, ctypes.byref(e), ctypes.byref(st)))
        return e.value, st.asdict()

    def fp64_peak(self, variant: int = 0, ms_target: float = 200.0) -> float:
        t = ctypes.c_double()
        self._check(self._L.fpt_fp64_peak(self._h, variant, ms_target, ctypes.byref(t)))
        return t.value

    def set_debug_flags(self, flags: int):
        self._check(self._L.fpt_set_debug_flags(self._h, flags))

    def set_triplet_window(self, t_begin: int = 0, t_end: int = -1):
        """Restrict the work list to positions [t_begin, t_end) of the reference's flattened i>=j>=k list (k fastest)."""
        self._check(self._L.fpt_set_triplet_window(self._h, t_begin, t_end))

    def set_item_order(self, order: int):
        self._check(self._L.fpt_set_item_order(self._h, order))

    def shard_items(self, rank: int, world: int):
        """Item range of part `rank` of `world` (contiguous, equal estimated cost) for `compute`."""
        b, e = ctypes.c_longlong(), ctypes.c_longlong()
        self._check(self._L.fpt_shard_items(self._h, rank, world, ctypes.byref(b), ctypes.byref(e)))
        return b.value, e.value

    def set_kernel_variant(self, variant: int):
        self._check(self._L.fpt_set_kernel_variant(self._h, variant))
        self._variant = variant

    @property
    def kernel_variant(self) -> int:
        return getattr(self, "_variant", DEFAULT_KERNEL_VARIANT)

    def set_profiling(self, on: bool):
        self._check(self._L.fpt_set_profiling(self._h, 1 if on else 0))

    def last_profile(self):
        buf = (ctypes.c_double * 24)()
        self._check(self._L.fpt_last_profile(self._h, buf))
        names = ["wait_item", "zero", "kloops", "rmw", "energy", "total", "token_wait", "rmw_pure", "full_wait", "ov_wait",
                 "bar_pre_energy", "bar_post_energy"]
        if self.kernel_variant == 2:
            names = ["wait_item", "setup", "kloops", "park", "energy", "total", "park_wait", "wdone_wait", "full_wait", "ov_wait",
                     "_", "bar_post_energy"]
        out = dict(zip(names, list(buf)[:12]))
        if self.kernel_variant == 2:
            enames = ["wait_item", "wfree_wait", "park_full_wait", "set_barrier", "batches", "total"]
            out.update({"g3_" + k: v for k, v in zip(enames, list(buf)[12:18])})   # "g3_" = the epilogue warp of quarter 0 here
        else:
            out.update({"g3_" + k: v for k, v in zip(names, list(buf)[12:])})
        return out

    def dmma_sweep(self, ilp: int, warps_per_sm: int) -> float:
        t = ctypes.c_double()
        self._check(self._L.fpt_dmma_sweep(self._h, ilp, warps_per_sm, ctypes.byref(t)))
        return t.value
