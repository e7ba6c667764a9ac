"""ORACLE (test infrastructure only): ctypes wrapper of oracle/ref_engine.c (libv2p_oracle.so)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libv2p_oracle.so")

REF_OK, REF_ERR_RES_OOB, REF_ERR_SRC_OOB, REF_ERR_BAD_STREAM, REF_ERR_NOT_CONTIGUOUS = 0, 1, 2, 3, 4
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "all"])


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB):
            build()
        lib = C.CDLL(LIB)
        P, U64 = C.c_void_p, C.c_uint64
        soa = [U64, P, P, P, P, P, U64, P, U64, P, U64, C.c_int, C.c_int, C.POINTER(U64)]
        lib.ref_gir_execute_u32.argtypes = soa
        lib.ref_gir_execute_u8.argtypes = soa
        lib.ref_batch_execute.argtypes = [U64, P, P, P, P, U64, P, P, P, P, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(U64), C.POINTER(U64)]
        lib.ref_batch_check.argtypes = [U64, P, P, P, P, U64, P, P, P, P, C.c_int, C.c_int, C.POINTER(U64)]
        lib.ref_batch_check.restype = U64
        lib.ref_widen_u8_to_u32.argtypes = [P, P, U64]
        lib.ref_widen_u8_to_u32.restype = None
        _lib = lib
    return _lib


def _p(a: Optional[np.ndarray]):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def gir_execute(tasks: Sequence[Tuple[int, int, int, int]], ref: np.ndarray, alt: np.ndarray, res: np.ndarray,
                fill_dot: bool, validate: bool = False) -> Tuple[int, int]:
    """Serial GIR::execute on UTF-32 (uint32) or 1-byte (uint8) tapes, in place on `res`.
    Returns (status, bad_index)."""
    lib = load()
    a = np.asarray(list(tasks), dtype=np.uint64).reshape(-1, 4)
    cols = [np.ascontiguousarray(a[:, i]) for i in range(4)]
    fn = lib.ref_gir_execute_u32 if res.dtype == np.uint32 else lib.ref_gir_execute_u8
    assert ref.dtype == res.dtype and alt.dtype == res.dtype
    bad = C.c_uint64(0)
    st = fn(len(a), _p(cols[0]), _p(cols[1]), _p(cols[2]), _p(cols[3]), _p(ref), len(ref), _p(alt), len(alt), _p(res),
            len(res), int(fill_dot), int(validate), C.byref(bad))
    return st, bad.value


def batch_execute(task_begin: np.ndarray, tasks: np.ndarray, ref: np.ndarray, alt: np.ndarray, alt_base: np.ndarray,
                  out: np.ndarray, out_base: np.ndarray, ref_base: Optional[np.ndarray] = None, validate: bool = False,
                  threads: int = 1, fill_dot: bool = True) -> Tuple[int, int, int]:
    """Batched layout (v2p_task16 rows = src_off,len,dst_off,stream).  width follows out.dtype (u8 or u32).
    Returns (status, bad_hap, bad_index)."""
    lib = load()
    width = out.dtype.itemsize
    assert width in (1, 4) and ref.dtype == out.dtype and alt.dtype == out.dtype
    task_begin = np.ascontiguousarray(task_begin, np.uint64)
    tasks = np.ascontiguousarray(tasks, np.uint32)
    alt_base = np.ascontiguousarray(alt_base, np.uint64)
    out_base = np.ascontiguousarray(out_base, np.uint64)
    if ref_base is not None:
        ref_base = np.ascontiguousarray(ref_base, np.uint64)
    bh, bi = C.c_uint64(0), C.c_uint64(0)
    st = lib.ref_batch_execute(len(task_begin) - 1, _p(task_begin), _p(tasks), _p(ref), _p(ref_base), len(ref), _p(alt),
                               _p(alt_base), _p(out), _p(out_base), width, int(fill_dot), int(validate), int(threads),
                               C.byref(bh), C.byref(bi))
    return st, bh.value, bi.value


def batch_check(task_begin: np.ndarray, tasks: np.ndarray, ref: np.ndarray, alt: np.ndarray, alt_base: np.ndarray,
                got: np.ndarray, out_base: np.ndarray, ref_base: Optional[np.ndarray] = None, threads: int = 1) -> Tuple[int, int]:
    """Whole-batch checker: every haplotype is executed into a thread-private scratch tape and compared with `got`
    (somebody else's result tape in the same layout; `got[out_base[h] - out_base[0] ...]` when the bases do not start
    at 0).  The base arrays may be slices of cohort-wide arrays: everything is rebased on entry [0].
    Returns (number of haplotypes that differ or fail, lowest such haplotype)."""
    lib = load()
    width = got.dtype.itemsize
    assert width in (1, 4) and ref.dtype == got.dtype and alt.dtype == got.dtype
    rb = lambda a: np.ascontiguousarray(a, np.uint64) - np.uint64(a[0]) if len(a) else np.ascontiguousarray(a, np.uint64)
    task_begin, alt_base, out_base = rb(task_begin), rb(alt_base), rb(out_base)
    tasks = np.ascontiguousarray(tasks, np.uint32)
    if ref_base is not None:
        ref_base = rb(ref_base)
    fb = C.c_uint64(0)
    n_bad = lib.ref_batch_check(len(task_begin) - 1, _p(task_begin), _p(tasks), _p(ref), _p(ref_base), len(ref), _p(alt),
                                _p(alt_base), _p(got), _p(out_base), width, int(threads), C.byref(fb))
    return int(n_bad), int(fb.value)
