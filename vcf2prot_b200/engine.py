"""Host-side mirror of the reference's engine interface, bound to the CUDA library through the C ABI.

Mirrors (paths relative to /root/reference/src/data_structures/InternalRep):
  engines.rs:15-30   enum Engine {ST, MT, GPU} + FromStr          -> Engine / Engine.from_str
  task.rs:2-19       struct Task {exe_code,start_pos,length,start_pos_res} -> Task
  gir.rs:15-46       struct GIR {g_rep, annotation, alt_stream, ref_stream, res_array} -> GIR
  gir.rs:197-241     GIR::execute(self, Engine) -> (res_array, annotation)   -> GIR.execute
The GPU arm (gir.rs:236-239, a panic in the reference's CPU build) is what this package implements; ST and
MT stay the caller's business, exactly as in the reference, so GIR.execute refuses them here.

This module is a thin ctypes harness for tests and bench.py -- a Rust caller binds include/v2p_engine.h
directly.  No result byte is ever computed in Python.
"""
from __future__ import annotations

import ctypes as C
import enum
from typing import Dict, List, NamedTuple, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L


class EngineError(RuntimeError):
    """Raised where the reference panics / returns Err on this path; carries the C-ABI status."""

    def __init__(self, status: int, msg: str = "", bad_hap: int = 0, bad_task: int = 0):
        super().__init__("%s (status %d %s)" % (msg, status, L.STATUS_NAMES.get(status, "?")))
        self.status = status
        self.bad_hap = bad_hap
        self.bad_task = bad_task


class Engine(enum.Enum):
    ST = L.ENGINE_ST
    MT = L.ENGINE_MT
    GPU = L.ENGINE_GPU

    @staticmethod
    def from_str(name: str) -> "Engine":
        """engines.rs:20-29 through the ABI's own parser (v2p_engine_from_str)."""
        kind = C.c_int(-1)
        st = L.load().v2p_engine_from_str(name.encode("utf-8"), C.byref(kind))
        if st != L.V2P_OK:
            raise EngineError(st, "%s is not a supported engine" % name)
        return Engine(kind.value)


class Task(NamedTuple):
    exe_code: int
    start_pos: int
    length: int
    start_pos_res: int


def _utf32(s) -> np.ndarray:
    if isinstance(s, np.ndarray):
        return np.ascontiguousarray(s, dtype=np.uint32)
    return np.frombuffer(s.encode("utf-32-le"), dtype=np.uint32).copy() if len(s) else np.zeros(0, np.uint32)


def pack_tasks(per_hap_tasks: Sequence[Sequence[Tuple[int, int, int, int]]]) -> Tuple[np.ndarray, np.ndarray]:
    """[(exe_code,start_pos,length,start_pos_res)] per haplotype -> (task_begin u64[n+1], tasks u32[n,4])
    in v2p_task16 field order (src_off, len, dst_off, stream)."""
    counts = np.array([len(t) for t in per_hap_tasks], dtype=np.uint64)
    task_begin = np.zeros(len(per_hap_tasks) + 1, dtype=np.uint64)
    np.cumsum(counts, out=task_begin[1:])
    flat = [t for hap in per_hap_tasks for t in hap]
    tasks = np.zeros((len(flat), 4), dtype=np.uint32)
    if flat:
        a = np.asarray(flat, dtype=np.uint64)
        tasks[:, 0] = a[:, 1]
        tasks[:, 1] = a[:, 2]
        tasks[:, 2] = a[:, 3]
        tasks[:, 3] = a[:, 0]
    return task_begin, tasks


class GpuEngine:
    """One v2p_engine context (one per GPU)."""

    def __init__(self, device: int = 0):
        self._lib = L.load()
        h = C.c_void_p()
        st = self._lib.v2p_engine_create(device, C.byref(h))
        if st != L.V2P_OK:
            raise EngineError(st, "v2p_engine_create(%d) failed: no usable CUDA device (there is no CPU fallback)" % device)
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.v2p_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ------------------------------------------------------------------ misc
    def last_error(self) -> str:
        return (self._lib.v2p_last_error(self._h) or b"").decode("utf-8", "replace")

    def launch_count(self) -> int:
        return int(self._lib.v2p_kernel_launch_count(self._h))

    def set_tuning(self, variant: int = -1, ctas_per_sm: int = 0):
        st = self._lib.v2p_engine_set_tuning(self._h, variant, ctas_per_sm)
        if st:
            raise EngineError(st, "set_tuning")

    def profile_warps(self, on: bool = True):
        st = self._lib.v2p_engine_profile_warps(self._h, int(on))
        if st:
            raise EngineError(st, "profile_warps")

    def read_warp_ns(self) -> np.ndarray:
        """Wall time (ns) of every warp of the copy grid in the last device-pointer launch group (profile_warps on)."""
        n = C.c_uint64(0)
        self._lib.v2p_engine_read_warp_ns(self._h, None, 0, C.byref(n))
        out = np.zeros(n.value, np.uint64)
        if n.value:
            st = self._lib.v2p_engine_read_warp_ns(self._h, out.ctypes.data_as(C.c_void_p), n.value, C.byref(n))
            if st:
                raise EngineError(st, self.last_error())
        return out

    def set_stream(self, cuda_stream: Optional[int]):
        st = self._lib.v2p_engine_set_stream(self._h, C.c_void_p(cuda_stream or 0))
        if st:
            raise EngineError(st, "set_stream")

    def set_reference(self, ref, mode: str = "replicas") -> None:
        """Register the proteome tape (numpy uint8 host array or a torch CUDA uint8 tensor); batches that pass
        ref=None then index it.  mode: "replicas" (16 shifted copies + TMA bulk copies, default) or "plain"
        (one copy, register path only)."""
        mflag = {"replicas": 0, "plain": L.REF_NO_TMA}[mode]
        if hasattr(ref, "data_ptr"):
            st = self._lib.v2p_engine_set_reference(self._h, C.c_void_p(ref.data_ptr()), int(ref.numel()),
                                                    L.FLAG_DEVICE_PTRS | mflag)
        else:
            ref = np.ascontiguousarray(ref, dtype=np.uint8)
            st = self._lib.v2p_engine_set_reference(self._h, ref.ctypes.data_as(C.c_void_p), len(ref), mflag)
        if st:
            raise EngineError(st, self.last_error())

    # ------------------------------------------------------------------ (i) GIR::execute(Engine::GPU)
    def execute_soa(self, tasks: Sequence[Tuple[int, int, int, int]], ref, alt, res, fill_dot: bool = True,
                    validate: bool = False, engine: Engine = Engine.GPU) -> np.ndarray:
        """`res` is either a length (with fill_dot) or an existing UTF-32 array / str that is updated in place
        semantics (returned as a new array)."""
        a = np.asarray(list(tasks), dtype=np.uint64).reshape(-1, 4)
        code = np.ascontiguousarray(a[:, 0])
        sp = np.ascontiguousarray(a[:, 1])
        ln = np.ascontiguousarray(a[:, 2])
        spr = np.ascontiguousarray(a[:, 3])
        r = _utf32(ref)
        al = _utf32(alt)
        if isinstance(res, (int, np.integer)):
            out = np.zeros(int(res), dtype=np.uint32)
        else:
            out = _utf32(res).copy()
        flags = (L.FLAG_FILL_DOT if fill_dot else 0) | (L.FLAG_VALIDATE if validate else 0)
        bad = C.c_uint64(0)
        p = lambda x: x.ctypes.data_as(C.c_void_p)
        st = self._lib.v2p_gir_execute(self._h, engine.value, len(code), p(code), p(sp), p(ln), p(spr), p(r), len(r),
                                       p(al), len(al), p(out), len(out), flags, C.byref(bad))
        if st != L.V2P_OK:
            raise EngineError(st, self.last_error(), 0, bad.value)
        return out

    # ------------------------------------------------------------------ (ii) batched native call
    def execute_batch(self, task_begin: np.ndarray, tasks: np.ndarray, ref: np.ndarray, alt: np.ndarray,
                      alt_base: np.ndarray, out_base: np.ndarray, ref_base: Optional[np.ndarray] = None,
                      validate: bool = False, out: Optional[np.ndarray] = None,
                      aligned_layout: bool = False) -> Tuple[np.ndarray, float]:
        """Host (numpy) buffers in, host result tape out.  Returns (out u8[out_base[-1]-out_base[0]...], kernel_ms).
        aligned_layout: V2P_FLAG_ALIGNED_LAYOUT, a performance hint for phase-aligned producers (never changes results)."""
        task_begin = np.ascontiguousarray(task_begin, dtype=np.uint64)
        tasks = np.ascontiguousarray(tasks, dtype=np.uint32)
        ref = None if ref is None else np.ascontiguousarray(ref, dtype=np.uint8)
        alt = np.ascontiguousarray(alt, dtype=np.uint8)
        alt_base = np.ascontiguousarray(alt_base, dtype=np.uint64)
        out_base = np.ascontiguousarray(out_base, dtype=np.uint64)
        n_hap = len(task_begin) - 1
        if out is None:
            out = np.zeros(int(out_base[-1]) if n_hap >= 0 and len(out_base) else 0, dtype=np.uint8)
        b = L.Batch()
        p = lambda x: x.ctypes.data_as(C.c_void_p) if x is not None and x.size else None
        b.task_begin, b.tasks, b.ref = p(task_begin), p(tasks), p(ref)
        if ref_base is not None:
            ref_base = np.ascontiguousarray(ref_base, dtype=np.uint64)
            b.ref_base = p(ref_base)
        b.n_ref = 0 if ref is None else len(ref)
        b.alt, b.alt_base, b.out, b.out_base = p(alt), p(alt_base), p(out), p(out_base)
        b.n_hap = max(n_hap, 0)
        res = L.Result()
        flags = (L.FLAG_VALIDATE if validate else 0) | (L.FLAG_ALIGNED_LAYOUT if aligned_layout else 0)
        st = self._lib.v2p_execute_batch(self._h, C.byref(b), flags, C.byref(res), None)
        if st != L.V2P_OK:
            raise EngineError(st, self.last_error(), res.bad_hap, res.bad_task)
        self.last_copy_ms = float(res.copy_ms)
        return out, float(res.kernel_ms)

    def execute_hap_range(self, h0: int, h1: int, task_begin: np.ndarray, tasks: np.ndarray, ref: np.ndarray,
                          alt: np.ndarray, alt_base: np.ndarray, out_base: np.ndarray, out_chunk: np.ndarray,
                          validate: bool = False, wait: bool = True, aligned_layout: bool = False):
        """Host-pointer call on haplotypes [h0,h1) of a larger host-resident cohort (the streaming shape: the
        caller's pinned staging buffer `out_chunk` receives just this range's result tapes).  The base arrays keep
        their cohort-absolute values; the library rebases on entry [0] of each slice."""
        b = L.Batch()
        addr = lambda x: x.ctypes.data
        b.task_begin = addr(task_begin) + 8 * h0
        b.tasks = addr(tasks)
        b.ref, b.ref_base, b.n_ref = (None, None, 0) if ref is None else (addr(ref), None, len(ref))
        b.alt = addr(alt)
        b.alt_base = addr(alt_base) + 8 * h0
        b.out_base = addr(out_base) + 8 * h0
        o0, o1 = int(out_base[h0]), int(out_base[h1])
        assert out_chunk.size >= o1 - o0 and out_chunk.dtype == np.uint8
        b.out = addr(out_chunk) - o0  # virtual base: out[o0] is out_chunk[0]
        b.n_hap = h1 - h0
        res = L.Result()
        flags = (L.FLAG_VALIDATE if validate else 0) | (L.FLAG_ALIGNED_LAYOUT if aligned_layout else 0)
        if not wait:  # returns an event; up to 3 host-pointer batches may be in flight (copy-back overlaps upload)
            ev = C.c_void_p()
            st = self._lib.v2p_execute_batch(self._h, C.byref(b), flags | L.FLAG_ASYNC, None, C.byref(ev))
            if st != L.V2P_OK:
                raise EngineError(st, self.last_error())
            return ev
        st = self._lib.v2p_execute_batch(self._h, C.byref(b), flags, C.byref(res), None)
        if st != L.V2P_OK:
            raise EngineError(st, self.last_error(), res.bad_hap, res.bad_task)
        self.last_copy_ms = float(res.copy_ms)
        return float(res.kernel_ms)

    def execute_batch_device(self, n_hap: int, task_begin, tasks, ref, alt, alt_base, out, out_base, n_tasks: int,
                             n_alt: int, n_out: int, ref_base=None, validate: bool = False, wait: bool = True,
                             aligned_layout: bool = False):
        """Device-resident buffers (torch CUDA tensors or raw device pointers).  Returns kernel_ms (wait=True)
        or an opaque event handle to pass to wait_event()."""
        ptr = lambda x: None if x is None else (x if isinstance(x, int) else x.data_ptr())
        b = L.Batch()
        b.task_begin, b.tasks, b.ref, b.ref_base = ptr(task_begin), ptr(tasks), ptr(ref), ptr(ref_base)
        b.n_ref = int(ref.numel()) if hasattr(ref, "numel") else 0  # ref=None -> the registered reference
        b.alt, b.alt_base, b.out, b.out_base = ptr(alt), ptr(alt_base), ptr(out), ptr(out_base)
        b.n_hap, b.n_tasks, b.n_alt, b.n_out = n_hap, n_tasks, n_alt, n_out
        flags = L.FLAG_DEVICE_PTRS | (L.FLAG_VALIDATE if validate else 0) | (L.FLAG_ALIGNED_LAYOUT if aligned_layout else 0)
        res = L.Result()
        if wait:
            st = self._lib.v2p_execute_batch(self._h, C.byref(b), flags, C.byref(res), None)
            if st != L.V2P_OK:
                raise EngineError(st, self.last_error(), res.bad_hap, res.bad_task)
            self.last_copy_ms = float(res.copy_ms)
            return float(res.kernel_ms)
        ev = C.c_void_p()
        st = self._lib.v2p_execute_batch(self._h, C.byref(b), flags | L.FLAG_ASYNC, None, C.byref(ev))
        if st != L.V2P_OK:
            raise EngineError(st, self.last_error())
        return ev

    def wait_event(self, ev) -> float:
        res = L.Result()
        st = self._lib.v2p_event_wait(self._h, ev, C.byref(res))
        if st != L.V2P_OK:
            raise EngineError(st, self.last_error(), res.bad_hap, res.bad_task)
        self.last_copy_ms = float(res.copy_ms)
        return float(res.kernel_ms)


class GIR:
    """gir.rs:15-46.  `execute` consumes the representation like the reference (gir.rs:197)."""

    def __init__(self, g_rep: List[Task], annotation: Dict[str, Tuple[int, int]], alt_stream, ref_stream, res_array):
        self.g_rep = [Task(*t) for t in g_rep]
        self.annotation = dict(annotation)
        self.alt_stream = alt_stream
        self.ref_stream = ref_stream
        self.res_array = res_array  # UTF-32 array / str pre-filled by the producer ('.' at haplotype_instruction.rs:78)

    def get_tasks(self) -> List[Task]:
        return self.g_rep

    def get_annotation(self) -> Dict[str, Tuple[int, int]]:
        return self.annotation

    def get_results_max(self) -> int:  # gir.rs:157-168
        return max([v[1] for v in self.annotation.values()], default=0)

    def execute(self, engine: Engine, gpu: Optional[GpuEngine] = None, validate: bool = False):
        """-> (res_array as UTF-32 np.uint32, annotation).  `validate` is the DEBUG_GPU / DEBUG_CPU_EXEC check."""
        if engine is not Engine.GPU:
            raise EngineError(L.ERR_NOT_GPU_ENGINE, "ST/MT engines are the caller's CPU path (gir.rs:201-235)")
        if gpu is None:
            raise EngineError(L.ERR_INVALID_ARG, "Engine.GPU needs a GpuEngine context")
        res = gpu.execute_soa(self.g_rep, self.ref_stream, self.alt_stream, self.res_array, fill_dot=False,
                              validate=validate)
        return res, self.annotation
