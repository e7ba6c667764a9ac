"""ctypes harness for the device gzip writer (include/v2p_gzip.h, csrc/v2p_gzip.cu): one complete gzip member per
output file from the FASTA image the engine leaves in HBM -- the `-c` path of the reference
(personalized_genome.rs:87-101, :135-170)."""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import numpy as np

from . import _lib as L
from .engine import EngineError

CHUNK = 16384


class DeviceGzip:
    def __init__(self, device: int = 0):
        self._lib = L.load()
        h = C.c_void_p()
        st = self._lib.v2p_gzip_create(device, C.byref(h))
        if st:
            raise EngineError(st, "v2p_gzip_create failed (no CUDA device? there is no CPU fallback)")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.v2p_gzip_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bound(self, in_bytes: int, n_files: int) -> int:
        return int(self._lib.v2p_gzip_bound(in_bytes, n_files))

    def _call(self, in_ptr, file_begin, out_ptr, cap, flags) -> Tuple[np.ndarray, L.GzipResult]:
        fb = np.ascontiguousarray(file_begin, np.uint64)
        ob = np.zeros(len(fb), np.uint64)
        res = L.GzipResult()
        st = self._lib.v2p_gzip_files(self._h, in_ptr, fb.ctypes.data_as(C.c_void_p), len(fb) - 1, out_ptr, cap,
                                      ob.ctypes.data_as(C.c_void_p), flags, C.byref(res))
        if st:
            raise EngineError(st, (self._lib.v2p_gzip_last_error(self._h) or b"").decode())
        return ob, res

    def compress(self, data: np.ndarray, file_begin, capacity: int = None) -> Tuple[List[bytes], L.GzipResult]:
        """Host buffers in, host buffers out: the list of gzip file images."""
        data = np.ascontiguousarray(data, np.uint8)
        fb = np.asarray(file_begin, np.uint64)
        cap = self.bound(int(fb[-1] - fb[0]), len(fb) - 1) if capacity is None else capacity
        out = np.zeros(max(cap, 1), np.uint8)
        ob, res = self._call(data.ctypes.data_as(C.c_void_p), fb, out.ctypes.data_as(C.c_void_p), cap, 0)
        return [out[int(ob[i]):int(ob[i + 1])].tobytes() for i in range(len(fb) - 1)], res

    def compress_device(self, in_ptr: int, file_begin, out_ptr: int, capacity: int) -> Tuple[np.ndarray, L.GzipResult]:
        """Device pointers (e.g. the result tape of execute_batch_device): -> (out_begin, stats); bytes stay in HBM."""
        return self._call(C.c_void_p(int(in_ptr)), file_begin, C.c_void_p(int(out_ptr)), capacity, L.FLAG_DEVICE_PTRS)
