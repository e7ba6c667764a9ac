"""ctypes harness for the whole-cohort pipeline (include/v2p_pipeline.h, csrc/v2p_pipeline.cu): per-haplotype site
lists (or the FORMAT/BCSQ mask matrix) in, per-sample .fasta / .fasta.gz file images out, everything in between on
the device -- what parts/exec.rs:27-41 + parts/io.rs:45-57 do per proband in the reference."""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Tuple

import numpy as np

from . import _lib as L
from .engine import EngineError, GpuEngine
from .taskgen import DeviceCatalogue


class DirWriter:
    """The reference's writer as a native sink (parts/io.rs:35-57): `{out_dir}/{proband}.fasta[.gz]`, one per sample."""

    def __init__(self, out_dir: str, proband_names: List[str], compressed: bool, threads: int = 8):
        self._lib = L.load()
        arr = (C.c_char_p * len(proband_names))(*[n.encode() for n in proband_names])
        h = C.c_void_p()
        st = self._lib.v2p_dir_writer_create(out_dir.encode(), arr, len(proband_names), int(compressed), threads, C.byref(h))
        if st:
            raise EngineError(st, "v2p_dir_writer_create failed")
        self._h = h
        self.sink = L.FILE_SINK(C.cast(self._lib.v2p_dir_writer_sink, C.c_void_p).value)

    @property
    def bytes_written(self) -> int:
        return int(self._lib.v2p_dir_writer_bytes(self._h))

    @property
    def files_written(self) -> int:
        return int(self._lib.v2p_dir_writer_files(self._h))

    def last_error(self) -> str:
        return (self._lib.v2p_dir_writer_last_error(self._h) or b"").decode()

    def close(self):
        if getattr(self, "_h", None):
            self._lib.v2p_dir_writer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DevicePipeline:
    """`eng` must have the proteome registered.  One DeviceCatalogue per lane (chunk in flight) is created here from
    (prot, cat) and `names` = (name_off[n_tx+1], pool), the transcript names of the FASTA headers -- or pass `cats`:
    ready catalogue objects (e.g. DeviceCatalogue.from_instructions, the general catalogue), one per lane, names set."""

    def __init__(self, eng: GpuEngine, prot=None, cat=None, names=None, lanes: int = 2, device: int = 0, cats=None):
        self._lib = L.load()
        self._eng = eng
        if cats is not None:
            self._cats, lanes = list(cats), len(cats)
        else:
            self._cats = [DeviceCatalogue(prot, cat, device) for _ in range(lanes)]
            if names is None:
                raise ValueError("DevicePipeline needs the transcript names (name_off, pool) of the FASTA headers")
            for c in self._cats:
                c.set_names(*names)
        arr = (C.c_void_p * lanes)(*[c._h for c in self._cats])
        h = C.c_void_p()
        st = self._lib.v2p_pipeline_create(eng._h, arr, lanes, C.byref(h))
        if st:
            raise EngineError(st, "v2p_pipeline_create failed")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.v2p_pipeline_destroy(self._h)
            self._h = None
            for c in self._cats:
                c.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def enable_all_records(self, proteome: np.ndarray, tx_offsets: np.ndarray, names) -> None:
        """Prepare all_records=True runs (the reference's `-a`): registers the extended reference tape on the engine."""
        pr, off = np.ascontiguousarray(proteome, np.uint8), np.ascontiguousarray(tx_offsets, np.uint64)
        no, pl = np.ascontiguousarray(names[0], np.uint64), np.ascontiguousarray(names[1], np.uint8)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        st = self._lib.v2p_pipeline_enable_all_records(self._h, p(pr), len(pr), len(off) - 1, p(off), p(no), p(pl))
        if st:
            raise EngineError(st, self._err())

    def _err(self) -> str:
        return (self._lib.v2p_pipeline_last_error(self._h) or b"").decode()

    @staticmethod
    def _sink(fn):
        """-> (function pointer, user pointer, object to keep alive)"""
        if fn is None:
            return L.FILE_SINK(), None, None
        if isinstance(fn, DirWriter):  # native sink, no Python in the loop
            return fn.sink, fn._h, fn

        def cb(_user, first, n, data, fb):
            try:
                begins = np.ctypeslib.as_array(fb, shape=(n + 1,))
                view = np.ctypeslib.as_array(C.cast(data, C.POINTER(C.c_uint8)), shape=(max(int(begins[n]), 1),))
                return int(fn(int(first), int(n), view[: int(begins[n])], begins) or 0)
            except Exception:  # never unwind through the C frame
                import traceback

                traceback.print_exc()
                return 1

        return L.FILE_SINK(cb), None, cb

    def _dest(self, out, n_samples):
        if out is None:
            return None, 0, None, None
        fb = np.zeros(n_samples + 1, np.uint64)
        return out.ctypes.data_as(C.c_void_p), out.nbytes, fb.ctypes.data_as(C.c_void_p), fb

    def run_lists(self, site_begin: np.ndarray, sites: np.ndarray, n_samples: int, chunk_samples: int = 128, gzip: bool = False,
                  out: Optional[np.ndarray] = None, sink: Optional[Callable] = None,
                  skip_aborts: bool = False, all_records: bool = False) -> Tuple[Optional[np.ndarray], L.PipelineResult]:
        """site_begin[2*n_samples+1], sites: the cohort's CSR site lists.  Either `out` (uint8 host array; returns the
        file_begin offsets into it) or `sink(first_sample, n, data, file_begin)` called per chunk in sample order."""
        sb, st_ = np.ascontiguousarray(site_begin, np.uint64), np.ascontiguousarray(sites, np.uint32)
        op, cap, fbp, fb = self._dest(out, n_samples)
        cb, user, keep = self._sink(sink)
        res = L.PipelineResult()
        st = self._lib.v2p_pipeline_run_lists(self._h, n_samples, sb.ctypes.data_as(C.c_void_p), st_.ctypes.data_as(C.c_void_p),
                                              chunk_samples, (L.PIPE_GZIP if gzip else 0) | (L.PIPE_SKIP_ABORTS if skip_aborts else 0) |
                                              (L.PIPE_ALL_RECORDS if all_records else 0),
                                              op, cap, fbp, cb, user, C.byref(res))
        del keep
        if st:
            raise EngineError(st, self._err())
        return fb, res

    def run_masks(self, masks, csq_begin: np.ndarray, csq_site: np.ndarray, shape=None, chunk_samples: int = 128, gzip: bool = False,
                  out: Optional[np.ndarray] = None, sink: Optional[Callable] = None) -> Tuple[Optional[np.ndarray], L.PipelineResult]:
        """masks[n_records, n_samples, W]: numpy array, or a device pointer (int) with shape=(n_records, n_samples, W)."""
        mflags = 0
        if isinstance(masks, np.ndarray):
            masks = np.ascontiguousarray(masks, np.uint32)
            n_rec, n_samp, w = masks.shape
            mp = masks.ctypes.data_as(C.c_void_p)
        else:
            n_rec, n_samp, w = shape
            mp, mflags = C.c_void_p(int(masks)), L.FLAG_DEVICE_PTRS
        cbeg, csite = np.ascontiguousarray(csq_begin, np.uint64), np.ascontiguousarray(csq_site, np.int32)
        op, cap, fbp, fb = self._dest(out, n_samp)
        cb, user, keep = self._sink(sink)
        res = L.PipelineResult()
        st = self._lib.v2p_pipeline_run_masks(self._h, n_rec, n_samp, w, mp, cbeg.ctypes.data_as(C.c_void_p),
                                              csite.ctypes.data_as(C.c_void_p), mflags, chunk_samples, L.PIPE_GZIP if gzip else 0,
                                              op, cap, fbp, cb, user, C.byref(res))
        del keep
        if st:
            raise EngineError(st, self._err())
        return fb, res


def csr_lists(hap: np.ndarray, site: np.ndarray, n_hap: int) -> Tuple[np.ndarray, np.ndarray]:
    """(hap, site) pairs sorted by (hap, site) -> (site_begin[n_hap+1], sites)."""
    site_begin = np.zeros(n_hap + 1, np.uint64)
    np.cumsum(np.bincount(hap, minlength=n_hap), out=site_begin[1:])
    return site_begin, np.ascontiguousarray(site, np.uint32)
