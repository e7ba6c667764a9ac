"""ctypes harness for the device-side Task generator (include/v2p_taskgen.h, csrc/v2p_taskgen.cu).

DeviceCatalogue uploads a cohort.Catalogue once; generate() turns per-haplotype site lists into a device-resident
batch that GpuEngine.execute_generated() runs without the Task arrays ever existing on the host."""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np

from . import _lib as L
from .engine import EngineError, GpuEngine


class DeviceCatalogue:
    def __init__(self, prot, cat, device: int = 0):
        self._lib = L.load()
        h = C.c_void_p()
        self._keep = [np.ascontiguousarray(prot.offsets, np.uint64), np.ascontiguousarray(cat.t, np.uint32),
                      np.ascontiguousarray(cat.p, np.uint32), np.ascontiguousarray(cat.cls, np.uint8),
                      np.ascontiguousarray(cat.rlen, np.uint32), np.ascontiguousarray(cat.doff, np.uint64),
                      np.ascontiguousarray(cat.dlen, np.uint32), np.ascontiguousarray(cat.pool, np.uint8)]
        p = [a.ctypes.data_as(C.c_void_p) for a in self._keep]
        st = self._lib.v2p_catalogue_create(device, prot.n_tx, p[0], cat.n, p[1], p[2], p[3], p[4], p[5], p[6], p[7],
                                            len(cat.pool), C.byref(h))
        if st:
            raise EngineError(st, "v2p_catalogue_create failed")
        self._h = h

    @classmethod
    def from_instructions(cls, tx_offsets, site_tx, code, flags, pos_ref, pos_res, length, doff, dlen, pool, device: int = 0):
        """The general catalogue (v2p_catalogue_create_ins): one reference `Instruction` per distinct mutation, sorted by
        (transcript, mutated position), computed with validate_s_state taken as true, plus the V2P_INS_* flags."""
        self = cls.__new__(cls)
        self._lib = L.load()
        arrs = [np.ascontiguousarray(tx_offsets, np.uint64), np.ascontiguousarray(site_tx, np.uint32),
                np.ascontiguousarray(code, np.uint8), np.ascontiguousarray(flags, np.uint8),
                np.ascontiguousarray(pos_ref, np.uint32), np.ascontiguousarray(pos_res, np.uint32),
                np.ascontiguousarray(length, np.uint32), np.ascontiguousarray(doff, np.uint64),
                np.ascontiguousarray(dlen, np.uint32), np.ascontiguousarray(pool, np.uint8)]
        self._keep = arrs
        p = [a.ctypes.data_as(C.c_void_p) for a in arrs]
        h = C.c_void_p()
        st = self._lib.v2p_catalogue_create_ins(device, len(arrs[0]) - 1, p[0], len(arrs[1]), p[1], p[2], p[3], p[4], p[5], p[6],
                                                p[7], p[8], p[9], len(arrs[9]), C.byref(h))
        if st:
            raise EngineError(st, "v2p_catalogue_create_ins failed")
        self._h = h
        return self

    def generate_lists(self, site_begin: np.ndarray, sites: np.ndarray, fasta: bool = False, skip_aborts: bool = False) -> L.Generated:
        """CSR site lists (site_begin[n_hap+1], sites) -> generated batch, packed layout.  skip_aborts: leave out (and
        count) transcripts the reference would abort on instead of raising V2P_ERR_TASKGEN."""
        sb, st_ = np.ascontiguousarray(site_begin, np.uint64), np.ascontiguousarray(sites, np.uint32)
        g = L.Generated()
        st = self._lib.v2p_generate_tasks(self._h, len(sb) - 1, sb.ctypes.data_as(C.c_void_p), st_.ctypes.data_as(C.c_void_p),
                                          self._flags(False, fasta) | (L.GEN_SKIP_ABORTS if skip_aborts else 0), C.byref(g))
        if st:
            raise EngineError(st, (self._lib.v2p_catalogue_last_error(self._h) or b"").decode())
        return g

    def close(self):
        if getattr(self, "_h", None):
            self._lib.v2p_catalogue_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_names(self, name_off: np.ndarray, pool: np.ndarray) -> None:
        """Transcript names for fasta=True generations (cohort.default_names gives the synthetic ones)."""
        no, pl = np.ascontiguousarray(name_off, np.uint64), np.ascontiguousarray(pool, np.uint8)
        st = self._lib.v2p_catalogue_set_names(self._h, no.ctypes.data_as(C.c_void_p), pl.ctypes.data_as(C.c_void_p))
        if st:
            raise EngineError(st, (self._lib.v2p_catalogue_last_error(self._h) or b"").decode())

    @staticmethod
    def _flags(aligned: bool, fasta: bool) -> int:
        return (L.GEN_ALIGNED if aligned else 0) | (L.GEN_FASTA if fasta else 0)

    def generate(self, hap: np.ndarray, site: np.ndarray, n_hap: int, aligned: bool = True, fasta: bool = False) -> L.Generated:
        """(hap, site) pairs sorted by (hap, site) -- what cohort.select_sites returns.  fasta=True: the result tape is
        the .fasta text (V2P_GEN_FASTA; packed layout only, set_names first)."""
        site_begin = np.zeros(n_hap + 1, np.uint64)
        np.cumsum(np.bincount(hap, minlength=n_hap), out=site_begin[1:])
        sites = np.ascontiguousarray(site, np.uint32)
        g = L.Generated()
        st = self._lib.v2p_generate_tasks(self._h, n_hap, site_begin.ctypes.data_as(C.c_void_p),
                                          sites.ctypes.data_as(C.c_void_p), self._flags(aligned, fasta), C.byref(g))
        if st:
            raise EngineError(st, (self._lib.v2p_catalogue_last_error(self._h) or b"").decode())
        return g

    def sites_from_masks(self, masks, csq_begin: np.ndarray, csq_site: np.ndarray, shape=None) -> L.SiteLists:
        """masks[n_records, n_samples, W] (the decimal FORMAT/BCSQ words) -> device CSR site lists per haplotype
        (MaskDecoder.rs:95-153 + the per-sample transpose, vcf_ds.rs:126-295).  `masks` is a numpy array, or a device
        pointer (int) together with shape=(n_records, n_samples, W)."""
        flags = 0
        if isinstance(masks, np.ndarray):
            masks = np.ascontiguousarray(masks, np.uint32)
            n_rec, n_samp, w = masks.shape
            mp = masks.ctypes.data_as(C.c_void_p)
        else:
            n_rec, n_samp, w = shape
            mp, flags = C.c_void_p(int(masks)), L.FLAG_DEVICE_PTRS
        cb = np.ascontiguousarray(csq_begin, np.uint64)
        cs = np.ascontiguousarray(csq_site, np.int32)
        out = L.SiteLists()
        st = self._lib.v2p_sites_from_masks(self._h, n_rec, n_samp, w, mp, cb.ctypes.data_as(C.c_void_p),
                                            cs.ctypes.data_as(C.c_void_p), flags, C.byref(out))
        if st:
            raise EngineError(st, (self._lib.v2p_catalogue_last_error(self._h) or b"").decode())
        return out

    def generate_from_lists(self, lists: L.SiteLists, aligned: bool = True, fasta: bool = False) -> L.Generated:
        g = L.Generated()
        st = self._lib.v2p_generate_tasks_from_lists(self._h, C.byref(lists), self._flags(aligned, fasta), C.byref(g))
        if st:
            raise EngineError(st, (self._lib.v2p_catalogue_last_error(self._h) or b"").decode())
        return g

    def read(self, dev_ptr, count: int, dtype) -> np.ndarray:
        out = np.zeros(count, dtype)
        if count:
            st = self._lib.v2p_device_read(out.ctypes.data_as(C.c_void_p), C.c_void_p(dev_ptr), out.nbytes)
            if st:
                raise EngineError(st, "v2p_device_read failed")
        return out


def execute_generated(eng: GpuEngine, g: L.Generated, validate: bool = False, aligned_layout: bool = False) -> Tuple[float, float]:
    """Run a generated (device-resident) batch; the result tape stays in g.batch.out.  -> (group_ms, copy_ms)
    aligned_layout: the batch was generated with aligned=True (V2P_FLAG_ALIGNED_LAYOUT hint)."""
    res = L.Result()
    flags = L.FLAG_DEVICE_PTRS | (L.FLAG_VALIDATE if validate else 0) | (L.FLAG_ALIGNED_LAYOUT if aligned_layout else 0)
    st = eng._lib.v2p_execute_batch(eng._h, C.byref(g.batch), flags, C.byref(res), None)
    if st:
        raise EngineError(st, eng.last_error(), res.bad_hap, res.bad_task)
    return float(res.kernel_ms), float(res.copy_ms)
