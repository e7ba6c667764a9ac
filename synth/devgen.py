"""Synthetic-cohort site lists on the GPU (synth/devgen.cu -> synth/libv2p_synth.so) and their numpy twin.

Bench / test infrastructure: this is how bench.py streams the 50,000-sample cohort of BASELINE.json configs[2]
(100,000 haplotypes x 280k catalogue sites = 2.8e10 Bernoulli draws) without 300 s of numpy.  The cohort is a pure
function of (seed, catalogue, haplotype index), see devgen.cu, so every rank generates exactly its own range."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Tuple

import numpy as np

from . import cohort as Co

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libv2p_synth.so")
K1, K2 = np.uint64(0x9E3779B97F4A7C15), np.uint64(0xBF58476D1CE4E5B9)
_lib = None


def build() -> None:
    """nvcc for sm_100a, in-tree (the .so travels to the GPU box with the snapshot)."""
    src = os.path.join(HERE, "devgen.cu")
    if os.path.isfile(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return
    subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                           "-Xcompiler", "-fPIC", "-shared", "-o", LIB, src])


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB):
            build()
        lib = C.CDLL(LIB)
        P, U64 = C.c_void_p, C.c_uint64
        lib.synth_gen_create.argtypes = [C.c_int, U64, P, P, P]
        lib.synth_gen_create.restype = P
        lib.synth_gen_destroy.argtypes = [P]
        lib.synth_gen_destroy.restype = None
        lib.synth_gen_lists.argtypes = [P, U64, U64, U64, P, P, U64, C.POINTER(U64)]
        lib.synth_probe_store_ms.argtypes = [C.c_int, P, U64, C.c_int, C.c_int, C.POINTER(C.c_float)]
        _lib = lib
    return _lib


def catalogue_tables(cat: Co.Catalogue) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(thresh u32, tx_first u32, trunc u8) of a catalogue: the generator's whole view of it."""
    thresh = np.floor(cat.af.astype(np.float64) * 4294967296.0).astype(np.uint32)
    newt = np.ones(cat.n, bool)
    newt[1:] = cat.t[1:] != cat.t[:-1]
    tx_first = np.flatnonzero(newt)[np.cumsum(newt) - 1].astype(np.uint32)
    trunc = np.isin(cat.cls, (Co.CLS_F, Co.CLS_G, Co.CLS_L, Co.CLS_0)).astype(np.uint8)
    return thresh, tx_first, trunc


def _draw(seed: int, h: np.ndarray, i: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + h.astype(np.uint64) * K1 + i.astype(np.uint64) * K2
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return (z >> np.uint64(32)).astype(np.uint32)


def site_lists_numpy(cat: Co.Catalogue, seed: int, h0: int, n_hap: int) -> Tuple[np.ndarray, np.ndarray]:
    """The twin: (hap, site) pairs (hap relative to h0) sorted by (hap, site) -- what cohort.build_batch takes.
    The truncation rule is applied here exactly as devgen.cu::kept does (build_batch would apply it again: idempotent)."""
    thresh, _, trunc = catalogue_tables(cat)
    haps, sites = [], []
    step = max(1, (16 << 20) // max(cat.n, 1))
    idx = np.arange(cat.n, dtype=np.uint64)
    for a in range(0, n_hap, step):
        b = min(n_hap, a + step)
        hh = (np.arange(a, b, dtype=np.uint64) + np.uint64(h0))[:, None]
        carried = _draw(seed, hh, idx[None, :]) < thresh[None, :]
        r, s = np.nonzero(carried)
        # truncating carried sites strictly before me on the same (haplotype, transcript)
        t = cat.t[s]
        newg = np.ones(len(s), bool)
        newg[1:] = (r[1:] != r[:-1]) | (t[1:] != t[:-1])
        tr = trunc[s].astype(np.int64)
        ct = np.cumsum(tr)
        before = ct - tr - (ct - tr)[np.flatnonzero(newg)][np.cumsum(newg) - 1]
        keep = before == 0
        haps.append(r[keep] + a)
        sites.append(s[keep])
    return np.concatenate(haps).astype(np.int64), np.concatenate(sites).astype(np.int64)


class DeviceCohort:
    """Site lists of any haplotype range of the cohort (seed, cat), produced in HBM."""

    def __init__(self, cat: Co.Catalogue, seed: int, device: int = 0):
        import torch

        self._lib = load()
        self._torch = torch
        self.seed, self.device = int(seed), device
        self._tables = catalogue_tables(cat)
        p = [a.ctypes.data_as(C.c_void_p) for a in self._tables]
        self._h = self._lib.synth_gen_create(device, cat.n, p[0], p[1], p[2])
        if not self._h:
            raise RuntimeError("synth_gen_create failed (no CUDA device?)")
        self._dev = torch.device("cuda", device)
        self._sites = torch.empty(1 << 20, dtype=torch.int32, device=self._dev)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.synth_gen_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def lists(self, h0: int, n_hap: int):
        """-> (site_begin: torch int64[n_hap+1] on the device, sites: torch int32 view on the device, n_sites).
        The buffers are reused by the next call."""
        torch = self._torch
        begin = torch.empty(n_hap + 1, dtype=torch.int64, device=self._dev)
        n = C.c_uint64(0)
        for _ in range(2):
            rc = self._lib.synth_gen_lists(self._h, self.seed, h0, n_hap, C.c_void_p(begin.data_ptr()),
                                           C.c_void_p(self._sites.data_ptr()), self._sites.numel(), C.byref(n))
            if rc == 1:  # grow and repeat
                self._sites = torch.empty(int(n.value * 1.2) + 1024, dtype=torch.int32, device=self._dev)
                continue
            break
        if rc != 0:
            raise RuntimeError("synth_gen_lists failed (rc %d)" % rc)
        return begin, self._sites[: n.value], int(n.value)


def store_ceiling_gbs(device: int, dev_ptr: int, nbytes: int, reps: int = 5):
    """Write-only GB/s of this GPU with our own store-only kernels over [dev_ptr, dev_ptr + nbytes) (contents are
    overwritten): {"memset": ..., "tma_bulk_store_8k": ...}.  The second is the instruction, tile size and grid the
    engine's copy kernel writes its result tape with (full sweep: profiles/dev/store_probe.cu)."""
    lib = load()
    nbytes = (nbytes // 8192) * 8192
    out = {}
    for name, mode in (("memset", 0), ("tma_bulk_store_8k", 1)):
        ms = C.c_float(0)
        if lib.synth_probe_store_ms(device, C.c_void_p(dev_ptr), nbytes, mode, reps, C.byref(ms)) != 0:
            raise RuntimeError("synth_probe_store_ms failed")
        out[name] = nbytes / (ms.value * 1e-3) / 1e9
    return out
