"""Shared test helpers (oracle side): golden fixtures -> per-haplotype inputs."""
from __future__ import annotations

import json
import os
from typing import Dict, List, Tuple

import numpy as np

from oracle import taskgen

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name: str):
    with open(os.path.join(GOLDEN, name)) as f:
        d = json.load(f)
    if isinstance(d, dict) and "cells_compact" in d:  # compact cohort fixture (oracle/make_golden.py::make_cohort_compact)
        d["records"] = [[[c], [[[0] if int(x) & 1 else [], [0] if int(x) & 2 else []] for x in cells]]
                        for c, cells in zip(d["csqs"], d["cells_compact"])]
    return d


def fasta_digest(records) -> str:
    """sha256 of a sample's record-sorted FASTA text -- how the large golden cohort pins the reference binary's output."""
    import hashlib

    return hashlib.sha256("".join(">%s\n%s\n" % (h, s) for h, s in sorted(tuple(r) for r in records)).encode()).hexdigest()


def assert_sample_matches_reference(cohort, smp: str, records) -> None:
    """`records` = [(header, sequence)] of one sample, any order, against what the reference binary wrote for it."""
    if smp in cohort.get("fasta", {}):
        assert sorted([list(r) for r in records]) == sorted([list(r) for r in cohort["fasta"][smp]]), smp
    if "fasta_sha256" in cohort:
        assert len(records) == cohort["fasta_records"][smp], smp
        assert fasta_digest(records) == cohort["fasta_sha256"][smp], smp


def cohort_haplotype_csqs(cohort) -> Dict[Tuple[str, int], List[str]]:
    """(sample, hap) -> csq strings in VCF record order (what decode_back hands to the grouping step)."""
    out: Dict[Tuple[str, int], List[str]] = {}
    for csqs, cells in cohort["records"]:
        for smp, (h1, h2) in zip(cohort["samples"], cells):
            for k in h1:
                out.setdefault((smp, 1), []).append(csqs[k])
            for k in h2:
                out.setdefault((smp, 2), []).append(csqs[k])
    return out


def hap_gir(csqs: List[str], refs: Dict[str, str]) -> taskgen.HaplotypeGIR:
    return taskgen.haplotype_g_rep(taskgen.haplotype_instructions(taskgen.group_muts_per_transcript(csqs), refs), refs)


def u8(s: str) -> np.ndarray:
    return np.frombuffer(s.encode("ascii"), dtype=np.uint8).copy()


def u32(s: str) -> np.ndarray:
    return np.frombuffer(s.encode("utf-32-le"), dtype=np.uint32).copy() if s else np.zeros(0, np.uint32)


def tape_to_str(a: np.ndarray) -> str:
    if a.dtype == np.uint8:
        return a.tobytes().decode("ascii")
    return a.astype("<u4").tobytes().decode("utf-32-le")


def batch_from_girs(girs: List[taskgen.HaplotypeGIR]):
    """Concatenate haplotype GIRs into the batched layout with per-haplotype ref tapes (ref_base given)."""
    from vcf2prot_b200.engine import pack_tasks

    task_begin, tasks = pack_tasks([g.tasks for g in girs])
    ref = u8("".join(g.ref for g in girs))
    alt = u8("".join(g.alt for g in girs))
    mk = lambda lens: np.concatenate([[0], np.cumsum(np.asarray(lens, dtype=np.uint64))]).astype(np.uint64)
    ref_base = mk([len(g.ref) for g in girs])
    alt_base = mk([len(g.alt) for g in girs])
    out_base = mk([g.res_len for g in girs])
    return dict(task_begin=task_begin, tasks=tasks, ref=ref, alt=alt, ref_base=ref_base, alt_base=alt_base,
                out_base=out_base)
