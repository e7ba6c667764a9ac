"""ORACLE (test infrastructure only): restatement of the reference's genotype bit-mask decode and transpose.

  data_structures/MaskDecoder.rs:95-121   parse_single_field   one u32: bit 2i -> haplotype 1 carries csq i,
                                                               bit 2i+1 -> haplotype 2 carries csq i  (i = 0..15)
  data_structures/MaskDecoder.rs:122-153  parse_concat_values  several words: csq index = 15*word + i
  data_structures/vcf_ds.rs:278-295       extract_effects      indices -> the record's csq strings
  data_structures/vcf_ds.rs:126-213       get_patient_fields / get_csq_per_patient: the per-sample transpose
  functions/vcf_tools.rs:82-96 + vcf_ds.rs:442-479             per transcript grouping, sort by position,
                                                               identical duplicates dropped

Pinned by the reference's own unit-test vectors (MaskDecoder.rs:160-400, tests/test_maskdecode.py) and, end to end,
by the cohort golden files (the reference binary decodes the same masks from the VCF text).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def get_indices(words: Sequence[int]) -> Tuple[List[int], List[int]]:
    """BitMask::get_indices (MaskDecoder.rs:57-153) for one sample cell."""
    h1: List[int] = []
    h2: List[int] = []
    if len(words) == 1:
        b, idx = int(words[0]), 0
        while b:
            if b & 1:
                h1.append(idx)
            if (b >> 1) & 1:
                h2.append(idx)
            b >>= 2
            idx += 1
        return h1, h2
    base = 0
    for w in words:
        b, idx = int(w), 0
        while b:
            if b & 1:
                h1.append(base + idx)
            if (b >> 1) & 1:
                h2.append(base + idx)
            b >>= 2
            idx += 1
        base += 15
    return h1, h2


def site_lists(masks: np.ndarray, csq_begin: np.ndarray, csq_site: np.ndarray):
    """masks[n_records, n_samples, W] -> per-haplotype ascending, duplicate-free catalogue-site lists in CSR form
    (haplotype = 2*sample + {0,1}).  csq k of record r maps to catalogue site csq_site[csq_begin[r] + k]; -1 = a csq
    the tool does not support (Constants::SUP_TYPE filter, vcf_ds.rs:249,262)."""
    n_rec, n_samp, _ = masks.shape
    per_hap = [[] for _ in range(2 * n_samp)]
    for r in range(n_rec):
        n_csq = int(csq_begin[r + 1] - csq_begin[r])
        for s in range(n_samp):
            if not masks[r, s].any():
                continue
            h1, h2 = get_indices(masks[r, s])
            for hapbit, idxs in ((0, h1), (1, h2)):
                for k in idxs:
                    if k >= n_csq:
                        raise IndexError("mask selects csq %d of a record with %d (vcf_ds.rs:287 panics)" % (k, n_csq))
                    site = int(csq_site[int(csq_begin[r]) + k])
                    if site >= 0:
                        per_hap[2 * s + hapbit].append(site)
    site_begin = np.zeros(2 * n_samp + 1, np.uint64)
    flat: List[int] = []
    for h, lst in enumerate(per_hap):
        u = sorted(set(lst))
        flat.extend(u)
        site_begin[h + 1] = len(flat)
    return site_begin, np.asarray(flat, np.uint32)
