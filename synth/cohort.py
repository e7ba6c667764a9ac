"""Synthetic phased cohorts -> packed Task batches (host-side producer, numpy-vectorised).

This is the host half of the hot path for benchmarks and parity tests: it plays the role of the reference's
instruction generation (TranscriptInstruction::get_g_rep, transcript_instructions.rs:335-427, :452-780) and
haplotype concatenation (HaplotypeInstruction::get_g_rep, haplotype_instruction.rs:75-158) for the seven csq
classes the synthetic cohorts of SURVEY.md section 8(d) use:

    missense ('M')  inframe_insertion ('I')  inframe_deletion ('D')  frameshift ('F')
    stop_gained ('G')  stop_lost ('L')  start_lost ('0')

and emits exactly the Task tuples the reference would (tests/test_cohort_taskgen.py checks them, tuple for tuple
and byte for byte of the alt tape, against the reference-pinned restatement on the same sites).  Only
reference-valid combinations are generated: unique, well separated positions per transcript; nothing after a
truncating class on the same haplotype+transcript (transcript_instructions.rs:486,496-499); a start_lost
transcript carries nothing else (transcript_instructions.rs:338-343).

Two tape layouts:
  ref_mode="per_hap"  the reference's own layout: each haplotype's ref tape is the concatenation of its altered
                      transcripts (haplotype_instruction.rs:118); Task.start_pos is relative to it.
  ref_mode="global"   B200 layout: all haplotypes share ONE proteome tape resident in HBM/L2 (v2p_batch.ref_base
                      == NULL); Task.start_pos = transcript offset in the proteome + local position.  The bytes
                      produced are identical; only the source offsets differ.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

AA = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
CLS_M, CLS_I, CLS_D, CLS_F, CLS_G, CLS_L, CLS_0 = range(7)
CLS_NAMES = ["missense", "inframe_insertion", "inframe_deletion", "frameshift", "stop_gained", "stop_lost", "start_lost"]

# class mixes of SURVEY.md section 8(d): (M, I, D, F, G, L, 0)
MIX_C2 = (0.96, 0.0075, 0.0075, 0.01, 0.008, 0.002, 0.005)
MIX_C4 = (0.35, 0.20, 0.0, 0.35, 0.0, 0.10, 0.0)


@dataclass
class Proteome:
    lengths: np.ndarray  # int64[n_tx]
    offsets: np.ndarray  # int64[n_tx+1]
    residues: np.ndarray  # uint8[total]

    @property
    def n_tx(self) -> int:
        return len(self.lengths)

    def name(self, t: int) -> str:
        return "ENST%011d" % t

    def seq(self, t: int) -> str:
        return self.residues[self.offsets[t]:self.offsets[t + 1]].tobytes().decode("ascii")


def make_proteome(seed: int = 0x5EED0001, n_tx: int = 20000, mu: float = 6.0, sigma: float = 0.75, lo: int = 30,
                  hi: int = 35000, giant: int = 0) -> Proteome:
    """P20k of SURVEY 8(d): lengths ~ clamp(LogNormal(mu, sigma)); residues uniform over 20 aa, first = 'M'.
    `giant` > 0 forces that many titin-scale (hi-residue) transcripts (skew stress)."""
    rng = np.random.default_rng(seed)
    lengths = np.clip(np.rint(rng.lognormal(mu, sigma, n_tx)), lo, hi).astype(np.int64)
    if giant:
        lengths[rng.choice(n_tx, giant, replace=False)] = hi
    offsets = np.zeros(n_tx + 1, np.int64)
    np.cumsum(lengths, out=offsets[1:])
    residues = AA[rng.integers(0, 20, int(offsets[-1]))]
    residues[offsets[:-1]] = ord("M")
    return Proteome(lengths, offsets, residues)


@dataclass
class Catalogue:
    """Variant sites sorted by (transcript, position)."""
    t: np.ndarray  # int64 transcript
    p: np.ndarray  # int64 0-based ref position (stop_lost: == ref_len; start_lost: 0)
    cls: np.ndarray  # int8
    rlen: np.ndarray  # int64 residues of the ref allele (D: deleted+anchor; others 1)
    doff: np.ndarray  # int64 offset of the instruction data in `pool`
    dlen: np.ndarray  # int64 length of the instruction data (M:1, I:anchor+ins, D:1, F/L: tail, G/0: 0)
    af: np.ndarray  # float32 allele frequency
    pool: np.ndarray  # uint8 payload bytes

    @property
    def n(self) -> int:
        return len(self.t)


def make_catalogue(prot: Proteome, n_sites: int, seed: int, mix=MIX_C2, ins_max: int = 10, del_max: int = 10,
                   fs_mean: float = 25.0, fs_max: int = 4000, sl_max: int = 100, long_ins_mean: float = 0.0,
                   long_ins_max: int = 5000, lognormal_tails: bool = False) -> Catalogue:
    rng = np.random.default_rng(seed)
    L = prot.lengths
    w = L / L.sum()
    t = rng.choice(prot.n_tx, size=n_sites, p=w)
    p = (1 + np.floor(rng.random(n_sites) * np.maximum(L[t] - 2, 1))).astype(np.int64)  # 0-based in [1, L-2]
    cls = rng.choice(7, size=n_sites, p=np.asarray(mix) / np.sum(mix)).astype(np.int8)
    p[cls == CLS_L] = L[t[cls == CLS_L]]
    p[cls == CLS_0] = 0
    # sort by (t, p); thin so that neighbours on one transcript are >= minsep apart (no overlap / engulfment)
    order = np.lexsort((p, t))
    t, p, cls = t[order], p[order], cls[order]
    minsep = del_max + 3
    keep = np.ones(n_sites, bool)
    last_t, last_p = -1, -10**9
    same = np.concatenate([[False], (t[1:] == t[:-1]) & (p[1:] - p[:-1] < minsep)])
    # greedy thinning in one pass would need a loop; the vectorised rule "drop a site closer than minsep to its
    # sorted predecessor" is conservative (may leave wider spacing) and always valid
    keep &= ~same
    t, p, cls = t[keep], p[keep], cls[keep]
    # a start_lost transcript carries nothing else: drop the other sites of those transcripts
    sl_tx = np.unique(t[cls == CLS_0])
    if len(sl_tx):
        drop = np.isin(t, sl_tx) & (cls != CLS_0)
        t, p, cls = t[~drop], p[~drop], cls[~drop]
        first = np.concatenate([[True], t[1:] != t[:-1]])
        dup0 = (cls == CLS_0) & ~first
        t, p, cls = t[~dup0], p[~dup0], cls[~dup0]
    # deletions must fit before the end of the transcript (anchor + >=1 deleted residue, something left after)
    n = len(t)
    rlen = np.ones(n, np.int64)
    d = cls == CLS_D
    room = L[t] - p - 2
    dl = 1 + rng.integers(0, del_max, n)
    bad = d & (room < 1)
    cls[bad] = CLS_M
    d = cls == CLS_D
    rlen[d] = 1 + np.minimum(dl[d], room[d])
    # instruction data lengths
    dlen = np.zeros(n, np.int64)
    dlen[cls == CLS_M] = 1
    dlen[cls == CLS_D] = 1
    i = cls == CLS_I
    if long_ins_mean > 0:
        ins = np.clip(np.rint(rng.lognormal(np.log(long_ins_mean) - 0.5, 1.0, n)), 1, long_ins_max).astype(np.int64)
    else:
        ins = 1 + rng.integers(0, ins_max, n)
    dlen[i] = 1 + ins[i]
    f = cls == CLS_F
    if lognormal_tails:
        tails = np.clip(np.rint(rng.lognormal(np.log(fs_mean) - 0.5, 1.0, n)), 1, fs_max).astype(np.int64)
    else:
        tails = np.clip(rng.geometric(1.0 / fs_mean, n), 1, fs_max).astype(np.int64)
    dlen[f] = tails[f]
    l_ = cls == CLS_L
    dlen[l_] = 1 + rng.integers(0, sl_max, n)[l_]
    doff = np.zeros(n, np.int64)
    np.cumsum(dlen[:-1], out=doff[1:])
    pool = AA[rng.integers(0, 20, int(dlen.sum()) + 1)]
    refaa = prot.residues[np.minimum(prot.offsets[t] + p, prot.offsets[-1] - 1)]
    # missense: a residue different from the reference one
    m = cls == CLS_M
    idx = np.searchsorted(AA, refaa[m])
    pool[doff[m]] = AA[(idx + 1 + rng.integers(0, 19, int(m.sum()))) % 20]
    # insertion data starts with the anchor residue, deletion data IS the anchor residue
    pool[doff[i]] = refaa[i]
    pool[doff[d]] = refaa[d]
    af = rng.choice(np.asarray([0.001, 0.01, 0.05, 0.2, 0.5], np.float32), size=n, p=[.5, .25, .15, .07, .03])
    return Catalogue(t, p, cls, rlen, doff, dlen, af.astype(np.float32), pool)


def instruction_arrays(prot: Proteome, cat: Catalogue):
    """The catalogue as reference `Instruction` values (instruction.rs:6-16) for the general device catalogue
    (v2p_catalogue_create_ins / DeviceCatalogue.from_instructions): what Instruction::from_mutation yields for the seven
    classes -- M: len 1; I: len = data length (anchor + inserted); D: len = deleted residues (ref allele - 1), data = anchor;
    F / L: len = data length; G / 0: nothing -- plus the V2P_INS_INVALIDATES flag of frameshift and stop_gained
    (instruction.rs:1083-1094).  -> positional arguments of DeviceCatalogue.from_instructions."""
    code = np.frombuffer(b"MIDFGL0", np.uint8)[cat.cls]
    length = np.select([cat.cls == CLS_M, cat.cls == CLS_D, (cat.cls == CLS_G) | (cat.cls == CLS_0)], [1, cat.rlen - 1, 0], cat.dlen)
    flags = np.where((cat.cls == CLS_F) | (cat.cls == CLS_G), 2, 0).astype(np.uint8)
    return (prot.offsets.astype(np.uint64), cat.t, code, flags, cat.p, cat.p, length, cat.doff, cat.dlen, cat.pool)


def site_csq(prot: Proteome, cat: Catalogue, i: int) -> str:
    """The bcftools/csq string of site i (for the oracle / reference-binary side of the tests)."""
    t, p, c = int(cat.t[i]), int(cat.p[i]), int(cat.cls[i])
    seq = prot.seq(t)
    data = cat.pool[cat.doff[i]:cat.doff[i] + cat.dlen[i]].tobytes().decode("ascii")
    q = p + 1
    if c == CLS_M:
        aa = "%d%s>%d%s" % (q, seq[p], q, data)
    elif c == CLS_I:
        aa = "%d%s>%d%s" % (q, seq[p], q, data)
    elif c == CLS_D:
        aa = "%d%s>%d%s" % (q, seq[p:p + int(cat.rlen[i])], q, data)
    elif c == CLS_F:
        aa = "%d%s*>%d%s*" % (q, seq[p:], q, data)
    elif c == CLS_G:
        aa = "%d%s>%d*" % (q, seq[p], q)
    elif c == CLS_L:
        aa = "%d*>%d%s*" % (q, q, data)
    else:
        aa = "1M>1K"
    return "%s|GENE|%s|protein_coding|-|%s|1936821C>T" % (CLS_NAMES[c], prot.name(t), aa)


@dataclass
class Batch:
    task_begin: np.ndarray  # u64[n_hap+1]
    tasks: np.ndarray  # u32[n_tasks,4]  (src_off, len, dst_off, stream)
    alt: np.ndarray  # u8
    alt_base: np.ndarray  # u64[n_hap+1]
    out_base: np.ndarray  # u64[n_hap+1]
    ref_base: Optional[np.ndarray]  # u64[n_hap+1] (per_hap mode) or None (global mode)
    ref: np.ndarray  # the tape the tasks index (proteome in global mode)
    # annotations: one row per altered transcript per haplotype, in tape order
    ann_hap: np.ndarray  # int64
    ann_tx: np.ndarray  # int64
    ann_start: np.ndarray  # int64 (haplotype-relative, like the reference's annotation map)
    ann_end: np.ndarray  # int64
    kept_hap: Optional[np.ndarray] = None  # the (hap, site) pairs that survived the truncation rule
    kept_site: Optional[np.ndarray] = None
    ann_ntasks: Optional[np.ndarray] = None  # tasks per annotation row (0 for a start_lost transcript)

    @property
    def n_hap(self) -> int:
        return len(self.task_begin) - 1

    @property
    def n_residues(self) -> int:
        return int(self.out_base[-1])


# ---- VCF-record view of a catalogue: the input side of v2p_sites_from_masks -------------------
@dataclass
class Records:
    """Synthetic VCF records over a catalogue: record r lists csq entries csq_begin[r]..csq_begin[r+1]; entry k names
    the catalogue site it stands for or -1 for a consequence class the tool drops (vcf_ds.rs:249,262).  site_rec /
    site_k locate every catalogue site's entry; words = FORMAT/BCSQ integers per cell (MaskDecoder.rs:122-153)."""
    csq_begin: np.ndarray  # u64 [n_rec+1]
    csq_site: np.ndarray   # i32 [n_csq]
    site_rec: np.ndarray   # i64 [cat.n]
    site_k: np.ndarray     # i64 [cat.n]
    words: int

    @property
    def n_rec(self) -> int:
        return len(self.csq_begin) - 1


def make_records(cat: Catalogue, seed: int, max_sites_per_record: int = 1, p_unsupported: float = 0.0,
                 wide_every: int = 0) -> Records:
    """Groups consecutive catalogue sites into records of 1..max_sites_per_record entries, sprinkles unsupported
    entries between them, and makes every `wide_every`-th record wider than one 15-entry mask word."""
    rng = np.random.default_rng(seed)
    csq_begin, csq_site = [0], []
    site_rec, site_k = np.zeros(cat.n, np.int64), np.zeros(cat.n, np.int64)
    i = r = 0
    while i < cat.n:
        take = int(rng.integers(1, max_sites_per_record + 1))
        if wide_every and r % wide_every == wide_every - 1:
            take = int(rng.integers(16, 40))
        entries = []
        for s in range(i, min(i + take, cat.n)):
            while p_unsupported and rng.random() < p_unsupported:
                entries.append(-1)
            site_rec[s], site_k[s] = r, len(entries)
            entries.append(s)
        i += take
        csq_site.extend(entries)
        csq_begin.append(len(csq_site))
        r += 1
    cb = np.asarray(csq_begin, np.uint64)
    widest = int(np.diff(cb.astype(np.int64)).max()) if r else 1
    words = 1 if widest <= 16 else (widest + 14) // 15
    return Records(cb, np.asarray(csq_site, np.int32), site_rec, site_k, words)


def encode_masks(rec: Records, n_samples: int, hap: np.ndarray, site: np.ndarray) -> np.ndarray:
    """(haplotype, site) carrier pairs -> masks[n_rec, n_samples, words]: what bcftools csq writes into FORMAT/BCSQ
    (bit 2k of the cell: haplotype 1 carries the record's csq k, bit 2k+1: haplotype 2; 15 csq per word when a
    record needs several words, MaskDecoder.rs:122-153)."""
    masks = np.zeros((rec.n_rec, n_samples, rec.words), np.uint32)
    k = rec.site_k[site]
    if rec.words == 1:
        w, bit = np.zeros(len(k), np.int64), 2 * k + (hap & 1)
    else:
        w, bit = k // 15, 2 * (k % 15) + (hap & 1)
    np.bitwise_or.at(masks, (rec.site_rec[site], hap >> 1, w), (np.uint32(1) << bit.astype(np.uint32)))
    return masks


def select_sites(cat: Catalogue, n_hap: int, rng: np.random.Generator) -> Tuple[np.ndarray, np.ndarray]:
    """Each haplotype carries site i with probability af[i].  Returns (hap, site), sorted by (hap, site)."""
    haps, sites = [], []
    step = max(1, (32 << 20) // max(cat.n, 1))
    for h0 in range(0, n_hap, step):
        h1 = min(n_hap, h0 + step)
        m = rng.random((h1 - h0, cat.n), dtype=np.float32) < cat.af[None, :]
        hh, ss = np.nonzero(m)
        haps.append(hh + h0)
        sites.append(ss)
    return np.concatenate(haps), np.concatenate(sites)


def build_batch(prot: Proteome, cat: Catalogue, hap: np.ndarray, site: np.ndarray, n_hap: int,
                ref_mode: str = "global", layout: str = "packed") -> Batch:
    """(hap, site) pairs sorted by (hap, transcript, position) -> Task batch.  Vectorised restatement of the
    reference's emission rules for the seven classes (file:line in the module docstring).

    layout="packed"   result tapes exactly as the reference lays them out (res_counter, haplotype_instruction.rs:132).
    layout="aligned"  B200 layout (global ref mode only): every transcript's result starts at an offset congruent
                      (mod 16) to its offset in the proteome tape, inside a 16-byte-multiple slot.  start_pos_res is
                      authoritative and uncovered bytes read '.', so the engine contract is unchanged; the consumer
                      slices by annotation and never sees the <= 30 pad bytes per transcript.  Runs of unshifted
                      (missense-only) transcripts then have source and destination in phase: one aligned TMA bulk
                      copy from replica 0, which stays L2-resident."""
    if layout == "aligned" and ref_mode != "global":
        raise ValueError("aligned layout needs the shared proteome tape")
    t, p, cls = cat.t[site], cat.p[site], cat.cls[site]
    # ---- truncation: nothing after F/G/L/0 on the same haplotype+transcript
    newg = np.ones(len(site), bool)
    newg[1:] = (hap[1:] != hap[:-1]) | (t[1:] != t[:-1])
    gid = np.cumsum(newg) - 1
    trunc = (cls == CLS_F) | (cls == CLS_G) | (cls == CLS_L) | (cls == CLS_0)
    ct = np.cumsum(trunc)
    g_first = np.flatnonzero(newg)
    before = ct - trunc - (ct - trunc)[g_first][gid]  # truncating sites strictly before me, in my group
    keep = before == 0
    hap, site, t, p, cls = hap[keep], site[keep], t[keep], p[keep], cls[keep]
    n = len(site)
    newg = np.ones(n, bool)
    newg[1:] = (hap[1:] != hap[:-1]) | (t[1:] != t[:-1])
    lastg = np.ones(n, bool)
    lastg[:-1] = newg[1:]
    gid = np.cumsum(newg) - 1
    g_first = np.flatnonzero(newg)
    n_groups = len(g_first)
    g_hap, g_tx = hap[g_first], t[g_first]
    Lr = prot.lengths[t]
    dlen, rlen, doff = cat.dlen[site], cat.rlen[site], cat.doff[site]
    p_next = np.empty(n, np.int64)
    p_next[:-1] = p[1:]
    p_next[lastg] = Lr[lastg]  # the tail runs to the end of the transcript

    # ---- ref tape origin of every group
    g_is0 = cls[g_first] == CLS_0  # start_lost groups contribute nothing (empty GIR)
    if ref_mode == "global":
        g_ref0 = prot.offsets[g_tx]
    else:
        contrib = np.where(g_is0, 0, prot.lengths[g_tx])
        g_ref0 = np.cumsum(contrib) - contrib
        hfirst = np.ones(n_groups, bool)
        hfirst[1:] = g_hap[1:] != g_hap[:-1]
        g_ref0 = g_ref0 - g_ref0[np.flatnonzero(hfirst)][np.cumsum(hfirst) - 1]
    ref0 = g_ref0[gid]

    # ---- alt tape: M pushes its residue twice (transcript_instructions.rs:659-660), G/0 push nothing
    acontrib = np.where(cls == CLS_M, 2, dlen)
    a_excl = np.cumsum(acontrib) - acontrib
    hstart = np.ones(n, bool)
    hstart[1:] = hap[1:] != hap[:-1]
    a_rel = a_excl - a_excl[np.flatnonzero(hstart)][np.cumsum(hstart) - 1]  # offset inside the haplotype's alt tape

    # ---- three candidate tasks per site: base (first site of a transcript), mutation, follow-up copy
    is0 = cls == CLS_0
    has_base = newg & ~is0
    base_len = p  # (0,0,first.pos_ref,0)  [stop_lost at pos_ref==ref_len: whole transcript, :725-728]
    has_mut = (cls == CLS_M) | (cls == CLS_I) | (cls == CLS_D) | (cls == CLS_F) | (cls == CLS_L)
    mut_len = np.where((cls == CLS_M) | (cls == CLS_D), 1, dlen)
    mut_src = a_rel + (cls == CLS_M)  # alt.len()-data.len() after the double push
    has_fol = (cls == CLS_M) | (cls == CLS_I) | (cls == CLS_D)
    fol_start = np.where(cls == CLS_D, p + rlen, p + 1)  # D: pos_ref+len+1 with len = rlen-1
    fol_len = p_next - fol_start
    if (fol_len[has_fol] < 0).any():
        raise ValueError("catalogue spacing violated (sites overlap)")
    # PHI rule of add_till_next_ins for D: pos_ref+len == next.pos_ref  (cannot happen with minsep, kept for fidelity)
    phi = (cls == CLS_D) & ~lastg & (p + rlen - 1 == p_next)
    has_fol &= ~phi

    cnt = has_base.astype(np.int64) + has_mut + has_fol
    first_slot = np.cumsum(cnt) - cnt
    n_tasks = int(cnt.sum())
    tasks = np.zeros((n_tasks, 4), np.uint32)
    ln = np.zeros(n_tasks, np.int64)
    # base
    s = first_slot[has_base]
    tasks[s, 0] = ref0[has_base]
    ln[s] = base_len[has_base]
    # mutation
    s = (first_slot + has_base)[has_mut]
    tasks[s, 0] = mut_src[has_mut]
    tasks[s, 3] = 1
    ln[s] = mut_len[has_mut]
    # follow-up
    s = (first_slot + has_base + has_mut)[has_fol]
    tasks[s, 0] = (ref0 + fol_start)[has_fol]
    ln[s] = fol_len[has_fol]
    tasks[:, 1] = ln

    # ---- per-haplotype bases; dst = running sum of lengths inside the haplotype (no gaps for these classes)
    task_hap = np.repeat(hap, cnt)
    task_gid = np.repeat(gid, cnt)
    task_begin = np.zeros(n_hap + 1, np.uint64)
    np.cumsum(np.bincount(task_hap, minlength=n_hap), out=task_begin[1:])
    l_excl = np.cumsum(ln) - ln
    g_len = np.bincount(task_gid, weights=ln, minlength=n_groups).astype(np.int64)
    hfirst = np.ones(n_groups, bool)
    hfirst[1:] = g_hap[1:] != g_hap[:-1]
    if layout == "aligned":
        # slot of a transcript: 16-byte aligned base, result at base + (proteome offset mod 16), slot = multiple of 16
        c = prot.offsets[g_tx] & 15
        slot = np.where(g_len > 0, (c + g_len + 15) & ~15, 0)
        g_base = np.cumsum(slot) - slot
        if n_groups:
            g_base = g_base - g_base[np.flatnonzero(hfirst)][np.cumsum(hfirst) - 1]
        g_start = g_base + np.where(g_len > 0, c, 0)
        res_per_hap = np.bincount(g_hap, weights=slot, minlength=n_hap).astype(np.int64)
        g_l0 = (np.cumsum(g_len) - g_len)  # l_excl of the group's first task
        tasks[:, 2] = g_start[task_gid] + (l_excl - g_l0[task_gid])
    else:
        res_per_hap = np.bincount(task_hap, weights=ln, minlength=n_hap).astype(np.int64)
        g_excl = np.cumsum(g_len) - g_len
        g_start = g_excl - g_excl[np.flatnonzero(hfirst)][np.cumsum(hfirst) - 1] if n_groups else g_excl
        tb = task_begin[:-1].astype(np.int64)
        nonempty = np.flatnonzero(task_begin[1:] > task_begin[:-1])
        hap_l0 = np.zeros(n_hap, np.int64)
        hap_l0[nonempty] = l_excl[tb[nonempty]]
        tasks[:, 2] = l_excl - hap_l0[task_hap]
    out_base = np.zeros(n_hap + 1, np.uint64)
    np.cumsum(res_per_hap, out=out_base[1:])

    # ---- alt tape.  packed: payloads in site order (the reference's alt_stream).  aligned: short payloads packed first,
    #      then every payload of >= 32 bytes (frameshift / stop-lost tails, long insertions) in a 16-byte-multiple slot
    #      at an offset congruent (mod 16) to its destination, so its full vectors are direct TMA bulk copies.
    tot = int(acontrib.sum())
    rep = np.repeat(np.arange(n), acontrib)
    within = np.arange(tot) - np.repeat(a_excl, acontrib)
    src = doff[rep] + np.where(cls[rep] == CLS_M, 0, within)
    if layout == "aligned":
        is_long = has_mut & (acontrib >= 32)
        mut_dst = np.zeros(n, np.int64)
        mut_dst[has_mut] = tasks[(first_slot + has_base)[has_mut], 2]
        c16 = mut_dst & 15
        short_c = np.where(is_long, 0, acontrib)
        sh_excl = np.cumsum(short_c) - short_c
        hidx = np.flatnonzero(hstart)
        hof = np.cumsum(hstart) - 1  # index of each site's haplotype among the non-empty ones
        sh_rel = sh_excl - sh_excl[hidx][hof]
        short_tot = np.bincount(hap, weights=short_c, minlength=n_hap).astype(np.int64)
        slot = np.where(is_long, (c16 + acontrib + 15) & ~15, 0)
        sl_excl = np.cumsum(slot) - slot
        sl_rel = sl_excl - sl_excl[hidx][hof]
        a_new = np.where(is_long, ((short_tot[hap] + 15) & ~15) + sl_rel + c16, sh_rel)
        alt_per_hap = ((short_tot + 15) & ~15) + np.bincount(hap, weights=slot, minlength=n_hap).astype(np.int64)
        tasks[(first_slot + has_base)[has_mut], 0] = (a_new + (cls == CLS_M))[has_mut]
        alt_base = np.zeros(n_hap + 1, np.uint64)
        np.cumsum(alt_per_hap, out=alt_base[1:])
        alt = np.full(int(alt_base[-1]), ord("."), np.uint8)
        if tot:
            alt[alt_base[:-1].astype(np.int64)[hap[rep]] + a_new[rep] + within] = cat.pool[src]
    else:
        alt_per_hap = np.bincount(hap, weights=acontrib, minlength=n_hap).astype(np.int64)
        alt_base = np.zeros(n_hap + 1, np.uint64)
        np.cumsum(alt_per_hap, out=alt_base[1:])
        alt = cat.pool[src] if tot else np.zeros(0, np.uint8)

    # ---- annotations: (start,end) of every altered transcript on its haplotype's result tape = (g_start, +g_len)

    if ref_mode == "global":
        ref_base, ref = None, prot.residues
    else:
        contrib = np.where(g_is0, 0, prot.lengths[g_tx])
        per = np.bincount(g_hap, weights=contrib, minlength=n_hap).astype(np.int64)
        ref_base = np.zeros(n_hap + 1, np.uint64)
        np.cumsum(per, out=ref_base[1:])
        live = ~g_is0
        lens = prot.lengths[g_tx][live]
        tot = int(lens.sum())
        rep = np.repeat(prot.offsets[g_tx][live], lens)
        within = np.arange(tot) - np.repeat(np.cumsum(lens) - lens, lens)
        ref = prot.residues[rep + within] if tot else np.zeros(0, np.uint8)
    g_ntasks = np.bincount(gid, weights=cnt, minlength=n_groups).astype(np.int64)
    return Batch(task_begin, tasks, alt, alt_base, out_base, ref_base, ref, g_hap, g_tx, g_start, g_start + g_len,
                 hap, site, g_ntasks)


def synth_batch(prot: Proteome, cat: Catalogue, n_hap: int, seed: int, ref_mode: str = "global",
                layout: str = "packed") -> Batch:
    rng = np.random.default_rng(seed)
    hap, site = select_sites(cat, n_hap, rng)
    return build_batch(prot, cat, hap, site, n_hap, ref_mode, layout)


def concat_batches(parts: List[Batch]) -> Batch:
    """Concatenate batches generated chunk by chunk (same proteome, global ref mode)."""
    assert all(b.ref_base is None for b in parts)
    def cat_base(name):
        out, acc = [np.zeros(1, np.uint64)], np.uint64(0)
        for b in parts:
            a = getattr(b, name)
            out.append(a[1:] + acc)
            acc = acc + a[-1]
        return np.concatenate(out)
    hap_off = np.cumsum([0] + [b.n_hap for b in parts])[:-1]
    return Batch(cat_base("task_begin"), np.concatenate([b.tasks for b in parts]), np.concatenate([b.alt for b in parts]),
                 cat_base("alt_base"), cat_base("out_base"), None, parts[0].ref,
                 np.concatenate([b.ann_hap + o for b, o in zip(parts, hap_off)]),
                 np.concatenate([b.ann_tx for b in parts]), np.concatenate([b.ann_start for b in parts]),
                 np.concatenate([b.ann_end for b in parts]),
                 np.concatenate([b.kept_hap + o for b, o in zip(parts, hap_off)]),
                 np.concatenate([b.kept_site for b in parts]),
                 np.concatenate([b.ann_ntasks for b in parts]) if parts[0].ann_ntasks is not None else None)


def fasta_records(prot: Proteome, batch: Batch, out: np.ndarray, h: int, hap_label: int) -> List[Tuple[str, str]]:
    """Consumer contract (sequence_tape.rs:77-89 + personalized_genome.rs:97): slice haplotype h's result tape by
    annotation -> [(">{transcript}_{1|2}" without '>', sequence)]."""
    o0 = int(batch.out_base[h])
    rows = np.flatnonzero(batch.ann_hap == h)
    recs = []
    for r in rows:
        s, e = int(batch.ann_start[r]), int(batch.ann_end[r])
        recs.append(("%s_%d" % (prot.name(int(batch.ann_tx[r])), hap_label), out[o0 + s:o0 + e].tobytes().decode("ascii")))
    return recs


def default_names(prot: Proteome) -> Tuple[np.ndarray, np.ndarray]:
    """Transcript names as a (name_off[n_tx+1], pool) tape: `ENST%011d` (Proteome.name)."""
    n = prot.n_tx
    pool = np.zeros((n, 15), np.uint8)
    pool[:, 0:4] = np.frombuffer(b"ENST", np.uint8)
    pool[:, 4:15] = (np.arange(n)[:, None] // 10 ** np.arange(10, -1, -1)[None, :]) % 10 + ord("0")
    return (15 * np.arange(n + 1)).astype(np.uint64), pool.reshape(-1)


def fasta_image(prot: Proteome, b: Batch, names: Optional[Tuple[np.ndarray, np.ndarray]] = None) -> Batch:
    """SURVEY 8(f) rank 1 -- FASTA record formatting on the device, with NO new kernel: the record framing of
    `write_altered_only` (personalized_genome.rs:97,107: `>{transcript}_{1|2}\n{seq}\n`) becomes two more copy
    segments per transcript, fed from a name tape appended to the haplotype's alt tape (one entry
    `>{name}_{1|2}\n\n` of len(name)+5 bytes per record, at offset e behind the alteration bytes):

        header task  (1, n_alt + e,               len(name)+4, record start)      ">ENST00000000042_1\n"
        ... the transcript's own tasks, shifted ...
        newline task (1, n_alt + e + len(name)+4, 1,           after the sequence)

    so haplotype h's result tape IS the text of its records and `out[out_base[2s] : out_base[2s+2]]` is sample s's
    .fasta file image (hap-1 records, then hap-2 records; the reference's own record order is HashMap-random).
    Needs the packed layout (a file image cannot contain pad bytes) and haplotype index = 2*sample + (hap-1).
    `names` = (name_off[n_tx+1], pool) gives the transcript names (default: Proteome.name)."""
    if b.ref_base is not None or b.ann_ntasks is None:
        raise ValueError("fasta_image needs a global-ref batch built by build_batch")
    name_off, name_pool = default_names(prot) if names is None else names
    name_off = np.asarray(name_off).astype(np.int64)
    G, n_hap, n_old = len(b.ann_hap), b.n_hap, len(b.tasks)
    c = b.ann_ntasks
    nlen = (name_off[1:] - name_off[:-1])[b.ann_tx]  # name length of every record
    elen = nlen + 5                                    # its entry on the name tape
    gstart = np.cumsum(c) - c
    gid_of_task = np.repeat(np.arange(G), c)
    new_idx = np.arange(n_old) + 2 * gid_of_task + 1
    hdr_idx = gstart + 2 * np.arange(G)
    nl_idx = gstart + c + 2 * np.arange(G) + 1
    # e = offset of the record's entry inside its haplotype's name tape
    hfirst = np.ones(G, bool)
    hfirst[1:] = b.ann_hap[1:] != b.ann_hap[:-1]
    e_excl = np.cumsum(elen) - elen
    e = e_excl - e_excl[np.flatnonzero(hfirst)[np.cumsum(hfirst) - 1]] if G else np.zeros(0, np.int64)
    n_alt_h = (b.alt_base[1:] - b.alt_base[:-1]).astype(np.int64)
    tasks = np.zeros((n_old + 2 * G, 4), np.uint32)
    tasks[new_idx] = b.tasks
    tasks[hdr_idx, 0] = n_alt_h[b.ann_hap] + e
    tasks[hdr_idx, 1] = nlen + 4
    tasks[hdr_idx, 3] = 1
    tasks[nl_idx, 0] = n_alt_h[b.ann_hap] + e + nlen + 4
    tasks[nl_idx, 1] = 1
    tasks[nl_idx, 3] = 1
    # destinations: running sum of lengths inside each haplotype, in the new order
    task_hap = np.empty(len(tasks), np.int64)
    task_hap[new_idx] = np.repeat(b.ann_hap, c)
    task_hap[hdr_idx] = b.ann_hap
    task_hap[nl_idx] = b.ann_hap
    ln = tasks[:, 1].astype(np.int64)
    l_excl = np.cumsum(ln) - ln
    task_begin = np.zeros(n_hap + 1, np.uint64)
    np.cumsum(np.bincount(task_hap, minlength=n_hap), out=task_begin[1:])
    nonempty = np.flatnonzero(task_begin[1:] > task_begin[:-1])
    hap_l0 = np.zeros(n_hap, np.int64)
    hap_l0[nonempty] = l_excl[task_begin[:-1].astype(np.int64)[nonempty]]
    tasks[:, 2] = l_excl - hap_l0[task_hap]
    out_base = np.zeros(n_hap + 1, np.uint64)
    np.cumsum(np.bincount(task_hap, weights=ln, minlength=n_hap).astype(np.int64), out=out_base[1:])
    # name tape: one entry per record appended behind the haplotype's alteration bytes
    names_per_hap = np.bincount(b.ann_hap, weights=elen, minlength=n_hap).astype(np.int64)
    alt_base = np.zeros(n_hap + 1, np.uint64)
    np.cumsum(n_alt_h + names_per_hap, out=alt_base[1:])
    alt = np.zeros(int(alt_base[-1]), np.uint8)
    old_pos = np.arange(len(b.alt)) + (alt_base[:-1].astype(np.int64) - b.alt_base[:-1].astype(np.int64))[
        np.repeat(np.arange(n_hap), n_alt_h)]
    alt[old_pos] = b.alt
    base = alt_base[:-1].astype(np.int64)[b.ann_hap] + n_alt_h[b.ann_hap] + e
    alt[base] = ord(">")
    tot = int(nlen.sum())
    if tot:  # the name bytes of every record, gathered from the pool
        rec_of = np.repeat(np.arange(G), nlen)
        within = np.arange(tot) - np.repeat(np.cumsum(nlen) - nlen, nlen)
        alt[base[rec_of] + 1 + within] = np.asarray(name_pool, np.uint8)[name_off[b.ann_tx][rec_of] + within]
    alt[base + 1 + nlen] = ord("_")
    alt[base + 2 + nlen] = ord("1") + (b.ann_hap & 1)
    alt[base + 3 + nlen] = ord("\n")
    alt[base + 4 + nlen] = ord("\n")
    seq_start = tasks[hdr_idx, 2].astype(np.int64) + nlen + 4
    return Batch(task_begin, tasks, alt, alt_base, out_base, None, b.ref, b.ann_hap, b.ann_tx, seq_start,
                 seq_start + (b.ann_end - b.ann_start), b.kept_hap, b.kept_site, c + 2)


def parse_fasta_image(image: np.ndarray) -> List[Tuple[str, str]]:
    """`>{name}\n{seq}\n` records of a file image (sequences may be empty), sorted like the test oracle does."""
    lines = image.tobytes().decode("ascii").split("\n")
    assert lines[-1] == ""
    recs = []
    for i in range(0, len(lines) - 1, 2):
        assert lines[i].startswith(">")
        recs.append((lines[i][1:], lines[i + 1]))
    return sorted(recs)
