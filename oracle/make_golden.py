#!/usr/bin/env python
"""ORACLE (test infrastructure only): harvest golden vectors from the reference itself.

Run in the authoring container (needs /root/reference and oracle/_ref/vcf2prot):
    python oracle/make_golden.py
Writes tests/golden/*.json (committed).  Nothing in tests/ reads /root/reference at run time.

Sources of truth
  unit_tests.json  the inputs of the reference's own unit tests
                   (src/data_structures/InternalRep/transcript_instructions.rs:884-1594,
                    test_correct_translation_1..30: csq strings + reference sequence) replayed through the
                   reference's prebuilt binary with RUN_SELECTED_TEST=1 DEBUG_TXP=<id>, capturing its
                   per-transcript Vec<Task> dump (transcript_instructions.rs:372-382) and its FASTA records.
                   Each case also stores the length the reference test asserts, when it asserts one.
  combos.json      hand-written multi-mutation / multi-transcript cases (SURVEY.md section 8c "combo" and
                   "concat" rows) incl. the DEBUG_CPU_EXEC haplotype table (gir.rs:212-222).
  cohort_*.json    seeded synthetic cohorts (several csq classes, many transcripts, both haplotypes):
                   record-sorted FASTA of every sample from the reference binary.
"""
from __future__ import annotations

import json
import os
import random
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import refbin  # noqa: E402

REF_TESTS = "/root/reference/src/data_structures/InternalRep/transcript_instructions.rs"
OUT = os.path.join(ROOT, "tests", "golden")
AA = "ACDEFGHIKLMNPQRSTVWY"


def harvest_unit_tests():
    src = open(REF_TESTS).read()
    cases = []
    for m in re.finditer(r"fn (test_correct_translation_\d+)\(\)\s*\{(.*?)\n    \}", src, re.S):
        name, body = m.group(1), m.group(2)
        muts = re.search(r"let mutations=vec!\[(.*?)\];", body, re.S)
        if not muts:
            continue
        # drop commented-out entries; the unit tests hand every csq to AltTranscript::new(name, ..) regardless of
        # the transcript id inside the string (vcf_ds.rs:366-373), so the replay rewrites field 2 to `name`.
        live = "\n".join(l for l in muts.group(1).split("\n") if not l.strip().startswith("//"))
        csqs = re.findall(r'"([^"]+)"\.to_string\(\)', live)
        refm = re.search(r'reference\.insert\("([^"]+)"\.to_string\(\),\s*"([^"]+)"\.to_string\(\)\)', body)
        if not refm:
            continue
        tname, tseq = refm.group(1), refm.group(2)
        csqs = ["|".join(f[:2] + [tname] + f[3:]) for f in (c.split("|") for c in csqs)]
        # the length the reference test asserts (several spellings)
        exp_len = None
        a = re.search(r"assert_eq!\((\d+) as usize,\s*res_string\.len\(\)\)", body)
        b = re.search(r"assert_eq!\(ref_string\.len\(\)\s*([+-])\s*(\d+) as usize,\s*res_string\.len\(\)\)", body)
        c = re.search(r"assert_eq!\(ref_string\.len\(\),\s*res_string\.len\(\)\)", body)
        if a:
            exp_len = int(a.group(1))
        elif b:
            exp_len = len(tseq) + (int(b.group(2)) if b.group(1) == "+" else -int(b.group(2)))
        elif c:
            exp_len = len(tseq)
        cases.append({"name": name, "transcript": tname, "ref": tseq, "csqs": csqs, "asserted_len": exp_len})
    return cases


def run_case(csqs, ref_seqs, debug_txp=None, extra_env=None):
    """All csqs on haplotype 1 of one sample, one VCF record per csq."""
    records = [([c], [([0], [])]) for c in csqs]
    vcf = refbin.vcf_text(["S1"], records)
    env = {"RUN_SELECTED_TEST": "1"}
    if debug_txp:
        env["DEBUG_TXP"] = debug_txp
    if extra_env:
        env.update(extra_env)
    recs, stdout, rc = refbin.run_reference(vcf, ref_seqs, "st", env=env)
    return recs.get("S1", []), stdout, rc


def make_unit_tests():
    out = []
    for case in harvest_unit_tests():
        recs, stdout, rc = run_case(case["csqs"], {case["transcript"]: case["ref"]}, debug_txp=case["transcript"])
        dumps = refbin.parse_task_dumps(stdout)
        case["returncode"] = rc
        case["tasks"] = dumps[0] if dumps else None
        case["records"] = recs
        out.append(case)
        print("%-28s rc=%d tasks=%s records=%s" % (case["name"], rc, case["tasks"], recs))
    return out


REF38 = "MEDLGENTMVLSTLRSLNNFISQRVEGGSGLEELERGG"


def csq(kind, t, aa):
    return "%s|GENE|%s|protein_coding|-|%s|1936821C>T" % (kind, t, aa)


def make_combos():
    combos = [
        ("missense_then_frameshift", {"T": REF38},
         [csq("missense", "T", "5G>5H"), csq("frameshift", "T", "20FISQRVEGGSGLEELERGG*>20LTESTLONGTAIL" + "X" * 34 + "*")]),
        ("frameshift_then_missense_is_skipped", {"T": REF38},
         [csq("frameshift", "T", "10V>10VTESTFRAMESHIFT"), csq("missense", "T", "30L>30K")]),
        ("start_lost_empty_record", {"T": REF38}, [csq("start_lost", "T", "1M>1K")]),
        ("dot_gap_then_next_transcript", {"TA": REF38, "TB": REF38},
         [csq("inframe_deletion&stop_retained", "TA", "38*>38*"), csq("missense", "TB", "5G>5H")]),
        ("three_transcripts_mixed", {"TA": REF38, "TB": REF38, "TC": REF38 + "KLMNPQRS"},
         [csq("inframe_insertion", "TA", "5G>5GTEST"), csq("inframe_deletion", "TB", "10VLSTLR>10V"),
          csq("missense", "TB", "30L>30K"), csq("stop_gained", "TC", "40M>40*"), csq("missense", "TC", "3D>3E")]),
        ("stop_lost_after_missense", {"T": REF38}, [csq("missense", "T", "5G>5H"), csq("stop_lost", "T", "39*>39TEST")]),
        ("deletion_then_insertion", {"T": REF38},
         [csq("inframe_deletion", "T", "10VLSTLR>10V"), csq("inframe_insertion", "T", "20F>20FAAA")]),
        ("missense_last_residue", {"T": REF38}, [csq("missense", "T", "38G>38A")]),
        ("missense_first_and_second", {"T": REF38}, [csq("missense", "T", "2E>2Q"), csq("missense", "T", "3D>3N")]),
        ("adjacent_missense_run", {"T": REF38},
         [csq("missense", "T", "%d%s>%d%s" % (i + 1, REF38[i], i + 1, "W")) for i in range(10, 16)]),
    ]
    out = []
    for name, refs, csqs in combos:
        item = {"name": name, "refs": refs, "csqs": csqs, "tasks": {}}
        recs, stdout, rc = run_case(csqs, refs)
        item["records"] = recs
        item["returncode"] = rc
        for t in refs:
            _, so, _ = run_case(csqs, refs, debug_txp=t)
            d = refbin.parse_task_dumps(so)
            item["tasks"][t] = d[0] if d else None
        _, so, rc2 = run_case(csqs, refs, extra_env={"DEBUG_CPU_EXEC": "1"})
        item["cpu_exec_table"] = refbin.parse_cpu_exec_table(so)
        item["cpu_exec_returncode"] = rc2
        out.append(item)
        print("%-36s rc=%d records=%s table=%s" % (name, rc, recs, item["cpu_exec_table"]))
    return out


def synth_cohort(seed, n_tx, n_samples, n_sites):
    """Seeded cohort over several csq classes; only reference-valid combinations:
    unique well-separated positions per transcript, truncating classes last per haplotype+transcript."""
    rng = random.Random(seed)
    refs = {}
    for i in range(n_tx):
        L = rng.randint(30, 400)
        refs["ENST%011d" % i] = "M" + "".join(rng.choice(AA) for _ in range(L - 1))
    names = sorted(refs)
    sites = []  # (transcript, pos0, csq string, truncating)
    used = {}
    for _ in range(n_sites):
        t = rng.choice(names)
        seq = refs[t]
        L = len(seq)
        p = rng.randint(1, L - 1)  # 0-based
        if any(abs(p - q) < 14 for q in used.get(t, [])):
            continue
        r = rng.random()
        trunc = False
        if r < 0.6:
            new = rng.choice([a for a in AA if a != seq[p]])
            s = csq("missense", t, "%d%s>%d%s" % (p + 1, seq[p], p + 1, new))
        elif r < 0.7:
            ins = "".join(rng.choice(AA) for _ in range(rng.randint(1, 9)))
            s = csq("inframe_insertion", t, "%d%s>%d%s" % (p + 1, seq[p], p + 1, seq[p] + ins))
        elif r < 0.8:
            dl = rng.randint(1, 8)
            if p + dl + 1 >= L:
                continue
            s = csq("inframe_deletion", t, "%d%s>%d%s" % (p + 1, seq[p:p + dl + 1], p + 1, seq[p]))
        elif r < 0.88:
            tail = "".join(rng.choice(AA) for _ in range(rng.randint(1, 60)))
            s = csq("frameshift", t, "%d%s*>%d%s*" % (p + 1, seq[p:], p + 1, tail))
            trunc = True
        elif r < 0.94:
            s = csq("stop_gained", t, "%d%s>%d*" % (p + 1, seq[p], p + 1))
            trunc = True
        elif r < 0.98:
            tail = "".join(rng.choice(AA) for _ in range(rng.randint(1, 40)))
            p = L
            if any(q == L for q in used.get(t, [])):
                continue
            s = csq("stop_lost", t, "%d*>%d%s*" % (L + 1, L + 1, tail))
            trunc = True
        else:
            if used.get(t):
                continue
            p = 0
            s = csq("start_lost", t, "1M>1K")
            trunc = True
            used.setdefault(t, []).extend(range(0, 100000, 7))  # nothing else on this transcript
        used.setdefault(t, []).append(p)
        sites.append((t, p, s, trunc))
    samples = ["S%03d" % i for i in range(n_samples)]
    # genotype: per sample+hap choose sites with prob; drop sites after a truncating one in the same transcript
    records = []
    per = {}
    for si, (t, p, s, trunc) in enumerate(sites):
        af = rng.choice([0.02, 0.05, 0.2, 0.5])
        for smp in range(n_samples):
            for hap in (0, 1):
                if rng.random() < af:
                    per.setdefault((smp, hap, t), []).append((p, si, trunc))
    keep = set()
    for (smp, hap, t), lst in per.items():
        lst.sort()
        for (p, si, trunc) in lst:
            keep.add((smp, hap, si))
            if trunc:
                break
    for si, (t, p, s, trunc) in enumerate(sites):
        cells = []
        for smp in range(n_samples):
            cells.append(([0] if (smp, 0, si) in keep else [], [0] if (smp, 1, si) in keep else []))
        records.append(([s], cells))
    return refs, samples, records


def make_cohort(seed, n_tx, n_samples, n_sites, engine="st"):
    refs, samples, records = synth_cohort(seed, n_tx, n_samples, n_sites)
    vcf = refbin.vcf_text(samples, records)
    recs, stdout, rc = refbin.run_reference(vcf, refs, engine)
    assert rc == 0, stdout[-2000:]
    n = sum(len(v) for v in recs.values())
    print("cohort seed=%d: %d samples, %d sites, %d records, rc=%d" % (seed, n_samples, len(records), n, rc))
    return {"seed": seed, "refs": refs, "samples": samples,
            "records": [[r[0], [[c[0], c[1]] for c in r[1]]] for r in records], "fasta": recs}


def canonical_fasta_digest(records):
    """sha256 of a sample's record-sorted FASTA text (record order in the reference is HashMap-random)."""
    import hashlib

    return hashlib.sha256("".join(">%s\n%s\n" % (h, s) for h, s in sorted(tuple(r) for r in records)).encode()).hexdigest()


def make_cohort_compact(seed, n_tx, n_samples, n_sites, engine="mt"):
    """The C1 substitute at the size SURVEY 8d planned (64 samples x 1,200 records): too large to commit as text, so the
    fixture keeps the inputs compactly (one character per genotype cell: bit 0 = haplotype 1 carries the record's csq,
    bit 1 = haplotype 2) and, of the reference binary's output, a sha256 of every sample's record-sorted FASTA plus the
    full records of the first and last sample."""
    refs, samples, records = synth_cohort(seed, n_tx, n_samples, n_sites)
    recs, stdout, rc = refbin.run_reference(refbin.vcf_text(samples, records), refs, engine, timeout=1800)
    assert rc == 0, stdout[-2000:]
    print("cohort seed=%d: %d samples, %d records in the VCF, %d FASTA records, rc=%d" %
          (seed, n_samples, len(records), sum(len(v) for v in recs.values()), rc))
    cells = ["".join(str((1 if c[0] else 0) | (2 if c[1] else 0)) for c in r[1]) for r in records]
    return {"seed": seed, "refs": refs, "samples": samples, "csqs": [r[0][0] for r in records], "cells_compact": cells,
            "fasta_sha256": {s: canonical_fasta_digest(recs.get(s, [])) for s in samples},
            "fasta_records": {s: len(recs.get(s, [])) for s in samples},
            "fasta": {s: recs.get(s, []) for s in (samples[0], samples[-1])}}


def main():
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "unit_tests.json"), "w") as f:
        json.dump(make_unit_tests(), f, indent=1)
    with open(os.path.join(OUT, "combos.json"), "w") as f:
        json.dump(make_combos(), f, indent=1)
    with open(os.path.join(OUT, "cohort_a.json"), "w") as f:
        json.dump(make_cohort(0x5EED0A, 40, 12, 160), f)
    with open(os.path.join(OUT, "cohort_b.json"), "w") as f:
        json.dump(make_cohort(0x5EED0B, 120, 24, 500, engine="mt"), f)
    with open(os.path.join(OUT, "cohort_c.json"), "w") as f:
        json.dump(make_cohort_compact(0x5EED0C, 300, 64, 1800), f)  # ~1,200 records survive the spacing rule


if __name__ == "__main__":
    main()
