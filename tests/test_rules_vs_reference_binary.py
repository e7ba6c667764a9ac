"""CPU suite: the device generator's rules (csrc/v2p_taskgen_rules.cuh, host build) against the REFERENCE BINARY itself --
not the oracle -- on seeded random cohorts over all 22 supported consequence classes: every FASTA record the reference's
own prebuilt tool writes (oracle/_ref/vcf2prot, `-g st`) must be what the rules' tasks produce, and every transcript the
rules skip / drop / leave empty must be absent / empty in its output.  The oracle only pre-filters mutation sets on which
the reference aborts (one abort would end the whole run of the binary)."""
import random
import subprocess

import pytest

from oracle import refbin
from oracle import taskgen as T
from tests.test_taskgen_rules import AA, OK, EMPTY, case_text, random_csq, rules_binary, static_instruction  # noqa: F401

pytestmark = pytest.mark.skipif(not refbin.available(), reason="oracle/_ref/vcf2prot not built (make -C oracle ref)")


def safe_for_the_binary(csqs, refs):
    """No abort anywhere in generation or execution (the oracle's view) -- otherwise the binary would die mid-run."""
    try:
        g = T.haplotype_g_rep(T.haplotype_instructions(T.group_muts_per_transcript(csqs), refs), refs)
        T.sequence_tape_records(T.execute_tasks(g.tasks, g.ref, g.alt, g.res_len), g.annotation, 1)
        return True
    except T.RefPanic:
        return False


def random_cohort(seed, n_tx=40, n_samples=120):
    rng = random.Random(seed)
    refs = {"ENST%05d" % i: "M" + "".join(rng.choice(AA) for _ in range(rng.randint(11, 90))) for i in range(n_tx)}
    pool = []
    for name, seq in refs.items():
        for _ in range(rng.randint(2, 6)):
            c = random_csq(rng, name, len(seq))
            try:
                static_instruction(T.Mutation.from_csq(c))
            except (T.TaskGenError, T.RefPanic):
                continue
            pool.append(c)
    pool = sorted(set(pool))
    haps = []
    while len(haps) < 2 * n_samples:
        csqs, used = [], set()
        for c in rng.sample(pool, rng.randint(0, 10)):
            m = T.Mutation.from_csq(c)
            if (m.transcript_name, m.mut_pos) in used:
                continue
            used.add((m.transcript_name, m.mut_pos))
            csqs.append(c)
        if safe_for_the_binary(csqs, refs):
            haps.append(csqs)
    return refs, pool, haps


def rules_records(binary, refs, csqs, hap_label):
    """FASTA records of one haplotype from the rules' output alone."""
    cases, muts_of = [], {}
    for name, muts in T.group_muts_per_transcript(csqs):
        muts_of[name] = muts
        cases.append(case_text(name, muts, len(refs[name])))
    if not cases:
        return []
    p = subprocess.run([binary], input="".join(cases), stdout=subprocess.PIPE, text=True, check=True)
    recs, cur = [], None
    blocks = {}
    for line in p.stdout.splitlines():
        f = line.split()
        if f[0] == "RESULT":
            cur = blocks[f[1]] = {"status": int(f[2]), "size": int(f[3]), "tasks": [], "alt": ""}
        elif f[0] == "TASK":
            cur["tasks"].append(tuple(int(x) for x in f[1:]))
        elif f[0] == "ALT":
            cur["alt"] = "" if f[1] == "-" else f[1]
    # HaplotypeInstruction::get_g_rep (haplotype_instruction.rs:75-158): the tape is sized from every transcript that has
    # instructions (skipped ones included), transcripts are laid down in name order, tasks run in that order on ONE tape
    # (a transcript whose tasks overrun its own result spills into its neighbour, as in the reference)
    SKIPPED = 3
    tape_len = sum(b["size"] for b in blocks.values() if b["status"] in (OK, EMPTY, SKIPPED))
    tape, res_c, ann = ["."] * tape_len, 0, []
    for name in sorted(blocks):
        b = blocks[name]
        if b["status"] == EMPTY:
            ann.append((name, res_c, res_c))
        elif b["status"] == OK:
            for (stream, src, ln, dst) in b["tasks"]:
                seg = (b["alt"] if stream else refs[name])[src:src + ln]
                assert len(seg) == ln and res_c + dst + ln <= tape_len, "the reference would abort here"
                tape[res_c + dst:res_c + dst + ln] = seg
            ann.append((name, res_c, res_c + b["size"]))
            res_c += b["size"]
    tape = "".join(tape)
    return [("%s_%d" % (name, hap_label), tape[s_:e_]) for name, s_, e_ in ann]


@pytest.mark.parametrize("seed", [101, 102, 103, 104])
def test_every_record_of_the_reference_binary(rules_binary, seed):
    refs, pool, haps = random_cohort(seed)
    n_samples = len(haps) // 2
    samples = ["S%03d" % i for i in range(n_samples)]
    records = []
    for c in pool:  # one VCF record per distinct consequence; bit 0 -> haplotype 1, bit 1 -> haplotype 2
        cells = [([0] if c in haps[2 * s] else [], [0] if c in haps[2 * s + 1] else []) for s in range(n_samples)]
        if any(h1 or h2 for h1, h2 in cells):
            records.append(([c], cells))
    got, stdout, rc = refbin.run_reference(refbin.vcf_text(samples, records), refs, "st")
    assert rc == 0, stdout[-2000:]
    n_rec = n_empty = 0
    for s, smp in enumerate(samples):
        want = sorted(rules_records(rules_binary, refs, haps[2 * s], 1) + rules_records(rules_binary, refs, haps[2 * s + 1], 2))
        assert [tuple(r) for r in got.get(smp, [])] == want, smp
        n_rec += len(want)
        n_empty += sum(1 for _, q in want if q == "")
    assert n_rec > 400 and n_empty > 8


def test_debug_cpu_exec_validator_against_the_binary(rules_binary):
    """gir.rs:203-229 (DEBUG_CPU_EXEC): on random haplotypes the reference binary either runs through or prints its
    execution table and panics at the first task that does not start where the previous one ended -- the same table and
    the same verdict as the oracle's validator, which is what V2P_FLAG_VALIDATE is tested against on the GPU."""
    import numpy as np

    from oracle import cengine
    from tests.helpers import u8

    refs, pool, haps = random_cohort(201, n_tx=20, n_samples=45)
    n_gap = n_clean = 0
    for csqs in haps:
        if not csqs:
            continue
        g = T.haplotype_g_rep(T.haplotype_instructions(T.group_muts_per_transcript(csqs), refs), refs)
        if len(g.tasks) < 2:
            continue
        st, bad = cengine.gir_execute(g.tasks, u8(g.ref), u8(g.alt), np.zeros(g.res_len, np.uint8), True, validate=True)
        records = [([c], [([0], [])]) for c in csqs]
        _, stdout, rc = refbin.run_reference(refbin.vcf_text(["S1"], records), refs, "st",
                                             env={"RUN_SELECTED_TEST": "1", "DEBUG_CPU_EXEC": "1"})
        table = refbin.parse_cpu_exec_table(stdout)
        if st == cengine.REF_ERR_NOT_CONTIGUOUS:
            assert rc != 0 and "gir.rs" in stdout, stdout[-600:]
            assert table == [tuple(t) for t in g.tasks], (table, g.tasks)  # the binary dumps the whole haplotype table
            assert g.tasks[bad][3] != g.tasks[bad - 1][3] + g.tasks[bad - 1][2]
            assert all(g.tasks[i][3] == g.tasks[i - 1][3] + g.tasks[i - 1][2] for i in range(1, bad))
            n_gap += 1
        else:
            assert st == 0 and rc == 0, stdout[-600:]
            n_clean += 1
    assert n_gap >= 3 and n_clean >= 10, (n_gap, n_clean)


def test_multiword_masks_decode_like_the_binary(rules_binary):
    """Records with up to 40 consequences each (FORMAT/BCSQ masks of one to three 15-csq words, MaskDecoder.rs:122-153):
    the VCF text goes to the reference binary; the same mask words go through oracle/maskdecode.py (the checker of the
    device decode) and the rules; the records must agree -- so the decode oracle is pinned to the binary beyond the
    reference's own unit vectors, multi-word cells included."""
    import numpy as np

    from oracle import maskdecode

    refs, pool, haps = random_cohort(301, n_tx=30, n_samples=40)
    n_samples = len(haps) // 2
    rng = random.Random(5)
    order = pool[:]
    rng.shuffle(order)
    groups, i = [], 0
    while i < len(order):  # records of 1 .. 40 consequences
        k = rng.choice([1, 2, 7, 15, 16, 31, 40])
        groups.append(order[i:i + k])
        i += k
    site_of = {c: j for j, c in enumerate(pool)}
    records = []
    for grp in groups:
        cells = [([k for k, c in enumerate(grp) if c in haps[2 * s]], [k for k, c in enumerate(grp) if c in haps[2 * s + 1]])
                 for s in range(n_samples)]
        records.append((grp, cells))
    samples = ["S%03d" % s for s in range(n_samples)]
    vcf = refbin.vcf_text(samples, records)
    got, stdout, rc = refbin.run_reference(vcf, refs, "st")
    assert rc == 0, stdout[-2000:]
    # the mask words exactly as the VCF text carries them
    W = max((len(g) + 14) // 15 for g in groups)
    masks = np.zeros((len(groups), n_samples, W), np.uint32)
    body = [l for l in vcf.split("\n") if l and not l.startswith("#")]
    for r, line in enumerate(body):
        for s, cell in enumerate(line.split("\t")[9:]):
            words = [int(x) for x in cell.split(":")[-1].split(",")]
            masks[r, s, :len(words)] = words
    assert (masks[:, :, 1:] != 0).any()  # multi-word cells are present
    csq_begin = np.cumsum([0] + [len(g) for g in groups]).astype(np.uint64)
    csq_site = np.array([site_of[c] for g in groups for c in g], np.int32)
    sb, sites = maskdecode.site_lists(masks, csq_begin, csq_site)
    for s, smp in enumerate(samples):
        want = []
        for k in (0, 1):
            h = 2 * s + k
            csqs = [pool[j] for j in sites[int(sb[h]):int(sb[h + 1])]]
            assert sorted(csqs) == sorted(haps[h])  # the decode gives back what was encoded ...
            want += rules_records(rules_binary, refs, csqs, k + 1)
        assert [tuple(r) for r in got.get(smp, [])] == sorted(want), smp  # ... and the binary read the same masks


def test_write_all_record_set_of_the_reference_binary():
    """`-a` (write_all, personalized_genome.rs:120-210), pinned on the reference binary itself: per haplotype the altered
    records (the annotation map's keys -- start_lost's empty record included) and then every OTHER transcript of the
    proteome unchanged with the same `_1` / `_2` suffix.  This is the record set V2P_PIPE_ALL_RECORDS builds on the
    device (tests/test_gpu_pipeline.py holds it to the same expectation)."""
    from oracle import refbin, taskgen
    from tests.helpers import cohort_haplotype_csqs, load_golden

    if not refbin.available():
        pytest.skip("reference binary not present (oracle/_ref/vcf2prot)")
    for name in ("cohort_a.json", "cohort_b.json"):
        co = load_golden(name)
        recs, _, rc = refbin.run_reference(refbin.vcf_text(co["samples"], co["records"]), co["refs"], "st", write_all=True)
        assert rc == 0
        per_hap = cohort_haplotype_csqs(co)
        for smp in co["samples"]:
            want = []
            for hap in (1, 2):
                altered = taskgen.haplotype_records(per_hap.get((smp, hap), []), co["refs"], hap)
                keys = {h[: -2] for h, _ in altered}
                want += altered + [("%s_%d" % (k, hap), v) for k, v in co["refs"].items() if k not in keys]
            assert sorted(recs[smp]) == sorted(want), (name, smp)
            assert len(recs[smp]) == 2 * len(co["refs"])
