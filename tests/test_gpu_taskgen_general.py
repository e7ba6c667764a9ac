"""GPU suite: the device task generator on the GENERAL catalogue (every instruction code of the reference, its skip /
absent / empty / abort outcomes) against the reference-pinned oracle, haplotype by haplotype: Task tuples, alteration
tapes, annotation maps, tape lengths -- and the executed tapes, and the FASTA text in V2P_GEN_FASTA mode."""
import random

import numpy as np
import pytest

from oracle import taskgen as T
from tests.helpers import cohort_haplotype_csqs, load_golden
from tests.test_taskgen_rules import random_csq, static_instruction
from vcf2prot_b200 import _lib as L
from vcf2prot_b200.engine import EngineError
from vcf2prot_b200.taskgen import DeviceCatalogue, execute_generated

pytestmark = pytest.mark.gpu


class World:
    """refs + per-haplotype csq lists -> proteome tape, general catalogue, CSR site lists (the host's part of the job)."""

    def __init__(self, refs, hap_csqs):
        self.refs, self.hap_csqs = refs, hap_csqs
        self.names = sorted(refs)
        tx = {n: i for i, n in enumerate(self.names)}
        lens = np.array([len(refs[n]) for n in self.names], np.int64)
        self.off = np.zeros(len(lens) + 1, np.uint64)
        np.cumsum(lens, out=self.off[1:])
        self.tape = np.frombuffer("".join(refs[n] for n in self.names).encode(), np.uint8).copy()
        muts = {}
        for csqs in hap_csqs:
            for c in csqs:
                if c not in muts:
                    muts[c] = T.Mutation.from_csq(c)
        order = sorted(muts, key=lambda c: (tx[muts[c].transcript_name], muts[c].mut_pos, c))
        self.site_of = {c: i for i, c in enumerate(order)}
        ins = [static_instruction(muts[c]) for c in order]
        pool = "".join(i.data for i, _ in ins)
        doff = np.cumsum([0] + [len(i.data) for i, _ in ins])[:-1]
        self.cat_args = (self.off, [tx[muts[c].transcript_name] for c in order], [ord(i.code) for i, _ in ins], [f for _, f in ins],
                         [i.pos_ref for i, _ in ins], [i.pos_res for i, _ in ins], [i.len for i, _ in ins], doff,
                         [len(i.data) for i, _ in ins], np.frombuffer(pool.encode(), np.uint8))
        lists = [sorted(self.site_of[c] for c in csqs) for csqs in hap_csqs]
        self.site_begin = np.cumsum([0] + [len(l) for l in lists]).astype(np.uint64)
        self.sites = np.array([s for l in lists for s in l], np.uint32)
        self.tx = tx

    def oracle_hap(self, h):
        """haplotype_instruction.rs:75-158 through the oracle, ref offsets taken in the shared proteome tape.
        -> (tasks, alt, [(tx, start, end)], res_len, n_skipped); raises T.RefPanic where the reference aborts."""
        instrs = T.haplotype_instructions(T.group_muts_per_transcript(self.hap_csqs[h]), self.refs)
        res_len = sum(t.expected_results_size() for t in instrs)
        tasks, ann, alt, alt_c, res_c, skipped = [], [], "", 0, 0, 0
        for ti in instrs:
            try:
                g = ti.get_g_rep(self.refs)
            except T.TaskGenError:
                skipped += 1
                continue
            t = self.tx[ti.name]
            for (code, sp, ln, spr) in g.tasks:
                tasks.append((code, sp + (int(self.off[t]) if code == 0 else alt_c), ln, spr + res_c))
            ann.append((t, g.annotation[0] + res_c, g.annotation[1] + res_c))
            alt += g.alt
            alt_c += len(g.alt)
            res_c += g.res_len
        return tasks, alt, ann, res_len, skipped


def random_world(seed, n_tx=40, n_hap=60, want="ok"):
    rng = random.Random(seed)
    refs = {"ENST%05d" % i: "M" + "".join(rng.choice("ACDEFGHIKLMNPQRSTVWY") for _ in range(rng.randint(11, 120))) for i in range(n_tx)}
    pool = []
    for name, seq in refs.items():  # a catalogue of candidate mutations per transcript
        for _ in range(rng.randint(1, 7)):
            c = random_csq(rng, name, len(seq))
            try:
                static_instruction(T.Mutation.from_csq(c))
            except (T.TaskGenError, T.RefPanic):
                continue
            pool.append(c)
    haps = []
    while len(haps) < n_hap:
        csqs, used = [], set()
        for c in rng.sample(pool, rng.randint(0, min(len(pool), 14))):
            m = T.Mutation.from_csq(c)
            if (m.transcript_name, m.mut_pos) in used:
                continue
            used.add((m.transcript_name, m.mut_pos))
            csqs.append(c)
        w = World(refs, [csqs])
        try:
            tasks, alt, _, res_len, _ = w.oracle_hap(0)
            kind = "ok"
        except T.RefPanic:
            kind = "gen_panic"  # the reference aborts while generating tasks
        if kind == "ok":
            try:
                T.execute_tasks(tasks, w.tape.tobytes().decode(), alt, res_len)
            except T.RefPanic:
                kind = "exec_panic"  # ... or later, in Task::execute (task.rs:44/48): the ENGINE's error to report
        if kind == want:
            haps.append(csqs)
    return World(refs, haps)


def check_world(w, gpu_engine, dc):
    n_hap = len(w.hap_csqs)
    g = dc.generate_lists(w.site_begin, w.sites)
    b = g.batch
    tb, ab, ob = (dc.read(x, n_hap + 1, np.uint64) for x in (b.task_begin, b.alt_base, b.out_base))
    tasks = dc.read(b.tasks, 4 * b.n_tasks, np.uint32).reshape(-1, 4)
    alt = dc.read(b.alt, b.n_alt, np.uint8)
    rows = list(zip(dc.read(g.ann_hap, g.n_rows, np.uint32), dc.read(g.ann_tx, g.n_rows, np.uint32),
                    dc.read(g.ann_start, g.n_rows, np.uint64), dc.read(g.ann_end, g.n_rows, np.uint64)))
    gpu_engine.set_reference(w.tape)
    execute_generated(gpu_engine, g)
    tape = dc.read(b.out, b.n_out, np.uint8)
    n_skipped, outcomes = 0, set()
    for h in range(n_hap):
        want_tasks, want_alt, want_ann, res_len, skipped = w.oracle_hap(h)
        n_skipped += skipped
        got = [(int(s), int(a), int(l), int(d)) for a, l, d, s in tasks[int(tb[h]):int(tb[h + 1])]]
        assert got == want_tasks, (h, got, want_tasks)
        assert alt[int(ab[h]):int(ab[h + 1])].tobytes().decode() == want_alt, h
        assert int(ob[h + 1] - ob[h]) == res_len, h
        assert [(int(t), int(s), int(e)) for hh, t, s, e in rows if hh == h] == want_ann, h
        alt_s = want_alt
        ref_s = w.tape.tobytes().decode()
        want_tape = T.execute_tasks(want_tasks, ref_s, alt_s, res_len)
        assert tape[int(ob[h]):int(ob[h + 1])].tobytes().decode() == want_tape, h
        outcomes.add("skipped" if skipped else "plain")
    assert g.n_skipped == n_skipped
    return outcomes


@pytest.mark.parametrize("name", ["cohort_a.json", "cohort_b.json"])
def test_golden_cohorts_through_the_general_catalogue(gpu_engine, name):
    """The reference binary's own FASTA for the seeded multi-class cohorts, from tasks generated on the device."""
    cohort = load_golden(name)
    per_hap = cohort_haplotype_csqs(cohort)
    keys = sorted(per_hap)
    csqs = []
    for k in keys:  # unparsable csq strings are dropped by AltTranscript::new (vcf_ds.rs:366-373): the host's part
        good = []
        for c in per_hap[k]:
            try:
                static_instruction(T.Mutation.from_csq(c))
                good.append(c)
            except T.TaskGenError:
                pass
        csqs.append(good)
    w = World(cohort["refs"], csqs)
    dc = DeviceCatalogue.from_instructions(*w.cat_args)
    check_world(w, gpu_engine, dc)
    # and the records the reference binary wrote
    g = dc.generate_lists(w.site_begin, w.sites)
    gpu_engine.set_reference(w.tape)
    execute_generated(gpu_engine, g)
    tape = dc.read(g.batch.out, g.batch.n_out, np.uint8)
    ob = dc.read(g.batch.out_base, len(keys) + 1, np.uint64)
    rows = list(zip(dc.read(g.ann_hap, g.n_rows, np.uint32), dc.read(g.ann_tx, g.n_rows, np.uint32),
                    dc.read(g.ann_start, g.n_rows, np.uint64), dc.read(g.ann_end, g.n_rows, np.uint64)))
    fasta = {}
    for h, t, s, e in rows:
        smp, hap = keys[int(h)]
        fasta.setdefault(smp, []).append(["%s_%d" % (w.names[int(t)], hap), tape[int(ob[h]) + int(s):int(ob[h]) + int(e)].tobytes().decode()])
    for smp, recs in cohort["fasta"].items():
        assert sorted(fasta.get(smp, [])) == sorted([list(r) for r in recs]), smp
    dc.close()


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_random_haplotypes_over_all_classes(gpu_engine, seed):
    w = random_world(seed)
    dc = DeviceCatalogue.from_instructions(*w.cat_args)
    outcomes = check_world(w, gpu_engine, dc)
    assert "skipped" in outcomes
    dc.close()


def test_fasta_mode_is_the_text_of_the_annotated_records(gpu_engine):
    w = random_world(21, n_hap=24)
    dc = DeviceCatalogue.from_instructions(*w.cat_args)
    name_off = (9 * np.arange(len(w.names) + 1)).astype(np.uint64)
    dc.set_names(name_off, np.frombuffer("".join(w.names).encode(), np.uint8))
    g = dc.generate_lists(w.site_begin, w.sites, fasta=True)
    gpu_engine.set_reference(w.tape)
    execute_generated(gpu_engine, g)
    n_hap = len(w.hap_csqs)
    ob = dc.read(g.batch.out_base, n_hap + 1, np.uint64)
    text = dc.read(g.batch.out, g.batch.n_out, np.uint8)
    ref_s = w.tape.tobytes().decode()
    for h in range(n_hap):
        tasks, alt, ann, res_len, _ = w.oracle_hap(h)
        tape = T.execute_tasks(tasks, ref_s, alt, res_len)
        want = "".join(">%s_%d\n%s\n" % (w.names[t], 1 + (h & 1), tape[s:e]) for t, s, e in ann)
        assert text[int(ob[h]):int(ob[h + 1])].tobytes().decode() == want, h
    dc.close()


def test_reference_aborts_are_reported_with_the_transcript(gpu_engine):
    w = random_world(31, n_hap=6, want="gen_panic")
    dc = DeviceCatalogue.from_instructions(*w.cat_args)
    with pytest.raises(EngineError) as ei:
        dc.generate_lists(w.site_begin, w.sites)
    assert ei.value.status == L.ERR_TASKGEN and "haplotype 0" in str(ei.value)
    with pytest.raises(EngineError):  # the packed layout only
        dc.generate(np.zeros(0, np.int64), np.zeros(0, np.int64), 2, aligned=True)
    dc.close()


def test_task_arrays_the_reference_panics_on_at_execution_are_rejected_by_the_engine(gpu_engine):
    """Some mutation sets generate fine and then slice out of range in Task::execute (task.rs:44/48): the generator emits
    the same tasks and the ENGINE reports the panic, as for host-built arrays."""
    w = random_world(41, n_hap=5, want="exec_panic")
    dc = DeviceCatalogue.from_instructions(*w.cat_args)
    g = dc.generate_lists(w.site_begin, w.sites)
    gpu_engine.set_reference(w.tape)
    with pytest.raises(EngineError) as ei:
        execute_generated(gpu_engine, g)
    assert ei.value.status in (L.ERR_RES_OOB, L.ERR_SRC_OOB) and ei.value.bad_hap == 0
    dc.close()


@pytest.mark.parametrize("gzip", [False, True])
def test_pipeline_over_the_general_catalogue(gpu_engine, gzip):
    """csq-level cohort -> per-sample .fasta(.gz) through v2p_pipeline_run_lists with general catalogues as lanes."""
    import zlib

    from vcf2prot_b200.pipeline import DevicePipeline

    w = random_world(51, n_hap=30)
    name_off = (9 * np.arange(len(w.names) + 1)).astype(np.uint64)
    cats = []
    for _ in range(2):
        dc = DeviceCatalogue.from_instructions(*w.cat_args)
        dc.set_names(name_off, np.frombuffer("".join(w.names).encode(), np.uint8))
        cats.append(dc)
    gpu_engine.set_reference(w.tape)
    pipe = DevicePipeline(gpu_engine, cats=cats)
    out = np.zeros(1 << 20, np.uint8)
    fb, res = pipe.run_lists(w.site_begin, w.sites, 15, 4, gzip, out=out)
    ref_s = w.tape.tobytes().decode()
    for s in range(15):
        want = ""
        for h in (2 * s, 2 * s + 1):
            tasks, alt, ann, res_len, _ = w.oracle_hap(h)
            tape = T.execute_tasks(tasks, ref_s, alt, res_len)
            want += "".join(">%s_%d\n%s\n" % (w.names[t], 1 + (h & 1), tape[a:b]) for t, a, b in ann)
        got = out[int(fb[s]):int(fb[s + 1])].tobytes()
        assert (zlib.decompress(got, wbits=31) if gzip else got).decode() == want, s
    pipe.close()


@pytest.mark.parametrize("fasta", [False, True])
def test_general_catalogue_equals_the_class_tables_on_a_synthetic_cohort(gpu_engine, fasta):
    """Two implementations of the same rules -- one thread per site over class tables, one thread per transcript over
    Instruction values -- give the same arrays on a cohort with all seven classes (packed layout, plain and FASTA)."""
    from synth import cohort as C

    prot = C.make_proteome(seed=71, n_tx=300, mu=5.3, sigma=0.7, lo=30, hi=3000)
    cat = C.make_catalogue(prot, 8000, seed=72, mix=(0.55, 0.10, 0.10, 0.08, 0.07, 0.05, 0.05), fs_mean=30, fs_max=600, sl_max=120)
    cat.af[:] = np.random.default_rng(8).choice([0.01, 0.05, 0.2, 0.5], size=cat.n)
    hap, site = C.select_sites(cat, 50, np.random.default_rng(9))
    kept = C.build_batch(prot, cat, hap, site, 50, "global", "packed")  # drops sites behind a truncating one (cohort rule)
    hap, site = kept.kept_hap, kept.kept_site
    a = DeviceCatalogue(prot, cat, 0)
    b = DeviceCatalogue.from_instructions(*C.instruction_arrays(prot, cat))
    for dc in (a, b):
        dc.set_names(*C.default_names(prot))
    ga = a.generate(hap, site, 50, aligned=False, fasta=fasta)
    sb = np.zeros(51, np.uint64)
    np.cumsum(np.bincount(hap, minlength=50), out=sb[1:])
    gb = b.generate_lists(sb, site, fasta=fasta)
    assert (ga.batch.n_tasks, ga.batch.n_alt, ga.batch.n_out, ga.n_rows) == (gb.batch.n_tasks, gb.batch.n_alt, gb.batch.n_out, gb.n_rows)
    for f, n, dt in (("task_begin", 51, np.uint64), ("alt_base", 51, np.uint64), ("out_base", 51, np.uint64),
                     ("tasks", 4 * ga.batch.n_tasks, np.uint32), ("alt", ga.batch.n_alt, np.uint8)):
        assert np.array_equal(a.read(getattr(ga.batch, f), n, dt), b.read(getattr(gb.batch, f), n, dt)), f
    for f, dt in (("ann_hap", np.uint32), ("ann_tx", np.uint32), ("ann_start", np.uint64), ("ann_end", np.uint64)):
        assert np.array_equal(a.read(getattr(ga, f), ga.n_rows, dt), b.read(getattr(gb, f), gb.n_rows, dt)), f
    assert gb.n_skipped == 0
    a.close()
    b.close()


def test_skip_aborts_leaves_the_transcript_out_and_counts_it(gpu_engine):
    """V2P_GEN_SKIP_ABORTS: haplotypes that would stop the reference are generated without the offending transcripts;
    every other transcript comes out as the oracle has it."""
    bad = random_world(31, n_hap=6, want="gen_panic")
    dc = DeviceCatalogue.from_instructions(*bad.cat_args)
    g = dc.generate_lists(bad.site_begin, bad.sites, skip_aborts=True)
    assert g.n_aborted >= 6
    rows = list(zip(dc.read(g.ann_hap, g.n_rows, np.uint32), dc.read(g.ann_tx, g.n_rows, np.uint32),
                    dc.read(g.ann_start, g.n_rows, np.uint64), dc.read(g.ann_end, g.n_rows, np.uint64)))
    tb = dc.read(g.batch.task_begin, 7, np.uint64)
    tasks = dc.read(g.batch.tasks, 4 * g.batch.n_tasks, np.uint32).reshape(-1, 4)
    n_checked = 0
    for h in range(6):
        # per transcript: the oracle either aborts (then the device left it out) or gives the tasks the device emitted
        want_rows, want_tasks, res_c, alt_c = [], [], 0, 0
        for ti in T.haplotype_instructions(T.group_muts_per_transcript(bad.hap_csqs[h]), bad.refs):
            try:
                size = ti.expected_results_size()
                gt = ti.get_g_rep(bad.refs)
            except T.RefPanic:
                continue
            except T.TaskGenError:
                res_c += 0
                continue
            t = bad.tx[ti.name]
            want_tasks += [(c, sp + (int(bad.off[t]) if c == 0 else alt_c), ln, spr + res_c) for (c, sp, ln, spr) in gt.tasks]
            want_rows.append((t, gt.annotation[0] + res_c, gt.annotation[1] + res_c))
            alt_c += len(gt.alt)
            res_c += gt.res_len
        assert [(int(t), int(s), int(e)) for hh, t, s, e in rows if hh == h] == want_rows, h
        assert [(int(s), int(a), int(l), int(d)) for a, l, d, s in tasks[int(tb[h]):int(tb[h + 1])]] == want_tasks, h
        n_checked += len(want_rows)
    assert n_checked > 0
    dc.close()


def test_pipeline_skip_aborts(gpu_engine):
    from vcf2prot_b200.pipeline import DevicePipeline

    bad = random_world(31, n_hap=6, want="gen_panic")
    name_off = (9 * np.arange(len(bad.names) + 1)).astype(np.uint64)
    cats = []
    for _ in range(2):
        dc = DeviceCatalogue.from_instructions(*bad.cat_args)
        dc.set_names(name_off, np.frombuffer("".join(bad.names).encode(), np.uint8))
        cats.append(dc)
    gpu_engine.set_reference(bad.tape)
    pipe = DevicePipeline(gpu_engine, cats=cats)
    out = np.zeros(1 << 20, np.uint8)
    with pytest.raises(EngineError) as ei:  # the reference's behaviour: the run stops
        pipe.run_lists(bad.site_begin, bad.sites, 3, 2, False, out=out)
    assert ei.value.status == L.ERR_TASKGEN
    fb, res = pipe.run_lists(bad.site_begin, bad.sites, 3, 2, False, out=out, skip_aborts=True)
    assert res.n_aborted >= 6 and res.n_records == out[: int(fb[-1])].tobytes().count(b">")
    pipe.close()
