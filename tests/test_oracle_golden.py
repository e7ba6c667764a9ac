"""CPU suite: the oracle (oracle/taskgen.py + oracle/ref_engine.c) is pinned against golden vectors that
oracle/make_golden.py harvested from the reference's own prebuilt binary (tests/golden/*.json), and against
the reference's in-file unit-test vectors (task.rs:118-144)."""
import numpy as np
import pytest

from oracle import cengine, taskgen
from tests.helpers import assert_sample_matches_reference, cohort_haplotype_csqs, hap_gir, load_golden, tape_to_str, u32, u8

UNIT = load_golden("unit_tests.json")
COMBOS = load_golden("combos.json")


@pytest.mark.parametrize("case", UNIT, ids=[c["name"] for c in UNIT])
def test_unit_test_replay_matches_reference_binary(case):
    refs = {case["transcript"]: case["ref"]}
    muts = taskgen.alt_transcript(case["transcript"], case["csqs"])
    try:
        ti = taskgen.TranscriptInstruction.from_alt_transcript(case["transcript"], muts, refs)
        g = ti.get_g_rep(refs)
    except taskgen.TaskGenError:
        assert case["tasks"] is None and case["records"] == []
        return
    assert [list(t) for t in g.tasks] == [list(t) for t in case["tasks"]]
    tape = taskgen.execute_tasks(g.tasks, g.ref, g.alt, g.res_len)
    seq = tape[g.annotation[0]:g.annotation[1]]
    assert [[case["transcript"] + "_1", seq]] == [list(r) for r in case["records"]]
    if case["asserted_len"] is not None:  # the reference test's own assert_eq! on the length
        assert len(seq) == case["asserted_len"]


@pytest.mark.parametrize("case", COMBOS, ids=[c["name"] for c in COMBOS])
def test_combo_matches_reference_binary(case):
    refs = case["refs"]
    g = hap_gir(case["csqs"], refs)
    tape = taskgen.execute_tasks(g.tasks, g.ref, g.alt, g.res_len)
    recs = sorted(taskgen.sequence_tape_records(tape, g.annotation, 1))
    assert [list(r) for r in recs] == [list(r) for r in case["records"]]
    # per-transcript Task vectors as dumped by DEBUG_TXP
    for tname, muts in taskgen.group_muts_per_transcript(case["csqs"]):
        want = case["tasks"].get(tname)
        try:
            ti = taskgen.TranscriptInstruction.from_alt_transcript(tname, muts, refs)
            got = ti.get_g_rep(refs).tasks
        except taskgen.TaskGenError:
            got = None
        if want is None:
            assert not got
        else:
            assert [list(t) for t in got] == [list(t) for t in want]
    # the concatenated haplotype table printed by the DEBUG_CPU_EXEC validator before it panics (gir.rs:212-224)
    if case["cpu_exec_table"]:
        assert [list(t) for t in g.tasks] == [list(t) for t in case["cpu_exec_table"]]
        st, bad = cengine.gir_execute(g.tasks, u8(g.ref), u8(g.alt), np.zeros(g.res_len, np.uint8), True, validate=True)
        assert st == cengine.REF_ERR_NOT_CONTIGUOUS and bad >= 1
        assert case["cpu_exec_returncode"] != 0  # the reference panicked


@pytest.mark.parametrize("name", ["cohort_a.json", "cohort_b.json", "cohort_c.json"])
def test_cohort_fasta_matches_reference_binary(name):
    cohort = load_golden(name)
    refs = cohort["refs"]
    per_hap = cohort_haplotype_csqs(cohort)
    n_records = 0
    for smp in cohort["samples"]:
        recs = []
        for hap in (1, 2):
            csqs = per_hap.get((smp, hap), [])
            if not csqs:
                continue
            g = hap_gir(csqs, refs)
            # python engine and C engine (both widths) agree
            tape = taskgen.execute_tasks(g.tasks, g.ref, g.alt, g.res_len)
            for conv, dt in ((u8, np.uint8), (u32, np.uint32)):
                res = np.zeros(g.res_len, dt)
                st, _ = cengine.gir_execute(g.tasks, conv(g.ref), conv(g.alt), res, True)
                assert st == cengine.REF_OK
                assert tape_to_str(res) == tape
            recs += taskgen.sequence_tape_records(tape, g.annotation, hap)
        assert_sample_matches_reference(cohort, smp, recs)  # full records, or count + digest for the large cohort (C1 substitute)
        n_records += len(recs)
    want_total = sum(cohort["fasta_records"].values()) if "fasta_records" in cohort else sum(len(v) for v in cohort["fasta"].values())
    assert n_records == want_total and n_records > 100


def test_task_rs_unit_vector():
    """task.rs:118-144: three ref copies into an 'x'-filled tape, applied in array order, no '.' fill."""
    ref = "ABCFEFGH"
    alt = ref[::-1]
    tasks = [(0, 1, 1, 8), (0, 4, 1, 4), (0, 6, 2, 6)]
    for conv, dt in ((u8, np.uint8), (u32, np.uint32)):
        res = conv("x" * 10)
        st, _ = cengine.gir_execute(tasks, conv(ref), conv(alt), res, fill_dot=False)
        assert st == cengine.REF_OK and tape_to_str(res) == "xxxxExGHBx"
    assert taskgen.execute_tasks(tasks, ref, alt, 10, fill="x") == "xxxxExGHBx"


def test_c_engine_error_classes():
    ref, alt = u8("ABCDEFGH"), u8("xyz")
    res = np.zeros(8, np.uint8)
    assert cengine.gir_execute([(0, 0, 9, 0)], ref, alt, res, True)[0] == cengine.REF_ERR_RES_OOB
    assert cengine.gir_execute([(1, 2, 2, 0)], ref, alt, res, True)[0] == cengine.REF_ERR_SRC_OOB
    assert cengine.gir_execute([(2, 0, 0, 0)], ref, alt, res, True)[0] == cengine.REF_ERR_BAD_STREAM
    st, bad = cengine.gir_execute([(0, 0, 2, 0), (0, 2, 2, 3)], ref, alt, res, True, validate=True)
    assert (st, bad) == (cengine.REF_ERR_NOT_CONTIGUOUS, 1)
    # overlap: later task wins (serial order, gir.rs:233)
    res = np.zeros(8, np.uint8)
    assert cengine.gir_execute([(0, 0, 6, 0), (1, 0, 3, 2)], ref, alt, res, True)[0] == cengine.REF_OK
    assert tape_to_str(res) == "ABxyzF.."


def test_batch_oracle_matches_single(tmp_path):
    from tests.helpers import batch_from_girs

    cohort = load_golden("cohort_a.json")
    per_hap = cohort_haplotype_csqs(cohort)
    girs = [hap_gir(c, cohort["refs"]) for _, c in sorted(per_hap.items())]
    girs.insert(3, taskgen.HaplotypeGIR([], {}, "", "", 0))  # an empty haplotype in the middle
    b = batch_from_girs(girs)
    out = np.zeros(int(b["out_base"][-1]), np.uint8)
    st, bh, bi = cengine.batch_execute(b["task_begin"], b["tasks"], b["ref"], b["alt"], b["alt_base"], out,
                                       b["out_base"], ref_base=b["ref_base"], threads=4)
    assert st == cengine.REF_OK
    want = "".join(taskgen.execute_tasks(g.tasks, g.ref, g.alt, g.res_len) for g in girs)
    assert tape_to_str(out) == want
