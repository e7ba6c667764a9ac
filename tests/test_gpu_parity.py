"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against the oracle -- bit-exact.

Oracle = oracle/ref_engine.c (restates task.rs:38-50 / gir.rs:197-241) and the golden vectors harvested from the
reference binary.  Edge cases follow the reference's own tests and panics: empty/ragged inputs, '.' gaps,
zero-length tasks, unsorted task arrays (task.rs:118-144), out-of-range slices, bad stream codes, the
DEBUG_CPU_EXEC contiguity validator.
"""
import numpy as np
import pytest

from oracle import cengine, taskgen
from tests.helpers import (assert_sample_matches_reference, batch_from_girs, cohort_haplotype_csqs, hap_gir, load_golden,
                           tape_to_str, u32, u8)
from tests.randtasks import chain_batch, random_batch
from vcf2prot_b200 import GIR, Engine, EngineError
from vcf2prot_b200 import _lib as L

pytestmark = pytest.mark.gpu

UNIT = load_golden("unit_tests.json")
COMBOS = load_golden("combos.json")


def oracle_batch(b, validate=False, dtype=np.uint8):
    out = np.zeros(int(b["out_base"][-1]) if len(b["out_base"]) else 0, dtype)
    st, bh, bi = cengine.batch_execute(b["task_begin"], b["tasks"], b["ref"].astype(dtype), b["alt"].astype(dtype),
                                       b["alt_base"], out, b["out_base"], ref_base=b.get("ref_base"),
                                       validate=validate, threads=4)
    return st, bh, bi, out


def gpu_batch(eng, b, validate=False):
    return eng.execute_batch(b["task_begin"], b["tasks"], b["ref"], b["alt"], b["alt_base"], b["out_base"],
                             ref_base=b.get("ref_base"), validate=validate)


# ---------------------------------------------------------------------------------------------- golden vectors
@pytest.mark.parametrize("case", [c for c in UNIT if c["tasks"]], ids=[c["name"] for c in UNIT if c["tasks"]])
def test_reference_unit_tests_through_gir_execute(gpu_engine, case):
    """transcript_instructions.rs:884-1594 inputs: golden Task vectors -> GIR::execute(Engine::GPU) -> golden FASTA."""
    refs = {case["transcript"]: case["ref"]}
    muts = taskgen.alt_transcript(case["transcript"], case["csqs"])
    g = taskgen.TranscriptInstruction.from_alt_transcript(case["transcript"], muts, refs).get_g_rep(refs)
    gir = GIR(case["tasks"], {case["transcript"]: g.annotation}, g.alt, g.ref, "." * g.res_len)
    res, ann = gir.execute(Engine.from_str("gpu"), gpu_engine)
    s, e = ann[case["transcript"]]
    assert tape_to_str(res)[s:e] == case["records"][0][1]
    if case["asserted_len"] is not None:
        assert len(res) == case["asserted_len"]


@pytest.mark.parametrize("case", COMBOS, ids=[c["name"] for c in COMBOS])
def test_combos_through_both_entries(gpu_engine, case):
    g = hap_gir(case["csqs"], case["refs"])
    res = gpu_engine.execute_soa(g.tasks, g.ref, g.alt, g.res_len, fill_dot=True)
    recs = sorted(taskgen.sequence_tape_records(tape_to_str(res), g.annotation, 1))
    assert [list(r) for r in recs] == [list(r) for r in case["records"]]
    out, _ = gpu_batch(gpu_engine, batch_from_girs([g]))
    assert tape_to_str(out) == tape_to_str(res)
    if case["cpu_exec_table"]:  # the reference panics in the DEBUG_CPU_EXEC validator (gir.rs:223)
        with pytest.raises(EngineError) as ei:
            gpu_engine.execute_soa(g.tasks, g.ref, g.alt, g.res_len, fill_dot=True, validate=True)
        st, bad = cengine.gir_execute(g.tasks, u8(g.ref), u8(g.alt), np.zeros(g.res_len, np.uint8), True, validate=True)
        assert ei.value.status == L.ERR_NOT_CONTIGUOUS and st == cengine.REF_ERR_NOT_CONTIGUOUS
        assert ei.value.bad_task == bad


@pytest.mark.parametrize("name", ["cohort_a.json", "cohort_b.json", "cohort_c.json"])
def test_cohort_fasta_byte_identical_to_reference_binary(gpu_engine, name):
    cohort = load_golden(name)
    per_hap = cohort_haplotype_csqs(cohort)
    keys = sorted(per_hap)
    girs = [hap_gir(per_hap[k], cohort["refs"]) for k in keys]
    b = batch_from_girs(girs)
    out, _ = gpu_batch(gpu_engine, b)
    st, _, _, want = oracle_batch(b)
    assert st == 0 and np.array_equal(out, want)
    fasta = {}
    for (smp, hap), g, o0 in zip(keys, girs, b["out_base"][:-1]):
        tape = tape_to_str(out[int(o0):int(o0) + g.res_len])
        fasta.setdefault(smp, []).extend(taskgen.sequence_tape_records(tape, g.annotation, hap))
    for smp in cohort["samples"]:  # cohort_c = the C1 substitute, 64 samples x ~1,200 records: count + sha256 per sample
        assert_sample_matches_reference(cohort, smp, fasta.get(smp, []))


def test_task_rs_unit_vector_keeps_uncovered_units(gpu_engine):
    """task.rs:118-144: unsorted tasks, 'x'-filled tape, no '.' fill -> serial-order kernel + keep_out."""
    res = gpu_engine.execute_soa([(0, 1, 1, 8), (0, 4, 1, 4), (0, 6, 2, 6)], "ABCFEFGH", "HGFEFCBA", "x" * 10,
                                 fill_dot=False)
    assert tape_to_str(res) == "xxxxExGHBx"
    # sorted variant takes the tile kernel with keep_out
    res = gpu_engine.execute_soa([(0, 4, 1, 4), (0, 6, 2, 6), (0, 1, 1, 8)], "ABCFEFGH", "HGFEFCBA", "x" * 10,
                                 fill_dot=False)
    assert tape_to_str(res) == "xxxxExGHBx"


def test_non_ascii_residues_ride_the_utf32_path(gpu_engine):
    ref = "MEDLé中\U0001F9EC" * 5
    res = gpu_engine.execute_soa([(0, 0, 20, 0), (1, 0, 2, 20), (0, 22, 13, 23)], ref, "αβ", 36, fill_dot=True)
    want = np.zeros(36, np.uint32)
    assert cengine.gir_execute([(0, 0, 20, 0), (1, 0, 2, 20), (0, 22, 13, 23)], u32(ref), u32("αβ"), want, True)[0] == 0
    assert np.array_equal(res, want) and tape_to_str(res)[22] == "."


# ---------------------------------------------------------------------------------------------- randomized parity
@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9])
@pytest.mark.parametrize("seed,n_hap,mean_res", [(1, 1, 100), (2, 7, 3000), (3, 40, 20000), (4, 300, 2000),
                                                 (5, 3, 3_000_000), (6, 64, 150_000)])
def test_random_batches_bit_exact(gpu_engine, variant, seed, n_hap, mean_res):
    gpu_engine.set_tuning(variant, 0)
    try:
        b = random_batch(seed, n_hap, mean_res)
        out, ms = gpu_batch(gpu_engine, b)
        st, _, _, want = oracle_batch(b)
        assert st == 0
        assert out.shape == want.shape and np.array_equal(out, want)
    finally:
        gpu_engine.set_tuning(-1, 0)


@pytest.mark.parametrize("variant", [0, 8])
@pytest.mark.parametrize("seed,n_hap,mean_res", [(61, 2, 50_000), (62, 97, 40_000), (63, 500, 9_000), (64, 33, 400_000)])
def test_tile_order_never_changes_results(gpu_engine, variant, seed, n_hap, mean_res):
    """The haplotype-interleaved tile order (default) and tape order (V2P_FLAG_ALIGNED_LAYOUT hint) are schedules of
    the same tiles: both give the oracle's bytes, with ragged haplotypes (exponential lengths leave most slots of the
    order table empty) and empty ones."""
    gpu_engine.set_tuning(variant, 0)
    try:
        b = random_batch(seed, n_hap, mean_res, n_ref=150_001, empty_hap_prob=0.15)
        st, _, _, want = oracle_batch(b)
        assert st == 0
        gpu_engine.set_reference(b["ref"], "replicas")
        for hint in (False, True):
            out, _ = gpu_engine.execute_batch(b["task_begin"], b["tasks"], None, b["alt"], b["alt_base"], b["out_base"],
                                              aligned_layout=hint)
            assert np.array_equal(out, want), hint
            out, _ = gpu_engine.execute_batch(b["task_begin"], b["tasks"], b["ref"], b["alt"], b["alt_base"], b["out_base"],
                                              aligned_layout=hint)
            assert np.array_equal(out, want), hint
    finally:
        gpu_engine.set_tuning(-1, 0)


@pytest.mark.parametrize("variant", [-1, 0, 3, 6, 8, 9])
@pytest.mark.parametrize("seed,n_hap,mean_res,run_mean", [(81, 3, 4000, 40.0), (82, 40, 60_000, 12.0), (83, 25, 200_000, 250.0),
                                                          (84, 200, 30_000, 60.0), (85, 6, 900_000, 400.0)])
def test_missense_chains_and_their_near_misses_bit_exact(gpu_engine, variant, seed, n_hap, mean_res, run_mean):
    """The copy kernel fuses `R A R` (same source - destination offset, 1-residue alteration in the hole) into one run
    plus a byte patch.  chain_batch is that pattern with every near-miss of it (randtasks.py); all tape modes."""
    gpu_engine.set_tuning(variant, 0)
    try:
        b = chain_batch(seed, n_hap, mean_res, run_mean=run_mean)
        st, _, _, want = oracle_batch(b)
        assert st == 0
        out, _ = gpu_batch(gpu_engine, b)  # caller-supplied tape: register path / in-phase TMA only
        assert np.array_equal(out, want)
        for mode in ("replicas", "plain"):
            gpu_engine.set_reference(b["ref"], mode)
            for hint in (False, True):
                out, _ = gpu_engine.execute_batch(b["task_begin"], b["tasks"], None, b["alt"], b["alt_base"], b["out_base"],
                                                  aligned_layout=hint)
                assert np.array_equal(out, want), (mode, hint)
    finally:
        gpu_engine.set_tuning(-1, 0)


def test_tile_order_falls_back_when_haplotypes_are_wildly_uneven(gpu_engine):
    """One 6 MB haplotype among 400 of a few bytes: the order table (max tiles per haplotype x haplotypes) would be
    ~100x the tile count, so the plan keeps tape order -- same bytes."""
    rng = np.random.default_rng(7)
    ref = rng.integers(65, 91, size=70_000, dtype=np.uint8)
    n_hap, big = 401, 200
    rows, tb, ob = [], [0], [0]
    for h in range(n_hap):
        n = 6_000_000 if h == big else int(rng.integers(0, 40))
        pos = 0
        while pos < n:
            ln = min(int(rng.integers(1, 60_000)), n - pos)
            rows.append((int(rng.integers(0, len(ref) - ln + 1)), ln, pos, 0))
            pos += ln
        tb.append(len(rows))
        ob.append(ob[-1] + n)
    u = lambda x: np.asarray(x, dtype=np.uint64)
    b = dict(task_begin=u(tb), tasks=np.asarray(rows, np.uint32).reshape(-1, 4), ref=ref, alt=np.zeros(0, np.uint8),
             alt_base=u([0] * (n_hap + 1)), out_base=u(ob), ref_base=None)
    out, _ = gpu_batch(gpu_engine, b)
    st, _, _, want = oracle_batch(b)
    assert st == 0 and np.array_equal(out, want)


@pytest.mark.parametrize("mode", ["replicas", "plain"])
@pytest.mark.parametrize("variant", [0, 2, 4, 5, 7, 8])
@pytest.mark.parametrize("seed,n_hap,mean_res", [(51, 5, 400), (52, 30, 30000), (53, 4, 2_000_000)])
def test_registered_reference_tma_path_bit_exact(gpu_engine, mode, variant, seed, n_hap, mean_res):
    """ref == NULL: tasks index the registered proteome; long reference runs are TMA bulk copies from the
    16 byte-shifted replicas, everything else takes the register path.  Same bytes as the oracle."""
    gpu_engine.set_tuning(variant, 0)
    try:
        b = random_batch(seed, n_hap, mean_res, n_ref=200_003)
        gpu_engine.set_reference(b["ref"], mode)
        out, _ = gpu_engine.execute_batch(b["task_begin"], b["tasks"], None, b["alt"], b["alt_base"], b["out_base"])
        st, _, _, want = oracle_batch(b)
        assert st == 0 and np.array_equal(out, want)
        # the generic (caller-supplied tape) path still works on the same engine afterwards
        out2, _ = gpu_batch(gpu_engine, b)
        assert np.array_equal(out2, want)
    finally:
        gpu_engine.set_tuning(-1, 0)


def test_registered_reference_cohort_and_errors(gpu_engine):
    from vcf2prot_b200 import GpuEngine
    from synth import cohort as C

    prot = C.make_proteome(seed=5, n_tx=400, mu=5.5, sigma=0.7, hi=6000)
    cat = C.make_catalogue(prot, 9000, seed=6, mix=(0.7, 0.06, 0.06, 0.08, 0.04, 0.03, 0.03), fs_mean=40, fs_max=900)
    b = C.synth_batch(prot, cat, 40, 7, ref_mode="global")
    gpu_engine.set_reference(prot.residues)
    out, _ = gpu_engine.execute_batch(b.task_begin, b.tasks, None, b.alt, b.alt_base, b.out_base, validate=True)
    want = np.zeros(b.n_residues, np.uint8)
    assert cengine.batch_execute(b.task_begin, b.tasks, prot.residues, b.alt, b.alt_base, want, b.out_base)[0] == 0
    assert np.array_equal(out, want)
    # a source slice beyond the registered tape is still the reference's slice panic
    bad = b.tasks.copy()
    k = int(np.flatnonzero(bad[:, 3] == 0)[5])
    bad[k, 0] = len(prot.residues) - 1
    bad[k, 1] = max(bad[k, 1], 2)
    with pytest.raises(EngineError) as ei:
        gpu_engine.execute_batch(b.task_begin, bad, None, b.alt, b.alt_base, b.out_base)
    assert ei.value.status in (L.ERR_SRC_OOB, L.ERR_RES_OOB)
    # no registered reference on a fresh engine -> loud failure, not a guess
    with GpuEngine(0) as fresh:
        with pytest.raises(EngineError) as ei:
            fresh.execute_batch(b.task_begin, b.tasks, None, b.alt, b.alt_base, b.out_base)
        assert ei.value.status == L.ERR_INVALID_ARG


@pytest.mark.parametrize("mix", [(1.0, 0, 0, 0), (0, 1.0, 0, 0), (0, 0, 0, 1.0), (0.5, 0.5, 0, 0)])
def test_extreme_length_mixes(gpu_engine, mix):
    """all 1-byte tasks (>32 tasks per tile -> multi-batch tiles), all short, all long (skew)."""
    b = random_batch(11, 6, 30000, len_mix=mix, gap_prob=0.1)
    out, _ = gpu_batch(gpu_engine, b)
    st, _, _, want = oracle_batch(b)
    assert st == 0 and np.array_equal(out, want)


def test_edge_shapes(gpu_engine):
    z64 = lambda *a: np.asarray(a, dtype=np.uint64)
    ref = u8("ABCDEFGHIJKLMNOPQRSTUVWXYZ")
    # no haplotypes at all
    out, _ = gpu_engine.execute_batch(z64(0), np.zeros((0, 4), np.uint32), ref, u8(""), z64(0), z64(0))
    assert out.size == 0
    # one haplotype, zero tasks, 5-residue tape -> all dots; one with zero-length tape
    out, _ = gpu_engine.execute_batch(z64(0, 0, 0), np.zeros((0, 4), np.uint32), ref, u8(""), z64(0, 0, 0), z64(0, 5, 5))
    assert tape_to_str(out) == "....."
    # zero-length tasks, a task ending exactly at the tape end, tapes whose sizes are not multiples of 16
    tasks = np.asarray([(0, 0, 0, 0), (3, 4, 0, 0), (0, 0, 7, 0), (25, 1, 6, 0), (0, 3, 7, 1), (3, 0, 10, 1)], np.uint32)
    out, _ = gpu_engine.execute_batch(z64(0, 2, 6), tasks, ref, u8("xyz"), z64(0, 0, 3), z64(0, 9, 9 + 23))
    b = dict(task_begin=z64(0, 2, 6), tasks=tasks, ref=ref, alt=u8("xyz"), alt_base=z64(0, 0, 3), out_base=z64(0, 9, 32))
    st, _, _, want = oracle_batch(b)
    assert st == 0 and np.array_equal(out, want)


def test_unsorted_and_overlapping_tasks_keep_serial_semantics(gpu_engine):
    rng = np.random.default_rng(5)
    b = random_batch(21, 12, 5000)
    # shuffle tasks inside each haplotype and add overlapping rewrites: later task must win (gir.rs:233)
    t = b["tasks"].copy()
    tb = b["task_begin"]
    for h in range(len(tb) - 1):
        s, e = int(tb[h]), int(tb[h + 1])
        if e - s > 2:
            t[s:e] = t[s:e][rng.permutation(e - s)]
            t[e - 1] = t[s]  # duplicate destination -> overlap
            t[e - 1, 0] = 0
            t[e - 1, 3] = 0
            t[e - 1, 1] = min(t[e - 1, 1], 1000)
    b["tasks"] = t
    out, _ = gpu_batch(gpu_engine, b)
    st, _, _, want = oracle_batch(b)
    assert st == 0 and np.array_equal(out, want)


def test_error_classes_match_the_reference_panics(gpu_engine):
    base = random_batch(31, 9, 4000, gap_prob=0.0, empty_hap_prob=0.0)
    tb = base["task_begin"]

    def corrupt(h, k, col, val):
        b = dict(base)
        b["tasks"] = base["tasks"].copy()
        b["tasks"][int(tb[h]) + k, col] = val
        return b

    # (haplotype, task, column, value, expected ABI status, expected oracle status)
    cases = [(4, 2, 3, 2, L.ERR_BAD_STREAM, cengine.REF_ERR_BAD_STREAM),
             (2, 1, 1, 0x7FFFFFFF, L.ERR_RES_OOB, cengine.REF_ERR_RES_OOB),
             (6, 0, 0, 0x7FFFFF00, L.ERR_SRC_OOB, cengine.REF_ERR_SRC_OOB)]
    for h, k, col, val, want_abi, want_ref in cases:
        b = corrupt(h, k, col, val)
        with pytest.raises(EngineError) as ei:
            gpu_batch(gpu_engine, b)
        st, bh, bi, _ = oracle_batch(b)
        assert st == want_ref and ei.value.status == want_abi
        assert (ei.value.bad_hap, ei.value.bad_task) == (bh, bi) == (h, k)
    # DEBUG_CPU_EXEC contiguity validator: introduce a 1-residue gap in haplotype 3
    b = dict(base)
    b["tasks"] = base["tasks"].copy()
    s = int(tb[3])
    k = next(i for i in range(1, int(tb[4]) - s - 1) if b["tasks"][s + i, 1] >= 2)
    b["tasks"][s + k, 1] -= 1  # shorten task k -> task k+1 no longer starts where it ends
    with pytest.raises(EngineError) as ei:
        gpu_batch(gpu_engine, b, validate=True)
    st, bh, bi, _ = oracle_batch(b, validate=True)
    assert st == cengine.REF_ERR_NOT_CONTIGUOUS and ei.value.status == L.ERR_NOT_CONTIGUOUS
    assert (ei.value.bad_hap, ei.value.bad_task) == (bh, bi) == (3, k + 1)
    # without VALIDATE the same input is legal: the gap reads '.'
    out, _ = gpu_batch(gpu_engine, b)
    st, _, _, want = oracle_batch(b)
    assert st == 0 and np.array_equal(out, want) and (want == ord(".")).any()
    # SoA entry: same classes, index reported
    for tasks, want_abi in (([(0, 0, 9, 0)], L.ERR_RES_OOB), ([(1, 2, 2, 0)], L.ERR_SRC_OOB), ([(2, 0, 0, 0)], L.ERR_BAD_STREAM),
                            ([(0, 0, 1, 0), (0, 2**40, 2**63, 1)], L.ERR_RES_OOB)):
        with pytest.raises(EngineError) as ei:
            gpu_engine.execute_soa(tasks, "ABCDEFGH", "xyz", 8, fill_dot=True)
        assert ei.value.status == want_abi and ei.value.bad_task == len(tasks) - 1
    with pytest.raises(EngineError) as ei:
        GIR([(0, 0, 1, 0)], {}, "", "A", ".").execute(Engine.ST, gpu_engine)
    assert ei.value.status == L.ERR_NOT_GPU_ENGINE


def test_device_pointer_entry_and_async(gpu_engine):
    import torch

    b = random_batch(41, 50, 60000)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)
    d = {k: t(v) for k, v in b.items() if v is not None}
    n_out = int(b["out_base"][-1])
    out = torch.empty(n_out + 16, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    args = (len(b["task_begin"]) - 1, d["task_begin"], d["tasks"], d["ref"], d["alt"], d["alt_base"], out, d["out_base"],
            len(b["tasks"]), len(b["alt"]), n_out)
    ms = gpu_engine.execute_batch_device(*args)
    st, _, _, want = oracle_batch(b)
    assert st == 0 and np.array_equal(out[:n_out].cpu().numpy(), want) and ms > 0
    out.zero_()
    ev = gpu_engine.execute_batch_device(*args, wait=False)
    ms2 = gpu_engine.wait_event(ev)
    assert np.array_equal(out[:n_out].cpu().numpy(), want) and ms2 > 0
    # inconsistent totals are rejected by the plan kernel, not trusted
    bad = list(args)
    bad[10] = n_out - 1
    with pytest.raises(EngineError) as ei:
        gpu_engine.execute_batch_device(*bad)
    assert ei.value.status == L.ERR_INVALID_ARG
    # findings are localised (haplotype, task within it) without a host copy of task_begin: every class
    tb = b["task_begin"]
    for h, code, field, value in ((7, L.ERR_BAD_STREAM, 3, 9), (11, L.ERR_SRC_OOB, 0, 0xFFFFFF00), (13, L.ERR_RES_OOB, 1, 0x7FFFFFFF)):
        k = int(tb[h]) + (int(tb[h + 1]) - int(tb[h])) // 2
        tk = b["tasks"].copy()
        tk[k, field] = value
        bad = list(args)
        bad[2] = t(tk)
        with pytest.raises(EngineError) as ei:
            gpu_engine.execute_batch_device(*bad)
        want_err = cengine.batch_execute(b["task_begin"], tk, b["ref"], b["alt"], b["alt_base"], np.zeros(n_out, np.uint8), b["out_base"])
        assert (ei.value.status, ei.value.bad_hap, ei.value.bad_task) == (code, h, k - int(tb[h])), (h, code)
        assert want_err[1] == h and want_err[2] == k - int(tb[h])
    # caller-owned stream (torch's current stream)
    side = torch.cuda.Stream()
    gpu_engine.set_stream(side.cuda_stream)
    try:
        out.zero_()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(side):
            t0.record()
            gpu_engine.execute_batch_device(*args)
            t1.record()
        torch.cuda.synchronize()
        assert np.array_equal(out[:n_out].cpu().numpy(), want) and t0.elapsed_time(t1) > 0
    finally:
        gpu_engine.set_stream(None)


def test_full_size_properties(gpu_engine):
    """At a size the scalar oracle would take too long for, use properties: every output byte is either '.' or
    equals the source byte its task names; identity Task arrays reproduce the reference tape (round trip)."""
    n_ref = 64 << 20
    rng = np.random.default_rng(7)
    ref = rng.integers(65, 91, size=n_ref, dtype=np.uint8)
    # identity: one haplotype = the whole tape cut into irregular chunks
    cuts = np.unique(np.concatenate([[0, n_ref], rng.integers(0, n_ref, size=200000)]))
    tasks = np.zeros((len(cuts) - 1, 4), np.uint32)
    tasks[:, 0] = cuts[:-1]
    tasks[:, 1] = np.diff(cuts)
    tasks[:, 2] = cuts[:-1]
    z = lambda *a: np.asarray(a, dtype=np.uint64)
    out, ms = gpu_engine.execute_batch(z(0, len(tasks)), tasks, ref, np.zeros(0, np.uint8), z(0, 0), z(0, n_ref))
    assert np.array_equal(out, ref)
    # shifted copy: dst = src + 3 everywhere except the first 3 bytes ('.')
    tasks[:, 2] = cuts[:-1] + 3
    tasks[-1, 1] -= 3
    out, ms = gpu_engine.execute_batch(z(0, len(tasks)), tasks, ref, np.zeros(0, np.uint8), z(0, 0), z(0, n_ref))
    assert (out[:3] == ord(".")).all() and np.array_equal(out[3:], ref[:-3])


def test_host_pointer_pipeline_three_chunks_in_flight(gpu_engine):
    """The streaming shape: per-chunk host-pointer calls with V2P_FLAG_ASYNC rotate over the engine's 3 staging slots
    (copy-back of chunk i overlaps upload + kernels of chunk i+1); a 4th un-waited batch is refused, not queued."""
    from synth import cohort as C

    prot = C.make_proteome(seed=15, n_tx=300, mu=5.5, sigma=0.7, hi=5000)
    cat = C.make_catalogue(prot, 6000, seed=16, mix=(0.8, 0.05, 0.05, 0.05, 0.02, 0.02, 0.01))
    cat.af[:] = 0.15
    b = C.synth_batch(prot, cat, 48, 17, ref_mode="global", layout="aligned")
    want = np.zeros(b.n_residues, np.uint8)
    assert cengine.batch_execute(b.task_begin, b.tasks, prot.residues, b.alt, b.alt_base, want, b.out_base)[0] == 0
    gpu_engine.set_reference(prot.residues)
    chunks = [(h, min(48, h + 7)) for h in range(0, 48, 7)]
    bufs = [np.zeros(int(max(b.out_base[c[1]] - b.out_base[c[0]] for c in chunks)) + 16, np.uint8) for _ in range(3)]
    got = np.zeros_like(want)
    pending = []

    def drain_one():
        ev, (a, z), buf = pending.pop(0)
        gpu_engine.wait_event(ev)
        o0, o1 = int(b.out_base[a]), int(b.out_base[z])
        got[o0:o1] = buf[:o1 - o0]

    for i, (a, z) in enumerate(chunks):
        if len(pending) == 3:
            drain_one()
        ev = gpu_engine.execute_hap_range(a, z, b.task_begin, b.tasks, None, b.alt, b.alt_base, b.out_base, bufs[i % 3],
                                          wait=False)
        pending.append((ev, (a, z), bufs[i % 3]))
    assert len(pending) == 3
    with pytest.raises(EngineError) as ei:  # all three slots busy
        gpu_engine.execute_hap_range(0, 1, b.task_begin, b.tasks, None, b.alt, b.alt_base, b.out_base,
                                     np.zeros(int(b.out_base[1]) + 16, np.uint8), wait=False)
    assert ei.value.status == L.ERR_INVALID_ARG
    while pending:
        drain_one()
    assert np.array_equal(got, want)


def test_fasta_file_image_on_device(gpu_engine):
    """record framing as copy segments -> the D2H buffer is the .fasta file image, byte for byte what
    the oracle produces, and it parses into exactly the records of the plain batch."""
    from synth import cohort as C

    prot = C.make_proteome(seed=25, n_tx=200, mu=5.2, sigma=0.6, hi=3000)
    cat = C.make_catalogue(prot, 5000, seed=26, mix=(0.7, 0.06, 0.06, 0.08, 0.04, 0.03, 0.03), fs_mean=30, fs_max=500)
    cat.af[:] = 0.2
    b = C.synth_batch(prot, cat, 20, 27, ref_mode="global")
    fb = C.fasta_image(prot, b)
    gpu_engine.set_reference(prot.residues)
    img, _ = gpu_engine.execute_batch(fb.task_begin, fb.tasks, None, fb.alt, fb.alt_base, fb.out_base, validate=True)
    want = np.zeros(fb.n_residues, np.uint8)
    assert cengine.batch_execute(fb.task_begin, fb.tasks, prot.residues, fb.alt, fb.alt_base, want, fb.out_base)[0] == 0
    assert np.array_equal(img, want)
    plain = np.zeros(b.n_residues, np.uint8)
    assert cengine.batch_execute(b.task_begin, b.tasks, prot.residues, b.alt, b.alt_base, plain, b.out_base)[0] == 0
    for s in range(10):
        image = img[int(fb.out_base[2 * s]):int(fb.out_base[2 * s + 2])]
        recs = sorted(C.fasta_records(prot, b, plain, 2 * s, 1) + C.fasta_records(prot, b, plain, 2 * s + 1, 2))
        assert C.parse_fasta_image(image) == recs and (len(image) == 0 or image[0] == ord(">"))


def test_concurrent_callers_share_one_context(gpu_engine):
    """GIR::execute is called from many rayon workers at once (parts/exec.rs:36-39): concurrent host threads on ONE
    context must each get their own haplotype's result (the context serialises them internally)."""
    import threading

    cohort = load_golden("cohort_a.json")
    per_hap = cohort_haplotype_csqs(cohort)
    girs = [hap_gir(c, cohort["refs"]) for _, c in sorted(per_hap.items())][:16]
    want = [taskgen.execute_tasks(g.tasks, g.ref, g.alt, g.res_len) for g in girs]
    got = [None] * len(girs)
    errs = []

    def worker(i):
        try:
            for _ in range(3):
                res, _ = GIR(girs[i].tasks, girs[i].annotation, girs[i].alt, girs[i].ref, "." * girs[i].res_len).execute(
                    Engine.GPU, gpu_engine)
                got[i] = tape_to_str(res)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(girs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs and got == want


def test_pinned_host_allocator_roundtrip(gpu_engine):
    import ctypes as C

    lib = L.load()
    p = C.c_void_p()
    assert lib.v2p_host_alloc(C.byref(p), 1 << 20) == 0 and p.value
    buf = (C.c_uint8 * (1 << 20)).from_address(p.value)
    out = np.frombuffer(buf, dtype=np.uint8)
    b = random_batch(77, 4, 50000)
    n = int(b["out_base"][-1])
    assert n <= out.size
    got, _ = gpu_engine.execute_batch(b["task_begin"], b["tasks"], b["ref"], b["alt"], b["alt_base"], b["out_base"], out=out[:n])
    st, _, _, want = oracle_batch(b)
    assert st == 0 and np.array_equal(got, want)
    assert lib.v2p_host_free(p) == 0


@pytest.mark.parametrize("variant", [-1, 0, 5, 8, 9])
def test_skew_stress_c4_mix_with_giant_transcripts(gpu_engine, variant):
    """SURVEY 8d C4 / BASELINE configs[3]: 35 % frameshift (log-normal tails up to 4,000), 20 % long inframe insertions
    (up to 5,000), 10 % stop_lost tails, 35 % missense, on a proteome with 35,000-residue transcripts -- the segments of
    transcript_instructions.rs:666-679 (frameshift tails) and :696-710 (stop_lost).  240 haplotypes, every tape mode,
    whole tapes against the oracle; and no warp of the persistent grid is pinned by a long segment."""
    from synth import cohort as C

    prot = C.make_proteome(seed=0x5EED0001, n_tx=1500, giant=12)
    cat = C.make_catalogue(prot, 9000, seed=0x5EED0004, mix=C.MIX_C4, fs_mean=150, fs_max=4000, sl_max=500, long_ins_mean=120,
                           long_ins_max=5000, lognormal_tails=True)  # (tails longer than C4's own, so that 9,000 sites reach the caps)
    cat.af[:] = np.random.default_rng(4).choice([0.02, 0.1, 0.3], size=cat.n).astype(np.float32)
    cat.af[cat.dlen > 400] = 0.4  # the long payloads are carried often
    b = C.synth_batch(prot, cat, 240, seed=0x5EED0004)
    lens, is_alt = b.tasks[:, 1], b.tasks[:, 3] == 1
    assert lens.max() > 5000 and lens[is_alt].max() > 1500  # giant reference runs and long alteration payloads
    want = np.zeros(b.n_residues, np.uint8)
    assert cengine.batch_execute(b.task_begin, b.tasks, prot.residues, b.alt, b.alt_base, want, b.out_base, threads=4)[0] == 0
    gpu_engine.set_tuning(variant, 0)
    try:
        out, _ = gpu_engine.execute_batch(b.task_begin, b.tasks, prot.residues, b.alt, b.alt_base, b.out_base, validate=True)
        assert np.array_equal(out, want)
        for mode in ("replicas", "plain"):
            gpu_engine.set_reference(prot.residues, mode)
            out, _ = gpu_engine.execute_batch(b.task_begin, b.tasks, None, b.alt, b.alt_base, b.out_base)
            assert np.array_equal(out, want), mode
        if variant == -1:  # load balance of the persistent grid on this input
            import torch

            dev = torch.device("cuda:0")
            up = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)
            d_out = torch.empty(b.n_residues + 16, dtype=torch.uint8, device=dev)
            args = (b.n_hap, up(b.task_begin), up(b.tasks), None, up(b.alt), up(b.alt_base), d_out, up(b.out_base), len(b.tasks),
                    len(b.alt), b.n_residues)
            gpu_engine.profile_warps(True)
            for _ in range(3):
                gpu_engine.execute_batch_device(*args)
            ns = gpu_engine.read_warp_ns().astype(np.float64)
            gpu_engine.profile_warps(False)
            assert np.array_equal(d_out[: b.n_residues].cpu().numpy(), want)
            # (a 100 MB batch is only 3-6 tiles per warp: the bound is one tile's granularity, the full-size figure is
            # bench.py's load_balance key: 1.005)
            assert ns.size > 0 and ns.max() / ns.mean() < 1.6, (ns.max(), ns.mean())
    finally:
        gpu_engine.set_tuning(-1, 0)
