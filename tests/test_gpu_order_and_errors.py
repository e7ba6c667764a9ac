"""GPU suite (-m gpu): what round 1's review found untested on the engine's error and order paths.

  * a rejected task with a garbage destination in FRONT of another haplotype's first task (plan_one's lb[] scatter
    must stay in range; the launch fails with the task's own status, nothing else is corrupted),
  * the reference's order of events: haplotype after haplotype; inside one, the stream-code panic of the Task-array
    construction (haplotype_instruction.rs:154) before the validator (gir.rs:203-229) before slice panics (task.rs:44/48),
  * per-haplotype serial order (gir.rs:233): ONE unsorted / overlapping haplotype among hundreds takes the span-parallel
    serial kernel, everybody else stays on the tile kernel -- same bytes as the oracle, and no cliff in time.
"""
import time

import numpy as np
import pytest

from oracle import cengine
from tests.randtasks import random_batch
from tests.test_gpu_parity import gpu_batch, oracle_batch
from vcf2prot_b200 import EngineError
from vcf2prot_b200 import _lib as L

pytestmark = pytest.mark.gpu


def test_garbage_destination_in_front_of_the_next_haplotype_is_a_clean_error(gpu_engine):
    base = random_batch(71, 40, 30_000, gap_prob=0.0, empty_hap_prob=0.0)
    tb = base["task_begin"]
    _, _, _, want = oracle_batch(base)
    for h in (0, 7, 38):  # the LAST task of haplotype h; haplotype h+1 is not empty
        for dst in (0xFFFFFFF0, 0x80000000, int(base["out_base"][-1])):
            b = dict(base)
            b["tasks"] = base["tasks"].copy()
            k = int(tb[h + 1]) - 1
            b["tasks"][k, 2] = dst
            with pytest.raises(EngineError) as ei:
                gpu_batch(gpu_engine, b)
            st, bh, bi, _ = oracle_batch(b)
            assert st == cengine.REF_ERR_RES_OOB and ei.value.status == L.ERR_RES_OOB
            assert (ei.value.bad_hap, ei.value.bad_task) == (bh, bi) == (h, k - int(tb[h]))
            # the context is intact: the same engine still produces the oracle's bytes
            out, _ = gpu_batch(gpu_engine, base)
            assert np.array_equal(out, want)


def test_error_precedence_follows_the_reference_order_of_events(gpu_engine):
    base = random_batch(72, 12, 6000, gap_prob=0.0, empty_hap_prob=0.0)
    tb = base["task_begin"]
    roomy = [h for h in range(12) if int(tb[h + 1]) - int(tb[h]) >= 8]  # haplotypes with at least 8 tasks
    assert len(roomy) >= 3
    ha, hb, hc = roomy[0], roomy[1], roomy[-1]  # ha < hb < hc

    def corrupt(*edits):
        b = dict(base)
        b["tasks"] = base["tasks"].copy()
        for h, k, col, val in edits:
            b["tasks"][int(tb[h]) + k, col] = val
        return b

    def both(b, validate=False):
        with pytest.raises(EngineError) as ei:
            gpu_batch(gpu_engine, b, validate=validate)
        st, bh, bi, _ = oracle_batch(b, validate=validate)
        return (st, bh, bi), (ei.value.status, ei.value.bad_hap, ei.value.bad_task)

    # same haplotype: slice error at task 3, bad stream at task 5 -> the construction panic (task 5) comes first
    ref, gpu = both(corrupt((hb, 3, 1, 0x7FFFFFFF), (hb, 5, 3, 7)))
    assert ref == (cengine.REF_ERR_BAD_STREAM, hb, 5) and gpu == (L.ERR_BAD_STREAM, hb, 5)
    # ... and with the validator on, a gap in front of it still loses to the stream code
    n_b = int(tb[hb + 1]) - int(tb[hb])
    k = next(i for i in range(1, n_b - 1) if base["tasks"][int(tb[hb]) + i, 1] >= 2)  # shortening it opens a gap
    last = n_b - 1
    ref, gpu = both(corrupt((hb, k, 1, int(base["tasks"][int(tb[hb]) + k, 1]) - 1), (hb, last, 3, 2)), validate=True)
    assert ref == (cengine.REF_ERR_BAD_STREAM, hb, last) and gpu == (L.ERR_BAD_STREAM, hb, last)
    # different haplotypes: the earlier haplotype's slice panic happens before the later one is even built
    ref, gpu = both(corrupt((ha, 1, 1, 0x7FFFFFFF), (hc, 0, 3, 2)))
    assert ref == (cengine.REF_ERR_RES_OOB, ha, 1) and gpu == (L.ERR_RES_OOB, ha, 1)
    # ... and the other way round
    ref, gpu = both(corrupt((ha, 2, 3, 9), (hc, 0, 1, 0x7FFFFFFF)))
    assert ref == (cengine.REF_ERR_BAD_STREAM, ha, 2) and gpu == (L.ERR_BAD_STREAM, ha, 2)
    # SoA entry (one haplotype): stream code at the end beats the slice error in front of it
    with pytest.raises(EngineError) as ei:
        gpu_engine.execute_soa([(0, 0, 9, 0), (0, 0, 1, 0), (2, 0, 0, 0)], "ABCDEFGH", "xyz", 8, fill_dot=True)
    assert ei.value.status == L.ERR_BAD_STREAM and ei.value.bad_task == 2


def _shuffle_haplotype(b, h, rng, overlap=True):
    tb = b["task_begin"]
    s, e = int(tb[h]), int(tb[h + 1])
    t = b["tasks"]
    t[s:e] = t[s:e][rng.permutation(e - s)]
    if overlap and e - s > 2:  # a later task rewrites an earlier one's bytes: the later one must win (gir.rs:233)
        t[e - 1] = t[s]
        t[e - 1, 0], t[e - 1, 3] = 0, 0
        t[e - 1, 1] = min(int(t[e - 1, 1]), 1000)


@pytest.mark.parametrize("variant", [0, 8])
@pytest.mark.parametrize("n_hap,mean_res,odd", [(60, 20_000, (17,)), (300, 9_000, (0, 299)), (41, 120_000, (5, 6, 7)),
                                                (9, 3000, tuple(range(9)))])
def test_only_the_unsorted_haplotypes_take_the_serial_kernel(gpu_engine, variant, n_hap, mean_res, odd):
    gpu_engine.set_tuning(variant, 0)
    try:
        rng = np.random.default_rng(n_hap)
        b = random_batch(200 + n_hap, n_hap, mean_res, empty_hap_prob=0.05)
        b["tasks"] = b["tasks"].copy()
        for h in odd:
            _shuffle_haplotype(b, h, rng)
        st, _, _, want = oracle_batch(b)
        assert st == 0
        out, _ = gpu_batch(gpu_engine, b)
        assert np.array_equal(out, want)
        gpu_engine.set_reference(b["ref"], "replicas")  # and with the registered tape (TMA path for everybody else)
        out, _ = gpu_engine.execute_batch(b["task_begin"], b["tasks"], None, b["alt"], b["alt_base"], b["out_base"])
        assert np.array_equal(out, want)
    finally:
        gpu_engine.set_tuning(-1, 0)


def test_one_unsorted_haplotype_among_500_costs_no_cliff(gpu_engine):
    """one odd haplotype used to send the whole batch to one-CTA-per-haplotype serial execution."""
    import torch

    rng = np.random.default_rng(500)
    good = random_batch(500, 500, 400_000, gap_prob=0.0, empty_hap_prob=0.0, len_mix=(0.3, 0.2, 0.45, 0.05))
    bad = dict(good)
    bad["tasks"] = good["tasks"].copy()
    _shuffle_haplotype(bad, 250, rng, overlap=False)
    dev = torch.device("cuda:0")
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)
    n_out = int(good["out_base"][-1])
    out = torch.empty(n_out + 16, dtype=torch.uint8, device=dev)
    times = {}
    for name, b in (("sorted", good), ("one_unsorted", bad)):
        d = {k: up(v) for k, v in b.items() if v is not None}
        args = (500, d["task_begin"], d["tasks"], d["ref"], d["alt"], d["alt_base"], out, d["out_base"], len(b["tasks"]),
                len(b["alt"]), n_out)
        for _ in range(3):
            gpu_engine.execute_batch_device(*args)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            gpu_engine.execute_batch_device(*args)  # synchronous: includes the serial kernel when one is needed
        torch.cuda.synchronize()
        times[name] = (time.perf_counter() - t0) / 10
        st, _, _, want = oracle_batch(b)
        assert st == 0 and np.array_equal(out[:n_out].cpu().numpy(), want)
    print("500 haplotypes, %.1f MB: sorted %.3f ms, one unsorted haplotype %.3f ms" %
          (n_out / 1e6, times["sorted"] * 1e3, times["one_unsorted"] * 1e3))
    assert times["one_unsorted"] < 1.5 * times["sorted"] + 0.3e-3


def test_gir_execute_from_many_threads_at_once(gpu_engine):
    """parts/exec.rs:36-39: GIR::execute is called from many rayon workers concurrently, one haplotype each.  Every
    caller owns a slot of the engine for the call (the entry used to hold the engine lock across
    its copies): 12 threads x 6 calls with different haplotypes, tapes, errors and a non-ASCII tape in between."""
    import threading

    from synth import cohort as C

    prot = C.make_proteome(seed=91, n_tx=500, mu=5.6, sigma=0.7, hi=7000)
    cat = C.make_catalogue(prot, 15000, seed=92, mix=(0.7, 0.06, 0.06, 0.08, 0.04, 0.03, 0.03))
    cat.af[:] = 0.2
    b = C.synth_batch(prot, cat, 24, 93, ref_mode="per_hap")
    haps = []
    for h in range(24):
        t0, t1 = int(b.task_begin[h]), int(b.task_begin[h + 1])
        tk = b.tasks[t0:t1]
        tasks = [(int(r[3]), int(r[0]), int(r[1]), int(r[2])) for r in tk]  # (exe_code, start_pos, length, start_pos_res)
        ref = b.ref[int(b.ref_base[h]):int(b.ref_base[h + 1])].astype(np.uint32)
        alt = b.alt[int(b.alt_base[h]):int(b.alt_base[h + 1])].astype(np.uint32)
        n_res = int(b.out_base[h + 1] - b.out_base[h])
        want = np.zeros(n_res, np.uint32)
        assert cengine.gir_execute(tasks, ref, alt, want, True)[0] == 0
        haps.append((tasks, ref, alt, n_res, want))
    errors = []

    def worker(w):
        try:
            for i in range(6):
                tasks, ref, alt, n_res, want = haps[(w * 5 + i * 7) % 24]
                if (w + i) % 5 == 0:  # a rejected call in between leaves the slot reusable
                    with pytest.raises(EngineError) as ei:
                        gpu_engine.execute_soa(tasks[:3] + [(2, 0, 0, 0)], ref, alt, n_res, fill_dot=True)
                    assert ei.value.status == L.ERR_BAD_STREAM and ei.value.bad_task == 3
                if (w + i) % 7 == 0:  # code points above 0xFF take the UTF-32 path on the same slot
                    r2 = ref.copy()
                    r2[::97] = 0x4E2D
                    w2 = np.zeros(n_res, np.uint32)
                    assert cengine.gir_execute(tasks, r2, alt, w2, True)[0] == 0
                    assert np.array_equal(gpu_engine.execute_soa(tasks, r2, alt, n_res, fill_dot=True), w2)
                got = gpu_engine.execute_soa(tasks, ref, alt, n_res, fill_dot=True)
                assert np.array_equal(got, want)
        except BaseException as ex:  # noqa: BLE001 -- reported by the main thread
            errors.append((w, repr(ex)))

    th = [threading.Thread(target=worker, args=(w,)) for w in range(12)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errors, errors


def test_dot_fill_never_depends_on_what_the_tile_buffers_held_before(gpu_engine):
    """The copy kernel prefills a tile only where the plan found a byte that no task covers.  Dirty every tile buffer of
    the GPU with letters first, then run batches whose '.' bytes come from every source of a gap: haplotypes without any
    task (with and without neighbours that have tasks, and a batch with no task at all), gaps between tasks, in front of
    the first and behind the last task of a haplotype, a haplotype left to the serial kernel."""
    z64 = lambda *a: np.asarray(a, dtype=np.uint64)
    dirty = random_batch(901, 64, 400_000, gap_prob=0.0, empty_hap_prob=0.0)
    ref = dirty["ref"]

    def run(b):
        gpu_batch(gpu_engine, dirty)  # every warp's tile buffer now holds letters
        out, _ = gpu_batch(gpu_engine, b)
        st, _, _, want = oracle_batch(b)
        assert st == 0 and np.array_equal(out, want)
        return want

    # no task at all: three haplotypes of 20,000 / 0 / 70,001 residues
    w = run(dict(task_begin=z64(0, 0, 0, 0), tasks=np.zeros((0, 4), np.uint32), ref=ref, alt=np.zeros(0, np.uint8),
                 alt_base=z64(0, 0, 0, 0), out_base=z64(0, 20_000, 20_000, 90_001), ref_base=None))
    assert (w == ord(".")).all()
    # a haplotype without tasks between two that have some; gaps in front of / between / behind the tasks of the others
    tasks = np.asarray([(5, 3000, 100, 0), (9000, 12_000, 3200, 0), (17, 40, 30_000, 0),   # haplotype 0 (40,000 residues)
                        (100, 9000, 0, 0), (20_000, 1, 9000, 0), (7, 5000, 30_000, 0)],   # haplotype 2 (36,000 residues)
                       np.uint32)
    w = run(dict(task_begin=z64(0, 3, 3, 6), tasks=tasks, ref=ref, alt=np.zeros(0, np.uint8), alt_base=z64(0, 0, 0, 0),
                 out_base=z64(0, 40_000, 65_000, 101_000), ref_base=None))
    assert (w[40_000:65_000] == ord(".")).all() and (w[:100] == ord(".")).all() and w[100] != ord(".") and (w[-1000:] == ord(".")).all()
    # the same with haplotype 0's tasks out of order (serial kernel) -- its gaps and its neighbours' stay '.'
    t2 = tasks.copy()
    t2[[0, 2]] = t2[[2, 0]]
    run(dict(task_begin=z64(0, 3, 3, 6), tasks=t2, ref=ref, alt=np.zeros(0, np.uint8), alt_base=z64(0, 0, 0, 0),
             out_base=z64(0, 40_000, 65_000, 101_000), ref_base=None))
    # SoA entry with FILL_DOT and no task
    gpu_batch(gpu_engine, dirty)
    res = gpu_engine.execute_soa([], "ABCDEFGH", "xyz", 50_000, fill_dot=True)
    assert (res == ord(".")).all()
