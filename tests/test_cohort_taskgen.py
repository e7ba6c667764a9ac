"""CPU suite: the vectorised host producer (vcf2prot_b200/cohort.py) emits exactly the Task arrays, alt tapes,
ref tapes and annotations of the reference-pinned restatement (oracle/taskgen.py) for the same variant sites."""
import numpy as np
import pytest

from oracle import cengine, taskgen
from tests.helpers import hap_gir, tape_to_str
from synth import cohort as C

RICH_MIX = (0.55, 0.10, 0.10, 0.08, 0.07, 0.05, 0.05)


@pytest.fixture(scope="module")
def small():
    prot = C.make_proteome(seed=11, n_tx=150, mu=5.0, sigma=0.6, lo=30, hi=2000)
    cat = C.make_catalogue(prot, 4000, seed=12, mix=RICH_MIX, fs_mean=20, fs_max=300)
    cat.af[:] = np.random.default_rng(3).choice([0.02, 0.1, 0.3], size=cat.n)
    return prot, cat


def test_catalogue_has_every_class(small):
    prot, cat = small
    assert set(np.unique(cat.cls)) == set(range(7))


@pytest.mark.parametrize("seed", [1, 2])
def test_tasks_match_oracle_tuple_for_tuple(small, seed):
    prot, cat = small
    n_hap = 24
    b = C.synth_batch(prot, cat, n_hap, seed, ref_mode="per_hap")
    refs = {prot.name(t): prot.seq(t) for t in range(prot.n_tx)}
    n_checked = 0
    for h in range(n_hap):
        sites = b.kept_site[b.kept_hap == h]
        csqs = [C.site_csq(prot, cat, int(i)) for i in sites]
        g = hap_gir(csqs, refs) if csqs else taskgen.HaplotypeGIR([], {}, "", "", 0)
        assert not g.skipped
        t0, t1 = int(b.task_begin[h]), int(b.task_begin[h + 1])
        got = [(int(r[3]), int(r[0]), int(r[1]), int(r[2])) for r in b.tasks[t0:t1]]
        assert got == [tuple(t) for t in g.tasks], "haplotype %d" % h
        assert b.alt[int(b.alt_base[h]):int(b.alt_base[h + 1])].tobytes().decode() == g.alt
        assert b.ref[int(b.ref_base[h]):int(b.ref_base[h + 1])].tobytes().decode() == g.ref
        assert int(b.out_base[h + 1] - b.out_base[h]) == g.res_len
        rows = np.flatnonzero(b.ann_hap == h)
        ann = {prot.name(int(b.ann_tx[r])): (int(b.ann_start[r]), int(b.ann_end[r])) for r in rows}
        assert ann == g.annotation
        n_checked += len(got)
    assert n_checked > 1000


def test_global_and_per_hap_layouts_produce_identical_bytes(small):
    prot, cat = small
    outs = []
    for mode in ("per_hap", "global"):
        b = C.synth_batch(prot, cat, 16, 5, ref_mode=mode)
        out = np.zeros(b.n_residues, np.uint8)
        st, _, _ = cengine.batch_execute(b.task_begin, b.tasks, b.ref, b.alt, b.alt_base, out, b.out_base,
                                         ref_base=b.ref_base, validate=True, threads=2)
        assert st == 0  # also: no '.' gaps for these classes (DEBUG_CPU_EXEC contiguity holds inside a haplotype)
        outs.append(out)
    assert np.array_equal(outs[0], outs[1]) and not (outs[0] == ord(".")).any()


def test_aligned_layout_yields_identical_records(small):
    """B200 layout: transcripts placed in phase with the proteome tape, pad bytes read '.', records unchanged."""
    prot, cat = small
    recs = []
    for layout in ("packed", "aligned"):
        b = C.synth_batch(prot, cat, 12, 21, ref_mode="global", layout=layout)
        out = np.zeros(b.n_residues, np.uint8)
        assert cengine.batch_execute(b.task_begin, b.tasks, b.ref, b.alt, b.alt_base, out, b.out_base)[0] == 0
        recs.append([C.fasta_records(prot, b, out, h, 1) for h in range(12)])
        if layout == "aligned":
            starts = b.ann_start[b.ann_end > b.ann_start]
            tx = b.ann_tx[b.ann_end > b.ann_start]
            assert ((starts - prot.offsets[tx]) % 16 == 0).all() and (b.out_base % 16 == 0).all()
            # long alteration payloads sit in phase with their destination, every tape base is 16-byte aligned
            longt = (b.tasks[:, 3] == 1) & (b.tasks[:, 1] >= 32)
            assert longt.any() and ((b.tasks[longt, 0].astype(np.int64) - b.tasks[longt, 2]) % 16 == 0).all()
            assert (b.alt_base % 16 == 0).all()
            # sorted, non-overlapping, inside the tape (what the engine's fast path needs)
            for h in range(12):
                t = b.tasks[int(b.task_begin[h]):int(b.task_begin[h + 1])].astype(np.int64)
                assert (t[1:, 2] >= t[:-1, 2] + t[:-1, 1]).all()
    assert recs[0] == recs[1] and sum(len(r) for r in recs[0]) > 100


def test_fasta_records_match_reference_binary(small, tmp_path):
    """End to end against the reference's own binary when it is present (authoring container)."""
    from oracle import refbin

    if not refbin.available():
        pytest.skip("reference binary not present")
    prot, cat = small
    n_samples = 6
    b = C.synth_batch(prot, cat, 2 * n_samples, 9, ref_mode="global")
    out = np.zeros(b.n_residues, np.uint8)
    assert cengine.batch_execute(b.task_begin, b.tasks, b.ref, b.alt, b.alt_base, out, b.out_base)[0] == 0
    refs = {prot.name(t): prot.seq(t) for t in range(prot.n_tx)}
    samples = ["S%02d" % i for i in range(n_samples)]
    used = np.unique(b.kept_site)
    records = []
    for i in used:
        cells = []
        for s in range(n_samples):
            h1 = bool(((b.kept_hap == 2 * s) & (b.kept_site == i)).any())
            h2 = bool(((b.kept_hap == 2 * s + 1) & (b.kept_site == i)).any())
            cells.append(([0] if h1 else [], [0] if h2 else []))
        records.append(([C.site_csq(prot, cat, int(i))], cells))
    recs, stdout, rc = refbin.run_reference(refbin.vcf_text(samples, records), refs, "st")
    assert rc == 0, stdout[-1500:]
    # device-side FASTA framing (headers/newlines as copy segments): the result tape is the file image
    fb = C.fasta_image(prot, b)
    img = np.zeros(fb.n_residues, np.uint8)
    assert cengine.batch_execute(fb.task_begin, fb.tasks, fb.ref, fb.alt, fb.alt_base, img, fb.out_base, validate=True)[0] == 0
    for s, name in enumerate(samples):
        mine = sorted(C.fasta_records(prot, b, out, 2 * s, 1) + C.fasta_records(prot, b, out, 2 * s + 1, 2))
        assert mine == [tuple(r) for r in recs.get(name, [])], name
        file_image = img[int(fb.out_base[2 * s]):int(fb.out_base[2 * s + 2])]
        assert C.parse_fasta_image(file_image) == [tuple(r) for r in recs.get(name, [])], name
