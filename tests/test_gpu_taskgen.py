"""GPU suite: the device-side Task generator is bit-exact against the host producer
(vcf2prot_b200/cohort.py, itself pinned to the reference through the oracle) -- every output array -- and the batch it
generates executes to the oracle's bytes."""
import numpy as np
import pytest

from oracle import cengine
from synth import cohort as C
from vcf2prot_b200.taskgen import DeviceCatalogue, execute_generated

pytestmark = pytest.mark.gpu

RICH_MIX = (0.55, 0.10, 0.10, 0.08, 0.07, 0.05, 0.05)


@pytest.fixture(scope="module")
def world():
    prot = C.make_proteome(seed=31, n_tx=400, mu=5.3, sigma=0.7, lo=30, hi=4000)
    cat = C.make_catalogue(prot, 12000, seed=32, mix=RICH_MIX, fs_mean=30, fs_max=600, sl_max=120)
    cat.af[:] = np.random.default_rng(5).choice([0.01, 0.05, 0.2, 0.5], size=cat.n)
    dc = DeviceCatalogue(prot, cat, 0)
    yield prot, cat, dc
    dc.close()


@pytest.mark.parametrize("aligned", [False, True])
@pytest.mark.parametrize("seed,n_hap", [(1, 1), (2, 9), (3, 64)])
def test_every_array_matches_the_host_producer(world, gpu_engine, aligned, seed, n_hap):
    prot, cat, dc = world
    hap, site = C.select_sites(cat, n_hap, np.random.default_rng(seed))
    if n_hap == 9:  # haplotypes without any site in the middle and at the end
        keep = ~np.isin(hap, (3, 8))
        hap, site = hap[keep], site[keep]
    want = C.build_batch(prot, cat, hap, site, n_hap, "global", "aligned" if aligned else "packed")
    g = dc.generate(hap, site, n_hap, aligned)
    b = g.batch
    assert (b.n_hap, b.n_tasks, b.n_alt, b.n_out) == (n_hap, len(want.tasks), len(want.alt), want.n_residues)
    assert np.array_equal(dc.read(b.task_begin, n_hap + 1, np.uint64), want.task_begin)
    assert np.array_equal(dc.read(b.out_base, n_hap + 1, np.uint64), want.out_base)
    assert np.array_equal(dc.read(b.alt_base, n_hap + 1, np.uint64), want.alt_base)
    assert np.array_equal(dc.read(b.tasks, 4 * b.n_tasks, np.uint32).reshape(-1, 4), want.tasks)
    assert np.array_equal(dc.read(b.alt, b.n_alt, np.uint8), want.alt)
    assert g.n_rows == len(want.ann_hap)
    assert np.array_equal(dc.read(g.ann_hap, g.n_rows, np.uint32), want.ann_hap)
    assert np.array_equal(dc.read(g.ann_tx, g.n_rows, np.uint32), want.ann_tx)
    assert np.array_equal(dc.read(g.ann_start, g.n_rows, np.uint64), want.ann_start)
    assert np.array_equal(dc.read(g.ann_end, g.n_rows, np.uint64), want.ann_end)
    # and the generated batch runs to the oracle's bytes, tasks never having existed on the host
    gpu_engine.set_reference(prot.residues)
    execute_generated(gpu_engine, g, validate=not aligned)
    got = dc.read(b.out, b.n_out, np.uint8)
    ref = np.zeros(want.n_residues, np.uint8)
    assert cengine.batch_execute(want.task_begin, want.tasks, prot.residues, want.alt, want.alt_base, ref, want.out_base)[0] == 0
    assert np.array_equal(got, ref)


def test_no_sites_at_all(world, gpu_engine):
    prot, cat, dc = world
    g = dc.generate(np.zeros(0, np.int64), np.zeros(0, np.int64), 5, True)
    assert (g.batch.n_tasks, g.batch.n_out, g.n_rows) == (0, 0, 0)
    assert not dc.read(g.batch.task_begin, 6, np.uint64).any()


def _variable_names(prot, seed):
    """Names of 1..24 bytes (one empty) so that every header length differs."""
    rng = np.random.default_rng(seed)
    nlen = rng.integers(1, 25, prot.n_tx)
    nlen[7] = 0
    off = np.zeros(prot.n_tx + 1, np.uint64)
    np.cumsum(nlen, out=off[1:])
    pool = rng.choice(np.frombuffer(b"ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789.-", np.uint8), int(off[-1]))
    return off, pool


@pytest.mark.parametrize("names", ["default", "variable"])
@pytest.mark.parametrize("seed,n_hap", [(4, 2), (5, 10), (6, 64)])
def test_fasta_framing_matches_the_host_producer(world, gpu_engine, names, seed, n_hap):
    """V2P_GEN_FASTA: header/newline segments and the name tape come out exactly as cohort.fasta_image builds them
    (itself checked record for record against the reference binary, tests/test_cohort_taskgen.py), and the executed
    tape is the .fasta text of every sample."""
    prot, cat, dc = world
    nm = C.default_names(prot) if names == "default" else _variable_names(prot, seed)
    dc.set_names(*nm)
    hap, site = C.select_sites(cat, n_hap, np.random.default_rng(seed))
    if n_hap == 10:  # haplotypes without any site: first, middle, last
        keep = ~np.isin(hap, (0, 4, 9))
        hap, site = hap[keep], site[keep]
    want = C.fasta_image(prot, C.build_batch(prot, cat, hap, site, n_hap, "global", "packed"), nm)
    g = dc.generate(hap, site, n_hap, aligned=False, fasta=True)
    b = g.batch
    assert (b.n_hap, b.n_tasks, b.n_alt, b.n_out) == (n_hap, len(want.tasks), len(want.alt), want.n_residues)
    assert np.array_equal(dc.read(b.task_begin, n_hap + 1, np.uint64), want.task_begin)
    assert np.array_equal(dc.read(b.out_base, n_hap + 1, np.uint64), want.out_base)
    assert np.array_equal(dc.read(b.alt_base, n_hap + 1, np.uint64), want.alt_base)
    assert np.array_equal(dc.read(b.tasks, 4 * b.n_tasks, np.uint32).reshape(-1, 4), want.tasks)
    assert np.array_equal(dc.read(b.alt, b.n_alt, np.uint8), want.alt)
    assert np.array_equal(dc.read(g.ann_start, g.n_rows, np.uint64), want.ann_start)
    assert np.array_equal(dc.read(g.ann_end, g.n_rows, np.uint64), want.ann_end)
    gpu_engine.set_reference(prot.residues)
    execute_generated(gpu_engine, g, validate=True)  # a file image has no gaps: gir.rs:208 holds
    got = dc.read(b.out, b.n_out, np.uint8)
    ref = np.zeros(want.n_residues, np.uint8)
    assert cengine.batch_execute(want.task_begin, want.tasks, prot.residues, want.alt, want.alt_base, ref, want.out_base)[0] == 0
    assert np.array_equal(got, ref)
    if names == "default":  # the text parses back into the records the consumer contract gives
        plain = C.build_batch(prot, cat, hap, site, n_hap, "global", "packed")
        tape = np.zeros(plain.n_residues, np.uint8)
        assert cengine.batch_execute(plain.task_begin, plain.tasks, prot.residues, plain.alt, plain.alt_base, tape, plain.out_base)[0] == 0
        for h in (0, n_hap - 1):
            image = got[int(want.out_base[h]):int(want.out_base[h + 1])]
            assert C.parse_fasta_image(image) == sorted(C.fasta_records(prot, plain, tape, h, 1 + (h & 1)))


def test_fasta_flag_errors(world):
    prot, cat, dc = world
    from vcf2prot_b200.engine import EngineError
    dc.set_names(*C.default_names(prot))
    hap, site = C.select_sites(cat, 2, np.random.default_rng(1))
    with pytest.raises(EngineError):  # a file image cannot contain pad bytes
        dc.generate(hap, site, 2, aligned=True, fasta=True)


def test_site_lists_are_validated_on_the_device(world):
    """a list entry that is not a catalogue site, or lists that are not strictly ascending inside a haplotype,
    are V2P_ERR_INVALID_ARG naming the entry -- never an out-of-range read of the catalogue tables."""
    from vcf2prot_b200 import EngineError
    from vcf2prot_b200 import _lib as L

    prot, cat, dc = world
    hap, site = C.select_sites(cat, 8, np.random.default_rng(77))
    dc.generate(hap, site, 8, False)  # fine as produced
    for bad_value, where in ((cat.n, 5), (0xFFFFFFF0, len(site) - 1)):
        s = site.copy()
        s[where] = bad_value
        with pytest.raises(EngineError) as ei:
            dc.generate(hap, s, 8, False)
        assert ei.value.status == L.ERR_INVALID_ARG and "entry %d " % where in str(ei.value)
    j = int(np.flatnonzero(hap[1:] == hap[:-1])[3]) + 1  # entries j-1, j belong to one haplotype
    swapped, dup = site.copy(), site.copy()
    swapped[j - 1], swapped[j] = site[j], site[j - 1]
    dup[j] = site[j - 1]
    for s in (swapped, dup):
        with pytest.raises(EngineError) as ei:
            dc.generate(hap, s, 8, False)
        assert ei.value.status == L.ERR_INVALID_ARG and "entry %d " % j in str(ei.value)
    dc.generate(hap, site, 8, False)  # the object is still usable


def test_catalogue_arguments_are_checked():
    from vcf2prot_b200 import EngineError

    prot = C.make_proteome(seed=3, n_tx=50, mu=5.0, sigma=0.5, hi=900)
    cat = C.make_catalogue(prot, 400, seed=4)
    bad = C.Catalogue(cat.t.copy(), cat.p, cat.cls, cat.rlen, cat.doff, cat.dlen, cat.af, cat.pool)
    bad.t[7] = prot.n_tx  # not a transcript
    with pytest.raises(EngineError):
        DeviceCatalogue(prot, bad, 0)
    bad = C.Catalogue(cat.t, cat.p, cat.cls, cat.rlen, cat.doff.copy(), cat.dlen, cat.af, cat.pool)
    bad.doff[-1] = len(cat.pool)  # payload beyond the pool
    with pytest.raises(EngineError):
        DeviceCatalogue(prot, bad, 0)
