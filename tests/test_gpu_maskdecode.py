"""GPU suite: genotype bit-masks -> per-haplotype site lists on the device (include/v2p_taskgen.h)
against the restatement of MaskDecoder.rs / vcf_ds.rs in oracle/maskdecode.py, and chained into the Task generator and
the engine without the lists, the tasks or the result ever visiting the host in between."""
import numpy as np
import pytest

from oracle import cengine, maskdecode
from synth import cohort as C
from vcf2prot_b200.engine import EngineError
from vcf2prot_b200.taskgen import DeviceCatalogue, execute_generated

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world():
    prot = C.make_proteome(seed=41, n_tx=300, mu=5.3, sigma=0.7, lo=30, hi=3000)
    cat = C.make_catalogue(prot, 6000, seed=42, mix=(0.55, 0.10, 0.10, 0.08, 0.07, 0.05, 0.05), fs_mean=30, fs_max=600,
                           sl_max=120)
    cat.af[:] = np.random.default_rng(6).choice([0.01, 0.05, 0.2, 0.5], size=cat.n)
    dc = DeviceCatalogue(prot, cat, 0)
    yield prot, cat, dc
    dc.close()


def _lists(dc, lst):
    return dc.read(lst.site_begin, lst.n_hap + 1, np.uint64), dc.read(lst.sites, lst.n_sites, np.uint32)


def test_reference_unit_vectors_on_device(world):
    """MaskDecoder.rs:160-400: "1" -> h1[0]; "3" -> h1[0] h2[0]; "1024" -> h1[5]; "3,3,3,3" -> [0,15,30,45] twice."""
    _, _, dc = world
    ident = np.arange(64, dtype=np.int32)  # csq k of the only record -> catalogue site k
    cb = np.array([0, 64], np.uint64)
    for words, h1, h2 in (([1], [0], []), ([3], [0], [0]), ([1024], [5], []), ([1, 1], [0, 15], []),
                          ([3, 3], [0, 15], [0, 15]), ([3, 3, 3, 3], [0, 15, 30, 45], [0, 15, 30, 45]), ([0], [], [])):
        lst = dc.sites_from_masks(np.array(words, np.uint32).reshape(1, 1, -1), cb, ident)
        sb, sites = _lists(dc, lst)
        assert sb.tolist() == [0, len(h1), len(h1) + len(h2)], words
        assert sites.tolist() == h1 + h2, words
        assert (maskdecode.get_indices(words)) == (h1, h2)


@pytest.mark.parametrize("seed,n_samp,per_rec,p_unsup,wide", [(1, 1, 1, 0.0, 0), (2, 7, 3, 0.2, 0), (3, 33, 2, 0.1, 5),
                                                             (4, 130, 1, 0.0, 0), (5, 5, 6, 0.3, 3)])
def test_lists_match_the_restatement(world, seed, n_samp, per_rec, p_unsup, wide):
    prot, cat, dc = world
    rec = C.make_records(cat, seed, per_rec, p_unsup, wide)
    hap, site = C.select_sites(cat, 2 * n_samp, np.random.default_rng(seed))
    masks = C.encode_masks(rec, n_samp, hap, site)
    want_sb, want_sites = maskdecode.site_lists(masks, rec.csq_begin, rec.csq_site)
    lst = dc.sites_from_masks(masks, rec.csq_begin, rec.csq_site)
    sb, sites = _lists(dc, lst)
    assert np.array_equal(sb, want_sb) and np.array_equal(sites, want_sites)
    # select_sites is already sorted by (haplotype, site): the decode is the inverse of the encode
    assert np.array_equal(sites, site.astype(np.uint32))


def test_duplicate_csq_entries_collapse(world):
    """Two csq entries of different records naming the same catalogue site (the same consequence reported twice) are
    one site for the haplotype: identical duplicates are dropped, vcf_ds.rs:442-479."""
    _, _, dc = world
    masks = np.zeros((3, 2, 1), np.uint32)
    masks[0, 0, 0], masks[0, 1, 0], masks[1, 0, 0], masks[2, 0, 0] = 0b0111, 0b1000, 0b01, 0b01
    cb, cs = np.array([0, 2, 3, 4], np.uint64), np.array([7, 3, -1, 3], np.int32)
    lst = dc.sites_from_masks(masks, cb, cs)
    sb, sites = _lists(dc, lst)
    want = maskdecode.site_lists(masks, cb, cs)
    assert sb.tolist() == want[0].tolist() == [0, 2, 3, 3, 4]
    assert sites.tolist() == want[1].tolist() == [3, 7, 7, 3]


def test_bit_beyond_the_records_csq_count_is_an_error(world):
    """vcf_ds.rs:287 indexes the record's csq vector with the decoded index: out of range panics there."""
    _, _, dc = world
    masks = np.zeros((4, 3, 1), np.uint32)
    masks[2, 1, 0] = 1 << 4  # csq 2 of a record with two
    cb, cs = np.array([0, 2, 4, 6, 8], np.uint64), np.arange(8, dtype=np.int32)
    with pytest.raises(IndexError):
        maskdecode.site_lists(masks, cb, cs)
    with pytest.raises(EngineError) as e:
        dc.sites_from_masks(masks, cb, cs)
    assert "record 2" in str(e.value)


def test_empty_matrix_and_all_zero_cells(world):
    _, _, dc = world
    cb, cs = np.array([0, 1, 2], np.uint64), np.array([0, 1], np.int32)
    lst = dc.sites_from_masks(np.zeros((2, 4, 1), np.uint32), cb, cs)
    assert (lst.n_hap, lst.n_sites) == (8, 0) and not dc.read(lst.site_begin, 9, np.uint64).any()
    lst = dc.sites_from_masks(np.zeros((0, 4, 1), np.uint32), np.zeros(1, np.uint64), np.zeros(0, np.int32))
    assert (lst.n_hap, lst.n_sites) == (8, 0)


@pytest.mark.parametrize("aligned", [False, True])
def test_masks_to_sequences_without_leaving_the_device(world, gpu_engine, aligned):
    prot, cat, dc = world
    n_samp = 24
    rec = C.make_records(cat, 9, 2, 0.15, 7)
    hap, site = C.select_sites(cat, 2 * n_samp, np.random.default_rng(9))
    masks = C.encode_masks(rec, n_samp, hap, site)
    lst = dc.sites_from_masks(masks, rec.csq_begin, rec.csq_site)
    g = dc.generate_from_lists(lst, aligned)
    want = C.build_batch(prot, cat, hap, site, 2 * n_samp, "global", "aligned" if aligned else "packed")
    b = g.batch
    assert (b.n_hap, b.n_tasks, b.n_out) == (2 * n_samp, len(want.tasks), want.n_residues)
    assert np.array_equal(dc.read(b.tasks, 4 * b.n_tasks, np.uint32).reshape(-1, 4), want.tasks)
    gpu_engine.set_reference(prot.residues)
    execute_generated(gpu_engine, g, validate=not aligned)
    ref = np.zeros(want.n_residues, np.uint8)
    assert cengine.batch_execute(want.task_begin, want.tasks, prot.residues, want.alt, want.alt_base, ref, want.out_base)[0] == 0
    assert np.array_equal(dc.read(b.out, b.n_out, np.uint8), ref)
