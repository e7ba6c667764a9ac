"""CPU suite: the mask-decode restatement against the reference's own unit-test vectors (MaskDecoder.rs:160-400)."""
import numpy as np

from oracle import maskdecode as M


def test_reference_unit_vectors():
    assert M.get_indices([1]) == ([0], [])            # test_get_indicies2  "1$"
    assert M.get_indices([3]) == ([0], [0])           # test_get_indicies3  "3$"
    assert M.get_indices([1024]) == ([5], [])         # test_get_indicies4  "1024$"
    assert M.get_indices([1, 1]) == ([0, 15], [])     # test_get_indicies5  "1,1"
    assert M.get_indices([3, 3]) == ([0, 15], [0, 15])                      # test_get_indicies6
    assert M.get_indices([3, 3, 3, 3]) == ([0, 15, 30, 45], [0, 15, 30, 45])  # test_get_indicies7
    assert M.get_indices([0]) == ([], [])             # "0$" -> no consequence


def test_site_lists_transpose_sort_dedup():
    # 3 records x 2 samples; record 0 has 2 csq (sites 7, 3), record 1 one unsupported csq, record 2 repeats site 3
    masks = np.zeros((3, 2, 1), np.uint32)
    masks[0, 0, 0] = 0b0111   # sample 0: hap1 csq0+csq1, hap2 csq0
    masks[0, 1, 0] = 0b1000   # sample 1: hap2 csq1
    masks[1, 0, 0] = 0b01     # unsupported csq -> dropped
    masks[2, 0, 0] = 0b01     # sample 0 hap1: site 3 again (duplicate)
    csq_begin = np.array([0, 2, 3, 4], np.uint64)
    csq_site = np.array([7, 3, -1, 3], np.int32)
    sb, sites = M.site_lists(masks, csq_begin, csq_site)
    assert sb.tolist() == [0, 2, 3, 3, 4]
    assert sites.tolist() == [3, 7, 7, 3]


def test_encode_masks_is_the_inverse_of_the_decode():
    """cohort.encode_masks writes what bcftools csq would put into FORMAT/BCSQ; decoding it gives the carrier lists back,
    with unsupported entries in between and records wider than one 15-entry word."""
    from synth import cohort as C

    prot = C.make_proteome(seed=41, n_tx=120, mu=5.0, sigma=0.6, lo=30, hi=1500)
    cat = C.make_catalogue(prot, 1500, seed=42)
    cat.af[:] = 0.2
    for seed, per_rec, p_unsup, wide in ((1, 1, 0.0, 0), (2, 4, 0.25, 0), (3, 2, 0.1, 4)):
        rec = C.make_records(cat, seed, per_rec, p_unsup, wide)
        hap, site = C.select_sites(cat, 10, np.random.default_rng(seed))
        masks = C.encode_masks(rec, 5, hap, site)
        assert masks.shape == (rec.n_rec, 5, rec.words) and (wide == 0 or rec.words > 1)
        sb, sites = M.site_lists(masks, rec.csq_begin, rec.csq_site)
        assert np.array_equal(sites, site.astype(np.uint32))
        assert np.array_equal(sb[1:], np.cumsum(np.bincount(hap, minlength=10)))
