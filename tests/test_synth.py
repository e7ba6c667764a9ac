"""The synthetic-cohort generators (test / bench infrastructure): the counter-based cohort definition of synth/devgen
is the same on the device and in numpy, and chunked / sharded generation yields the same cohort as one call."""
import numpy as np
import pytest

from synth import cohort as C
from synth import devgen

MIX = (0.6, 0.08, 0.08, 0.1, 0.06, 0.04, 0.04)


def small_world():
    prot = C.make_proteome(seed=5, n_tx=300, mu=5.5, sigma=0.7, hi=5000)
    cat = C.make_catalogue(prot, 6000, seed=6, mix=MIX)
    cat.af[:] = np.random.default_rng(9).choice([0.01, 0.05, 0.2, 0.5], size=cat.n).astype(np.float32)
    return prot, cat


def test_numpy_twin_is_a_function_of_the_haplotype_index_and_obeys_the_truncation_rule():
    prot, cat = small_world()
    h, s = devgen.site_lists_numpy(cat, 0xC0FFEE, 0, 40)
    # any sub-range generated on its own is the same cohort (what lets every rank produce just its range)
    h2, s2 = devgen.site_lists_numpy(cat, 0xC0FFEE, 13, 9)
    sel = (h >= 13) & (h < 22)
    assert np.array_equal(h[sel] - 13, h2) and np.array_equal(s[sel], s2)
    # nothing follows a truncating variant on the same haplotype + transcript, so build_batch keeps every pair
    b = C.build_batch(prot, cat, h, s, 40)
    assert len(b.kept_site) == len(s) and np.array_equal(b.kept_site, s)
    trunc = np.isin(cat.cls[s], (C.CLS_F, C.CLS_G, C.CLS_L, C.CLS_0))
    same = (h[1:] == h[:-1]) & (cat.t[s][1:] == cat.t[s][:-1])
    assert not (trunc[:-1] & same).any()
    # carrier frequencies follow af
    hh, ss = devgen.site_lists_numpy(cat, 7, 0, 400)
    freq = np.bincount(ss, minlength=cat.n) / 400.0
    first = np.ones(cat.n, bool)
    first[1:] = cat.t[1:] != cat.t[:-1]  # (first site of a transcript: never dropped by the truncation rule)
    assert abs(freq[first].mean() - cat.af[first].mean()) < 0.02


@pytest.mark.gpu
def test_device_lists_equal_the_numpy_twin_and_generate_the_host_producers_tasks(gpu_engine):
    from vcf2prot_b200 import _lib as L
    from vcf2prot_b200.taskgen import DeviceCatalogue

    prot, cat = small_world()
    gen = devgen.DeviceCohort(cat, 0xC0FFEE, 0)
    dc = DeviceCatalogue(prot, cat, 0)
    try:
        for h0, n in ((0, 1), (3, 64), (1000, 257)):
            begin, sites, n_sites = gen.lists(h0, n)
            want_h, want_s = devgen.site_lists_numpy(cat, 0xC0FFEE, h0, n)
            want_begin = np.zeros(n + 1, np.int64)
            np.cumsum(np.bincount(want_h, minlength=n), out=want_begin[1:])
            assert n_sites == len(want_s)
            assert np.array_equal(begin.cpu().numpy(), want_begin)
            assert np.array_equal(sites.cpu().numpy().astype(np.int64), want_s)
            lists = L.SiteLists()
            lists.n_hap, lists.n_sites, lists.site_begin, lists.sites = n, n_sites, begin.data_ptr(), sites.data_ptr()
            g = dc.generate_from_lists(lists, aligned=False)
            want = C.build_batch(prot, cat, want_h, want_s, n)
            assert g.batch.n_tasks == len(want.tasks) and g.batch.n_out == want.n_residues
            assert np.array_equal(dc.read(g.batch.tasks, 4 * g.batch.n_tasks, np.uint32).reshape(-1, 4), want.tasks)
            assert np.array_equal(dc.read(g.batch.out_base, n + 1, np.uint64), want.out_base)
    finally:
        gen.close()
        dc.close()


@pytest.mark.gpu
def test_store_ceiling_probe_reports_sane_rates():
    import torch

    buf = torch.empty(1 << 30, dtype=torch.uint8, device="cuda:0")
    r = devgen.store_ceiling_gbs(0, buf.data_ptr(), buf.numel(), reps=3)
    assert set(r) == {"memset", "tma_bulk_store_8k"} and all(1000 < v < 12000 for v in r.values())
    assert int(buf[123456].item()) == 0x2E and int(buf[-1].item()) == 0x2E
