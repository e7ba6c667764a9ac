"""GPU suite: the single-process multi-GPU cohort runner (include/v2p_cohort.h) -- parts/exec.rs:34-40's proband loop
with one worker per device.  Uses two devices when the box has them, otherwise two (or three) workers share GPU 0;
either way every sample's file must be the oracle's text and every sample must be delivered exactly once."""
import os
import tempfile
import threading
import zlib

import numpy as np
import pytest

from synth import cohort as C
from tests.test_gpu_pipeline import RICH_MIX, cohort_sites, oracle_files
from vcf2prot_b200.cohort_run import CohortRunner
from vcf2prot_b200.pipeline import DirWriter, csr_lists

pytestmark = pytest.mark.gpu


def devices(n):
    import torch

    have = torch.cuda.device_count()
    return [i % have for i in range(n)]


@pytest.fixture(scope="module")
def world():
    prot = C.make_proteome(seed=71, n_tx=300, mu=5.3, sigma=0.7, lo=30, hi=3000)
    cat = C.make_catalogue(prot, 6000, seed=72, mix=RICH_MIX, fs_mean=30, fs_max=600, sl_max=120)
    cat.af[:] = np.random.default_rng(8).choice([0.01, 0.05, 0.2, 0.5], size=cat.n)
    return prot, cat


def runner(prot, cat, devs, lanes=2):
    return CohortRunner(devs, prot.residues, prot.offsets, C.default_names(prot), cat.t, cat.p, cat.cls, cat.rlen, cat.doff,
                        cat.dlen, cat.pool, lanes=lanes)


@pytest.mark.parametrize("n_workers,chunk,gzip", [(1, 4, False), (2, 3, False), (2, 5, True), (3, 2, False)])
def test_every_file_equals_the_oracle_text(world, n_workers, chunk, gzip):
    prot, cat = world
    n_samples = 23
    hap, site = cohort_sites(cat, n_samples, 300 + n_workers, drop=(0, 1, 17))
    want, _ = oracle_files(prot, cat, hap, site, n_samples)
    sb, sites = csr_lists(hap, site, 2 * n_samples)
    got, calls, tids = {}, [], set()

    def sink(first, n, data, begins):
        tids.add(threading.get_ident())
        calls.append((first, n))
        for i in range(n):
            assert first + i not in got, "sample delivered twice"
            got[first + i] = bytes(data[int(begins[i]):int(begins[i + 1])])
        return 0

    r = runner(prot, cat, devices(n_workers))
    try:
        res = r.run_lists(sb, sites, n_samples, chunk, gzip, sink=sink)
    finally:
        r.close()
    un = (lambda b: zlib.decompress(b, wbits=31)) if gzip else (lambda b: b)
    assert sorted(got) == list(range(n_samples))
    for s in range(n_samples):
        assert un(got[s]) == want[s], s
    # contiguous ranges, one per worker, covering the cohort; chunks of one range arrive in order
    fs = [int(res.first_sample[g]) for g in range(n_workers + 1)]
    assert fs[0] == 0 and fs[-1] == n_samples and fs == sorted(fs) and res.n_devices == n_workers
    for g in range(n_workers):
        mine = [c for c in calls if fs[g] <= c[0] < fs[g + 1]]
        assert [c[0] for c in mine] == sorted(c[0] for c in mine)
        assert sum(c[1] for c in mine) == fs[g + 1] - fs[g] == int(res.per_device[g].n_samples)
    assert int(res.total.n_samples) == n_samples and int(res.total.n_records) == sum(w.count(b">") for w in want)
    # ranges are balanced by variant sites: no worker carries more than its share plus one sample's worth
    per = [int(sb[2 * fs[g + 1]] - sb[2 * fs[g]]) for g in range(n_workers)]
    biggest_sample = int(np.max(sb[2::2] - sb[:-2:2]))
    assert max(per) <= sum(per) / n_workers + biggest_sample


def test_directory_writer_from_two_workers_and_error_reporting(world):
    from vcf2prot_b200 import EngineError
    from vcf2prot_b200 import _lib as L

    prot, cat = world
    n_samples = 16
    hap, site = cohort_sites(cat, n_samples, 400)
    want, _ = oracle_files(prot, cat, hap, site, n_samples)
    sb, sites = csr_lists(hap, site, 2 * n_samples)
    names = ["P%03d" % i for i in range(n_samples)]
    r = runner(prot, cat, devices(2))
    try:
        with tempfile.TemporaryDirectory() as d:
            w = DirWriter(d, names, compressed=False, threads=2)
            res = r.run_lists(sb, sites, n_samples, 3, False, sink=w, concurrent_sink=True)  # the writer is thread-safe
            assert w.files_written == n_samples and int(res.total.out_bytes) == w.bytes_written == sum(len(x) for x in want)
            for s in (0, 7, 15):
                assert open(os.path.join(d, names[s] + ".fasta"), "rb").read() == want[s]
            w.close()
        # a bad list entry on the second worker's range: its status and message come back, nothing hangs
        bad = sites.copy()
        bad[-1] = cat.n + 5
        with pytest.raises(EngineError) as ei:
            r.run_lists(sb, bad, n_samples, 3, False, sink=lambda *a: 0)
        assert ei.value.status == L.ERR_INVALID_ARG and "device" in str(ei.value)
        # and the runner is still usable afterwards
        ok = r.run_lists(sb, sites, n_samples, 8, False, sink=lambda *a: 0)
        assert int(ok.total.n_samples) == n_samples and r.launch_count() > 0
    finally:
        r.close()


@pytest.mark.parametrize("n_workers,gzip", [(2, False), (3, True)])
def test_all_records_flag_on_every_worker(world, n_workers, gzip):
    """`-a` (personalized_genome.rs:120-210) through the cohort runner: every file holds 2 x n_tx records, oracle text."""
    from tests.test_gpu_pipeline import oracle_files_all
    from vcf2prot_b200 import EngineError

    prot, cat = world
    n_samples = 9
    hap, site = cohort_sites(cat, n_samples, 640 + n_workers, drop=(4, 5))  # sample 2 carries nothing
    want = oracle_files_all(prot, cat, hap, site, n_samples)
    sb, sites = csr_lists(hap, site, 2 * n_samples)
    got = {}

    def sink(first, n, data, begins):
        for i in range(n):
            assert first + i not in got
            got[first + i] = bytes(data[int(begins[i]):int(begins[i + 1])])
        return 0

    r = runner(prot, cat, devices(n_workers))
    try:
        with pytest.raises(EngineError) as ei:  # not prepared yet: the worker's message comes back
            r.run_lists(sb, sites, n_samples, 2, gzip, sink=lambda *a: 0, all_records=True)
        assert "enable_all_records" in str(ei.value)
        r.enable_all_records()
        res = r.run_lists(sb, sites, n_samples, 2, gzip, sink=sink, all_records=True)
        un = (lambda b: zlib.decompress(b, wbits=31)) if gzip else (lambda b: b)
        assert sorted(got) == list(range(n_samples))
        for s in range(n_samples):
            assert un(got[s]) == want[s], s
        assert int(res.total.n_records) == 2 * n_samples * prot.n_tx
        # the altered-only files still come out of the same runner
        want_altered, _ = oracle_files(prot, cat, hap, site, n_samples)
        got.clear()
        r.run_lists(sb, sites, n_samples, 4, False, sink=sink)
        assert [got[s] for s in range(n_samples)] == want_altered
    finally:
        r.close()


@pytest.mark.parametrize("n_workers,gzip", [(2, False), (3, True)])
def test_masks_to_files_over_the_workers(world, n_workers, gzip):
    """FORMAT/BCSQ matrix in, files out (MaskDecoder.rs:95-153 + parts/exec.rs:34-40): decoded once, run by every worker."""
    from vcf2prot_b200 import EngineError

    prot, cat = world
    n_samples = 29
    rec = C.make_records(cat, 5, 3, 0.2, 4)
    hap, site = cohort_sites(cat, n_samples, 77 + n_workers)
    masks = C.encode_masks(rec, n_samples, hap, site)
    want, _ = oracle_files(prot, cat, hap, site, n_samples)
    got = {}

    def sink(first, n, data, begins):
        for i in range(n):
            assert first + i not in got
            got[first + i] = bytes(data[int(begins[i]):int(begins[i + 1])])
        return 0

    r = runner(prot, cat, devices(n_workers))
    try:
        res = r.run_masks(masks, rec.csq_begin, rec.csq_site, chunk_samples=4, gzip=gzip, sink=sink)
        un = (lambda b: zlib.decompress(b, wbits=31)) if gzip else (lambda b: b)
        assert sorted(got) == list(range(n_samples))
        for s in range(n_samples):
            assert un(got[s]) == want[s], s
        assert int(res.total.n_samples) == n_samples and res.total.decode_ms > 0 and int(res.total.n_sites) == len(site)
        assert int(res.total.h2d_bytes) >= masks.nbytes
        # a mask bit beyond the record's consequences: the decoder's error comes back through the runner
        bad = masks.copy()
        r0 = int(np.argmin(np.diff(rec.csq_begin)))
        bad[r0, 0, -1] |= np.uint32(1 << 29)
        with pytest.raises(EngineError) as ei:
            r.run_masks(bad, rec.csq_begin, rec.csq_site, chunk_samples=4, sink=lambda *a: 0)
        assert "mask decode" in str(ei.value)
    finally:
        r.close()
