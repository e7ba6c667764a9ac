"""CPU suite: the N>1 path (sample sharding + max/sum-over-ranks accounting) under torch.distributed/gloo, world 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from synth import cohort as C
from vcf2prot_b200 import shard


def test_sample_ranges_partition_exactly():
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 8, 2504, 50000):
            r = [shard.sample_range(k, world, n) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def test_balanced_ranges_by_output_bytes():
    w = [10] * 50 + [1000] * 5 + [10] * 45
    rs = shard.balanced_ranges(w, 4)
    assert rs[0][0] == 0 and rs[-1][1] == len(w) and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
    sums = [sum(w[lo:hi]) for lo, hi in rs]
    assert max(sums) <= 2.2 * (sum(w) / 4)


def _worker(rank, world, port, n_samples, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard.sample_range(rank, world, n_samples)
        prot = C.make_proteome(seed=3, n_tx=60, mu=5.0, sigma=0.5, hi=1500)
        cat = C.make_catalogue(prot, 1500, seed=4)
        cat.af[:] = 0.1
        # every rank generates ONLY its own samples; seeds are per sample so the union is independent of world size
        parts = [C.synth_batch(prot, cat, 2, seed=1000 + s) for s in range(lo, hi)]
        residues = sum(b.n_residues for b in parts)
        digest = sum(int(b.tasks[:, 1].astype(np.int64).sum()) * (s + 1) for s, b in zip(range(lo, hi), parts))
        total = shard.sum_over_ranks(residues)
        slowest = shard.max_over_ranks(float(rank + 1) * 1.5)
        q.put((rank, lo, hi, residues, digest, total, slowest))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_world2_sharding_matches_single_rank():
    n_samples = 9
    ctx = mp.get_context("spawn")
    results = {}
    for world in (1, 2):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, n_samples, q)) for r in range(world)]
        for p in procs:
            p.start()
        got = [q.get(timeout=120) for _ in range(world)]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        results[world] = sorted(got)
    one, two = results[1], results[2]
    assert [(r[1], r[2]) for r in two] == [(0, 5), (5, 9)]
    # no sample lost or duplicated: residues and the per-sample digest add up to the single-rank run
    assert sum(r[3] for r in two) == one[0][3] and sum(r[4] for r in two) == one[0][4]
    # the all-reduces every rank sees: total units and the slowest rank's time
    assert all(r[5] == one[0][3] for r in two) and all(r[6] == 3.0 for r in two)


def _c3_worker(rank, world, port, n_samples, q):
    """The host logic of bench.py's c3 section on CPU: ONE counter-based cohort (synth/devgen), every rank plans the same
    contiguous ranges from per-sample result-tape bytes, builds and 'executes' (oracle) only its own range, and the
    parity counters are summed over ranks."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import cengine
        from synth import devgen

        prot = C.make_proteome(seed=3, n_tx=80, mu=5.0, sigma=0.6, hi=2500)
        cat = C.make_catalogue(prot, 2500, seed=4, mix=(0.7, 0.06, 0.06, 0.08, 0.04, 0.03, 0.03))
        cat.af[:] = np.random.default_rng(1).choice([0.02, 0.1, 0.3], size=cat.n).astype(np.float32)
        seed = 0x5EED0003
        # planning pass (every rank, whole cohort): result-tape bytes per sample -> ranges balanced by bytes
        hap, site = devgen.site_lists_numpy(cat, seed, 0, 2 * n_samples)
        whole = C.build_batch(prot, cat, hap, site, 2 * n_samples)
        ob = whole.out_base.astype(np.int64)
        weights = (ob[2::2] - ob[:-2:2]).tolist()
        ranges = shard.balanced_ranges(weights, world)
        lo, hi = ranges[rank]
        # this rank's range, generated on its own from the haplotype indices alone
        h2, s2 = devgen.site_lists_numpy(cat, seed, 2 * lo, 2 * (hi - lo))
        mine = C.build_batch(prot, cat, h2, s2, 2 * (hi - lo))
        out = np.zeros(mine.n_residues, np.uint8)
        assert cengine.batch_execute(mine.task_begin, mine.tasks, prot.residues, mine.alt, mine.alt_base, out, mine.out_base)[0] == 0
        bad, _ = cengine.batch_check(mine.task_begin, mine.tasks, prot.residues, mine.alt, mine.alt_base, out, mine.out_base, threads=2)
        digest = int(out.astype(np.int64).sum()) + 31 * int(mine.n_residues)
        q.put((rank, lo, hi, int(mine.n_residues), digest, shard.sum_over_ranks(int(mine.n_residues)),
               shard.sum_over_ranks(bad), shard.sum_over_ranks(2 * (hi - lo)), int(whole.n_residues), max(weights)))
    finally:
        dist.destroy_process_group()


def test_world2_one_cohort_sharded_by_bytes_like_the_c3_section():
    n_samples = 21
    ctx = mp.get_context("spawn")
    results = {}
    for world in (1, 2):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_c3_worker, args=(r, world, port, n_samples, q)) for r in range(world)]
        for p in procs:
            p.start()
        got = [q.get(timeout=180) for _ in range(world)]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        results[world] = sorted(got)
    one, two = results[1][0], results[2]
    assert (two[0][1], two[1][2]) == (0, n_samples) and two[0][2] == two[1][1]  # contiguous, covering, no overlap
    # the union of the ranks' ranges is the single-rank cohort: residues and content digests add up
    assert sum(r[3] for r in two) == one[3] == one[8] and sum(r[4] for r in two) == one[4]
    # what every rank sees after the all-reduces: all residues, no mismatching haplotype, every haplotype checked
    assert all(r[5] == one[3] and r[6] == 0 and r[7] == 2 * n_samples for r in two)
    # balanced by bytes: no rank is more than one sample's worth above its share
    assert max(r[3] for r in two) <= one[3] / 2 + one[9]
