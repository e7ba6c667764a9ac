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
