"""GPU suite: gzip file images written on the device (include/v2p_gzip.h).  The judge is an
independent inflater (Python zlib): every file must be exactly one complete gzip member that inflates to the bytes the
reference would have compressed (personalized_genome.rs:87-101); byte equality with oracle/gzip_twin.py additionally
pins the encoder's format decisions."""
import os
import random
import zlib

import numpy as np
import pytest

from oracle import cengine, gzip_twin as G
from synth import cohort as C
from vcf2prot_b200.engine import EngineError
from vcf2prot_b200.gzipdev import DeviceGzip

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gz():
    z = DeviceGzip(0)
    yield z
    z.close()


def _fasta(n, seed):
    rnd = random.Random(seed)
    out = []
    while sum(map(len, out)) < n:
        out.append(">ENST%011d_%d\n" % (rnd.randrange(10**6), 1 + rnd.randrange(2)) +
                   "".join(rnd.choice("ACDEFGHIKLMNPQRSTVWY") for _ in range(rnd.randrange(30, 900))) + "\n")
    return "".join(out).encode()[:n]


def _skewed():
    fib = [1, 1]
    while len(fib) < 20:
        fib.append(fib[-1] + fib[-2])
    return b"".join(bytes([65 + i]) * f for i, f in enumerate(fib))[:16384]


def _run(gz, files):
    data = np.frombuffer(b"".join(files), np.uint8) if files else np.zeros(0, np.uint8)
    fb = np.zeros(len(files) + 1, np.uint64)
    np.cumsum([len(f) for f in files], out=fb[1:])
    return gz.compress(data, fb)


def test_every_file_is_one_member_and_equals_the_twin(gz):
    files = [b"", b"A", b"AAAA" * 100, bytes(range(256)) * 3, os.urandom(5000), _fasta(50000, 1), _fasta(16384, 2),
             _fasta(16385, 3), b"", b"ab" * 9000, _skewed(), _fasta(100001, 4), os.urandom(40000), b"\n"]
    out, res = _run(gz, files)
    assert res.in_bytes == sum(map(len, files)) and res.out_bytes == sum(map(len, out))
    assert res.n_stored_chunks >= 4  # the random files cannot be entropy coded
    for f, g in zip(files, out):
        G.check_member(g, f)
        assert g == G.encode_file(f)


def test_unaligned_file_boundaries_and_odd_sizes(gz):
    rnd = random.Random(5)
    files = [_fasta(rnd.randrange(1, 40000), 100 + i) for i in range(40)]
    out, _ = _run(gz, files)
    for f, g in zip(files, out):
        G.check_member(g, f)


def test_file_range_inside_a_larger_buffer(gz):
    blob = _fasta(70000, 9)
    out, res = gz.compress(np.frombuffer(blob, np.uint8), [1001, 1001 + 33333, 69999])
    G.check_member(out[0], blob[1001:1001 + 33333])
    G.check_member(out[1], blob[1001 + 33333:69999])
    assert res.in_bytes == 69999 - 1001


def test_capacity_too_small_is_reported(gz):
    blob = np.frombuffer(_fasta(50000, 10), np.uint8)
    with pytest.raises(EngineError) as e:
        gz.compress(blob, [0, 50000], capacity=1000)
    assert e.value.status == 5 and "capacity" in str(e.value)


def test_no_files(gz):
    out, res = gz.compress(np.zeros(0, np.uint8), [0])
    assert out == [] and res.out_bytes == 0


def test_sample_fasta_gz_from_the_device_image(gz, gpu_engine):
    """Engine -> FASTA image in HBM -> one .fasta.gz per sample (both haplotypes), nothing but compressed bytes leaving
    the device; inflated, each file is the reference's uncompressed file for that sample."""
    import torch

    prot = C.make_proteome(seed=51, n_tx=300, mu=5.3, sigma=0.7, lo=30, hi=3000)
    cat = C.make_catalogue(prot, 6000, seed=52)
    cat.af[:] = np.random.default_rng(7).choice([0.01, 0.05, 0.2], size=cat.n)
    n_samp = 6
    img = C.fasta_image(prot, C.synth_batch(prot, cat, 2 * n_samp, 53, "global", "packed"))
    dev = torch.device("cuda", 0)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)
    gpu_engine.set_reference(prot.residues)
    d_out = torch.empty(max(img.n_residues, 1), dtype=torch.uint8, device=dev)
    d = [up(img.task_begin), up(img.tasks), up(img.alt), up(img.alt_base), up(img.out_base)]
    gpu_engine.execute_batch_device(2 * n_samp, d[0], d[1], None, d[2], d[3], d_out, d[4], len(img.tasks), len(img.alt),
                                    img.n_residues)
    file_begin = img.out_base[::2]  # sample s = haplotypes 2s and 2s+1
    cap = gz.bound(img.n_residues, n_samp)
    d_gz = torch.zeros(cap, dtype=torch.uint8, device=dev)
    ob, res = gz.compress_device(d_out.data_ptr(), file_begin, d_gz.data_ptr(), cap)
    host = d_gz[: int(ob[-1])].cpu().numpy()
    want = np.zeros(img.n_residues, np.uint8)
    assert cengine.batch_execute(img.task_begin, img.tasks, prot.residues, img.alt, img.alt_base, want, img.out_base)[0] == 0
    for s in range(n_samp):
        member = host[int(ob[s]):int(ob[s + 1])].tobytes()
        G.check_member(member, want[int(file_begin[s]):int(file_begin[s + 1])].tobytes())
    assert res.out_bytes < 0.62 * res.in_bytes  # ~4.2 bits of entropy per residue
