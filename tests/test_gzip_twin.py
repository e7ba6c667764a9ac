"""CPU suite: the gzip checker (oracle/gzip_twin.py).  zlib is the independent inflater; the twin restates the device
encoder's format decisions so they can be verified without a GPU."""
import os
import random
import zlib

from oracle import gzip_twin as G


def _fasta(n, seed):
    rnd = random.Random(seed)
    out = []
    while sum(map(len, out)) < n:
        out.append(">ENST%011d_%d\n" % (rnd.randrange(10**6), 1 + rnd.randrange(2)) +
                   "".join(rnd.choice("ACDEFGHIKLMNPQRSTVWY") for _ in range(rnd.randrange(30, 900))) + "\n")
    return "".join(out).encode()[:n]


def _skewed():  # Fibonacci frequencies: unrestricted Huffman depth 19 > 15, exercises the Kraft repair
    fib = [1, 1]
    while len(fib) < 20:
        fib.append(fib[-1] + fib[-2])
    return b"".join(bytes([65 + i]) * f for i, f in enumerate(fib))[:16384]


CASES = [b"", b"A", b"AAAA" * 100, bytes(range(256)) * 3, os.urandom(5000), _fasta(50000, 1), _fasta(16384, 2),
         _fasta(16385, 3), b"ab" * 9000, _skewed()]


def test_twin_members_inflate_to_the_input():
    for data in CASES:
        G.check_member(G.encode_file(data), data)


def test_entropy_coding_alone_is_not_worse_than_deflate_best_on_protein_fasta():
    data = _fasta(200000, 7)
    assert len(G.encode_file(data)) <= len(zlib.compress(data, 9)) + 18


def test_lengths_are_a_complete_prefix_code_within_15_bits():
    for data in CASES[1:]:
        for i in range(0, len(data), 16384):
            freq = [0] * 257
            for b in data[i:i + 16384]:
                freq[b] += 1
            freq[256] = 1
            lens = G.code_lengths(freq)
            assert max(lens) <= 15 and all((l > 0) == (f > 0) for l, f in zip(lens, freq))
            assert sum(2 ** (15 - l) for l in lens if l) == 2 ** 15  # Kraft equality: zlib rejects incomplete codes


def test_crc_concatenation_algebra():
    a, b = os.urandom(1000), os.urandom(777)
    assert G.crc_concat(zlib.crc32(a), zlib.crc32(b), len(b)) == zlib.crc32(a + b)
    assert G.crc_concat(zlib.crc32(a), 0, 0) == zlib.crc32(a)
    assert G.crc_concat(0, zlib.crc32(b), len(b)) == zlib.crc32(b)
    parts = [os.urandom(n) for n in (0, 1, 64, 16384, 5)]
    crc = 0
    for p in parts:
        crc = G.crc_concat(crc, zlib.crc32(p), len(p))
    assert crc == zlib.crc32(b"".join(parts))
