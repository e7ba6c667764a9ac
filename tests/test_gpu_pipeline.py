"""GPU suite: the whole-cohort pipeline (include/v2p_pipeline.h) -- site lists or bit-masks in, per-sample .fasta /
.fasta.gz file images out -- against the oracle: every file equals the text the reference's writer produces from the
oracle-executed tapes (personalized_genome.rs:97,107), chunk boundaries, lanes and destinations notwithstanding."""
import zlib

import numpy as np
import pytest

from oracle import cengine
from synth import cohort as C
from vcf2prot_b200.engine import EngineError
from vcf2prot_b200.pipeline import DevicePipeline, csr_lists

pytestmark = pytest.mark.gpu

RICH_MIX = (0.55, 0.10, 0.10, 0.08, 0.07, 0.05, 0.05)


@pytest.fixture(scope="module")
def world():
    prot = C.make_proteome(seed=61, n_tx=300, mu=5.3, sigma=0.7, lo=30, hi=3000)
    cat = C.make_catalogue(prot, 6000, seed=62, mix=RICH_MIX, fs_mean=30, fs_max=600, sl_max=120)
    cat.af[:] = np.random.default_rng(7).choice([0.01, 0.05, 0.2, 0.5], size=cat.n)
    return prot, cat


def oracle_files(prot, cat, hap, site, n_samples):
    """Per-sample file text from the ORACLE's tapes: records of hap 1 then hap 2, in tape order."""
    b = C.build_batch(prot, cat, hap, site, 2 * n_samples, "global", "packed")
    tape = np.zeros(b.n_residues, np.uint8)
    assert cengine.batch_execute(b.task_begin, b.tasks, prot.residues, b.alt, b.alt_base, tape, b.out_base)[0] == 0
    files = []
    for s in range(n_samples):
        txt = []
        for k in (0, 1):
            for name, seq in C.fasta_records(prot, b, tape, 2 * s + k, k + 1):
                txt.append(">%s\n%s\n" % (name, seq))
        files.append("".join(txt).encode("ascii"))
    return files, int((b.ann_end - b.ann_start).sum())


def cohort_sites(cat, n_samples, seed, drop=()):
    hap, site = C.select_sites(cat, 2 * n_samples, np.random.default_rng(seed))
    if drop:
        keep = ~np.isin(hap, drop)
        hap, site = hap[keep], site[keep]
    return hap, site


@pytest.mark.parametrize("lanes,chunk", [(1, 4), (2, 3), (3, 1), (2, 64)])
@pytest.mark.parametrize("gzip", [False, True])
def test_files_equal_the_oracle_text(world, gpu_engine, lanes, chunk, gzip):
    prot, cat = world
    n_samples = 11
    hap, site = cohort_sites(cat, n_samples, 100 + lanes, drop=(0, 1, 9, 21))  # sample 0 has no variant at all
    want, n_res = oracle_files(prot, cat, hap, site, n_samples)
    assert want[0] == b""
    gpu_engine.set_reference(prot.residues)
    pipe = DevicePipeline(gpu_engine, prot, cat, C.default_names(prot), lanes=lanes)
    sb, sites = csr_lists(hap, site, 2 * n_samples)
    out = np.zeros(sum(len(w) for w in want) + 64 * n_samples + 1024, np.uint8)
    fb, res = pipe.run_lists(sb, sites, n_samples, chunk, gzip, out=out)
    got = [out[int(fb[s]):int(fb[s + 1])].tobytes() for s in range(n_samples)]
    if gzip:
        for s in range(n_samples):
            d = zlib.decompressobj(wbits=31)
            assert d.decompress(got[s]) == want[s] and d.eof and d.unused_data == b"", s
    else:
        assert got == want
    assert res.n_samples == n_samples and res.n_chunks == -(-n_samples // chunk)
    assert res.image_bytes == sum(len(w) for w in want) and res.out_bytes == int(fb[-1])
    assert res.n_records == sum(w.count(b">") for w in want) and res.n_sites == len(site)
    # the same through the sink (pinned ring inside the pipeline): chunks arrive in sample order
    seen = []

    def sink(first, n, data, begins):
        assert first == sum(k for _, k in seen)
        seen.append((bytes(data), n))
        assert int(begins[0]) == 0 and int(begins[n]) == len(data)
        return 0

    _, res2 = pipe.run_lists(sb, sites, n_samples, chunk, gzip, sink=sink)
    assert b"".join(d for d, _ in seen) == b"".join(got) and sum(k for _, k in seen) == n_samples
    assert (res2.out_bytes, res2.n_tasks) == (res.out_bytes, res.n_tasks)
    pipe.close()


@pytest.mark.parametrize("gzip", [False, True])
def test_masks_to_files(world, gpu_engine, gzip):
    """FORMAT/BCSQ matrix -> files, nothing visiting the host in between except the per-haplotype list offsets."""
    prot, cat = world
    n_samples = 37
    rec = C.make_records(cat, 5, 3, 0.2, 4)
    hap, site = cohort_sites(cat, n_samples, 7)
    masks = C.encode_masks(rec, n_samples, hap, site)
    want, _ = oracle_files(prot, cat, hap, site, n_samples)
    gpu_engine.set_reference(prot.residues)
    pipe = DevicePipeline(gpu_engine, prot, cat, C.default_names(prot), lanes=2)
    out = np.zeros(sum(len(w) for w in want) + 64 * n_samples + 1024, np.uint8)
    fb, res = pipe.run_masks(masks, rec.csq_begin, rec.csq_site, chunk_samples=8, gzip=gzip, out=out)
    for s in range(n_samples):
        got = out[int(fb[s]):int(fb[s + 1])].tobytes()
        assert (zlib.decompress(got, wbits=31) if gzip else got) == want[s], s
    assert res.n_chunks == 5 and res.decode_ms > 0
    pipe.close()


def test_pipeline_errors(world, gpu_engine):
    prot, cat = world
    gpu_engine.set_reference(prot.residues)
    pipe = DevicePipeline(gpu_engine, prot, cat, C.default_names(prot), lanes=2)
    hap, site = cohort_sites(cat, 4, 3)
    sb, sites = csr_lists(hap, site, 8)
    with pytest.raises(EngineError) as ei:  # destination too small: V2P_ERR_RES_OOB, like v2p_gzip_files
        pipe.run_lists(sb, sites, 4, 2, False, out=np.zeros(100, np.uint8))
    assert ei.value.status == 5
    with pytest.raises(EngineError):  # neither out nor sink
        pipe.run_lists(sb, sites, 4, 2, False)
    with pytest.raises(EngineError):  # a sink that refuses
        pipe.run_lists(sb, sites, 4, 2, False, sink=lambda *a: 1)
    bad = sb.copy()
    bad[3] = bad[2] - 1 if bad[2] else 5
    with pytest.raises(EngineError):
        pipe.run_lists(bad, sites, 4, 2, False, out=np.zeros(1 << 20, np.uint8))
    # and it still works afterwards
    fb, _ = pipe.run_lists(sb, sites, 4, 2, False, out=np.zeros(1 << 22, np.uint8))
    assert fb[-1] > 0
    pipe.close()


@pytest.mark.parametrize("gzip", [False, True])
def test_files_on_disk_through_the_native_writer(world, gpu_engine, tmp_path, gzip):
    """lists -> `{out_dir}/{proband}.fasta[.gz]` (parts/io.rs:35-57), no Python between the GPU and the file system."""
    from vcf2prot_b200.pipeline import DirWriter

    prot, cat = world
    n_samples = 23
    hap, site = cohort_sites(cat, n_samples, 11, drop=(6, 7))
    want, _ = oracle_files(prot, cat, hap, site, n_samples)
    gpu_engine.set_reference(prot.residues)
    pipe = DevicePipeline(gpu_engine, prot, cat, C.default_names(prot), lanes=2)
    names = ["NA%05d" % (7 * i) for i in range(n_samples)]
    w = DirWriter(str(tmp_path), names, compressed=gzip, threads=4)
    sb, sites = csr_lists(hap, site, 2 * n_samples)
    _, res = pipe.run_lists(sb, sites, n_samples, 5, gzip, sink=w)
    assert w.files_written == n_samples and w.bytes_written == res.out_bytes
    for s, name in enumerate(names):
        raw = (tmp_path / (name + (".fasta.gz" if gzip else ".fasta"))).read_bytes()
        assert (zlib.decompress(raw, wbits=31) if gzip else raw) == want[s], name
    assert want[3] == b"" and (tmp_path / ("NA00021" + (".fasta.gz" if gzip else ".fasta"))).exists()
    w.close()
    pipe.close()


def oracle_files_all(prot, cat, hap, site, n_samples):
    """`-a` file text from the ORACLE's tapes: per haplotype the altered records (tape order), then every other
    transcript of the proteome unchanged, proteome order (the record set of personalized_genome.rs:120-210, pinned on
    the reference binary by tests/test_rules_vs_reference_binary.py::test_write_all_record_set_of_the_reference_binary)."""
    b = C.build_batch(prot, cat, hap, site, 2 * n_samples, "global", "packed")
    tape = np.zeros(b.n_residues, np.uint8)
    assert cengine.batch_execute(b.task_begin, b.tasks, prot.residues, b.alt, b.alt_base, tape, b.out_base)[0] == 0
    files = []
    for s in range(n_samples):
        txt = []
        for k in (0, 1):
            recs = C.fasta_records(prot, b, tape, 2 * s + k, k + 1)
            altered = {n[:-2] for n, _ in recs}
            txt += [">%s\n%s\n" % r for r in recs]
            txt += [">%s_%d\n%s\n" % (prot.name(t), k + 1, prot.seq(t)) for t in range(prot.n_tx) if prot.name(t) not in altered]
        files.append("".join(txt).encode("ascii"))
    return files


@pytest.mark.parametrize("lanes,chunk,gzip", [(1, 3, False), (2, 2, False), (2, 5, True)])
def test_all_records_flag_writes_the_unaltered_reference_too(world, lanes, chunk, gzip):
    from vcf2prot_b200 import GpuEngine

    prot, cat = world
    n_samples = 7
    hap, site = cohort_sites(cat, n_samples, 500 + lanes, drop=(2, 3))  # sample 1 carries nothing: its file is the proteome twice
    want = oracle_files_all(prot, cat, hap, site, n_samples)
    assert want[1].count(b">") == 2 * prot.n_tx and all(w.count(b">") == 2 * prot.n_tx for w in want)
    sb, sites = csr_lists(hap, site, 2 * n_samples)
    with GpuEngine(0) as eng:  # (its registered reference is replaced by the extended tape)
        eng.set_reference(prot.residues)
        pipe = DevicePipeline(eng, prot, cat, C.default_names(prot), lanes=lanes)
        with pytest.raises(EngineError):  # not prepared yet
            pipe.run_lists(sb, sites, n_samples, chunk, gzip, sink=lambda *a: 0, all_records=True)
        pipe.enable_all_records(prot.residues, prot.offsets, C.default_names(prot))
        got = {}

        def sink(first, n, data, begins):
            for i in range(n):
                got[first + i] = bytes(data[int(begins[i]):int(begins[i + 1])])
            return 0

        _, res = pipe.run_lists(sb, sites, n_samples, chunk, gzip, sink=sink, all_records=True)
        un = (lambda b: zlib.decompress(b, wbits=31)) if gzip else (lambda b: b)
        for s in range(n_samples):
            assert un(got[s]) == want[s], s
        assert int(res.n_records) == 2 * n_samples * prot.n_tx
        # the same pipeline still writes the altered-only files (tasks index the proteome part of the extended tape)
        want_altered, _ = oracle_files(prot, cat, hap, site, n_samples)
        got.clear()
        pipe.run_lists(sb, sites, n_samples, chunk, False, sink=sink)
        assert [got[s] for s in range(n_samples)] == want_altered
        pipe.close()
