"""compute-sanitizer target: a few small batches through every path (registered/plain, aligned/packed, serial)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import cengine
from tests.randtasks import random_batch
from vcf2prot_b200 import GpuEngine
from synth import cohort as C

eng = GpuEngine(0)
ok = True
for seed, n_hap, mean in ((1, 3, 3000), (2, 6, 20000)):
    b = random_batch(seed, n_hap, mean, n_ref=30011)
    want = np.zeros(int(b["out_base"][-1]), np.uint8)
    assert cengine.batch_execute(b["task_begin"], b["tasks"], b["ref"], b["alt"], b["alt_base"], want, b["out_base"])[0] == 0
    for mode in ("replicas", "plain", None):
        for variant in (-1, 0, 8):
            eng.set_tuning(variant, 0)
            if mode:
                eng.set_reference(b["ref"], mode)
            out, _ = eng.execute_batch(b["task_begin"], b["tasks"], None if mode else b["ref"], b["alt"], b["alt_base"], b["out_base"])
            ok &= bool(np.array_equal(out, want))
prot = C.make_proteome(seed=5, n_tx=80, mu=5.0, sigma=0.6, hi=2000)
cat = C.make_catalogue(prot, 2500, seed=6, mix=(0.6, 0.1, 0.1, 0.1, 0.04, 0.03, 0.03), fs_mean=40, fs_max=400)
cat.af[:] = 0.2
eng.set_tuning(-1, 0)
eng.set_reference(prot.residues)
for layout in ("aligned", "packed"):
    bb = C.synth_batch(prot, cat, 6, 7, layout=layout)
    want = np.zeros(bb.n_residues, np.uint8)
    assert cengine.batch_execute(bb.task_begin, bb.tasks, prot.residues, bb.alt, bb.alt_base, want, bb.out_base)[0] == 0
    for hint in (False, True):  # interleaved tile order, then tape order (V2P_FLAG_ALIGNED_LAYOUT)
        out, _ = eng.execute_batch(bb.task_begin, bb.tasks, None, bb.alt, bb.alt_base, bb.out_base, aligned_layout=hint)
        ok &= bool(np.array_equal(out, want))
res = eng.execute_soa([(0, 1, 1, 8), (0, 4, 1, 4), (0, 6, 2, 6)], "ABCFEFGH", "HGFEFCBA", "x" * 10, fill_dot=False)
ok &= res.astype(np.uint8).tobytes() == b"xxxxExGHBx"
# ---- the widened path: device task generation (class tables + FASTA framing, general catalogue), pipeline, gzip
import zlib
from tests.test_gpu_taskgen_general import random_world
from vcf2prot_b200.pipeline import DevicePipeline, csr_lists
from vcf2prot_b200.taskgen import DeviceCatalogue, execute_generated

hap, site = C.select_sites(cat, 8, np.random.default_rng(3))
plain = C.build_batch(prot, cat, hap, site, 8, "global", "packed")
img = C.fasta_image(prot, plain)
want = np.zeros(img.n_residues, np.uint8)
assert cengine.batch_execute(img.task_begin, img.tasks, prot.residues, img.alt, img.alt_base, want, img.out_base)[0] == 0
pipe = DevicePipeline(eng, prot, cat, C.default_names(prot), lanes=2)
sb, sites = csr_lists(hap, site, 8)
for gz in (False, True):
    out = np.zeros(len(want) + 4096, np.uint8)
    fb, _ = pipe.run_lists(sb, sites, 4, 1, gz, out=out)
    for s_ in range(4):
        got = out[int(fb[s_]):int(fb[s_ + 1])].tobytes()
        ok &= (zlib.decompress(got, wbits=31) if gz else got) == want[int(img.out_base[2 * s_]):int(img.out_base[2 * s_ + 2])].tobytes()
pipe.close()
w = random_world(11, n_tx=12, n_hap=10)
dc = DeviceCatalogue.from_instructions(*w.cat_args)
g = dc.generate_lists(w.site_begin, w.sites)
eng.set_reference(w.tape)
execute_generated(eng, g)
tape = dc.read(g.batch.out, g.batch.n_out, np.uint8)
ob = dc.read(g.batch.out_base, 11, np.uint64)
from oracle import taskgen as T
for h in range(10):
    tasks, alt, ann, res_len, _ = w.oracle_hap(h)
    ok &= tape[int(ob[h]):int(ob[h + 1])].tobytes().decode() == T.execute_tasks(tasks, w.tape.tobytes().decode(), alt, res_len)
dc.close()
# ---- round 2: fused missense chains and their near-misses, a haplotype in serial order among sorted ones, the skew mix
#      (warp-cooperative copy of out-of-phase payloads), error paths with garbage destinations, `-a`, the cohort runner
from tests.randtasks import chain_batch
from vcf2prot_b200 import EngineError

eng.set_tuning(-1, 0)
for seed, n_hap, mean, run in ((81, 4, 30000, 30.0), (82, 5, 60000, 300.0)):
    b = chain_batch(seed, n_hap, mean, run_mean=run)
    t = b["tasks"].copy()
    s0, s1 = int(b["task_begin"][1]), int(b["task_begin"][2])
    t[s0:s1] = t[s0:s1][np.random.default_rng(seed).permutation(s1 - s0)]  # haplotype 1 needs serial order
    b["tasks"] = t
    want = np.zeros(int(b["out_base"][-1]), np.uint8)
    assert cengine.batch_execute(b["task_begin"], b["tasks"], b["ref"], b["alt"], b["alt_base"], want, b["out_base"])[0] == 0
    for mode in ("replicas", "plain", None):
        if mode:
            eng.set_reference(b["ref"], mode)
        out, _ = eng.execute_batch(b["task_begin"], b["tasks"], None if mode else b["ref"], b["alt"], b["alt_base"], b["out_base"])
        ok &= bool(np.array_equal(out, want))
    bad = b["tasks"].copy()
    bad[int(b["task_begin"][3]) - 1, 2] = 0xFFFFFFF0  # garbage destination in front of the next haplotype
    try:
        eng.execute_batch(b["task_begin"], bad, b["ref"], b["alt"], b["alt_base"], b["out_base"])
        ok = False
    except EngineError:
        pass
prot4 = C.make_proteome(seed=0x5EED0001, n_tx=300, giant=3)
cat4 = C.make_catalogue(prot4, 2500, seed=0x5EED0004, mix=C.MIX_C4, fs_mean=150, fs_max=4000, sl_max=500, long_ins_mean=120,
                        long_ins_max=5000, lognormal_tails=True)
cat4.af[:] = 0.15
b4 = C.synth_batch(prot4, cat4, 12, seed=4)
want = np.zeros(b4.n_residues, np.uint8)
assert cengine.batch_execute(b4.task_begin, b4.tasks, prot4.residues, b4.alt, b4.alt_base, want, b4.out_base)[0] == 0
eng.set_reference(prot4.residues)
out, _ = eng.execute_batch(b4.task_begin, b4.tasks, None, b4.alt, b4.alt_base, b4.out_base)
ok &= bool(np.array_equal(out, want))
from vcf2prot_b200.cohort_run import CohortRunner

eng.set_reference(prot.residues)
pipe = DevicePipeline(eng, prot, cat, C.default_names(prot), lanes=2)
pipe.enable_all_records(prot.residues, prot.offsets, C.default_names(prot))
n_all = []
pipe.run_lists(sb, sites, 4, 3, False, sink=lambda f, n, data, begins: n_all.append(bytes(data[: int(begins[n])]).count(b">")) or 0, all_records=True)
ok &= sum(n_all) == 4 * 2 * prot.n_tx
pipe.close()
r = CohortRunner([0, 0], prot.residues, prot.offsets, C.default_names(prot), cat.t, cat.p, cat.cls, cat.rlen, cat.doff, cat.dlen, cat.pool, lanes=1)
res = r.run_lists(sb, sites, 4, 1, False, sink=lambda *a: 0)
ok &= int(res.total.n_samples) == 4
r.close()
print("SANITIZE_PROBE", "OK" if ok else "MISMATCH")
