"""compute-sanitizer target: a few small batches through every path (registered/plain, aligned/packed, serial)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import cengine
from tests.randtasks import random_batch
from vcf2prot_b200 import GpuEngine
from vcf2prot_b200 import cohort as C

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
    out, _ = eng.execute_batch(bb.task_begin, bb.tasks, None, bb.alt, bb.alt_base, bb.out_base)
    ok &= bool(np.array_equal(out, want))
res = eng.execute_soa([(0, 1, 1, 8), (0, 4, 1, 4), (0, 6, 2, 6)], "ABCFEFGH", "HGFEFCBA", "x" * 10, fill_dot=False)
ok &= res.astype(np.uint8).tobytes() == b"xxxxExGHBx"
print("SANITIZE_PROBE", "OK" if ok else "MISMATCH")
