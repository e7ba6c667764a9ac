"""Single-process multi-GPU: v2p_cohort_run_lists over every visible GPU (no torchrun), C2-mix cohort, null sink and the
directory writer.  python profiles/dev/cohort_run_probe.py [samples] [devices, e.g. 0,1]"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from synth import cohort as C  # noqa: E402
from synth import devgen  # noqa: E402
from vcf2prot_b200.cohort_run import CohortRunner  # noqa: E402
from vcf2prot_b200.pipeline import DirWriter  # noqa: E402

n_samples = int(sys.argv[1]) if len(sys.argv) > 1 else 2504
devs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else list(range(torch.cuda.device_count()))
prot = C.make_proteome(seed=0x5EED0001)
cat = C.make_catalogue(prot, 280000, seed=0x5EED0003)
# the cohort's site lists, made on GPU 0 by the synthetic generator and brought to the host (the runner takes host lists)
gen = devgen.DeviceCohort(cat, 0x5EED0003, devs[0])
begins, sites = [np.zeros(1, np.uint64)], []
for h0 in range(0, 2 * n_samples, 1024):
    n = min(1024, 2 * n_samples - h0)
    b, s, _ = gen.lists(h0, n)
    begins.append(b.cpu().numpy().astype(np.uint64)[1:] + begins[-1][-1])
    sites.append(s.cpu().numpy().astype(np.uint32))
gen.close()
sb, st = np.concatenate(begins), np.concatenate(sites)
t0 = time.perf_counter()
r = CohortRunner(devs, prot.residues, prot.offsets, C.default_names(prot), cat.t, cat.p, cat.cls, cat.rlen, cat.doff, cat.dlen, cat.pool, lanes=2)
t_create = time.perf_counter() - t0
out = {"devices": devs, "samples": n_samples, "sites": int(len(st)), "create_s": round(t_create, 2)}
for gz in (False, True):
    nwarm = min(n_samples, 128 * 2 * len(devs))  # warm-up: one chunk per lane of every worker (device + pinned allocations)
    r.run_lists(sb[: 2 * nwarm + 1], st[: int(sb[2 * nwarm])], nwarm, 128, gz, sink=None)
    res = r.run_lists(sb, st, n_samples, 128, gz, sink=None)
    n_res = int(res.total.image_bytes)  # file text incl. headers; residues = image - framing
    out["fasta_gz" if gz else "fasta"] = {
        "wall_s": res.total.wall_s, "image_bytes": int(res.total.image_bytes), "out_bytes": int(res.total.out_bytes),
        "records": int(res.total.n_records), "image_gbs": res.total.image_bytes / res.total.wall_s / 1e9,
        "per_device_wall_s": [res.per_device[g].wall_s for g in range(len(devs))],
        "per_device_exec_ms": [res.per_device[g].exec_ms for g in range(len(devs))],
        "per_device_host_wall_s": [{k: round(getattr(res.per_device[g], k + "_wall_s"), 3) for k in ("gen", "exec", "gzip", "wait", "sink")}
                                   for g in range(len(devs))],
        "first_sample": [int(res.first_sample[g]) for g in range(len(devs) + 1)]}
with tempfile.TemporaryDirectory() as d:
    nw = min(n_samples, 512)
    w = DirWriter(d, ["S%06d" % i for i in range(n_samples)], compressed=False, threads=8)
    res = r.run_lists(sb[: 2 * nw + 1], st[: int(sb[2 * nw])], nw, 128, False, sink=w, concurrent_sink=True)
    out["written"] = {"samples": nw, "files": w.files_written, "bytes": w.bytes_written, "wall_s": res.total.wall_s,
                      "file_gbs": w.bytes_written / res.total.wall_s / 1e9}
    w.close()
r.close()
print(json.dumps(out))
