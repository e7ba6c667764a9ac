"""Where does the copy grid's tail come from?  Wall time of every warp of k_copy_tiles on the C2 cohort, grouped by CTA and by
the SM the CTA most likely ran on (CTA b of a 3-CTAs/SM persistent grid).  python profiles/dev/warp_time_probe.py [samples]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench_support import make_workload  # noqa: E402
from vcf2prot_b200 import GpuEngine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2504
prot, cat, b = make_workload("c2", n, 0)
dev = torch.device("cuda:0")
up = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)
eng = GpuEngine(0)
eng.set_reference(up(prot.residues))
out = torch.empty(b.n_residues + 64, dtype=torch.uint8, device=dev)
args = (b.n_hap, up(b.task_begin), up(b.tasks), None, up(b.alt), up(b.alt_base), out, up(b.out_base), len(b.tasks), len(b.alt), b.n_residues)
for _ in range(3):
    eng.execute_batch_device(*args)
eng.profile_warps(True)
runs = []
for _ in range(5):
    eng.execute_batch_device(*args)
    runs.append(eng.read_warp_ns().astype(np.float64))
w = np.mean(runs, axis=0)
cta = w.reshape(-1, 8)
per_cta = cta.max(axis=1)
res = {"warps": int(w.size), "warp_ms": {"min": w.min() / 1e6, "mean": w.mean() / 1e6, "max": w.max() / 1e6},
       "cta_ms_max": {"min": per_cta.min() / 1e6, "mean": per_cta.mean() / 1e6, "max": per_cta.max() / 1e6},
       "within_cta_spread_ms_mean": float((cta.max(axis=1) - cta.min(axis=1)).mean() / 1e6),
       "run_to_run_corr_of_warp_times": float(np.corrcoef(runs[0], runs[-1])[0, 1]),
       "slowest_ctas": [int(i) for i in np.argsort(-per_cta)[:12]], "fastest_ctas": [int(i) for i in np.argsort(per_cta)[:12]],
       "cta_ms_by_index_mod_148_spread": float(np.ptp([per_cta[i::148].mean() for i in range(148)]) / 1e6),
       "cta_ms_first_wave_vs_later": [float(per_cta[:148].mean() / 1e6), float(per_cta[148:296].mean() / 1e6), float(per_cta[296:].mean() / 1e6)]}
print(json.dumps(res))
