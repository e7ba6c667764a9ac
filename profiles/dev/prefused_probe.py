"""What would the copy kernel do if the plan pass handed it missense chains already fused?  (Upper bound, no parity.)

`R A R A R ...` -- reference runs with one (source - destination) offset, separated by 1-residue alterations that
exactly fill the holes -- is what 2/3 of the C2 cohort's tasks are.  The copy kernel fuses them per 32-task batch and per
tile, every time a tile is assembled.  This probe fuses them ONCE, on the host, over whole haplotypes (no 32-task
window), drops the 1-residue patches (so the bytes under them are the reference's: the output is NOT the oracle's,
only the copy structure is what a plan-level fusion would leave), and times the same kernel on the shorter task array.
One JSON line: kernel ms with the cohort as it is / pre-fused, tasks per tile of both.

  python profiles/dev/prefused_probe.py [samples]
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402

from bench_support import make_workload  # noqa: E402
from vcf2prot_b200 import GpuEngine  # noqa: E402


def prefuse(task_begin, tasks):
    t = tasks.astype(np.int64)
    src, ln, dst, st = t[:, 0], t[:, 1], t[:, 2], t[:, 3]
    n = len(t)
    hap = np.repeat(np.arange(len(task_begin) - 1), np.diff(task_begin).astype(np.int64))
    i = np.arange(2, n)
    link = np.zeros(n, bool)  # task i continues the chain of reference run i-2 through the patch i-1
    link[2:] = ((st[i] == 0) & (st[i - 2] == 0) & (st[i - 1] == 1) & (ln[i - 1] == 1) & (dst[i - 1] == dst[i - 2] + ln[i - 2]) &
                (dst[i] == dst[i - 1] + 1) & (src[i] - dst[i] == src[i - 2] - dst[i - 2]) & (hap[i] == hap[i - 2]))
    patch = np.zeros(n, bool)
    patch[:-1] = link[1:]
    # head of every task's chain: last index of the same parity at or before it that is not a continuation
    idx = np.arange(n)
    head = np.empty(n, np.int64)
    for p in (0, 1):
        sel = idx[p::2]
        head[p::2] = np.maximum.accumulate(np.where(link[p::2], -1, sel))
    end = dst + ln
    last = np.ones(n, bool)  # the chain's last member: nobody two behind it continues it
    last[:-2] = ~link[2:]
    new_end = end.copy()
    new_end[head[last]] = end[last]
    keep = ~link & ~patch
    out = tasks[keep].copy()
    out[:, 1] = (new_end[keep] - dst[keep]).astype(np.uint32)
    kept_before = np.concatenate([[0], np.cumsum(keep)])
    return kept_before[task_begin.astype(np.int64)].astype(np.uint64), out, int(patch.sum())


def main():
    samples = int(sys.argv[1]) if len(sys.argv) > 1 else 2504
    prot, cat, b = make_workload("c2", samples, 0)
    t0 = time.time()
    f_begin, f_tasks, n_patch = prefuse(b.task_begin, b.tasks)
    t_fuse = time.time() - t0
    dev = torch.device("cuda", 0)
    to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)
    eng = GpuEngine(0)
    eng.set_reference(to_dev(b.ref))
    n_out = b.n_residues
    d_out = torch.empty(n_out + 64, dtype=torch.uint8, device=dev)
    d_alt, d_alt_base, d_out_base = to_dev(b.alt), to_dev(b.alt_base), to_dev(b.out_base)
    res = {}
    # floor: every haplotype's tape as ONE reference run (a plain copy through the same kernel: one bulk load per tile;
    # the plan pass is slow here for an unrelated reason -- 5,008 tasks are ten warps, each lane filling ~500 lb[] entries)
    n_res_h = np.diff(b.out_base).astype(np.int64)
    one = np.zeros((b.n_hap, 4), np.uint32)
    one[:, 0] = np.arange(b.n_hap) % 16  # every source phase, one band of the proteome (the cohort's own access pattern)
    one[:, 1] = n_res_h
    one_begin = np.arange(b.n_hap + 1, dtype=np.uint64)
    for name, tb, tk in (("as_is", b.task_begin, b.tasks), ("prefused", f_begin, f_tasks), ("one_run_per_haplotype", one_begin, one)):
        d_tb, d_tk = to_dev(tb), to_dev(tk)
        args = (b.n_hap, d_tb, d_tk, None, d_alt, d_alt_base, d_out, d_out_base, len(tk), len(b.alt), n_out)
        for _ in range(5):
            eng.execute_batch_device(*args)
        grp, cp = [], []
        for _ in range(40):
            grp.append(eng.execute_batch_device(*args))
            cp.append(eng.last_copy_ms)
        res[name] = {"tasks": int(len(tk)), "tasks_per_8k_tile": round(len(tk) / (n_out / 8192), 2),
                     "launch_group_ms": round(float(np.mean(grp)), 4), "copy_kernel_ms": round(float(np.mean(cp)), 4)}
        del d_tb, d_tk
    res["patches_dropped"] = n_patch
    res["host_fusion_s"] = round(t_fuse, 1)
    res["note"] = "prefused: output differs from the oracle at the dropped 1-residue patches by construction (upper bound probe)"
    print(json.dumps(res))


if __name__ == "__main__":
    main()
