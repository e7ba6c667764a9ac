"""Seeded random Task arrays in the batched layout (test input generator, no reference semantics needed)."""
from __future__ import annotations

import numpy as np


def random_batch(seed: int, n_hap: int, mean_res: int, n_ref: int = 50000, gap_prob: float = 0.05,
                 shared_ref: bool = True, empty_hap_prob: float = 0.05, len_mix=(0.35, 0.25, 0.3, 0.1)):
    """Sorted, non-overlapping tasks with occasional '.' gaps, arbitrary (mis)alignments.
    len_mix = probabilities of task length classes: 1 byte, 2-15, 16-600, 600-40000."""
    rng = np.random.default_rng(seed)
    ref = rng.integers(65, 91, size=n_ref, dtype=np.uint8)
    tasks, task_begin, alt_parts, alt_base, out_base, ref_base = [], [0], [], [0], [0], [0]
    for h in range(n_hap):
        if rng.random() < empty_hap_prob:
            res_len = 0 if rng.random() < 0.5 else int(rng.integers(1, 40))  # no tasks: empty or all-dots tape
            task_begin.append(task_begin[-1])
            alt_base.append(alt_base[-1])
            out_base.append(out_base[-1] + res_len)
            ref_base.append(ref_base[-1] + (0 if shared_ref else 0))
            continue
        res_len = max(1, int(rng.exponential(mean_res)))
        n_alt = int(rng.integers(1, 4000))
        alt = rng.integers(97, 123, size=n_alt, dtype=np.uint8)
        pos = 0
        rows = []
        while pos < res_len:
            if rng.random() < gap_prob:
                pos += int(rng.integers(1, 40))
                continue
            c = rng.choice(4, p=len_mix)
            ln = [1, int(rng.integers(2, 16)), int(rng.integers(16, 600)), int(rng.integers(600, 40000))][c]
            if rng.random() < 0.02:
                ln = 0
            ln = min(ln, res_len - pos)
            stream = 1 if (ln <= n_alt and rng.random() < 0.4) else 0
            cap = n_alt if stream else n_ref
            ln = min(ln, cap)
            src = int(rng.integers(0, cap - ln + 1))
            rows.append((src, ln, pos, stream))
            pos += ln
        tasks += rows
        task_begin.append(task_begin[-1] + len(rows))
        alt_parts.append(alt)
        alt_base.append(alt_base[-1] + n_alt)
        out_base.append(out_base[-1] + res_len)
        ref_base.append(ref_base[-1])
    t = np.asarray(tasks, dtype=np.uint32).reshape(-1, 4)
    alt = np.concatenate(alt_parts) if alt_parts else np.zeros(0, np.uint8)
    u = lambda x: np.asarray(x, dtype=np.uint64)
    return dict(task_begin=u(task_begin), tasks=t, ref=ref, alt=alt, alt_base=u(alt_base), out_base=u(out_base),
                ref_base=None)
