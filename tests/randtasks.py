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


def chain_batch(seed: int, n_hap: int, mean_res: int, n_ref: int = 60000, run_mean: float = 40.0):
    """Task arrays shaped like missense-dominated transcripts -- `R A R A R ...` with the reference runs of one
    transcript sharing one (source - destination) offset and 1-residue alterations filling the holes -- plus every
    near-miss of that pattern: 2-residue alterations, holes nobody fills ('.' gap), a follow-up run shifted by one
    (deletion) or sourced one further (insertion), reference-sourced 1-residue "patches", runs of 1-3 residues between
    patches, transcripts back to back with the same offset by chance, zero-length tasks in between.  Sorted, disjoint."""
    rng = np.random.default_rng(seed)
    ref = rng.integers(65, 91, size=n_ref, dtype=np.uint8)
    tasks, task_begin, alt_parts, alt_base, out_base = [], [0], [], [0], [0]
    for h in range(n_hap):
        res_len = max(8, int(rng.exponential(mean_res)))
        alt = rng.integers(97, 123, size=int(rng.integers(8, 3000)), dtype=np.uint8)
        rows, pos = [], 0
        a_cur = 0
        while pos < res_len:
            # one transcript: source offset src0 for destination pos
            t_len = min(res_len - pos, max(1, int(rng.exponential(500))))
            src = int(rng.integers(0, n_ref - t_len - 8))
            if rows and rng.random() < 0.05:
                src = min(rows[-1][0] + rows[-1][1], n_ref - t_len - 8) if rows[-1][3] == 0 else src  # same offset as the run before
            end = pos + t_len
            while pos < end:
                run = min(end - pos, int(rng.choice([1, 2, 3, max(1, int(rng.exponential(run_mean)))], p=[.08, .06, .06, .8])))
                rows.append((src, run, pos, 0))
                pos += run
                src += run
                if pos >= end:
                    break
                kind = rng.choice(8, p=[.62, .06, .06, .06, .05, .05, .05, .05])
                a_off = int(a_cur % (len(alt) - 4))
                a_cur += 2
                if kind == 0:    # the missense patch: 1 residue from alt, reference continues unshifted
                    rows.append((a_off, 1, pos, 1)); pos += 1; src += 1
                elif kind == 1:  # 2-residue substitution
                    n = min(2, end - pos); rows.append((a_off, n, pos, 1)); pos += n; src += n
                elif kind == 2:  # a hole nobody fills (reads '.'), reference continues unshifted
                    pos += 1; src += 1
                elif kind == 3:  # patch, then the reference continues one residue further on (deletion)
                    rows.append((a_off, 1, pos, 1)); pos += 1; src += 2
                elif kind == 4:  # patch, then the reference repeats the patched position (insertion)
                    rows.append((a_off, 1, pos, 1)); pos += 1
                elif kind == 5:  # a reference-sourced single residue from elsewhere
                    rows.append((int(rng.integers(0, n_ref - 1)), 1, pos, 0)); pos += 1; src += 1
                elif kind == 6:  # patch with a zero-length task behind it
                    rows.append((a_off, 1, pos, 1)); pos += 1; src += 1
                    rows.append((0, 0, pos, int(rng.integers(0, 2))))
                else:            # patch whose follow-up run leaves a one-residue hole
                    rows.append((a_off, 1, pos, 1)); pos += 2; src += 2
            pos = max(pos, end)
        rows = [(s_, l_, d_, st) for (s_, l_, d_, st) in rows if d_ + l_ <= res_len and s_ + l_ <= (len(alt) if st else n_ref)]
        tasks += rows
        task_begin.append(task_begin[-1] + len(rows))
        alt_parts.append(alt)
        alt_base.append(alt_base[-1] + len(alt))
        out_base.append(out_base[-1] + res_len)
    t = np.asarray(tasks, dtype=np.uint32).reshape(-1, 4)
    u = lambda x: np.asarray(x, dtype=np.uint64)
    return dict(task_begin=u(task_begin), tasks=t, ref=ref, alt=np.concatenate(alt_parts), alt_base=u(alt_base),
                out_base=u(out_base), ref_base=None)
