#!/usr/bin/env python
"""bench.py -- throughput of the sequence-generation engine (Task-array execution) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c4] [--samples S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path (validate + plan + copy launch group) over one cohort of synthetic,
device-resident Task arrays: every haplotype's result tape is materialised once.
Workload (config.workload):
  c2  BASELINE.json configs[1]: 2,504 phased samples (5,008 haplotypes) x 20k-transcript proteome,
      missense-dominated csq mix.  Default.  ~17 GB of residues per step per GPU.
  c4  skewed-segment stress (configs[3] mix), same sample count unless --samples is given.
N > 1: every rank owns its own contiguous sample range of the same size (weak scaling, no data-path collective);
value = residues produced by all ranks / max-over-ranks device time.

Prints ONE JSON line (rank 0).  `--impl reference` times the reference-equivalent CPU engine (the oracle's C
restatement of gir.rs:230-234 on UTF-32 tapes, haplotypes spread over all host threads like rayon par_iter,
exec.rs:34-40) on a bounded sample of the same workload; the reference itself is Rust and cannot be rebuilt here
(its prebuilt whole-tool binary is timed beside it when oracle/_ref/vcf2prot is present).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c4"])
    ap.add_argument("--samples", type=int, default=2504, help="phased samples per GPU (2 haplotypes each)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-chunk-haps", type=int, default=256)
    ap.add_argument("--e2e-depth", type=int, default=2, help="host-pointer chunks in flight (1 = no overlap)")
    ap.add_argument("--cpu-sample-haps", type=int, default=64)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", type=int, default=-1, help="copy-kernel variant (-1: engine's automatic choice)")
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--ref-mode", default="replicas", choices=["replicas", "plain"])
    ap.add_argument("--layout", default="packed", choices=["packed", "aligned"],
                    help="result-tape layout: packed = the reference's own (res_counter, haplotype_instruction.rs:132), "
                         "what a drop-in caller hands over; aligned = transcripts in phase with the proteome tape (<=30 "
                         "'.' pad bytes per transcript, never seen by the consumer) -- measured beside it (other_layout)")
    ap.add_argument("--no-registered-ref", action="store_true",
                    help="pass the proteome with every call (generic register path) instead of registering it once")
    ap.add_argument("--fasta-image", action="store_true",
                    help="emit `>{transcript}_{hap}\\n{seq}\\n` framing as extra copy segments (packed layout only): "
                         "the result tape is the FASTA file image")
    ap.add_argument("--maskdecode-samples", type=int, default=512,
                    help="samples of the cohort whose FORMAT/BCSQ mask matrix is decoded on the device (0 = skip)")
    ap.add_argument("--gzip-samples", type=int, default=256,
                    help="samples whose FASTA file image is gzip-compressed on the device (0 = skip)")
    ap.add_argument("--no-taskgen", action="store_true", help="skip the device-side Task generation measurement")
    ap.add_argument("--pipeline-samples", type=int, default=-1,
                    help="samples run through v2p_pipeline_run_lists (site lists -> .fasta / .fasta.gz images in pinned host "
                         "memory); -1 = the whole cohort, 0 = skip")
    ap.add_argument("--pipeline-chunk", type=int, default=128, help="samples per pipeline chunk")
    ap.add_argument("--written-samples", type=int, default=256,
                    help="samples whose files the pipeline also WRITES, {tmpdir}/{proband}.fasta and .fasta.gz (0 = skip)")
    ap.add_argument("--ref-binary-samples", type=int, default=0,
                    help="also time the reference's prebuilt whole-tool binary on this many samples (slow)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ workload
def make_workload(kind: str, n_samples: int, rank: int, layout: str = "packed", fasta: bool = False):
    from vcf2prot_b200 import cohort as C

    prot = C.make_proteome(seed=0x5EED0001, giant=(20 if kind == "c4" else 0))
    if kind == "c2":
        cat = C.make_catalogue(prot, 280000, seed=0x5EED0002)
    else:
        cat = C.make_catalogue(prot, 120000, seed=0x5EED0004, mix=C.MIX_C4, fs_mean=60, fs_max=4000, sl_max=500,
                               long_ins_mean=50, long_ins_max=5000, lognormal_tails=True)
    n_hap = 2 * n_samples
    parts = []
    step = 256
    for i, h0 in enumerate(range(0, n_hap, step)):
        parts.append(C.synth_batch(prot, cat, min(step, n_hap - h0), seed=(0x5EED0002 + 7919 * rank) * 1000 + i,
                                   layout=layout))
    if fasta:  # record framing as copy segments: the result tape is the .fasta file image
        parts = [C.fasta_image(prot, b) for b in parts]
    return prot, cat, C.concat_batches(parts)


def alg_bytes(batch) -> int:
    """sum(len) read + sum(len) written + '.' gap bytes + 16 B per packed task."""
    covered = int(batch.tasks[:, 1].astype(np.int64).sum())
    n_out = batch.n_residues
    return covered + n_out + 16 * len(batch.tasks)  # n_out = covered + gap bytes


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons, pw = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                clk, mxc, p = float(f[1]), float(f[2]), float(f[3])
            except ValueError:
                continue
            mx = mxc
            if t0 - 0.05 <= ts <= t1 + 0.1:
                sm.append(clk)
                pw.append(p)
                for n, v in zip(names, f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw) if pw else None}


# ------------------------------------------------------------------------------------------------ CPU side
def oracle_check_range(batch, prot, h0: int, h1: int, gpu_bytes: np.ndarray) -> bool:
    """Oracle (1-byte port) on haplotypes [h0,h1) of the cohort vs the GPU's bytes for the same range."""
    from oracle import cengine

    t0, t1 = int(batch.task_begin[h0]), int(batch.task_begin[h1])
    a0, a1 = int(batch.alt_base[h0]), int(batch.alt_base[h1])
    o0, o1 = int(batch.out_base[h0]), int(batch.out_base[h1])
    out = np.zeros(o1 - o0, np.uint8)
    rebase = lambda a, x: (a[h0:h1 + 1] - np.uint64(x)).astype(np.uint64)
    st, _, _ = cengine.batch_execute(rebase(batch.task_begin, t0), batch.tasks[t0:t1], prot.residues, batch.alt[a0:a1],
                                     rebase(batch.alt_base, a0), out, rebase(batch.out_base, o0), threads=os.cpu_count() or 1)
    return st == 0 and bool(np.array_equal(out, gpu_bytes))


def cpu_engine_rate(batch, prot, n_haps: int, seconds: float, threads: int, width: int, check_against=None):
    """Reference-equivalent CPU engine (oracle port) on the first n_haps haplotypes; returns residues/s."""
    from oracle import cengine

    n_haps = min(n_haps, batch.n_hap)
    t1 = int(batch.task_begin[n_haps])
    a1, o1 = int(batch.alt_base[n_haps]), int(batch.out_base[n_haps])
    dt = np.uint32 if width == 4 else np.uint8
    ref = prot.residues.astype(dt)
    alt = batch.alt[:a1].astype(dt)
    out = np.zeros(o1, dt)
    args = (batch.task_begin[:n_haps + 1], batch.tasks[:t1], ref, alt, batch.alt_base[:n_haps + 1], out,
            batch.out_base[:n_haps + 1])
    st, _, _ = cengine.batch_execute(*args, threads=threads)  # warm-up + page-in
    assert st == 0
    reps, t0 = 0, time.perf_counter()
    while True:
        st, _, _ = cengine.batch_execute(*args, threads=threads)
        reps += 1
        el = time.perf_counter() - t0
        if el >= seconds or reps >= 1000:
            break
    ok = None
    if check_against is not None:
        ok = bool(np.array_equal(out.astype(np.uint8), check_against[:o1]))
    return o1 * reps / el, el, reps, n_haps, o1, ok


def reference_binary_rate(prot, cat, batch, n_samples: int):
    """Whole-tool timing of the reference's own prebuilt binary on the first n_samples of the cohort."""
    from oracle import refbin
    from vcf2prot_b200 import cohort as C

    if not refbin.available() or n_samples <= 0:
        return None
    n_samples = min(n_samples, batch.n_hap // 2)
    sel = batch.kept_hap < 2 * n_samples if batch.kept_hap is not None else None
    return None if sel is None else _ref_binary_run(prot, cat, batch, n_samples, sel, refbin, C)


def _ref_binary_run(prot, cat, batch, n_samples, sel, refbin, C):
    hap, site = batch.kept_hap[sel], batch.kept_site[sel]
    used, inv = np.unique(site, return_inverse=True)
    mask = np.zeros((len(used), n_samples), np.uint8)
    np.bitwise_or.at(mask, (inv, hap // 2), (1 << (hap % 2)).astype(np.uint8))  # both haplotypes may carry a site
    refs = {prot.name(t): prot.seq(t) for t in range(prot.n_tx)}
    samples = ["S%05d" % i for i in range(n_samples)]
    lines = [refbin.VCF_HEADER, "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(samples) + "\n"]
    cell = ["0|0:0", "1|0:1", "0|1:2", "1|1:3"]
    for r, i in enumerate(used):
        lines.append("1\t%d\t.\tC\tT\t.\t.\tAC=1;BCSQ=%s\tGT:BCSQ\t%s\n" %
                     (100 + r, C.site_csq(prot, cat, int(i)), "\t".join(cell[m] for m in mask[r])))
    t0 = time.perf_counter()
    recs, stdout, rc = refbin.run_reference("".join(lines), refs, "mt", verbose=True, timeout=1200)
    wall = time.perf_counter() - t0
    st = refbin.stage_seconds(stdout)
    n_res = sum(len(s) for v in recs.values() for _, s in v)
    if rc != 0 or not st:
        return {"error": "reference binary rc=%d" % rc}
    return {"samples": n_samples, "residues": n_res, "wall_s": round(wall, 2), "parse_s": round(st["parse"], 2),
            "exec_stage_s": round(st["exec"], 3), "write_s": round(st["write"], 2),
            "exec_stage_residues_per_s": n_res / max(st["exec"], 1e-9), "whole_tool_residues_per_s": n_res / st["total"],
            "engine": "mt", "version": "0.1.2 (bins/Linux/vcf2prot)"}


def gzip_measure(args, prot, cat, eng, dev, local_rank, torch):
    """FASTA image of `--gzip-samples` samples produced on the device, then v2p_gzip_files device -> device; only the
    compressed bytes cross PCIe.  Beside it: zlib level 9 (what flate2 Compression::best amounts to) on one host core."""
    import zlib

    from vcf2prot_b200 import cohort as C
    from vcf2prot_b200.gzipdev import DeviceGzip

    ns = args.gzip_samples
    parts = [C.fasta_image(prot, C.synth_batch(prot, cat, min(256, 2 * ns - h0), seed=0x5EED0011 * 1000 + i, layout="packed"))
             for i, h0 in enumerate(range(0, 2 * ns, 256))]
    img = C.concat_batches(parts)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)
    d = [up(img.task_begin), up(img.tasks), up(img.alt), up(img.alt_base), up(img.out_base)]
    d_img = torch.empty(img.n_residues + 64, dtype=torch.uint8, device=dev)
    eng.execute_batch_device(img.n_hap, d[0], d[1], None, d[2], d[3], d_img, d[4], len(img.tasks), len(img.alt), img.n_residues)
    torch.cuda.synchronize()
    gz = DeviceGzip(local_rank)
    file_begin = np.ascontiguousarray(img.out_base[::2])
    cap = gz.bound(img.n_residues, ns)
    d_gz = torch.empty(cap, dtype=torch.uint8, device=dev)
    gz.compress_device(d_img.data_ptr(), file_begin, d_gz.data_ptr(), cap)  # warm-up (allocations)
    runs = [gz.compress_device(d_img.data_ptr(), file_begin, d_gz.data_ptr(), cap) for _ in range(5)]
    ob, res = runs[-1]
    ms = sorted(r.ms for _, r in runs)
    h_gz = torch.empty(int(ob[-1]), dtype=torch.uint8).pin_memory()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    h_gz.copy_(d_gz[: int(ob[-1])], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    d2h_ms = e0.elapsed_time(e1)
    # the judge: zlib inflates the first and the last sample's member back to the image the device produced
    image = d_img[: img.n_residues].cpu().numpy()
    ok = True
    for s_ in (0, ns - 1):
        dec = zlib.decompressobj(wbits=31)
        got = dec.decompress(h_gz.numpy()[int(ob[s_]):int(ob[s_ + 1])].tobytes())
        ok = ok and dec.eof and got == image[int(file_begin[s_]):int(file_begin[s_ + 1])].tobytes()
    one = image[int(file_begin[0]):int(file_begin[1])].tobytes()
    t0 = time.perf_counter()
    z9 = len(zlib.compress(one, 9))
    t_z9 = time.perf_counter() - t0
    gz.close()
    return {"samples": ns, "image_bytes": int(res.in_bytes), "gz_bytes": int(res.out_bytes), "ratio": res.in_bytes / max(1, res.out_bytes),
            "chunks": int(res.n_chunks), "stored_chunks": int(res.n_stored_chunks), "ms_median": ms[len(ms) // 2], "ms_best": ms[0],
            "image_gbs": res.in_bytes / (ms[len(ms) // 2] * 1e-3) / 1e9, "d2h_ms_of_gz_bytes": d2h_ms,
            "d2h_ms_if_uncompressed": d2h_ms * res.in_bytes / max(1, res.out_bytes), "inflates_to_the_image": bool(ok),
            "cpu_zlib9_one_core": {"mbs": len(one) / t_z9 / 1e6, "ratio": len(one) / z9, "sample_bytes": len(one)},
            "what": "v2p_gzip_files, device -> device: one gzip member per sample file, 16 KiB dynamic-Huffman chunks"}


def oracle_file_text(batch, prot, s: int) -> bytes:
    """Sample s's .fasta text from the ORACLE's tapes (hap-1 records, then hap-2 records, tape order)."""
    from oracle import cengine
    from vcf2prot_b200 import cohort as C

    h0, h1 = 2 * s, 2 * s + 2
    t0, t1 = int(batch.task_begin[h0]), int(batch.task_begin[h1])
    a0, a1 = int(batch.alt_base[h0]), int(batch.alt_base[h1])
    o0, o1 = int(batch.out_base[h0]), int(batch.out_base[h1])
    tape = np.zeros(o1 - o0, np.uint8)
    rebase = lambda a, x: (a[h0:h1 + 1] - np.uint64(x)).astype(np.uint64)
    st, _, _ = cengine.batch_execute(rebase(batch.task_begin, t0), batch.tasks[t0:t1], prot.residues, batch.alt[a0:a1],
                                     rebase(batch.alt_base, a0), tape, rebase(batch.out_base, o0))
    assert st == 0
    txt = []
    for k in (0, 1):
        base = int(batch.out_base[h0 + k]) - o0
        lo, hi = np.searchsorted(batch.ann_hap, [h0 + k, h0 + k + 1])
        for r in range(lo, hi):
            seq = tape[base + int(batch.ann_start[r]): base + int(batch.ann_end[r])].tobytes()
            txt.append(b">" + prot.name(int(batch.ann_tx[r])).encode() + b"_%d\n" % (k + 1) + seq + b"\n")
    return b"".join(txt)


def pipeline_measure(args, prot, cat, batch, eng, local_rank, barrier, shard, dev):
    """v2p_pipeline_run_lists on the timed cohort: the per-haplotype site lists go up (4 B/site), every sample's .fasta
    (then .fasta.gz) image lands in the pipeline's pinned ring and is handed to a sink; tasks, tapes and images never
    exist on the host.  First and last file are compared with the oracle's text."""
    import zlib

    from vcf2prot_b200.pipeline import DevicePipeline, csr_lists

    ns = batch.n_hap // 2 if args.pipeline_samples < 0 else min(args.pipeline_samples, batch.n_hap // 2)
    sel = batch.kept_hap < 2 * ns
    sb, sites = csr_lists(batch.kept_hap[sel], batch.kept_site[sel], 2 * ns)
    n_res = int((batch.ann_end - batch.ann_start)[batch.ann_hap < 2 * ns].sum())
    want_first, want_last = oracle_file_text(batch, prot, 0), oracle_file_text(batch, prot, ns - 1)
    pipe = DevicePipeline(eng, prot, cat, lanes=2, device=local_rank)
    out = {"samples": ns, "chunk_samples": args.pipeline_chunk, "lanes": 2, "residues": n_res,
           "api": "v2p_pipeline_run_lists (host site lists in, file images to a sink through the pipeline's pinned ring)"}
    warm = min(ns, 2 * args.pipeline_chunk)
    for gz in (False, True):
        got = {}

        def sink(first, n, data, begins):
            if first == 0:
                got["first"] = bytes(data[: int(begins[1])])
            if first + n == ns:
                got["last"] = bytes(data[int(begins[n - 1]): int(begins[n])])
            return 0

        pipe.run_lists(sb[: 2 * warm + 1], sites[: int(sb[2 * warm])], warm, args.pipeline_chunk, gz, sink=lambda *a: 0)  # allocations
        barrier()
        _, r = pipe.run_lists(sb, sites, ns, args.pipeline_chunk, gz, sink=sink)
        un = (lambda b: zlib.decompress(b, wbits=31)) if gz else (lambda b: b)
        wall = shard.max_over_ranks(r.wall_s, dev)  # all ranks run their own sample range at once (weak scaling)
        out["fasta_gz" if gz else "fasta"] = {
            "residues_per_s": shard.sum_over_ranks(n_res, dev) / wall, "wall_s": wall, "h2d_bytes": int(r.h2d_bytes), "d2h_bytes": int(r.out_bytes),
            "image_bytes": int(r.image_bytes), "records": int(r.n_records), "tasks": int(r.n_tasks), "chunks": int(r.n_chunks),
            "gen_ms": r.gen_ms, "exec_ms": r.exec_ms, "gzip_ms": r.gzip_ms,
            "first_and_last_file_equal_oracle_text": bool(un(got["first"]) == want_first and un(got["last"]) == want_last)}
    # ---- and onto the file system: {tmpdir}/{proband}.fasta[.gz] through the native directory writer (rank 0 only)
    nw = min(args.written_samples, ns)
    if nw > 0 and int(os.environ.get("RANK", "0")) == 0:
        import shutil
        import tempfile

        from vcf2prot_b200.pipeline import DirWriter

        n_res_w = int((batch.ann_end - batch.ann_start)[batch.ann_hap < 2 * nw].sum())
        names = ["S%06d" % i for i in range(nw)]
        out["written"] = {"samples": nw, "residues": n_res_w, "writer_threads": 8,
                          "what": "v2p_pipeline_run_lists -> v2p_dir_writer_sink: write(2) of the file images, one file per proband"}
        for gz in (False, True):
            tmpdir = tempfile.mkdtemp(prefix="v2p_written_")
            try:
                w = DirWriter(tmpdir, names, compressed=gz, threads=8)
                _, r = pipe.run_lists(sb[: 2 * nw + 1], sites[: int(sb[2 * nw])], nw, args.pipeline_chunk, gz, sink=w)
                one = open(os.path.join(tmpdir, names[0] + (".fasta.gz" if gz else ".fasta")), "rb").read()
                ok = (zlib.decompress(one, wbits=31) if gz else one) == want_first and w.files_written == nw
                out["written"]["fasta_gz" if gz else "fasta"] = {
                    "residues_per_s": n_res_w / r.wall_s, "wall_s": r.wall_s, "bytes": w.bytes_written, "files": w.files_written,
                    "file_gbs": w.bytes_written / r.wall_s / 1e9, "first_file_equals_oracle_text": bool(ok)}
                w.close()
            finally:
                shutil.rmtree(tmpdir, ignore_errors=True)
    pipe.close()
    return out


# ------------------------------------------------------------------------------------------------ main
def main():
    args = parse_args()
    if args.fasta_image:
        args.layout = "packed"  # a file image cannot contain pad bytes
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    from vcf2prot_b200 import GpuEngine, shard

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    t_gen = time.perf_counter()
    prot, cat, batch = make_workload(args.workload, args.samples, rank, args.layout, args.fasta_image)
    t_gen = time.perf_counter() - t_gen
    n_hap, n_out, n_tasks = batch.n_hap, batch.n_residues, len(batch.tasks)
    # residues produced: the aligned layout also writes '.' pads and the FASTA image also writes headers -- not counted
    n_res = int((batch.ann_end - batch.ann_start).sum())
    b_alg = alg_bytes(batch)

    to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)
    d_task_begin, d_tasks, d_ref = to_dev(batch.task_begin), to_dev(batch.tasks), to_dev(batch.ref)
    d_alt, d_alt_base, d_out_base = to_dev(batch.alt), to_dev(batch.alt_base), to_dev(batch.out_base)
    d_out = torch.empty(n_out + 64, dtype=torch.uint8, device=dev)

    eng = GpuEngine(local_rank)
    eng.set_tuning(args.variant, args.ctas_per_sm)
    if not args.no_registered_ref:
        eng.set_reference(d_ref, args.ref_mode)  # proteome registered once, as the FASTA is loaded once
        d_ref_arg = None
    else:
        d_ref_arg = d_ref
    side = torch.cuda.Stream(device=dev)
    eng.set_stream(side.cuda_stream)
    dargs = (n_hap, d_task_begin, d_tasks, d_ref_arg, d_alt, d_alt_base, d_out, d_out_base, n_tasks, len(batch.alt), n_out)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dkw = {"aligned_layout": args.layout == "aligned"}  # the producer's hint (V2P_FLAG_ALIGNED_LAYOUT)
    for _ in range(max(args.warmup, 3)):
        eng.execute_batch_device(*dargs, **dkw)
    barrier()

    sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else
                           int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    group_ms, copy_ms = [], []
    barrier()
    w0 = time.time()
    with torch.cuda.stream(side):
        ev0.record()
        for _ in range(args.steps):
            group_ms.append(eng.execute_batch_device(*dargs, **dkw))
            copy_ms.append(eng.last_copy_ms)
        ev1.record()
    barrier()
    w1 = time.time()
    launches = eng.launch_count() - launches0
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    dev_ms = ev0.elapsed_time(ev1)  # CUDA events on the launching stream, all K steps
    max_ms = shard.max_over_ranks(dev_ms, dev)  # the slowest rank's device time
    total_res = shard.sum_over_ranks(n_res, dev)  # residues produced by all ranks in one step
    total_haps = shard.sum_over_ranks(n_hap, dev)
    total_alg = shard.sum_over_ranks(b_alg, dev)
    ms_per_step = max_ms / args.steps

    # ---- end to end through the C ABI with HOST buffers: per step, every chunk's tasks/alt go H2D from pinned
    #      memory and every result tape comes back D2H into a pinned staging buffer (the FASTA writer's input)
    eng.set_stream(None)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    h_tasks, h_alt, h_ref = pin(batch.tasks), pin(batch.alt), pin(batch.ref)
    h_task_begin, h_alt_base, h_out_base = (np.ascontiguousarray(batch.task_begin), np.ascontiguousarray(batch.alt_base),
                                            np.ascontiguousarray(batch.out_base))
    chunks = [(h0, min(n_hap, h0 + args.e2e_chunk_haps)) for h0 in range(0, n_hap, args.e2e_chunk_haps)]
    max_chunk = max(int(batch.out_base[b] - batch.out_base[a]) for a, b in chunks)
    depth = max(1, min(args.e2e_depth, 3))  # host-pointer batches in flight (engine has 3 staging slots)
    h_outs = [torch.empty(max_chunk + 64, dtype=torch.uint8).pin_memory().numpy() for _ in range(depth)]
    h_out = h_outs[0]
    h2d = sum(int(batch.task_begin[b] - batch.task_begin[a]) * 16 + int(batch.alt_base[b] - batch.alt_base[a]) +
              3 * 8 * (b - a + 1) + (len(batch.ref) if args.no_registered_ref else 0) for a, b in chunks)

    def e2e_step():
        # chunk i's copy-back overlaps chunk i+1's upload + kernels; each in-flight chunk has its own pinned buffer
        pending = []
        for i, (a, b) in enumerate(chunks):
            ev = eng.execute_hap_range(a, b, h_task_begin, h_tasks, None if not args.no_registered_ref else h_ref, h_alt,
                                       h_alt_base, h_out_base, h_outs[i % depth], wait=False, **dkw)
            pending.append(ev)
            if len(pending) >= depth:
                eng.wait_event(pending.pop(0))
        for ev in pending:
            eng.wait_event(ev)

    if args.e2e_steps > 0:
        e2e_step()  # warm-up (allocates the engine's device staging)
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = max(time.perf_counter() - e0, 1e-9)
    e2e_s = shard.max_over_ranks(e2e_s, dev)
    # the last chunk sits in h_out: keep it for the parity spot-check against the device-resident result
    a, b = chunks[-1]
    o0, o1 = int(batch.out_base[a]), int(batch.out_base[b])
    h_last = h_outs[(len(chunks) - 1) % depth]
    e2e_matches_device = bool(np.array_equal(h_last[:o1 - o0], d_out[o0:o1].cpu().numpy())) if args.e2e_steps > 0 else None

    # ---- all of it behind one call: site lists -> .fasta / .fasta.gz images in pinned host memory (every rank its range)
    pipeline_line = None
    if (args.pipeline_samples != 0 and not args.no_registered_ref and not args.fasta_image and
            batch.kept_hap is not None and not args.no_cpu_baseline):
        pipeline_line = pipeline_measure(args, prot, cat, batch, eng, local_rank, barrier, shard, dev)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- CPU baseline (rank 0, N == 1 only): oracle port on a bounded sample, also the parity checker
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        nh = min(args.cpu_sample_haps, n_hap)
        o1 = int(batch.out_base[nh])
        gpu_sample = d_out[:o1].cpu().numpy()
        rate32, el, reps, nh, nres, ok = cpu_engine_rate(batch, prot, nh, args.cpu_seconds, threads, 4, gpu_sample)
        rate8, _, _, _, _, ok8 = cpu_engine_rate(batch, prot, nh, min(3.0, args.cpu_seconds), threads, 1, gpu_sample)
        # the tail of the cohort sits beyond 4 GiB of result tape: check those offsets against the oracle too
        nt = min(16, n_hap)
        t_o0, t_o1 = int(batch.out_base[n_hap - nt]), int(batch.out_base[n_hap])
        tail_ok = oracle_check_range(batch, prot, n_hap - nt, n_hap, d_out[t_o0:t_o1].cpu().numpy())
        parity = {"checked_haplotypes": nh + nt, "residues": nres + (t_o1 - t_o0), "gpu_equals_oracle": bool(ok and ok8 and tail_ok),
                  "head_haplotypes": nh, "tail_haplotypes": nt, "tail_result_tape_offset": t_o0,
                  "e2e_equals_device_path": e2e_matches_device}
        cpu = {"value": rate32, "unit": "residues/s", "cores": threads, "kind": "port",
               "sample": "first %d haplotypes (%d residues) of the same cohort, UTF-32 tapes like the reference "
                         "(gir.rs:18-22), %d passes in %.1f s, haplotypes over %d threads (exec.rs:34-40)" %
                         (nh, nres, reps, el, threads),
               "value_u8_tapes": rate8}
        rb = reference_binary_rate(prot, cat, batch, args.ref_binary_samples)
        if rb:
            cpu["reference_binary"] = rb

    # ---- the same cohort's Task arrays generated ON the device from the per-haplotype site lists
    taskgen = None
    other_line = None
    if world == 1 and not args.no_taskgen and not args.fasta_image and batch.kept_hap is not None:
        from vcf2prot_b200 import cohort as C
        from vcf2prot_b200.taskgen import DeviceCatalogue

        dc = DeviceCatalogue(prot, cat, local_rank)
        dc.generate(batch.kept_hap, batch.kept_site, n_hap, args.layout == "aligned")  # warm-up (allocations)
        g = dc.generate(batch.kept_hap, batch.kept_site, n_hap, args.layout == "aligned")
        same = (g.batch.n_tasks == n_tasks and g.batch.n_out == n_out and
                bool(np.array_equal(dc.read(g.batch.tasks, 4 * min(n_tasks, 1 << 22), np.uint32).reshape(-1, 4),
                                    batch.tasks[: min(n_tasks, 1 << 22)])) and
                bool(np.array_equal(dc.read(g.batch.out_base, n_hap + 1, np.uint64), batch.out_base)))
        if n_tasks > (1 << 22):  # and the tail of the task array
            tail = dc.read(g.batch.tasks + 16 * (n_tasks - (1 << 20)), 4 << 20, np.uint32).reshape(-1, 4)
            same = same and bool(np.array_equal(tail, batch.tasks[n_tasks - (1 << 20):]))
        n_sites = int(len(batch.kept_site))
        taskgen = {"gen_ms": g.gen_ms, "sites": n_sites, "tasks": n_tasks, "tasks_per_s": n_tasks / (g.gen_ms * 1e-3),
                   "h2d_bytes": 4 * n_sites + 8 * (n_hap + 1), "h2d_bytes_if_tasks_were_uploaded": 16 * n_tasks + len(batch.alt),
                   "equals_host_producer": same, "host_producer_seconds": round(t_gen, 1),
                   "what": "v2p_generate_tasks: per-haplotype site lists -> packed Task batch on the GPU (incl. the H2D of the lists)"}
        # ---- the same cohort in the OTHER result-tape layout, generated on the device and executed from there
        other = "aligned" if args.layout == "packed" else "packed"
        from vcf2prot_b200.taskgen import execute_generated

        go = dc.generate(batch.kept_hap, batch.kept_site, n_hap, other == "aligned")
        runs = [execute_generated(eng, go, aligned_layout=(other == "aligned")) for _ in range(3 + min(args.steps, 20))][3:]
        o_group, o_copy = float(np.mean([r[0] for r in runs])), float(np.mean([r[1] for r in runs]))
        # its records must be the primary layout's records (which the oracle checked): first and last haplotype
        rec_ok = True
        o_base = dc.read(go.batch.out_base, n_hap + 1, np.uint64)
        rows_of = lambda h: np.searchsorted(batch.ann_hap, [h, h + 1])
        for h in (0, n_hap - 1):
            lo, hi = rows_of(h)
            o_s = dc.read(go.ann_start + 8 * int(lo), int(hi - lo), np.uint64)
            o_e = dc.read(go.ann_end + 8 * int(lo), int(hi - lo), np.uint64)
            o_tape = dc.read(go.batch.out + int(o_base[h]), int(o_base[h + 1] - o_base[h]), np.uint8)
            p_tape = d_out[int(batch.out_base[h]):int(batch.out_base[h + 1])].cpu().numpy()
            for r in range(int(hi - lo)):
                a, b_ = p_tape[int(batch.ann_start[lo + r]):int(batch.ann_end[lo + r])], o_tape[int(o_s[r]):int(o_e[r])]
                rec_ok = rec_ok and a.shape == b_.shape and bool((a == b_).all())
        n_alg_o = int(go.batch.n_out) + n_res + 16 * int(go.batch.n_tasks)  # written + read + tasks
        other_line = {"layout": other, "value": n_res / (o_group * 1e-3), "ms_per_step": o_group, "kernel_ms": o_copy,
                      "alg_gbs_kernel": n_alg_o / (o_copy * 1e-3) / 1e9, "result_tape_bytes": int(go.batch.n_out),
                      "records_equal_primary_layout": bool(rec_ok),
                      "what": "same cohort, Task batch generated on the device in the other layout (v2p_generate_tasks) and "
                              "executed in place; per-call CUDA events, %d calls" % len(runs)}
        # ---- the general catalogue (every instruction code; one thread per transcript) on the same lists
        dg = DeviceCatalogue.from_instructions(*C.instruction_arrays(prot, cat), device=local_rank)
        sb_all = np.zeros(n_hap + 1, np.uint64)
        np.cumsum(np.bincount(batch.kept_hap, minlength=n_hap), out=sb_all[1:])
        dg.generate_lists(sb_all, batch.kept_site)  # warm-up (allocations)
        gg = dg.generate_lists(sb_all, batch.kept_site)
        g_same = int(gg.batch.n_tasks) == (n_tasks if args.layout == "packed" else int(go.batch.n_tasks))
        if args.layout == "packed":
            m = min(n_tasks, 1 << 22)
            g_same = (g_same and int(gg.batch.n_out) == n_out and
                      bool(np.array_equal(dg.read(gg.batch.tasks, 4 * m, np.uint32).reshape(-1, 4), batch.tasks[:m])) and
                      bool(np.array_equal(dg.read(gg.batch.tasks + 16 * (n_tasks - m), 4 * m, np.uint32).reshape(-1, 4), batch.tasks[n_tasks - m:])) and
                      bool(np.array_equal(dg.read(gg.batch.out_base, n_hap + 1, np.uint64), batch.out_base)))
        taskgen["general_catalogue"] = {"gen_ms": gg.gen_ms, "tasks_per_s": int(gg.batch.n_tasks) / (gg.gen_ms * 1e-3),
                                        "equals_host_producer": bool(g_same), "skipped_transcripts": int(gg.n_skipped),
                                        "what": "v2p_catalogue_create_ins + v2p_generate_tasks: the reference's Instruction values, "
                                                "all 22 codes' rules, one thread per transcript-on-haplotype (packed layout)"}
        dg.close()
        # ---- FORMAT/BCSQ bit-mask matrix -> per-haplotype site lists -> Task batch, on the device
        md_samples = min(n_hap // 2, args.maskdecode_samples)
        if md_samples > 0:
            sel = batch.kept_hap < 2 * md_samples
            mh, ms_ = batch.kept_hap[sel], batch.kept_site[sel]
            rec = C.make_records(cat, 0x5EED0009, 1)
            masks = C.encode_masks(rec, md_samples, mh, ms_)
            d_masks = torch.from_numpy(masks.view(np.int32)).to(dev)
            shape = masks.shape
            lst = dc.sites_from_masks(masks, rec.csq_begin, rec.csq_site)  # host matrix: pageable H2D inside decode_ms
            host_ms = lst.decode_ms
            dc.sites_from_masks(d_masks.data_ptr(), rec.csq_begin, rec.csq_site, shape)
            lst = dc.sites_from_masks(d_masks.data_ptr(), rec.csq_begin, rec.csq_site, shape)
            want_begin = np.zeros(2 * md_samples + 1, np.uint64)
            np.cumsum(np.bincount(mh, minlength=2 * md_samples), out=want_begin[1:])
            ok = (lst.n_sites == len(ms_) and bool(np.array_equal(dc.read(lst.sites, lst.n_sites, np.uint32), ms_.astype(np.uint32)))
                  and bool(np.array_equal(dc.read(lst.site_begin, 2 * md_samples + 1, np.uint64), want_begin)))
            g2 = dc.generate_from_lists(lst, args.layout == "aligned")
            n_t2 = int(batch.task_begin[2 * md_samples])
            ok = ok and g2.batch.n_tasks == n_t2 and bool(np.array_equal(
                dc.read(g2.batch.tasks, 4 * min(n_t2, 1 << 20), np.uint32).reshape(-1, 4), batch.tasks[: min(n_t2, 1 << 20)]))
            cells = shape[0] * shape[1]
            taskgen["maskdecode"] = {
                "records": int(shape[0]), "samples": int(md_samples), "words_per_cell": int(shape[2]), "carrier_bits": int(len(ms_)),
                "decode_ms": lst.decode_ms, "cells_per_s": cells / (lst.decode_ms * 1e-3),
                "matrix_gbs": masks.nbytes * 2 / (lst.decode_ms * 1e-3) / 1e9, "decode_ms_from_pageable_host": host_ms,
                "then_generate_ms": g2.gen_ms, "equals_host_lists_and_tasks": ok,
                "what": "v2p_sites_from_masks (matrix resident in HBM; streamed twice) -> v2p_generate_tasks_from_lists"}
            del d_masks
        dc.close()

    # ---- the -c path.  FASTA image in HBM -> one .fasta.gz per sample, compressed on the device
    gzip_line = None
    if world == 1 and args.gzip_samples > 0 and not args.no_registered_ref:
        gzip_line = gzip_measure(args, prot, cat, eng, dev, local_rank, torch)

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
    copy_avg_ms = float(np.mean(copy_ms))
    achieved = b_alg / (copy_avg_ms * 1e-3) / 1e9
    # the kernel must WRITE every residue once: measure this GPU's write-only ceiling (torch fill_, best of 5) beside it
    wbuf = d_out[: min(n_out, 8 << 30)]
    wbest = 1e9
    for _ in range(6):
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        wbuf.fill_(46)
        w1.record()
        torch.cuda.synchronize()
        wbest = min(wbest, w0.elapsed_time(w1))
    write_peak = wbuf.numel() / (wbest * 1e-3) / 1e9
    write_rate = n_out / (copy_avg_ms * 1e-3) / 1e9
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.isfile(tpath) and not args.fasta_image:
        mode = "plain" if (args.no_registered_ref or args.ref_mode == "plain") else "replicas"
        for ent in json.load(open(tpath)).get("entries", []):
            if (ent["workload"], ent["mode"], ent["layout"]) == (args.workload, mode, args.layout) and ent["dram_bytes_per_launch"]:
                if ent["samples"] == args.samples:
                    traffic, traffic_note = ent["dram_bytes_per_launch"], "ncu --set full capture of this configuration (%s)" % ent["capture"]
                else:  # same mix, reference mode and layout at another cohort size: DRAM bytes scale with residues
                    traffic = int(ent["dram_bytes_per_launch"] * (n_res / ent["residues"]))
                    traffic_note = "scaled by residues from the %d-sample ncu --set full capture (%s)" % (ent["samples"], ent["capture"])
                break

    value = total_res / (ms_per_step * 1e-3)  # every rank holds its own same-sized sample range (weak scaling)
    line = {
        "metric": "generated residues/sec", "value": value, "unit": "residues/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "%s: %d phased samples/GPU (%d haplotypes) x 20k-transcript proteome (%d residues), "
                               "%s csq mix" % (args.workload, args.samples, n_hap, len(batch.ref),
                                               "missense-dominated" if args.workload == "c2" else "skewed frameshift/insertion"),
                   "haplotypes_per_gpu": n_hap, "tasks_per_gpu": n_tasks, "residues_per_gpu": n_res, "result_tape_bytes_per_gpu": n_out,
                   "mean_task_bytes": n_res / max(n_tasks, 1), "l2_policy": "inputs_larger_than_l2 (output %.1f GB, tasks %.2f GB "
                   "per step; the %.1f MB proteome is L2-resident by design)" % (n_out / 1e9, n_tasks * 16 / 1e9, len(batch.ref) / 1e6),
                   "tile_variant": args.variant, "layout": args.layout,
                   "tile_order": "tape (V2P_FLAG_ALIGNED_LAYOUT)" if args.layout == "aligned" else "haplotype-interleaved", "fasta_image": bool(args.fasta_image), "reference_tape": "caller-supplied per call" if args.no_registered_ref else
                   "registered once (v2p_engine_set_reference, mode %s)" % args.ref_mode, "parallelism": "sample-sharded x%d, no collective" % world},
        "haplotypes_per_s": total_haps / (ms_per_step * 1e-3),
        "alg_gbs": total_alg / (ms_per_step * 1e-3) / 1e9,
        "clocks": clocks,
        "e2e": {"value": total_res * args.e2e_steps / e2e_s, "unit": "residues/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": n_out, "steps": args.e2e_steps, "chunk_haplotypes": args.e2e_chunk_haps,
                "chunks_in_flight": depth,
                "api": "v2p_execute_batch (host pointers, pinned, ASYNC), one call per chunk"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_note": traffic_note, "kernel": "k_copy_tiles", "peak_source": peak_src,
                     "alg_bytes_per_launch": b_alg, "kernel_ms": copy_avg_ms, "launch_group_ms": float(np.mean(group_ms)),
                     "kernel_share_of_step": copy_avg_ms / ms_per_step,
                     "write_only": {"achieved_gbs": write_rate, "peak_gbs": write_peak, "frac": write_rate / write_peak,
                                    "note": "result-tape bytes written / kernel time vs torch fill_ on the same GPU: the "
                                            "hard floor of this path is one DRAM write per residue"}},
        "cpu_baseline": cpu, "parity": parity, "other_layout": other_line, "taskgen": taskgen, "gzip": gzip_line, "pipeline": pipeline_line, "gen_seconds": round(t_gen, 1),
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_reference_arm(args):
    """Reference arm: the reference's CPU implementation of the path on this box's host cores."""
    threads = os.cpu_count() or 1
    nh = args.cpu_sample_haps
    prot, cat, batch = make_workload(args.workload, max(1, (nh + 1) // 2), 0)
    steps = max(1, args.steps)
    rates = []
    # each step = one bounded pass budget; keep the whole run within a few minutes
    per_step = min(args.cpu_seconds, max(1.0, 120.0 / (steps + max(args.warmup, 0))))
    for i in range(max(args.warmup, 0) + steps):
        rate, el, reps, nhh, nres, _ = cpu_engine_rate(batch, prot, nh, per_step, threads, 4)
        if i >= args.warmup:
            rates.append(rate)
    value = float(np.mean(rates))
    cpu = {"value": value, "unit": "residues/s", "cores": threads, "kind": "port",
           "sample": "%d haplotypes (%d residues) of the %s cohort, UTF-32 tapes (gir.rs:18-22), haplotypes over %d "
                     "threads (exec.rs:34-40); the Rust reference cannot be compiled in this image" %
                     (nhh, nres, args.workload, threads)}
    rb = reference_binary_rate(prot, cat, batch, args.ref_binary_samples or min(32, batch.n_hap // 2))
    if rb:
        cpu["reference_binary"] = rb
    line = {"impl": "reference", "metric": "generated residues/sec", "value": value, "unit": "residues/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": nres / value * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 (Rust char)",
            "data": "synthetic",
            "config": {"workload": "%s: bounded sample of %d haplotypes of the same synthetic cohort" % (args.workload, nhh)},
            "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": "residues/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
