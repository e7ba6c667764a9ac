"""ORACLE (test infrastructure only): harness around the reference's own prebuilt binary.

`oracle/_ref/vcf2prot` is an unmodified copy of /root/reference/bins/Linux/vcf2prot (v0.1.2, CPU-only),
put there by `make -C oracle ref` (it cannot be rebuilt here: no cargo/rustc in the image).  It is the
full VCF -> FASTA reference for `-g st|mt`; this module writes the minimal VCF dialect it accepts
(SURVEY.md section 8c), runs it, and parses

  * the per-sample FASTA files (record order is HashMap order, personalized_genome.rs:95,105 -> we sort), and
  * with RUN_SELECTED_TEST=1 DEBUG_TXP=<id>, the per-transcript `Vec<Task>` dump
    (transcript_instructions.rs:372-382), and with DEBUG_CPU_EXEC the haplotype table (gir.rs:212-222).

Nothing here is reachable from the product path, and nothing here is used on the GPU box except the
bounded whole-binary timing in bench.py's `--impl reference` / `cpu_baseline` legs.
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import tempfile
from typing import Dict, List, Optional, Sequence, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(HERE, "_ref", "vcf2prot")

VCF_HEADER = "##fileformat=VCFv4.2\n"


def available() -> bool:
    return os.path.isfile(REF_BIN) and os.access(REF_BIN, os.X_OK)


def vcf_text(samples: Sequence[str], records: Sequence[Tuple[Sequence[str], Sequence[Tuple[Sequence[int], Sequence[int]]]]]) -> str:
    """records: [(csq_list, per_sample[(hap1_idx_list, hap2_idx_list)])].

    Bit 2k of the FORMAT/BCSQ mask selects csq k on haplotype 1, bit 2k+1 on haplotype 2
    (MaskDecoder.rs:95-121); masks wider than 30 bits are split into comma-joined words of 15 csq
    (MaskDecoder.rs:122-153).
    """
    out = [VCF_HEADER, "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(samples) + "\n"]
    for pos, (csqs, per_sample) in enumerate(records):
        cells = []
        for (h1, h2) in per_sample:
            nwords = max(1, (len(csqs) + 14) // 15)
            words = [0] * nwords
            for k in h1:
                words[k // 15] |= 1 << (2 * (k % 15))
            for k in h2:
                words[k // 15] |= 1 << (2 * (k % 15) + 1)
            gt = "%d|%d" % (1 if h1 else 0, 1 if h2 else 0)
            cells.append(gt + ":" + ",".join(str(w) for w in words))
        out.append("1\t%d\t.\tC\tT\t.\t.\tAC=1;BCSQ=%s\tGT:BCSQ\t%s\n" % (100 + pos, ",".join(csqs), "\t".join(cells)))
    return "".join(out)


def fasta_text(ref_seqs: Dict[str, str]) -> str:
    return "".join(">%s\n%s\n" % (k, v) for k, v in ref_seqs.items())


def parse_fasta_records(text: str) -> List[Tuple[str, str]]:
    """`>{name}\\n{seq}\\n` records (personalized_genome.rs:97); sequences may be empty."""
    lines = text.split("\n")
    recs = []
    i = 0
    while i < len(lines):
        if lines[i].startswith(">"):
            seq = lines[i + 1] if i + 1 < len(lines) and not lines[i + 1].startswith(">") else ""
            recs.append((lines[i][1:], seq))
            i += 2 if (i + 1 < len(lines) and not lines[i + 1].startswith(">")) else 1
        else:
            i += 1
    return sorted(recs)


_TASK_RE = re.compile(r"exe_code:\s*(\d+),\s*start_pos:\s*(\d+),\s*length:\s*(\d+),\s*start_pos_res:\s*(\d+)")
_STAMP_RE = re.compile(r"(\d{4}-\d\d-\d\d \d\d:\d\d:\d\d\.\d+) UTC")


def parse_task_dumps(stdout: str) -> List[List[Tuple[int, int, int, int]]]:
    """Every `Vector of tasks is: [...]` block printed under DEBUG_TXP, in order of appearance."""
    out = []
    for block in stdout.split("Vector of tasks is:")[1:]:
        end = block.find("]\n")
        body = block[: end if end >= 0 else len(block)]
        out.append([tuple(int(x) for x in m.groups()) for m in _TASK_RE.finditer(body)])
    return out


def parse_cpu_exec_table(stdout: str) -> List[Tuple[int, int, int, int]]:
    """The 'CPU Execution Table' printed by the DEBUG_CPU_EXEC validator before it panics (gir.rs:212-222)."""
    rows = []
    seen = False
    for line in stdout.split("\n"):
        if line.startswith("index\tstream"):
            seen = True
            continue
        if seen:
            f = line.strip().split("\t")
            if len(f) >= 5 and all(x.isdigit() for x in f[:5]):
                rows.append((int(f[1]), int(f[2]), int(f[3]), int(f[4])))
            elif rows:
                break
    return rows


def stage_seconds(stdout: str) -> Optional[Dict[str, float]]:
    """Stage durations from the -v stamps (main.rs:17-60): parse / exec (fasta+instr+engine+collect) / write."""
    import datetime as dt

    stamps = {}
    keys = [("Reading and loading the VCF file", "t0"), ("VCF file have been parsed", "t1"),
            ("Loading the Reference file", "t2"), ("Personalized proteomes have been generated", "t3"),
            ("Write the generated results", "t4"), ("Execution finished", "t5")]
    for line in stdout.split("\n"):
        for needle, k in keys:
            if needle in line and k not in stamps:
                m = _STAMP_RE.search(line)
                if m:
                    s = m.group(1)
                    head, frac = s.split(".")
                    stamps[k] = dt.datetime.strptime(head, "%Y-%m-%d %H:%M:%S").timestamp() + float("0." + frac)
    if not all(k in stamps for _, k in keys):
        return None
    return {"parse": stamps["t1"] - stamps["t0"], "exec": stamps["t3"] - stamps["t2"],
            "write": stamps["t5"] - stamps["t4"], "total": stamps["t5"] - stamps["t0"]}


def run_reference(vcf: str, ref_seqs: Dict[str, str], engine: str = "st", env: Optional[Dict[str, str]] = None,
                  verbose: bool = False, keep_dir: Optional[str] = None, timeout: float = 600.0, write_all: bool = False):
    """Run the reference binary; returns (records_by_sample, stdout, returncode).  write_all: its `-a` flag."""
    if not available():
        raise RuntimeError("reference binary missing: run `make -C oracle ref` in the authoring container")
    d = keep_dir or tempfile.mkdtemp(prefix="v2p_ref_")
    try:
        os.makedirs(os.path.join(d, "out"), exist_ok=True)
        with open(os.path.join(d, "in.vcf"), "w") as f:
            f.write(vcf)
        with open(os.path.join(d, "ref.fa"), "w") as f:
            f.write(fasta_text(ref_seqs))
        e = {k: v for k, v in os.environ.items() if k not in
             ("DEBUG_GPU", "DEBUG_CPU_EXEC", "DEBUG_TXP", "INSPECT_TXP", "INSPECT_INS_GEN", "PANIC_INSPECT_ERR",
              "NO_TEST", "RUN_SELECTED_TEST")}
        e.update(env if env is not None else {"NO_TEST": "1"})
        cmd = [REF_BIN, "-f", os.path.join(d, "in.vcf"), "-r", os.path.join(d, "ref.fa"), "-g", engine,
               "-o", os.path.join(d, "out")]
        if verbose:
            cmd.append("-v")
        if write_all:
            cmd.append("-a")
        p = subprocess.run(cmd, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout)
        stdout = p.stdout.decode("utf-8", "replace")
        recs = {}
        for fn in sorted(os.listdir(os.path.join(d, "out"))):
            if fn.endswith(".fasta"):
                with open(os.path.join(d, "out", fn)) as f:
                    recs[fn[:-6]] = parse_fasta_records(f.read())
        return recs, stdout, p.returncode
    finally:
        if keep_dir is None:
            shutil.rmtree(d, ignore_errors=True)
