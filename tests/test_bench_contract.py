"""bench.py keeps its contract: ONE JSON line with the keys the driver reads, on a tiny cohort.
CPU: the reference arm (`--impl reference`).  GPU: our arm, every section switched on, all parity flags true."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "e2e"}


def run_bench(*args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample-haps", "4", "--cpu-seconds", "0.5",
                  "--ref-binary-samples", "2")
    assert BASE <= set(d) and d["impl"] == "reference" and d["metric"] == "generated residues/sec" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"] == {"value": d["value"], "unit": "residues/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "haplotypes" in cb["sample"]
    assert "workload" in d["config"]


@pytest.mark.gpu
def test_our_arm_line_on_a_tiny_cohort():
    d = run_bench("--samples", "24", "--steps", "3", "--warmup", "3", "--e2e-steps", "1", "--e2e-chunk-haps", "16", "--cpu-sample-haps", "8",
                  "--cpu-seconds", "0.5", "--maskdecode-samples", "8", "--gzip-samples", "8", "--pipeline-chunk", "5",
                  "--written-samples", "6", "--c3-samples", "40", "--c3-chunk-samples", "16", "--dropin-haps", "6")
    assert BASE <= set(d) and d["n_gpus"] == 1 and d["scaling"] == "weak" and d["dtype"] == "u8" and d["vs_baseline"] is None
    assert d["gpu_launches"] == 7 * d["steps"] and d["warmup"] >= 3  # init, plan x4, copy, status hand-off
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] == d["config"]["result_tape_bytes_per_gpu"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["kernel"] == "k_copy_tiles"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert r["dram_frac"] is None or 0 < r["dram_frac"] < 1.2
    assert r["write_only"]["peak_gbs"] == r["write_only"]["ceilings_gbs"]["tma_bulk_store_8k"]
    p = d["parity"]
    assert p["gpu_equals_oracle"] is True and p["e2e_equals_device_path"] is True
    assert p["checked_haplotypes"] == d["config"]["haplotypes_per_gpu"] == p["haplotypes_per_gpu"] and p["all_ranks"] is True
    assert d["dropin"]["gpu_equals_oracle"] is True and d["dropin"]["haplotypes_per_s"] > 0 and d["dropin"]["host_threads"] >= 1
    c3 = d["c3"]
    assert c3["scaling"] == "strong" and c3["samples"] == 40 and c3["value"] > 0 and c3["chunks_this_rank"] == 3
    assert c3["parity"]["gpu_equals_oracle"] is True and c3["parity"]["checked_haplotypes"] == 80 == c3["parity"]["haplotypes"]
    assert d["other_layout"]["records_equal_primary_layout"] is True
    t = d["taskgen"]
    assert t["equals_host_producer"] and t["general_catalogue"]["equals_host_producer"] and t["maskdecode"]["equals_host_lists_and_tasks"]
    assert d["gzip"]["inflates_to_the_image"] is True
    for k in ("fasta", "fasta_gz"):
        assert d["pipeline"][k]["first_and_last_file_equal_oracle_text"] is True
        assert d["pipeline"]["written"][k]["first_file_equals_oracle_text"] is True
