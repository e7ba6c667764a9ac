#!/bin/bash
# SURVEY 8(d) C5 / BASELINE configs[4]: ONE cohort of S samples (C2 mix, seed 0x5EED0003) at S = 1k ... 100k, sharded by
# contiguous sample ranges over N = 1/2/4/8 GPUs and streamed (bench.py --only-c3: the c3 section for that S) -- strong
# scaling per cohort size, with every haplotype of every point checked against the oracle.  One JSON line per point
# into gpurun_out/c5_sweep_n$N.jsonl.  Usage (on a box with >= N GPUs):  bash profiles/dev/c5_sweep.sh "1 2 4 8" "1000 2500 ..."
# Beside it (N = 1 only) the reference's own protocol, automation_scripts/performance_benchmark.py:60-91: whole-tool wall
# clock of the prebuilt binary at 1 ... 128 samples (bench.py --impl reference --ref-binary-samples S).
set -u
gpus=${1:-"1 2 4 8"}
sizes=${2:-"1000 2500 5000 10000 25000 50000 100000"}
for n in $gpus; do
  out=gpurun_out/c5_sweep_n$n.jsonl
  : > "$out"
  for s in $sizes; do
    chunk=1024; [ "$s" -lt 4096 ] && chunk=256
    if [ "$n" -eq 1 ]; then
      python bench.py --only-c3 --c3-samples "$s" --c3-chunk-samples $chunk 2>/dev/null | grep '^{' | tail -1 >> "$out"
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29600 + n)) \
        bench.py --gpus "$n" --only-c3 --c3-samples "$s" --c3-chunk-samples $chunk 2>/dev/null | grep '^{' | tail -1 >> "$out"
    fi
  done
done
if [ "${REF_BINARY:-0}" = "1" ]; then
  : > gpurun_out/c5_reference_binary.jsonl
  for s in 1 2 4 8 16 32 64 128; do
    python bench.py --impl reference --steps 1 --warmup 0 --cpu-seconds 1 --cpu-sample-haps $((2 * s)) --ref-binary-samples "$s" 2>/dev/null \
      | grep '^{' | tail -1 >> gpurun_out/c5_reference_binary.jsonl
  done
fi
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/c5_sweep_n*.jsonl")):
    for l in open(f):
        d = json.loads(l)
        p = d["parity"] or {}
        print("%7d samples x %d GPU(s): exec %.3g residues/s (%.1f ms), with task generation %.3g, wall %.2f s, parity %s (%s haplotypes)" %
              (d["samples"], d["n_gpus"], d["value"], d["exec_ms_max_rank"], d["value_with_task_generation"], d["wall_s_max_rank"],
               p.get("gpu_equals_oracle"), p.get("checked_haplotypes")))
try:
    for l in open("gpurun_out/c5_reference_binary.jsonl"):
        rb = json.loads(l)["cpu_baseline"].get("reference_binary")
        if rb:
            print("reference binary, %4d samples: whole tool %.2f s (parse %.2f, exec stage %.3f, write %.2f) = %.3g residues/s" %
                  (rb["samples"], rb["wall_s"], rb["parse_s"], rb["exec_stage_s"], rb["write_s"], rb["whole_tool_residues_per_s"]))
except OSError:
    pass
PY
