#!/bin/bash
# C2 mix at 1k ... 100k samples on 1/2/4/8 GPUs (weak scaling: --samples is PER GPU, so a cohort of S samples on
# N GPUs runs with --samples S/N).  One JSON line per point into gpurun_out/c5_sweep.jsonl.  Run on a box with 8 GPUs:
#     gpurun --gpus 8 --timeout 3000 -- bash profiles/dev/c5_sweep.sh
# (~18 s of cohort synthesis per 2,504 samples per rank dominates the wall time; the extras are switched off).
set -u
out=gpurun_out/c5_sweep.jsonl
: > "$out"
common="--steps 20 --warmup 3 --e2e-steps 1 --no-taskgen --gzip-samples 0 --written-samples 0 --no-cpu-baseline --pipeline-samples 0"
for total in 1000 2500 5000 10000 25000 50000 100000; do
  for n in 1 2 4 8; do
    per=$(( total / n ))
    [ "$per" -gt 12500 ] && continue        # 12,500 samples = 100 GB of tape + tasks per GPU: the largest resident share
    [ "$per" -lt 100 ] && continue
    if [ "$n" -eq 1 ]; then
      python bench.py --samples "$per" $common 2>/dev/null | tail -1 >> "$out"
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29600 + n)) \
        bench.py --gpus "$n" --samples "$per" $common 2>/dev/null | grep '^{' | tail -1 >> "$out"
    fi
  done
done
python - <<'PY'
import json
for l in open("gpurun_out/c5_sweep.jsonl"):
    d = json.loads(l)
    print("%6d samples/GPU x %d GPUs: %.3g residues/s  (%.3f ms/step, e2e %.3g)" %
          (d["config"]["haplotypes_per_gpu"] // 2, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"]))
PY
