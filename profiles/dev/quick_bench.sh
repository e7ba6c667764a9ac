#!/bin/bash
# quick A/B of the copy kernel on the C2 cohort: kernel ms + whole-tape parity, nothing else (about 40 s on the box)
tag=${1:-x}
timeout 600 python bench.py --no-cpu-baseline --c3-samples 0 --no-taskgen --gzip-samples 0 --pipeline-samples 0 --e2e-steps 0 "${@:2}" > gpurun_out/qb_$tag.json 2> gpurun_out/qb_$tag.err
tail -c 300 gpurun_out/qb_$tag.err
python -c "
import json; d=json.load(open('gpurun_out/qb_$tag.json')); r=d['roofline']
print('$tag value %.3f T  step %.3f ms  kernel %.3f ms  parity %s bad %s  write_frac %.3f' % (d['value']/1e12, d['ms_per_step'], r['kernel_ms'], d['parity']['gpu_equals_oracle'], d['parity']['mismatching_haplotypes'], r['write_only']['frac']))"
