#!/bin/bash
# quick A/B on the GPU box: parity tests (predict subset) + a short resident-only bench line
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline --no-e2e --sketch-genomes 0 --steps 3 --warmup 3 "$@" 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d['roofline']
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'], 2), 'kernel_ms', round(r['avg_launch_ms'], 4), 'frac', round(r['frac'], 3), 'passes', d['predict_stats']['passes'])"
