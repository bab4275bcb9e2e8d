#!/bin/bash
# tools/ab.sh <variant>...  : short resident-only bench line per variant library ("base" = the in-tree library);
# result_crc32 must be the same for every variant (same answer for all 100,000 reads)
for v in "$@"; do
  if [ "$v" = base ]; then unset SKB_LIB; else export SKB_LIB=$PWD/sketchy_b200/build/variants/lib_$v.so; fi
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --sketch-genomes 0 --steps 3 --warmup 3 $AB_ARGS 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d['roofline']
print('$v', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 2), 'kernel_ms', round(r['avg_launch_ms'], 4), 'frac', round(r['frac'], 3), 'passes', d['predict_stats']['passes'], 'crc', d['result_crc32'], 'kms', {k: round(v, 2) for k, v in d['kernel_ms_per_step'].items()})" || echo "$v FAILED"
done
