timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --no-extras --steps 5 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d['roofline']
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value']), 'e2e_ms', round(d['e2e']['ms_per_step'],2), 'kernel_ms', round(r['avg_launch_ms'], 4), 'frac', round(r['frac'], 3), 'passes', d['predict_stats']['passes'], 'B', d['config']['reads_per_pass_max'], 'crc', d['result_crc32'], 'kms', {k: round(v, 2) for k, v in d['kernel_ms_per_step'].items()})
print(d['e2e'])"
nproc; lscpu | grep -E "Model name|Thread|Core|Socket" 
