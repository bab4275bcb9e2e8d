# one-call GPU check of the tree as it stands: parity tests, smoke, a short bench line
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'], 3), 'crc', d['result_crc32'], 'launches', d['gpu_launches'])"
