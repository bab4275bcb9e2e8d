SKB_PIPELINE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
run() { timeout 300 python bench.py --no-cpu-baseline --no-extras --no-e2e --steps 5 --warmup 3 $AB_ARGS 2>gpurun_out/err_$1.log | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d['roofline']
print('$1', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 2), 'kernel_ms', round(r['avg_launch_ms'], 4), 'frac', round(r['frac'], 3), 'passes', d['predict_stats']['passes'], 'B', d['config']['reads_per_pass_max'], 'crc', d['result_crc32'], 'launches', d['gpu_launches'], 'cands', d['predict_stats']['candidates'], 'kms', {k: round(v, 2) for k, v in d['kernel_ms_per_step'].items()})" || echo "$1 FAILED"; }
run base
SKB_PIPELINE=1 run pipe
AB_ARGS="--refs 5000" run r5000
AB_ARGS="--refs 5000" SKB_PIPELINE=1 run r5000pipe
AB_ARGS="--config c4" SKB_PIPELINE=1 run c4pipe
