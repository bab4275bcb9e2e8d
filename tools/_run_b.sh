timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
run() { timeout 300 python bench.py --no-cpu-baseline --no-extras --steps 5 --warmup 3 $AB_ARGS 2>gpurun_out/err_$1.log | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d['roofline']
print('$1', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value']), 'e2e_ms', round(d['e2e']['ms_per_step'],2), 'pack', round(d['e2e']['host_pack_ms_per_step'],1), d['e2e']['timeline_ms'], 'crc', d['result_crc32'], 'kms', {k: round(v, 2) for k, v in d['kernel_ms_per_step'].items()})" || echo "$1 FAILED"; }
run base
SKB_BENCH_PACK_ALL=1 run packall
SKB_TRACE_PASSES=1 run trace
grep "predict call" gpurun_out/err_trace.log | tail -12
