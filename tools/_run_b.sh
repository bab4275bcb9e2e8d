tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 ${@:4} > gpurun_out/$3.json 2> gpurun_out/$3.log; tail -1 gpurun_out/$3.log | cut -c1-200; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$3.json').read().strip().splitlines()[-1])
    print('$3 value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']) if d.get('e2e') else None, 'e2e_ms', round(d['e2e']['ms_per_step'],2) if d.get('e2e') else None, 'kms', {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()}, 'fused_ms', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3), 'crc', d['result_crc32'], 'passes', d['predict_stats']['passes'], 'B', d['config']['reads_per_pass_max'])
except Exception as e: print('$3 FAILED', e)
PY
}
tr 8 29521 r02_bench_8gpu --steps 5 --warmup 3
tr 4 29522 r02_bench_4gpu --steps 5 --warmup 3
tr 8 29523 r02_bench_c5_8gpu --config c5 --steps 2 --warmup 3 --no-cpu-baseline --no-extras
nvidia-smi --query-gpu=memory.used --format=csv | head -3
