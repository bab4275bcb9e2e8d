timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -x 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02d_n2.json 2> gpurun_out/r02d_n2.log
tail -2 gpurun_out/r02d_n2.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02d_n2.json').read().strip().splitlines()[-1])
print('n2 value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'e2e_ms', round(d['e2e']['ms_per_step'],2), 'pack_ms', round(d['e2e']['host_pack_ms_per_step'],1), 'kms', {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()}, 'fused_ms', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3), 'crc', d['result_crc32'], 'passes', d['predict_stats']['passes'], 'B', d['config']['reads_per_pass_max'], d['e2e']['timeline_ms'])
PY
