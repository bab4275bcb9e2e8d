python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -8
for n in 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_p_n$n.json 2> gpurun_out/r02_p_n$n.log
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_p_n$n.json').read().strip().splitlines()[-1])
print($n, 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'e2e_ms', round(d['e2e']['ms_per_step'],2), 'pack_ms', round(d['e2e']['host_pack_ms_per_step'],1), 'kms', {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()}, 'fused_ms', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3), 'crc', d['result_crc32'], 'passes', d['predict_stats']['passes'], 'launches', d['gpu_launches'])
PY
done
