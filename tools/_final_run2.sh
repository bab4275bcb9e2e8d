timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02h_bench_2gpu.json 2> gpurun_out/r02h_bench_2gpu.log; tail -c 300 gpurun_out/r02h_bench_2gpu.log
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02h_bench_2gpu.json').read().strip().splitlines() if l.startswith('{')][-1])
print('value',round(d['value']),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']),'frac',round(d['roofline']['frac'],3),'crc',d['result_crc32'],'n',d['n_gpus'], d['kernel_ms_per_step'])
P
