timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02h_bench_1gpu.json 2> gpurun_out/r02h_bench_1gpu.log; tail -c 200 gpurun_out/r02h_bench_1gpu.log
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02h_bench_1gpu.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']),'pack_ms',round(d['e2e']['host_pack_ms_per_step'],2),'frac',round(d['roofline']['frac'],3),'crc',d['result_crc32'])
print(json.dumps(d['sketch'])[:1500])
P
