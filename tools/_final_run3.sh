timeout 600 python bench.py --config c4 --steps 5 --warmup 3 > gpurun_out/r02h_bench_c4_1gpu.json 2> gpurun_out/r02h_bench_c4_1gpu.log; tail -c 300 gpurun_out/r02h_bench_c4_1gpu.log
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02h_bench_c4_1gpu.json').read().strip().splitlines() if l.startswith('{')][-1])
print('value',round(d['value']),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']),'pack_ms',round(d['e2e']['host_pack_ms_per_step'],1),'frac',round(d['roofline']['frac'],3),'crc',d['result_crc32'])
P
