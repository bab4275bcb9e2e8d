timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02i_bench_short.json 2> gpurun_out/r02i_bench_short.log; tail -c 200 gpurun_out/r02i_bench_short.log
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02i_bench_short.json').read().strip().splitlines() if l.startswith('{')][-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'frac',round(d['roofline']['frac'],3),'crc',d['result_crc32']); print(d['config']); print(d['passes'])
P
