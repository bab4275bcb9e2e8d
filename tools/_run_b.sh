timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_l_bench.json 2> gpurun_out/r02_l_bench.log; tail -3 gpurun_out/r02_l_bench.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_l_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','roofline','roofline_zero_hit','kernel_ms_per_step','parity','cpu_baseline','sketch','result_crc32','gpu_launches'):
    print(k, json.dumps(d.get(k))[:700])
PY
