for pr in 0 8192; do
timeout 900 python bench.py --config c4 --steps 3 --warmup 3 --pass-reads $pr > gpurun_out/r02_c4_$pr.json 2> gpurun_out/r02_c4_$pr.log; tail -1 gpurun_out/r02_c4_$pr.log; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_c4_$pr.json').read().strip().splitlines()[-1])
print($pr, 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'e2e_ms', round(d['e2e']['ms_per_step'],2), 'pack_ms', round(d['e2e']['host_pack_ms_per_step'],1), 'kms', {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()}, 'fused_ms', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3), 'crc', d['result_crc32'], 'passes', d['predict_stats']['passes'], d['consensus'])
PY
done
