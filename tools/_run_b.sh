grep -m1 "model name" /proc/cpuinfo; grep -m1 flags /proc/cpuinfo | tr ' ' '\n' | grep -cE "^avx512vbmi$"
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02g_bench_1gpu.json 2> gpurun_out/r02g_bench_1gpu.log; tail -c 300 gpurun_out/r02g_bench_1gpu.log
SKB_NO_AVX512=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02g_bench_1gpu_avx2.json 2> gpurun_out/r02g_bench_1gpu_avx2.log
python - <<'P'
import json
for f in ('gpurun_out/r02g_bench_1gpu.json','gpurun_out/r02g_bench_1gpu_avx2.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f,'value',round(d['value']),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']),'pack_ms',round(d['e2e']['host_pack_ms_per_step'],2),'frac',round(d['roofline']['frac'],3),'crc',d['result_crc32'])
    sk=d.get('sketch')
    if sk: print('sketch kernel',round(sk['kernel_gbp_per_s'],1),'host_call',round(sk['host_call_gbp_per_s'],2),'cli',sk['cli_fasta_to_msh'])
P
