#!/bin/bash
# tools/prof_round.sh <tag> : one GPU-box call that produces the round's evidence under gpurun_out/:
#   <tag>_bench_1gpu.json   full default bench line (not under a profiler)
#   <tag>_launches.csv      ncu launch list (gpu__time_duration.sum, --clock-control none) of a short bench run
#   <tag>_fused.ncu-rep     ncu --set full of one steady-state fused_kernel launch
#   <tag>_hash.ncu-rep      ncu --set full of one hash_kernel launch of the sketch leg (2.8 Mbp assemblies)
# tools/summarize_profiles.py turns them into the tracked summaries under profiles/.
tag=${1:-r02}
o=gpurun_out
mkdir -p $o
timeout 600 python bench.py --steps 5 --warmup 3 > $o/${tag}_bench_1gpu.json 2> $o/${tag}_bench_1gpu.log
tail -c 400 $o/${tag}_bench_1gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $o/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $o/${tag}_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 40 -c 1 -o $o/${tag}_fused -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $o/${tag}_fused_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hash_kernel -s 1 -c 1 -o $o/${tag}_hash -f \
  python bench.py --sketch-only --sketch-genomes 64 > $o/${tag}_hash_ncu.log 2>&1
ls -la $o | grep ${tag}_
