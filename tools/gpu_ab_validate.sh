#!/bin/bash
# One GPU-box call: A/B the fused-kernel build variants on the C3 bench, pick the fastest one whose answer checksum
# equals the base build's, then run the whole GPU test suite, a full bench line and one `ncu --set full` capture of the fused kernel with that library
# (the per-launch list costs minutes under ncu: it is a separate call, see profiles/r01_launches_summary.md).
# usage (on the box): bash tools/gpu_ab_validate.sh base w24 w24r6 ...      outputs under gpurun_out/ab/
out=gpurun_out/ab; mkdir -p $out
bash tools/ab.sh "$@" 2>&1 | tee $out/ab.txt
win=$(python - <<'PY'
import re
rows = {}
for l in open("gpurun_out/ab/ab.txt"):
    m = re.match(r"(\S+) value (\d+) .* crc (\w+)", l)
    if m: rows[m.group(1)] = (int(m.group(2)), m.group(3))
base = rows.get("base")
best = "base"
if base:
    for k, (v, c) in rows.items():
        if c == base[1] and v > rows[best][0] * 1.02: best = k
print(best)
PY
)
echo "winner: $win" | tee $out/winner.txt
if [ "$win" != base ]; then export SKB_LIB=$PWD/sketchy_b200/build/variants/lib_$win.so; fi
timeout 240 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $out/pytest.txt
timeout 120 python bench.py --steps 5 --warmup 3 > $out/bench_full.json 2> $out/bench_full.log; tail -c 600 $out/bench_full.json
timeout 150 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 40 -c 1 -o $out/fused_full -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sketch-genomes 0 > $out/fused_full.log 2>&1
ls -la $out
