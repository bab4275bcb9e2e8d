bash tools/ab.sh base w24 w32 abl2 abl6 abl4 2>&1 | tee gpurun_out/r02_b_ab.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 40 -c 1 -o gpurun_out/r02_b_fused -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sketch-genomes 0 > gpurun_out/r02_b_ncu.log 2>&1
tail -3 gpurun_out/r02_b_ncu.log
