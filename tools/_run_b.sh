timeout 300 ncu --set full --clock-control none --import-source on -k regex:hash_kernel -s 1 -c 1 -o gpurun_out/r02b_hash -f python bench.py --sketch-only --sketch-genomes 64 > gpurun_out/r02b_hash_ncu.log 2>&1
tail -2 gpurun_out/r02b_hash_ncu.log
