python -m pytest tests -m gpu -x -q 2>&1 | tail -4
SKB_PIPELINE=0 bash tools/ab.sh base 2>&1 | tee gpurun_out/r02_j_ab.txt
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:'fused_kernel|dense_topk|merge_topn' -s 81 -c 8 --csv --log-file gpurun_out/r02_j_fused.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sketch-genomes 0 > gpurun_out/r02_j_ncu.log 2>&1
grep -E "dense_topk|merge_topn" gpurun_out/r02_j_fused.csv | cut -d, -f1,5,12- | head
