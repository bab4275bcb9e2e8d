AB_ARGS="--no-extras --pass-reads 8192" bash tools/ab.sh base 2>&1 | tee -a gpurun_out/r02_r_ab.txt
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size or large_scale" 2>&1 | tail -3
