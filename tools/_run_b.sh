python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
AB_ARGS="--no-extras" bash tools/ab.sh base b8k 2>&1 | tee gpurun_out/r02_o_ab.txt
SKB_LIB=$PWD/sketchy_b200/build/variants/lib_b8k.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
