python -m pytest tests -m gpu -x -q 2>&1 | tail -5
AB_ARGS="--no-extras" bash tools/ab.sh base 2>&1 | tee gpurun_out/r02_u_ab.txt
AB_ARGS="--no-extras --pass-reads 8192" bash tools/ab.sh base 2>&1 | tee -a gpurun_out/r02_u_ab.txt
