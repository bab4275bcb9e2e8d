timeout 600 python tools/sketch_scale.py 256 1,2 > gpurun_out/r02_sketch_scale.json 2> gpurun_out/r02_sketch_scale.log; cat gpurun_out/r02_sketch_scale.json; tail -3 gpurun_out/r02_sketch_scale.log
timeout 600 python -m pytest tests/test_host_cli.py tests/test_multi_gpu.py -m gpu -q -x 2>&1 | tail -2
