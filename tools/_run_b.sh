timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_host_cli.py -m gpu -q -x 2>&1 | tail -3
SKB_TRACE_SKETCH=1 timeout 600 python tools/sketch_scale.py 2048 1,2 > gpurun_out/r02_sketch_scale.json 2> gpurun_out/r02_sketch_scale.log; cat gpurun_out/r02_sketch_scale.json; grep -v "files " gpurun_out/r02_sketch_scale.log
