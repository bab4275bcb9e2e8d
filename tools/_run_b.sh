timeout 600 python tools/predict_cli_rate.py 100000 10 > gpurun_out/r02_predict_cli.json 2> gpurun_out/r02_predict_cli.log; cat gpurun_out/r02_predict_cli.json; tail -5 gpurun_out/r02_predict_cli.log
