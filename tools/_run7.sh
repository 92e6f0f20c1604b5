timeout 1500 python tools/cli_config1.py > gpurun_out/cli_config1_r01.json 2> gpurun_out/cli_config1_r01.err; cat gpurun_out/cli_config1_r01.json; tail -5 gpurun_out/cli_config1_r01.err
