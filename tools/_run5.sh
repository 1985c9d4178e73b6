mkdir -p gpurun_out/r2e
timeout 300 python -m pytest tests/test_tc_gpu.py -x -q -m gpu 2>&1 | tail -25
timeout 120 python tools/tc_time.py > gpurun_out/r2e/tc_time_cfg2.json 2>gpurun_out/r2e/tc_time_cfg2.err; cat gpurun_out/r2e/tc_time_cfg2.json; tail -3 gpurun_out/r2e/tc_time_cfg2.err
timeout 200 python tools/tc_time.py --D 4 --T 127 --log2n 22 --channels 256 --reps 5 > gpurun_out/r2e/tc_time_cfg4.json 2>gpurun_out/r2e/tc_time_cfg4.err; cat gpurun_out/r2e/tc_time_cfg4.json; tail -3 gpurun_out/r2e/tc_time_cfg4.err
