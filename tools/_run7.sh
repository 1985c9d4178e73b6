mkdir -p gpurun_out/r2g
GSDR_TC_DEBUG=1 timeout 120 python tools/tc_time.py > gpurun_out/r2g/tc_time_cfg2.json 2>gpurun_out/r2g/tc_time_cfg2.err; cat gpurun_out/r2g/tc_time_cfg2.json; tail -n 3 gpurun_out/r2g/tc_time_cfg2.err
GSDR_TC_DEBUG=1 timeout 200 python tools/tc_time.py --D 4 --T 127 --log2n 22 --channels 256 --reps 5 > gpurun_out/r2g/tc_time_cfg4.json 2>gpurun_out/r2g/tc_time_cfg4.err; cat gpurun_out/r2g/tc_time_cfg4.json; tail -n 3 gpurun_out/r2g/tc_time_cfg4.err
timeout 300 python -m pytest tests/test_tc_gpu.py -x -q -m gpu 2>&1 | tail -n 3
