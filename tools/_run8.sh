mkdir -p gpurun_out/r2h
timeout 300 python -m pytest tests/test_tc_gpu.py -x -q -m gpu 2>&1 | tail -n 3
timeout 120 python tools/tc_time.py > gpurun_out/r2h/tc_time_cfg2.json 2>/dev/null; cat gpurun_out/r2h/tc_time_cfg2.json
timeout 120 python tools/tc_time.py --D 16 --T 511 --log2n 26 > gpurun_out/r2h/tc_time_d16.json 2>/dev/null; cat gpurun_out/r2h/tc_time_d16.json
python tools/one_launch.py --D 8 --T 255 --log2n 26 --variant -4 > gpurun_out/r2h/plain_tc.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:firTc -s 2 -c 1 -o gpurun_out/r2h/tc_d8 python tools/one_launch.py --D 8 --T 255 --log2n 26 --variant -4 > gpurun_out/r2h/ncu_tc.log 2>&1
tail -n 2 gpurun_out/r2h/ncu_tc.log
