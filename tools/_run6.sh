mkdir -p gpurun_out/r2f
python tools/one_launch.py --D 8 --T 255 --log2n 26 --variant -4 > gpurun_out/r2f/plain_tc.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:firTc -s 2 -c 1 -o gpurun_out/r2f/tc_d8 python tools/one_launch.py --D 8 --T 255 --log2n 26 --variant -4 > gpurun_out/r2f/ncu_tc.log 2>&1
tail -3 gpurun_out/r2f/plain_tc.log gpurun_out/r2f/ncu_tc.log
