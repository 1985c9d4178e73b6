mkdir -p gpurun_out/r2p
python tools/one_launch.py --D 10 --T 255 --log2n 27 --nco > gpurun_out/r2p/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:firTma -s 2 -c 1 -o gpurun_out/r2p/nco_d10 python tools/one_launch.py --D 10 --T 255 --log2n 27 --nco > gpurun_out/r2p/ncu.log 2>&1
python tools/one_launch.py --D 8 --T 255 --log2n 26 --nco > gpurun_out/r2p/plain8.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:firTma -s 2 -c 1 -o gpurun_out/r2p/nco_d8 python tools/one_launch.py --D 8 --T 255 --log2n 26 --nco > gpurun_out/r2p/ncu8.log 2>&1
tail -n 1 gpurun_out/r2p/ncu.log gpurun_out/r2p/ncu8.log
