set -x
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests/test_fir_gpu.py -x -q -m gpu -k "nco or tma_kernel or shards or capturable" 2>&1 | tail -5
timeout 300 python tools/sweep.py --D 32 --T 1023 --log2n 28 --nco > gpurun_out/r2a/sweep_nco_d32.jsonl 2>&1
timeout 300 python tools/sweep.py --D 32 --T 1023 --log2n 28 --split > gpurun_out/r2a/sweep_fc_d32.jsonl 2>&1
timeout 300 python tools/sweep.py --D 10 --T 255 --log2n 28 --nco > gpurun_out/r2a/sweep_nco_d10.jsonl 2>&1
timeout 300 python tools/sweep.py --D 8 --T 255 --log2n 26 --nco > gpurun_out/r2a/sweep_nco_d8.jsonl 2>&1
timeout 200 python bench.py --workload cfg3 --no-cpu --steps 20 > gpurun_out/r2a/bench_cfg3.json 2> gpurun_out/r2a/bench_cfg3.err
timeout 200 python bench.py --workload cfg2 --no-cpu --steps 20 > gpurun_out/r2a/bench_cfg2.json 2> gpurun_out/r2a/bench_cfg2.err
python tools/one_launch.py --D 32 --T 1023 --log2n 27 --nco > gpurun_out/r2a/plain_nco.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:firTma -s 2 -c 1 -o gpurun_out/r2a/nco_d32 python tools/one_launch.py --D 32 --T 1023 --log2n 27 --nco > gpurun_out/r2a/ncu_nco.log 2>&1
python tools/one_launch.py --D 32 --T 1023 --log2n 27 > gpurun_out/r2a/plain_fc.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:firTma -s 2 -c 1 -o gpurun_out/r2a/fc_d32 python tools/one_launch.py --D 32 --T 1023 --log2n 27 > gpurun_out/r2a/ncu_fc.log 2>&1
cat gpurun_out/r2a/bench_cfg3.json gpurun_out/r2a/bench_cfg2.json | cut -c1-400
