mkdir -p gpurun_out/r2k
timeout 900 python -m pytest tests/test_demod_gpu.py tests/test_host_gpu.py -x -q -m gpu 2>&1 | tail -n 15
timeout 300 python bench.py --workload cfg5 --no-cpu --no-others --steps 20 > gpurun_out/r2k/bench_cfg5.json 2> gpurun_out/r2k/bench_cfg5.err; cut -c1-260 gpurun_out/r2k/bench_cfg5.json; tail -n 3 gpurun_out/r2k/bench_cfg5.err
