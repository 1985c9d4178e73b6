mkdir -p gpurun_out/r2n
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 8
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2n/bench_ours.json 2> gpurun_out/r2n/bench_ours.err; tail -n 2 gpurun_out/r2n/bench_ours.err; cut -c1-200 gpurun_out/r2n/bench_ours.json
