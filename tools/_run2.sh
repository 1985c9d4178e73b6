set -x
mkdir -p gpurun_out/r2b
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2b/bench_ref.json 2> gpurun_out/r2b/bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2b/bench_ours.json 2> gpurun_out/r2b/bench_ours.err
tail -3 gpurun_out/r2b/bench_ours.err
cut -c1-600 gpurun_out/r2b/bench_ref.json
cut -c1-300 gpurun_out/r2b/bench_ours.json
