set -x
mkdir -p gpurun_out/r2c
nvidia-smi -L
timeout 600 python -m pytest tests/test_host_gpu.py -x -q -m gpu 2>&1 | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2c/bench_n2.json 2> gpurun_out/r2c/bench_n2.err
tail -5 gpurun_out/r2c/bench_n2.err
cut -c1-300 gpurun_out/r2c/bench_n2.json
