set -x
mkdir -p gpurun_out/r2i
nvidia-smi -L | wc -l
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2i/bench_n$n.json 2> gpurun_out/r2i/bench_n$n.err
tail -n 3 gpurun_out/r2i/bench_n$n.err
cut -c1-200 gpurun_out/r2i/bench_n$n.json
done
timeout 300 python -m pytest tests/test_host_gpu.py -x -q -m gpu 2>&1 | tail -n 3
