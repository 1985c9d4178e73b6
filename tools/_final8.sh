cd $GRAFT_REPO_ROOT
run() { n=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n "$@"; }
run 8 > gpurun_out/bench_n8_final.json 2> gpurun_out/bench_n8_final.err
tail -c 400 gpurun_out/bench_n8_final.err
run 4 --no-others > gpurun_out/bench_n4_final.json 2> gpurun_out/bench_n4_final.err
run 2 --no-others > gpurun_out/bench_n2_final.json 2> gpurun_out/bench_n2_final.err
timeout 600 python -m pytest tests/test_host_gpu.py tests/test_tc_gpu.py -m gpu -q 2>&1 | tail -n 4
timeout 600 python tools/multigpu_bench.py > gpurun_out/multigpu_bench_final.json 2> gpurun_out/multigpu_bench_final.err; tail -c 300 gpurun_out/multigpu_bench_final.err
