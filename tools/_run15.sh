mkdir -p gpurun_out/r2o
timeout 600 python -m pytest tests/test_channelizer_gpu.py -x -q -m gpu 2>&1 | tail -n 12
timeout 400 python tools/chan_time.py > gpurun_out/r2o/channelizer.jsonl 2> gpurun_out/r2o/channelizer.err; cat gpurun_out/r2o/channelizer.jsonl | cut -c1-260; tail -n 3 gpurun_out/r2o/channelizer.err
