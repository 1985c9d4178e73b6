timeout 600 python -m pytest tests/test_channelizer_gpu.py tests/test_tc_gpu.py -x -q -m gpu 2>&1 | tail -n 5
