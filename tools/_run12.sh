mkdir -p gpurun_out/r2l
timeout 900 python -m pytest tests/test_demod_gpu.py tests/test_abi.py -x -q -m gpu 2>&1 | tail -n 4
for d in "10 255 28" "8 255 27" "32 1023 28" "4 127 26"; do set -- $d; timeout 200 python tools/demod_time.py --D $1 --T $2 --log2n $3 >> gpurun_out/r2l/demod_time.jsonl 2>>gpurun_out/r2l/demod_time.err; done
cat gpurun_out/r2l/demod_time.jsonl; tail -n 3 gpurun_out/r2l/demod_time.err
