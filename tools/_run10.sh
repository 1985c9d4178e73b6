mkdir -p gpurun_out/r2j
timeout 600 python -m pytest tests/test_fir_gpu.py -x -q -m gpu -k "wide_row or nco or shards or tma_kernel" 2>&1 | tail -n 6
timeout 300 python tools/sweep.py --D 32 --T 1023 --log2n 28 --nco > gpurun_out/r2j/sweep_nco_d32.jsonl 2>&1
timeout 300 python tools/sweep.py --D 32 --T 1023 --log2n 28 > gpurun_out/r2j/sweep_fc_d32.jsonl 2>&1
grep -h '"variant": \(4[89]\|50\|22\|25\|-1\)' gpurun_out/r2j/sweep_nco_d32.jsonl gpurun_out/r2j/sweep_fc_d32.jsonl | cut -c1-330
