timeout 300 python bench.py --no-others --no-cpu --no-e2e --steps 50 2>/dev/null | cut -c1-250
timeout 300 python bench.py --no-others --no-cpu --no-e2e --steps 50 2>/dev/null | cut -c150-250
timeout 600 python -m pytest tests/test_fir_gpu.py -x -q -m gpu -k "shards or every_kernel or config2 or batched or reference" 2>&1 | tail -n 3
