mkdir -p gpurun_out/r2r
timeout 900 python -m pytest tests -x -q -m gpu -k "nco or stream or shard or demod or channel or wide or host or int8 or fuzz" 2>&1 | tail -n 6
for d in "10 255 28" "8 255 27" "32 1023 28" "4 127 26"; do set -- $d; timeout 200 python tools/demod_time.py --D $1 --T $2 --log2n $3 >> gpurun_out/r2r/demod_time.jsonl 2>>gpurun_out/r2r/demod_time.err; done
cut -c1-420 gpurun_out/r2r/demod_time.jsonl
timeout 300 python bench.py --workload cfg5 --no-cpu --no-others --steps 20 2>/dev/null | cut -c1-200
