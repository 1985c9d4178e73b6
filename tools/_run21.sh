mkdir -p gpurun_out/r2t
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2t/bench_plain.json 2> gpurun_out/r2t/bench_plain.err && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2t/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2t/ncu_launches.log 2>&1
tail -n 2 gpurun_out/r2t/ncu_launches.log | cut -c1-200
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-others > gpurun_out/r2t/bench_short.json 2>/dev/null && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:firTmaKernel -s 3 -c 1 -o gpurun_out/r2t/cfg2_headline python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-others > gpurun_out/r2t/ncu_full.log 2>&1
tail -n 2 gpurun_out/r2t/ncu_full.log | cut -c1-200
timeout 300 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu --no-e2e --no-others > gpurun_out/r2t/bench_cfg3_short.json 2>/dev/null && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:firTmaWideKernel -s 3 -c 1 -o gpurun_out/r2t/cfg3_wide_nco python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu --no-e2e --no-others > gpurun_out/r2t/ncu_full3.log 2>&1
tail -n 2 gpurun_out/r2t/ncu_full3.log | cut -c1-200
python tools/call_overhead.py > gpurun_out/r2t/call_overhead.txt 2>&1; tail -n 6 gpurun_out/r2t/call_overhead.txt
