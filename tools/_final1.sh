set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 4
timeout 300 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -n 2
timeout 900 python bench.py --impl reference > gpurun_out/bench_n1_reference_final.json 2> gpurun_out/bench_n1_reference_final.err
timeout 1200 python bench.py > gpurun_out/bench_n1_ours_final.json 2> gpurun_out/bench_n1_ours_final.err
tail -c 300 gpurun_out/bench_n1_ours_final.err
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-others --no-e2e > /dev/null 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:firTc -c 1 --launch-skip 4 -o gpurun_out/r02_cfg2_tensor_core -f python bench.py --steps 5 --warmup 3 --no-cpu --no-others --no-e2e > gpurun_out/ncu_tc.log 2>&1
bash tools/_sweep_tc.sh > gpurun_out/tc_f16_sweep.txt 2>&1
