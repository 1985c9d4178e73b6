set -x
cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --impl reference > gpurun_out/bench_n1_reference_final.json 2> gpurun_out/bench_n1_reference_final.err
timeout 1200 python bench.py > gpurun_out/bench_n1_ours_final.json 2> gpurun_out/bench_n1_ours_final.err
tail -c 300 gpurun_out/bench_n1_ours_final.err
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 120 python tools/tc_time.py --D 8 --T 255 --reps 2 > /dev/null && timeout 600 ncu --set full --clock-control none --import-source on -k regex:firTc -c 1 -o gpurun_out/r02_cfg2_tensor_core -f python tools/tc_time.py --D 8 --T 255 --reps 1 > gpurun_out/ncu_tc.log 2>&1
bash tools/_sweep_tc.sh > gpurun_out/tc_f16_sweep.txt 2>&1
