mkdir -p gpurun_out/r2d
timeout 120 ./tools/tc_probe > gpurun_out/r2d/tc_probe.log 2>&1; echo "rc=$?" >> gpurun_out/r2d/tc_probe.log
cat gpurun_out/r2d/tc_probe.log
