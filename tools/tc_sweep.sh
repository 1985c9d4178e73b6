# Kernel-choice sweep of the tensor-core FIR (tools/tc_time.py per shape): writes gpurun_out/tc_sweep.jsonl and prints the table
# kept as profiles/r02/tc_f16_sweep.txt.  Run on a B200: bash tools/tc_sweep.sh
for T in 63 127 129 144 159 191 223 255 264; do for L in 26; do timeout 120 python tools/tc_time.py --D 8 --T $T --log2n $L --reps 10 >> gpurun_out/tc_sweep.jsonl; done; done
for L in 18 20 21 22 23 24; do timeout 120 python tools/tc_time.py --D 8 --T 255 --log2n $L --reps 20 >> gpurun_out/tc_sweep.jsonl; done
timeout 120 python tools/tc_time.py --D 8 --T 255 --log2n 20 --channels 64 --reps 10 >> gpurun_out/tc_sweep.jsonl
timeout 120 python tools/tc_time.py --D 8 --T 255 --log2n 16 --channels 1024 --reps 10 >> gpurun_out/tc_sweep.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/tc_sweep.jsonl'):
    r=json.loads(l); print(r['D'],r['T'],r['n_in'],r['channels'],'ffma2 %.4f tc %.4f ratio %.3f diff %.2e'%(r['ffma2']['ms'],r['tensor_core']['ms'],r['ffma2']['ms']/r['tensor_core']['ms'],r['max_abs_diff_tc_vs_ffma2']))
PY
