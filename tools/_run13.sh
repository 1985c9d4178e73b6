mkdir -p gpurun_out/r2m
timeout 300 python tools/sweep.py --D 4 --T 127 --log2n 26 --nco > gpurun_out/r2m/sweep_nco_d4.jsonl 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/r2m/sweep_nco_d4.jsonl"):
    try: d=json.loads(l)
    except: continue
    if "skipped" in d: continue
    print(d["variant"], d["threads"], d["smem"], "ms=%.4f"%d["ms_median"])
PY
