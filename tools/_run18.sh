mkdir -p gpurun_out/r2q
timeout 300 python tools/sweep.py --D 8 --T 255 --log2n 26 --split > gpurun_out/r2q/sweep_fc_d8.jsonl 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/r2q/sweep_fc_d8.jsonl"):
    try: d=json.loads(l)
    except: continue
    if "skipped" in d: continue
    print(d["variant"], d["threads"], d["smem"], "ms=%.4f"%d["ms_median"], {k:round(v,4) for k,v in d.items() if k.startswith("ms_") and k not in ("ms_median","ms_best")})
PY
