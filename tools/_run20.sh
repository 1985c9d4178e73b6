mkdir -p gpurun_out/r2s
timeout 300 python tools/sweep.py --D 10 --T 255 --log2n 28 --nco > gpurun_out/r2s/sweep_nco_d10.jsonl 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/r2s/sweep_nco_d10.jsonl"):
    try: d=json.loads(l)
    except: continue
    if "skipped" in d: continue
    print(d["variant"], d["threads"], d["smem"], "ms=%.4f"%d["ms_median"], "maxdiff", d["maxdiff_vs_first"])
PY
timeout 600 python -m pytest tests/test_fir_gpu.py -x -q -m gpu -k "every_kernel_variant or tma_kernel or nco" 2>&1 | tail -n 4
