#!/usr/bin/env python
"""Single-process multi-GPU executor (gsdrMultiGpu, include/gsdr/b200.h) on BASELINE config 2, weak scaling:
every device holds 2^26 samples (+ overlap) of an N * 2^26-sample capture.  Times, per N:
  compute        every device filters its shard into its own output buffer
  fused_gather   every device's kernel stores its block straight into the gather buffer on devices[0] (peer stores)
  gather_after   compute, then gsdrMultiGpuGather (cudaMemcpyPeerAsync, one copy per shard to its final offset)
and checks the gathered outputs of both forms against each other.  Run with `gpurun --gpus N`."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import gsdr_b200 as g  # noqa: E402
from gsdr_b200 import synth  # noqa: E402

D, T, n_gpu = 8, 255, 1 << 26
taps = synth.lowpass_taps(T, D)
avail = torch.cuda.device_count()
for n in (1, 2, 4, 8):
    if n > avail:
        break
    devs = list(range(n))
    n_out = g.fir_num_outputs(n_gpu * n, T, D)
    shards = [g.shard_plan_time(n_out, D, T, 0, n, s) for s in range(n)]
    xs = [synth.tone_plus_noise(sh.firstInput, sh.numInputs, seed=0x5EED0002, device=torch.device("cuda", d))
          for sh, d in zip(shards, devs)]
    tp = [torch.from_numpy(taps).to(torch.device("cuda", d)) for d in devs]
    outs = [torch.zeros(sh.numOutputs, dtype=torch.complex64, device=torch.device("cuda", d)) for sh, d in zip(shards, devs)]
    gathered = torch.zeros(n_out, dtype=torch.complex64, device="cuda:0")
    later = torch.zeros(n_out, dtype=torch.complex64, device="cuda:0")
    mg = g.MultiGpu(devs)
    peer = all(mg.peer_ok(i) for i in range(n))
    for _ in range(2):
        mg.gsdrFirFCMultiGpu(0.0, 0.0, 0, D, tp, T, xs, outs, None, n_out, repeats=3)
    compute_ms = mg.gsdrFirFCMultiGpu(0.0, 0.0, 0, D, tp, T, xs, outs, None, n_out, repeats=20)
    res = {"n_gpus": n, "input_samples_total": n_gpu * n, "compute_ms": compute_ms,
           "compute_msamples_s": n_gpu * n / compute_ms / 1e3, "peer_access": peer}
    if peer:
        mg.gsdrFirFCMultiGpu(0.0, 0.0, 0, D, tp, T, xs, None, gathered, n_out, repeats=3)
        fused_ms = mg.gsdrFirFCMultiGpu(0.0, 0.0, 0, D, tp, T, xs, None, gathered, n_out, repeats=20)
        mg.gsdrMultiGpuGather(D, T, outs, later, n_out)
        gather_ms = min(mg.gsdrMultiGpuGather(D, T, outs, later, n_out) for _ in range(5))
        torch.cuda.synchronize()
        res.update({"fused_gather_ms": fused_ms, "fused_gather_msamples_s": n_gpu * n / fused_ms / 1e3,
                    "gather_after_ms": gather_ms, "compute_plus_gather_after_ms": compute_ms + gather_ms,
                    "bytes_to_device0": 8 * (n_out - shards[0].numOutputs),
                    "fused_equals_gather_after": bool(torch.equal(gathered, later))})
    print(json.dumps(res), flush=True)
    mg.close()
    del xs, outs, gathered, later
    torch.cuda.empty_cache()
