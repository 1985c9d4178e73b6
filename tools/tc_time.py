#!/usr/bin/env python
"""Times gsdrFirFC with the tensor-core kernel forced on / off (tuning build) on one shape."""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import gsdr_b200 as g  # noqa: E402
from gsdr_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--D", type=int, default=8)
ap.add_argument("--T", type=int, default=255)
ap.add_argument("--log2n", type=int, default=26)
ap.add_argument("--channels", type=int, default=1)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda:0")
n_in = 1 << a.log2n
n_out = g.fir_num_outputs(n_in, a.T, a.D)
C = a.channels
x = synth.tone_plus_noise(0, n_in * min(C, 8), seed=1, device=dev).view(min(C, 8), n_in)
if C > 8:
    x = x.repeat((C + 7) // 8, 1)[:C].contiguous()
taps = torch.from_numpy(synth.lowpass_taps(a.T, a.D)).to(dev)
y = torch.zeros((C, n_out), dtype=torch.complex64, device=dev)
stream = torch.cuda.Stream()
res = {}
outs = {}
for name, v in (("ffma2", -3), ("tensor_core", -4), ("default_release", -1)):
    g.set_kernel_variant(v)
    info = g.describe_kernel(0, a.D, a.T, n_out)
    def call():
        if C == 1:
            g.gsdrFirFC(a.D, taps, a.T, x[0], y[0], n_out, 0, stream)
        else:
            g.gsdrFirFCBatched(a.D, taps, a.T, 0, x, n_in, y, n_out, n_out, C, 0, stream)
    y.zero_()
    for _ in range(3):
        call()
    stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.reps):
        call()
    e1.record(stream)
    stream.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    outs[name] = y.clone()
    res[name] = {"variant": info.variant, "ms": ms, "msamples_s": C * n_in / ms / 1e3,
                 "hbm_gbs": C * (8 * n_in + 8 * n_out) / ms / 1e6}
res["max_abs_diff_tc_vs_ffma2"] = float((outs["tensor_core"] - outs["ffma2"]).abs().max())
res["default_equals_tc"] = bool(torch.equal(outs["default_release"], outs["tensor_core"]))
print(json.dumps({"D": a.D, "T": a.T, "n_in": n_in, "channels": C, **res}))
