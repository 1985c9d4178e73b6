#!/usr/bin/env python
"""Times the fused output stages against the kernel chains they replace:
FM: gsdrFmDemodFused vs gsdrFmDemod (fused NCO + FIR, then the demodulator kernel);
AM: gsdrAmDemod (envelope in the FIR's store path) vs gsdrAdjustFrequencyFirFC + gsdrQuadAmDemod."""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import gsdr_b200 as g  # noqa: E402
from gsdr_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--D", type=int, default=10)
ap.add_argument("--T", type=int, default=255)
ap.add_argument("--log2n", type=int, default=28)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda:0")
n_in = 1 << a.log2n
n_out = (n_in - a.T) // a.D  # one fewer than the FIR could give: the FM stage needs numOutputs + 1 low-pass values
x = synth.tone_plus_noise(0, n_in, seed=5, device=dev, tone_cycles_per_sample=0.125)
taps = torch.from_numpy(synth.lowpass_taps(a.T, a.D)).to(dev)
y = torch.zeros(n_out, dtype=torch.float32, device=dev)
lp = torch.zeros(n_out + 1, dtype=torch.complex64, device=dev)
s = torch.cuda.Stream()
fs, tun, ch, dv = 2.4e6, 100.0e6, 100.3e6, 75e3


def timeit(fn):
    for _ in range(3):
        fn()
    s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(a.reps):
        fn()
    e1.record(s)
    s.synchronize()
    return e0.elapsed_time(e1) / a.reps


res = {"D": a.D, "T": a.T, "n_in": n_in}
res["fm_two_kernels_ms"] = timeit(lambda: g.gsdrFmDemod(fs, tun, ch, dv, a.D, 0, taps, a.T, x, y, n_out, 0, s))
ref = y.clone()
res["fm_fused_ms"] = timeit(lambda: g.gsdrFmDemodFused(fs, tun, ch, dv, a.D, 0, taps, a.T, x, y, n_out, 0, s))
res["fm_max_abs_diff"] = float((y - ref).abs().max())
res["nco_fir_only_ms"] = timeit(lambda: g.gsdrAdjustFrequencyFirFC(fs, tun - ch, 0, a.D, taps, a.T, x, lp, n_out + 1, 0, s))
res["quad_fm_kernel_only_ms"] = timeit(lambda: g.gsdrQuadFmDemod(lp, y, 5.0, n_out, 0, s))


def am_chain():
    g.gsdrAdjustFrequencyFirFC(fs, tun - ch, 0, a.D, taps, a.T, x, lp, n_out, 0, s)
    g.gsdrQuadAmDemod(lp, y, n_out, 0, s)


res["am_two_kernels_ms"] = timeit(am_chain)
ref = y.clone()
res["am_fused_ms"] = timeit(lambda: g.gsdrAmDemod(fs, tun, ch, a.D, 0, taps, a.T, x, y, n_out, 0, s))
res["am_max_abs_diff"] = float((y - ref).abs().max())
print(json.dumps(res))
