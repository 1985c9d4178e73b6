#!/usr/bin/env python
"""Times gsdrChannelizeFC (one launch, window fetched once per tile) against K separate gsdrAdjustFrequencyFirFC calls
on the same input.  One JSON line per shape."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import gsdr_b200 as g  # noqa: E402
from gsdr_b200 import synth  # noqa: E402

dev = torch.device("cuda:0")
s = torch.cuda.Stream()


def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(reps):
        fn()
    e1.record(s)
    s.synchronize()
    return e0.elapsed_time(e1) / reps


for D, T, log2n in ((8, 255, 26), (4, 127, 26), (10, 255, 26), (8, 63, 26), (8, 31, 26), (4, 15, 26)):
    n_in = 1 << log2n
    n_out = g.fir_num_outputs(n_in, T, D)
    x = synth.tone_plus_noise(0, n_in, seed=3, device=dev)
    taps = torch.from_numpy(synth.lowpass_taps(T, D)).to(dev)
    for K in (4, 8, 16):
        shifts = [1.0e5 * (k - K / 2 + 0.5) for k in range(K)]
        out = torch.zeros((K, n_out), dtype=torch.complex64, device=dev)
        g.set_kernel_variant(-5)  # the fused kernel lives in the tuning build
        fused = timeit(lambda: g.gsdrChannelizeFC(2.4e6, shifts, 0, D, taps, T, x, out, n_out, n_out, 0, s))
        g.set_kernel_variant(-1)
        ref = out.clone()

        def loop():
            for k in range(K):
                g.gsdrAdjustFrequencyFirFC(2.4e6, shifts[k], 0, D, taps, T, x, out[k], n_out, 0, s)
        sep = timeit(loop)
        print(json.dumps({"D": D, "T": T, "n_in": n_in, "K": K, "channelizer_ms": fused, "separate_calls_ms": sep,
                          "speedup": sep / fused, "max_abs_diff": float((out - ref).abs().max()),
                          "input_gb_s_fused": 8 * n_in / fused / 1e6}), flush=True)
