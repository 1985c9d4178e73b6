#!/usr/bin/env python
"""Host-side cost of one entry-point call (tiny problem, stream kept busy by nothing): wall time per call over many
calls, against a no-op ctypes call for scale."""
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import gsdr_b200 as g
from gsdr_b200 import synth
dev = torch.device("cuda:0")
D, T, n_out = 8, 255, 4096
n_in = (n_out - 1) * D + T
x = synth.tone_plus_noise(0, n_in, seed=1, device=dev)
taps = torch.from_numpy(synth.lowpass_taps(T, D)).to(dev)
y = torch.zeros(n_out, dtype=torch.complex64, device=dev)
s = torch.cuda.Stream()
for name, fn in (("gsdrFirFC", lambda: g.gsdrFirFC(D, taps, T, x, y, n_out, 0, s)),
                 ("gsdrAdjustFrequencyFirFC", lambda: g.gsdrAdjustFrequencyFirFC(2.4e6, 1e5, 0, D, taps, T, x, y, n_out, 0, s)),
                 ("gsdrFirNumOutputs (no CUDA)", lambda: g.fir_num_outputs(n_in, T, D))):
    for _ in range(200):
        fn()
    s.synchronize()
    t0 = time.perf_counter()
    N = 5000
    for _ in range(N):
        fn()
    t1 = time.perf_counter()
    s.synchronize()
    t2 = time.perf_counter()
    print(f"{name}: {1e6 * (t1 - t0) / N:.2f} us per call on the host, {1e6 * (t2 - t0) / N:.2f} us per call including the GPU drain")
