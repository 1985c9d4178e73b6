#!/usr/bin/env python
"""What the fused NCO costs per shape: gsdrFirFC against gsdrAdjustFrequencyFirFC on the same buffers."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import gsdr_b200 as g  # noqa: E402
from gsdr_b200 import synth  # noqa: E402

dev = torch.device("cuda:0")
stream = torch.cuda.Stream()
for D, T, log2n in ((10, 255, 28), (32, 1023, 28), (8, 255, 28), (4, 127, 28), (16, 511, 28)):
    n_in = 1 << log2n
    n_out = g.fir_num_outputs(n_in, T, D)
    x = synth.tone_plus_noise(0, n_in, seed=1, device=dev)
    taps = torch.from_numpy(synth.lowpass_taps(T, D)).to(dev)
    y = torch.zeros(n_out, dtype=torch.complex64, device=dev)
    res = {"D": D, "T": T, "n_in": n_in}
    for name, fn in (("fir", lambda: g.gsdrFirFC(D, taps, T, x, y, n_out, 0, stream)),
                     ("nco_fir", lambda: g.gsdrAdjustFrequencyFirFC(2.4e6, 29520.0, 77, D, taps, T, x, y, n_out, 0, stream))):
        for _ in range(3):
            fn()
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10):
            fn()
        e1.record(stream)
        stream.synchronize()
        res[name + "_ms"] = e0.elapsed_time(e1) / 10
        res[name + "_variant"] = g.describe_kernel(4 if name == "nco_fir" else 0, D, T, n_out).variant
    res["nco_cost_ms"] = res["nco_fir_ms"] - res["fir_ms"]
    print(json.dumps(res))
    del x, y
