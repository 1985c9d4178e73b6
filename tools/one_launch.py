#!/usr/bin/env python
"""A short hot-path run for ncu: a few launches of gsdrFirFC (or the fused NCO) on a BASELINE workload."""
import argparse
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
ap.add_argument("--variant", type=int, default=-1)
ap.add_argument("--launches", type=int, default=4)
ap.add_argument("--nco", action="store_true")
ap.add_argument("--kind", default="fc")
ap.add_argument("--dbg", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda:0")
real = a.kind in ("ff", "cf")
n_in = 1 << a.log2n
n_out = g.fir_num_outputs(n_in, a.T, a.D)
x = synth.tone_plus_noise(0, n_in, seed=1, device=dev, real=real)
ctaps = a.kind in ("cc", "cf")
taps = torch.from_numpy(synth.random_taps(a.T, 3, complex_taps=True) if ctaps else synth.lowpass_taps(a.T, a.D)).to(dev)
y = torch.zeros(n_out, dtype=torch.float32 if a.kind == "ff" else torch.complex64, device=dev)
if a.kind == "i8":
    x = torch.view_as_real(x).mul(127.0).round().clamp(-127, 127).to(torch.int8).reshape(-1).contiguous()
g.set_kernel_variant(a.variant)
g.set_debug_flags(a.dbg)
for _ in range(a.launches):
    if a.kind == "i8":
        (g.gsdrAdjustFrequencyFirFCInt8(2.4e6, 29520.0, 0, a.D, taps, a.T, x, y, n_out, 0, None) if a.nco
         else g.gsdrFirFCInt8(a.D, taps, a.T, x, y, n_out, 0, None))
    elif a.nco:
        g.gsdrAdjustFrequencyFirFC(2.4e6, 29520.0, 0, a.D, taps, a.T, x, y, n_out, 0, None)
    elif a.kind == "cc":
        g.gsdrFirCC(a.D, taps, a.T, x, y, n_out, 0, None)
    elif a.kind == "cf":
        g.gsdrFirCF(a.D, taps, a.T, x, y, n_out, 0, None)
    elif real:
        g.gsdrFirFF(a.D, taps, a.T, x, y, n_out, 0, None)
    else:
        g.gsdrFirFC(a.D, taps, a.T, x, y, n_out, 0, None)
torch.cuda.synchronize()
print("ok", float(y.abs().sum()))
