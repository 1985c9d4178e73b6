"""Seeded random shapes through the automatic kernel choice of every entry point, against the C oracle.  The point is
the dispatch edges: tiny output counts (fewer tiles than CTAs), tapCount below the decimation, windows that end
exactly at the input's end, odd output counts on the pair kernels, 8-byte-aligned pointers that force the fallbacks."""
import os
import random

import numpy as np
import pytest
import torch

import gsdr_b200 as g
from gsdr_b200 import synth
from oracle import oracle

pytestmark = pytest.mark.gpu

KINDS = ["fc", "ff", "cc", "cf", "nco", "i8", "i8nco"]


def _case(seed):
    rng = random.Random(seed)
    kind = KINDS[seed % len(KINDS)]
    D = rng.choice([1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 32, 48, 64, rng.randrange(1, 90)])
    T = rng.choice([1, 2, D, D + 1, 8 * D, 16 * D, 16 * D + 1, 63, 255, rng.randrange(1, 1200)])
    n_out = rng.choice([1, 2, 7, 255, 256, 257, 2049, rng.randrange(1, 30000)])
    offset = rng.choice([0, 0, 0, 1, 2, 3])  # elements skipped at the front of the input allocation
    return kind, D, T, n_out, offset


@pytest.mark.timeout(120)
@pytest.mark.parametrize("seed", list(range(int(os.environ.get("GSDR_FUZZ_CASES", "84")))))
def test_random_shape(seed, cuda_device):
    kind, D, T, n_out, offset = _case(seed)
    fs, f, first = 1.0e6, -123456.0, 2 ** 33 + seed
    complex_taps = kind in ("cc", "cf")
    real_in = kind in ("ff", "cf")
    taps = synth.random_taps(T, 1000 + seed, complex_taps=complex_taps)
    n_in = (n_out - 1) * D + T
    dt = torch.from_numpy(taps).to(cuda_device)
    out_dtype = torch.float32 if kind == "ff" else torch.complex64
    dy = torch.full((n_out + 3,), 2.0, dtype=out_dtype, device=cuda_device)
    if kind in ("i8", "i8nco"):
        rng = np.random.default_rng(seed)
        iq = rng.integers(-128, 128, size=2 * (n_in + offset), dtype=np.int64).astype(np.int8)
        f32 = np.maximum(np.float32(-1.0), iq.astype(np.float32) / np.float32(127.0))
        x = (f32[0::2] + 1j * f32[1::2]).astype(np.complex64)[offset:]
        di = torch.from_numpy(iq).to(cuda_device)[2 * offset:]
        if kind == "i8":
            g.gsdrFirFCInt8(D, dt, T, di, dy, n_out, 0, None)
        else:
            g.gsdrAdjustFrequencyFirFCInt8(fs, f, first, D, dt, T, di, dy, n_out, 0, None)
    else:
        xa = synth.tone_plus_noise(0, n_in + offset, seed=2000 + seed, real=real_in)
        x = xa[offset:]
        dx = torch.from_numpy(xa).to(cuda_device)[offset:]
        if kind == "nco":
            g.gsdrAdjustFrequencyFirFC(fs, f, first, D, dt, T, dx, dy, n_out, 0, None)
        else:
            {"fc": g.gsdrFirFC, "ff": g.gsdrFirFF, "cc": g.gsdrFirCC, "cf": g.gsdrFirCF}[kind](D, dt, T, dx, dy, n_out, 0, None)
    torch.cuda.synchronize()
    y = dy.cpu().numpy()
    assert (y[n_out:] == 2.0).all(), "wrote past the last output"
    n_chk = min(n_out, 2500)
    if kind in ("nco", "i8nco"):
        want = oracle.adjust_frequency_fir_fc(oracle.NCO_EXACT, fs, f, first, D, taps, x, n_chk, f64=True)
    else:
        okind = "fc" if kind == "i8" else kind
        want = oracle.fir(okind, D, taps, x, n_chk, f64=True)
    tol = 1e-5 * float(np.abs(taps).sum()) * max(float(np.abs(x).max()), 1e-30)
    err = float(np.abs(y[:n_chk] - want).max())
    assert err <= tol, f"{kind} D={D} T={T} n_out={n_out} offset={offset}: max|err| {err} > {tol}"
    assert np.isfinite(y[:n_out].view(np.float32)).all()
