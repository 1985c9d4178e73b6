"""gsdrFirStream (include/gsdr/stream.h) on the GPU: blocks of any length in, and the concatenated outputs are the
outputs of one stateless call over the concatenated input — bit for bit when the blocks keep 16-byte alignment
(same kernel, same per-output summation order, NCO phase a pure function of the absolute sample index), within the
FP32 tolerance otherwise (an unaligned block falls back to the cp.async kernel, whose order differs)."""
import random

import numpy as np
import pytest
import torch

import gsdr_b200 as g
from gsdr_b200 import synth
from oracle import oracle

pytestmark = pytest.mark.gpu

FS, SHIFT, FIRST = 2.4e6, 29520.0, 2 ** 33 + 12345


def _tol(taps, x):
    return 1e-5 * float(np.abs(taps).sum()) * float(np.abs(x).max())


def _one_shot(kind, D, dt, T, dx, n_out, dev):
    y = torch.zeros(n_out, dtype=torch.float32 if kind == "ff" else torch.complex64, device=dev)
    if kind == "fc":
        g.gsdrFirFC(D, dt, T, dx, y, n_out, 0, None)
    elif kind == "ff":
        g.gsdrFirFF(D, dt, T, dx, y, n_out, 0, None)
    else:
        g.gsdrAdjustFrequencyFirFC(FS, SHIFT, FIRST, D, dt, T, dx, y, n_out, 0, None)
    torch.cuda.synchronize()
    return y


def _streamed(kind, D, dt, T, dx, blocks, dev):
    ftype = {"fc": g.FirStream.FC, "ff": g.FirStream.FF, "nco": g.FirStream.FC_NCO}[kind]
    st = g.FirStream(ftype, D, dt, T, FS, SHIFT, FIRST, 0)
    outs, pos = [], 0
    for n in blocks:
        want = st.num_outputs(n)
        y = torch.full((want + 4,), 3.0, dtype=torch.float32 if kind == "ff" else torch.complex64, device=dev)
        got = st.push(dx[pos:pos + n] if n else None, n, y, None)
        assert got == want
        torch.cuda.synchronize()
        assert (y[want:].real == 3.0).all(), "wrote past the announced output count"
        outs.append(y[:want].clone())
        pos += n
    st.close()
    return torch.cat(outs) if outs else torch.zeros(0, device=dev)


@pytest.mark.parametrize("kind", ["fc", "ff", "nco"])
@pytest.mark.parametrize("D,T", [(8, 255), (10, 255), (5, 63), (1, 63), (32, 1023), (7, 5), (4, 2)])
def test_aligned_blocks_reproduce_the_one_shot_bits(kind, D, T, cuda_device):
    rng = random.Random(7 * D + T)
    blocks = [rng.choice([4096, 8192, 65536, 16, 256, 4, 0, 20000]) for _ in range(24)]
    total = sum(blocks)
    taps = synth.random_taps(T, 11 + D)
    x = synth.tone_plus_noise(0, total, seed=200 + D, real=(kind == "ff"))
    dt, dx = torch.from_numpy(taps).to(cuda_device), torch.from_numpy(x).to(cuda_device)
    n_out = g.fir_num_outputs(total, T, D)
    ref = _one_shot(kind, D, dt, T, dx, n_out, cuda_device)
    got = _streamed(kind, D, dt, T, dx, blocks, cuda_device)
    assert got.numel() == n_out
    assert torch.equal(got, ref), f"max diff {float((got - ref).abs().max())}"


@pytest.mark.parametrize("kind", ["fc", "ff", "nco"])
@pytest.mark.parametrize("D,T", [(8, 255), (5, 63), (3, 100), (16, 9)])
def test_ragged_blocks_match_the_oracle(kind, D, T, cuda_device):
    rng = random.Random(99 * D + T)
    blocks = [rng.choice([1, 3, 7, T - 1, T, T + 1, 1001, 4097, 30011, 0, D]) for _ in range(40)]
    total = sum(blocks)
    taps = synth.random_taps(T, 13 + D)
    x = synth.tone_plus_noise(0, total, seed=300 + D, real=(kind == "ff"))
    dt, dx = torch.from_numpy(taps).to(cuda_device), torch.from_numpy(x).to(cuda_device)
    n_out = g.fir_num_outputs(total, T, D)
    got = _streamed(kind, D, dt, T, dx, blocks, cuda_device).cpu().numpy()
    assert got.shape[0] == n_out
    n_chk = min(n_out, 4000)
    if kind == "nco":
        want = oracle.adjust_frequency_fir_fc(oracle.NCO_EXACT, FS, SHIFT, FIRST, D, taps, x, n_chk, f64=True)
    else:
        want = oracle.fir(kind, D, taps, x, n_chk, f64=True)
    assert np.abs(got[:n_chk] - want).max() <= _tol(taps, x)
    ref = _one_shot(kind, D, dt, T, dx, n_out, cuda_device).cpu().numpy()
    assert np.abs(got - ref).max() <= 2 * _tol(taps, x)


def test_reset_restarts_the_sample_count(cuda_device):
    D, T, n = 8, 255, 50_000
    taps = synth.random_taps(T, 5)
    x = synth.tone_plus_noise(0, n, seed=17)
    dt, dx = torch.from_numpy(taps).to(cuda_device), torch.from_numpy(x).to(cuda_device)
    st = g.FirStream(g.FirStream.FC_NCO, D, dt, T, FS, SHIFT, FIRST, 0)
    ya = torch.zeros(st.num_outputs(n), dtype=torch.complex64, device=cuda_device)
    st.push(dx, n, ya)
    st.reset()
    yb = torch.zeros_like(ya)
    assert st.push(dx, n, yb) == ya.numel()
    torch.cuda.synchronize()
    assert torch.equal(ya, yb)


@pytest.mark.parametrize("nco", [False, True])
@pytest.mark.parametrize("D,T", [(8, 255), (10, 255), (32, 1023), (5, 63)])
def test_int8_blocks(D, T, nco, cuda_device):
    """int8 I/Q blocks (2 bytes per sample): blocks that are multiples of 8 samples reproduce the one-shot
    gsdrFirFCInt8 / gsdrAdjustFrequencyFirFCInt8 bits; ragged blocks stay within the tolerance."""
    rng = random.Random(5 * D + T + int(nco))
    taps = synth.random_taps(T, 17 + D)
    dt = torch.from_numpy(taps).to(cuda_device)
    ftype = g.FirStream.FC_NCO_INT8 if nco else g.FirStream.FC_INT8
    for ragged in (False, True):
        blocks = [rng.choice([4096, 8192, 65536, 16, 256, 8, 0, 20000]) for _ in range(20)]
        if ragged:
            blocks = [b + rng.choice([0, 1, 3, 5]) for b in blocks]
        total = sum(blocks)
        iq = np.random.default_rng(total).integers(-128, 128, size=2 * total, dtype=np.int64).astype(np.int8)
        di = torch.from_numpy(iq).to(cuda_device)
        n_out = g.fir_num_outputs(total, T, D)
        ref = torch.zeros(n_out, dtype=torch.complex64, device=cuda_device)
        if nco:
            g.gsdrAdjustFrequencyFirFCInt8(FS, SHIFT, FIRST, D, dt, T, di, ref, n_out, 0, None)
        else:
            g.gsdrFirFCInt8(D, dt, T, di, ref, n_out, 0, None)
        st = g.FirStream(ftype, D, dt, T, FS, SHIFT, FIRST, 0)
        outs, pos = [], 0
        for n in blocks:
            want = st.num_outputs(n)
            y = torch.zeros(want + 2, dtype=torch.complex64, device=cuda_device)
            assert st.push(di[2 * pos:2 * (pos + n)] if n else None, n, y, None) == want
            outs.append(y[:want])
            pos += n
        torch.cuda.synchronize()
        got = torch.cat(outs)
        st.close()
        assert got.numel() == n_out
        if ragged:
            x = np.maximum(np.float32(-1), iq.astype(np.float32) / np.float32(127))
            tol = 2 * _tol(taps, x)
            assert float((got - ref).abs().max()) <= tol
        else:
            assert torch.equal(got, ref), f"max diff {float((got - ref).abs().max())}"
