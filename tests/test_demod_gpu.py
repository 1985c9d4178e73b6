"""NEXT row: quadrature demodulators and the FM receive stage, through the C ABI."""
import math

import numpy as np
import pytest
import torch

import gsdr_b200 as g
from gsdr_b200 import synth
from oracle import oracle, ref_cuda

pytestmark = pytest.mark.gpu


def _fm_signal(n, fs, dev_hz, audio_hz, offset_hz, seed):
    """Complex FM: a tone of audio_hz deviating +-dev_hz, carrier offset_hz away from the tuning frequency."""
    t = np.arange(n, dtype=np.float64) / fs
    phase = 2 * math.pi * offset_hz * t - (dev_hz / audio_hz) * np.cos(2 * math.pi * audio_hz * t)
    noise = 0.01 * synth.tone_plus_noise(0, n, seed=seed, amp=0.0, sigma=1.0)
    return (0.7 * np.exp(1j * phase)).astype(np.complex64) + noise.astype(np.complex64)


@pytest.mark.parametrize("n,offset", [(1, 0), (5, 1), (4096, 0), (100_003, 0), (100_003, 3)])
def test_quad_fm_demod(n, offset, cuda_device):
    x = synth.tone_plus_noise(0, n + 1 + offset, seed=70, tone_cycles_per_sample=0.05)
    dx = torch.from_numpy(x).to(cuda_device)[offset:]  # offset != 0: 8-byte-aligned input, scalar path
    dy = torch.full((n + 8,), float("nan"), dtype=torch.float32, device=cuda_device)
    g.gsdrQuadFmDemod(dx, dy[4:], 1.75, n, 0, None)
    torch.cuda.synchronize()
    y = dy.cpu().numpy()
    assert np.isnan(y[:4]).all() and np.isnan(y[4 + n:]).all()
    want = oracle.quad_fm_demod(x[offset:], 1.75, n)
    assert np.abs(y[4:4 + n] - want).max() <= 1.75 * 4e-7 * math.pi  # atan2f: CUDA vs libm, a few ulp of pi
    if ref_cuda.available():
        dr = torch.zeros(n, dtype=torch.float32, device=cuda_device)
        rc = ref_cuda.lib().gsdrQuadFmDemod(dx.data_ptr(), dr.data_ptr(), 1.75, n, 0, 0)
        torch.cuda.synchronize()
        assert rc == 0
        assert torch.equal(dr, dy[4:4 + n]), "must reproduce the reference kernel's bits (ref: src/quad_demod.cu:23-37)"


def test_quad_am_demod(cuda_device):
    n = 50_001
    x = (synth.tone_plus_noise(0, n, seed=71) * 1.7).astype(np.complex64)
    dx = torch.from_numpy(x).to(cuda_device)
    dy = torch.zeros(n, dtype=torch.float32, device=cuda_device)
    g.gsdrQuadAmDemod(dx, dy, n, 0, None)
    torch.cuda.synchronize()
    want = 2.0 * np.clip(np.hypot(x.real.astype(np.float64), x.imag.astype(np.float64)), 0.0, 1.0) - 1.0
    assert np.abs(dy.cpu().numpy() - want).max() <= 5e-7


@pytest.mark.parametrize("D,T", [(10, 255), (8, 63), (5, 33)])
def test_fm_demod_stage_against_oracle_chain(D, T, cuda_device):
    """gsdrFmDemod == oracle(mix + FIR, numOutputs + 1 values) -> oracle(quad demod)."""
    fs, tuning, channel, dev_hz = 2.4e6, 100.0e6, 100.3e6, 75e3
    n_out, first = 20_000, 2 ** 33 + 77
    n_in = n_out * D + T
    x = _fm_signal(n_in, fs, dev_hz, 1000.0, channel - tuning, seed=72)
    # the NCO must undo the +300 kHz offset: frequencyShift = tuning - channel = -300 kHz
    # (the signal is generated from absolute index 0; start the NCO at `first` to exercise large indices — the
    #  constant phase offset does not change the FM demodulation)
    taps = synth.lowpass_taps(T, D, cutoff=0.45)
    dx, dt = torch.from_numpy(x).to(cuda_device), torch.from_numpy(taps).to(cuda_device)
    dy = torch.zeros(n_out, dtype=torch.float32, device=cuda_device)
    g.gsdrFmDemod(fs, tuning, channel, dev_hz, D, first, dt, T, dx, dy, n_out, 0, None)
    torch.cuda.synchronize()
    lp = oracle.adjust_frequency_fir_fc(oracle.NCO_EXACT, fs, tuning - channel, first, D, taps, x, n_out + 1)
    gain = np.float32(fs) / (np.float32(2.0) * np.float32(math.pi) * np.float32(dev_hz))
    want = oracle.quad_fm_demod(lp, float(gain), n_out)
    y = dy.cpu().numpy()
    # |lp| ~ 0.7, FIR error <= 1e-5 * sum|h| * max|x| per sample => phase error <= ~4e-5 rad, times gain
    assert np.abs(y - want).max() <= float(gain) * 6e-5
    # and it really is the modulating tone: 1 kHz at the decimated rate, amplitude gain * 2*pi*dev/fs_out... = D
    seg = y[T // D + 5:]
    assert abs(float(seg.max()) - D) < 0.05 * D and abs(float(seg.min()) + D) < 0.05 * D


def test_fm_chain_config5_shape_small(cuda_device):
    """BASELINE config 5 at a small size: mix -> 255-tap FIR decim 10 -> quad demod -> 63-tap audio FIR decim 5,
    composed from the C-ABI calls, against the same chain of oracle stages."""
    fs, shift, dev_hz = 2.4e6, -300e3, 75e3
    D1, T1, D3, T3 = 10, 255, 5, 63
    n3 = 3000
    n2 = g.fir_num_inputs(n3, T3, D3)       # demod samples needed
    n1 = n2 + 1                             # low-pass samples needed
    n_in = g.fir_num_inputs(n1, T1, D1)
    assert n_in == (n3 - 1) * D1 * D3 + 885  # window 885, stride 50
    x = _fm_signal(n_in, fs, dev_hz, 1500.0, 300e3, seed=73)
    h1, h3 = synth.lowpass_taps(T1, D1, cutoff=0.45), synth.lowpass_taps(T3, D3)
    dx = torch.from_numpy(x).to(cuda_device)
    d1, d3 = torch.from_numpy(h1).to(cuda_device), torch.from_numpy(h3).to(cuda_device)
    lp = torch.zeros(n1, dtype=torch.complex64, device=cuda_device)
    dm = torch.zeros(n2, dtype=torch.float32, device=cuda_device)
    au = torch.zeros(n3, dtype=torch.float32, device=cuda_device)
    gain = fs / (2 * math.pi * dev_hz)
    s = torch.cuda.Stream()
    g.gsdrAdjustFrequencyFirFC(fs, shift, 0, D1, d1, T1, dx, lp, n1, 0, s)
    g.gsdrQuadFmDemod(lp, dm, gain, n2, 0, s)
    g.gsdrFirFF(D3, d3, T3, dm, au, n3, 0, s)
    s.synchronize()
    o_lp = oracle.adjust_frequency_fir_fc(oracle.NCO_EXACT, fs, shift, 0, D1, h1, x, n1)
    o_dm = oracle.quad_fm_demod(o_lp, gain, n2)
    o_au = oracle.fir("ff", D3, h3, o_dm, n3)
    assert np.abs(au.cpu().numpy() - o_au).max() <= gain * 6e-5 * float(np.abs(h3).sum())


# ---- fused output stages: quad demod / AM envelope in the FIR's store path -----------------

@pytest.mark.parametrize("D,T,n_out", [(10, 255, 300_001), (8, 255, 100_000), (4, 127, 65_537), (32, 1023, 160_000),
                                       (6, 100, 50_000), (16, 500, 40_003), (10, 255, 7)])
def test_am_demod_against_oracle_chain(D, T, n_out, cuda_device):
    """gsdrAmDemod (ref: include/gsdr/am.h:25-37, src/am.cu:41-49) == oracle(mix + FIR) -> 2*sat(|y|) - 1."""
    fs, tuning, channel, first = 2.4e6, 100.0e6, 100.3e6, 2 ** 34 + 5
    n_in = g.fir_num_inputs(n_out, T, D)
    x = (1.6 * synth.tone_plus_noise(0, n_in, seed=80 + D, tone_cycles_per_sample=0.125)).astype(np.complex64)
    taps = synth.lowpass_taps(T, D, cutoff=0.45)
    dx, dt = torch.from_numpy(x).to(cuda_device), torch.from_numpy(taps).to(cuda_device)
    dy = torch.full((n_out + 8,), float("nan"), dtype=torch.float32, device=cuda_device)
    g.gsdrAmDemod(fs, tuning, channel, D, first, dt, T, dx, dy[4:], n_out, 0, None)
    torch.cuda.synchronize()
    y = dy.cpu().numpy()
    assert np.isnan(y[:4]).all() and np.isnan(y[4 + n_out:]).all()
    n_chk = min(n_out, 4000)
    for o0 in sorted({0, (n_out - n_chk) // 2, n_out - n_chk}):
        lp = oracle.adjust_frequency_fir_fc(oracle.NCO_EXACT, fs, tuning - channel, first + o0 * D, D, taps,
                                            x[o0 * D:], n_chk)
        want = oracle.quad_am_demod(lp)
        tol = 2.0 * 1e-5 * float(np.abs(taps).sum()) * float(np.abs(x).max()) + 5e-7
        assert np.abs(y[4 + o0: 4 + o0 + n_chk] - want).max() <= tol
    assert y[4:4 + n_out].min() >= -1.0 and y[4:4 + n_out].max() <= 1.0
    # the unfused path (direct kernel: no fused stage, pool scratch + envelope kernel) gives the same within tolerance
    g.set_kernel_variant(-2)
    dz = torch.zeros(n_out, dtype=torch.float32, device=cuda_device)
    g.gsdrAmDemod(fs, tuning, channel, D, first, dt, T, dx, dz, min(n_out, 20_000), 0, None)
    torch.cuda.synchronize()
    m = min(n_out, 20_000)
    assert np.abs(dz[:m].cpu().numpy() - y[4:4 + m]).max() <= 2.0 * 1e-5 * float(np.abs(taps).sum()) * float(np.abs(x).max()) + 5e-7


@pytest.mark.parametrize("D,T,n_out", [(10, 255, 300_001), (8, 255, 100_003), (4, 127, 70_000), (32, 1023, 50_000),
                                       (6, 100, 20_001), (10, 255, 1), (10, 255, 503), (10, 255, 504), (10, 255, 505)])
def test_fm_demod_fused_stage_equals_unfused_chain(D, T, n_out, cuda_device):
    """The quadrature demodulator in the FIR's store path (tiles overlap by one row group) against the two-kernel
    chain composed from the C-ABI calls: phase steps agree to the FIR tolerance, tile seams included."""
    fs, tuning, channel, dev_hz, first = 2.4e6, 100.0e6, 100.3e6, 75e3, 2 ** 33 + 77
    n_in = n_out * D + T
    x = _fm_signal(n_in, fs, dev_hz, 1000.0, channel - tuning, seed=85)
    taps = synth.lowpass_taps(T, D, cutoff=0.45)
    dx, dt = torch.from_numpy(x).to(cuda_device), torch.from_numpy(taps).to(cuda_device)
    dy = torch.full((n_out + 8,), float("nan"), dtype=torch.float32, device=cuda_device)
    g.gsdrFmDemodFused(fs, tuning, channel, dev_hz, D, first, dt, T, dx, dy[4:], n_out, 0, None)
    lp = torch.zeros(n_out + 1, dtype=torch.complex64, device=cuda_device)
    dm = torch.zeros(n_out, dtype=torch.float32, device=cuda_device)
    g.gsdrAdjustFrequencyFirFC(fs, tuning - channel, first, D, dt, T, dx, lp, n_out + 1, 0, None)
    gain = fs / (2 * math.pi * dev_hz)
    g.gsdrQuadFmDemod(lp, dm, gain, n_out, 0, None)
    torch.cuda.synchronize()
    y = dy.cpu().numpy()
    assert np.isnan(y[:4]).all() and np.isnan(y[4 + n_out:]).all()
    assert np.abs(y[4:4 + n_out] - dm.cpu().numpy()).max() <= gain * 6e-5


def test_fm_demod_fused_reports_unsupported_shapes(cuda_device):
    D, T, n_out = 5, 63, 1000  # odd decimation: no TMA-fed kernel, hence no fused stage
    x = synth.tone_plus_noise(0, n_out * D + T, seed=86, device=cuda_device)
    dt = torch.from_numpy(synth.lowpass_taps(T, D)).to(cuda_device)
    dy = torch.zeros(n_out, dtype=torch.float32, device=cuda_device)
    with pytest.raises(g.CudaError) as e:
        g.gsdrFmDemodFused(2.4e6, 100.0e6, 100.3e6, 75e3, D, 0, dt, T, x, dy, n_out, 0, None)
    assert e.value.code == 801  # cudaErrorNotSupported
    g.gsdrFmDemod(2.4e6, 100.0e6, 100.3e6, 75e3, D, 0, dt, T, x, dy, n_out, 0, None)  # the default form serves it
    torch.cuda.synchronize()
    assert float(dy.abs().max()) > 0.0
