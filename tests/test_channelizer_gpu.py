"""gsdrChannelizeFC: one input, K frequency shifts — the fused kernel (window fetched once per tile)
and the per-shift fallback, against separate gsdrAdjustFrequencyFirFC calls and the oracle."""
import numpy as np
import pytest
import torch

import gsdr_b200 as g
from gsdr_b200 import synth
from oracle import oracle

pytestmark = pytest.mark.gpu


def _tol(taps, x):
    return 1e-5 * float(np.abs(taps).sum()) * float(np.abs(x).max())


@pytest.mark.parametrize("D,T,n_out,shifts", [
    (8, 255, 70_003, [0.0, 29520.0, -300e3, 1.1e6, 7.0]),
    (4, 127, 50_000, [100e3, -100e3, 250e3, 0.5]),
    (10, 255, 33_333, [300e3, -75e3, 12345.0]),
    (8, 31, 90_001, [float(f) for f in np.linspace(-1.0e6, 1.0e6, 20)]),   # 20 shifts: two launches of <= 16
    (8, 255, 9, [1.0e5, 2.0e5]),                                            # less than one tile
    (32, 1023, 20_000, [29520.0, -1.0e5]),                                  # no fused kernel for this decimation: per-shift calls
    (5, 63, 10_000, [29520.0, 400.0, -5.0e5]),                              # odd decimation: cp.async kernel per shift
])
@pytest.mark.parametrize("fused", [False, True])
def test_channelizer_equals_separate_calls_and_oracle(D, T, n_out, shifts, fused, cuda_device):
    """fused=True: firTmaChannelizerKernel through the tuning build's override (-5), where its shapes allow."""
    if fused:
        g.set_kernel_variant(-5)
    fs, first = 2.4e6, 2 ** 35 + 123
    n_in = g.fir_num_inputs(n_out, T, D)
    taps = synth.random_taps(T, 400 + D)
    x = synth.tone_plus_noise(0, n_in, seed=410 + T)
    dx, dt = torch.from_numpy(x).to(cuda_device), torch.from_numpy(taps).to(cuda_device)
    K = len(shifts)
    stride = n_out + 5
    out = torch.full((K, stride), float("nan"), dtype=torch.complex64, device=cuda_device)
    g.gsdrChannelizeFC(fs, shifts, first, D, dt, T, dx, out, stride, n_out, 0, None)
    torch.cuda.synchronize()
    assert torch.isnan(out[:, n_out:].real).all(), "wrote past numOutputs"
    tol = _tol(taps, x)
    g.set_kernel_variant(-1)
    for k, f in enumerate(shifts):
        one = torch.zeros(n_out, dtype=torch.complex64, device=cuda_device)
        g.gsdrAdjustFrequencyFirFC(fs, f, first, D, dt, T, dx, one, n_out, 0, None)
        torch.cuda.synchronize()
        assert float((out[k, :n_out] - one).abs().max()) <= tol, f"shift {k}"
    n_chk = min(n_out, 1500)
    for k in (0, K - 1):
        o0 = n_out - n_chk
        want = oracle.adjust_frequency_fir_fc(oracle.NCO_EXACT, fs, shifts[k], first + o0 * D, D, taps, x[o0 * D:], n_chk,
                                              f64=True)
        assert np.abs(out[k, o0:n_out].cpu().numpy() - want).max() <= tol


def test_channelizer_is_capturable_and_handles_empty_calls(cuda_device):
    D, T, n_out = 8, 255, 20_000
    dx = synth.tone_plus_noise(0, g.fir_num_inputs(n_out, T, D), seed=420, device=cuda_device)
    dt = torch.from_numpy(synth.lowpass_taps(T, D)).to(cuda_device)
    out = torch.zeros((3, n_out), dtype=torch.complex64, device=cuda_device)
    g.gsdrChannelizeFC(2.4e6, [], 0, D, dt, T, dx, out, n_out, n_out, 0, None)       # no shifts: nothing to do
    g.gsdrChannelizeFC(2.4e6, [1e5], 0, D, dt, T, dx, out, n_out, 0, 0, None)        # no outputs
    s = torch.cuda.Stream()
    g.gsdrChannelizeFC(2.4e6, [1e5, 2e5, 3e5], 0, D, dt, T, dx, out, n_out, n_out, 0, s)
    s.synchronize()
    want = out.clone()
    out.zero_()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        g.gsdrChannelizeFC(2.4e6, [1e5, 2e5, 3e5], 0, D, dt, T, dx, out, n_out, n_out, 0, s)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, want)
