"""int8 IQ input (include/gsdr/conversion.h): gsdrInt8ToNormFloat is bit-identical to the reference's
conversion; gsdrFirFCInt8 / gsdrAdjustFrequencyFirFCInt8 equal convert-then-filter within the FP32 tolerance."""
import numpy as np
import pytest
import torch

import gsdr_b200 as g
from gsdr_b200 import synth
from oracle import oracle

pytestmark = pytest.mark.gpu


def _iq_int8(n, seed):
    rng = np.random.default_rng(seed)
    iq = rng.integers(-128, 128, size=2 * n, dtype=np.int64).astype(np.int8)
    iq[:8] = np.array([-128, -127, 127, 0, 1, -1, 126, -126], np.int8)  # the conversion's corner values
    return iq


def _conv(iq):
    """ref: src/conversion.cu:26 — max(-1, x / 127) in float32 (IEEE division)."""
    f = np.maximum(np.float32(-1.0), iq.astype(np.float32) / np.float32(127.0)).astype(np.float32)
    return (f[0::2] + 1j * f[1::2]).astype(np.complex64)


def _tol(taps, x):
    return 1e-5 * float(np.abs(taps).sum()) * float(np.abs(x).max())


def test_int8_to_norm_float_is_bit_exact(cuda_device):
    iq = np.arange(-128, 128, dtype=np.int64).astype(np.int8)
    iq = np.concatenate([iq, _iq_int8(5000, 1)])
    d = torch.from_numpy(iq).to(cuda_device)
    out = torch.full((iq.size + 1,), 9.0, dtype=torch.float32, device=cuda_device)
    g.gsdrInt8ToNormFloat(d, out, iq.size, 0, None)
    torch.cuda.synchronize()
    want = np.maximum(np.float32(-1.0), iq.astype(np.float32) / np.float32(127.0)).astype(np.float32)
    assert out[:-1].cpu().numpy().tobytes() == want.tobytes()
    assert float(out[-1]) == 9.0, "element numElements must not be written (the reference's off-by-one)"
    assert want[0] == -1.0 and want[1] == -1.0 and want[255] == 1.0 and want[128] == 0.0


@pytest.mark.parametrize("nco", [False, True])
@pytest.mark.parametrize("D,T,n_out", [(8, 255, 40_001), (10, 255, 30_000), (32, 1023, 9_000), (4, 127, 20_000),
                                       (2, 9, 5_000), (16, 100, 6_000), (6, 64, 7_777), (5, 63, 4_000), (1, 31, 3_000)])
def test_int8_fir_equals_convert_then_filter(D, T, n_out, nco, cuda_device):
    fs, f, first = 2.4e6, 29520.0, 2 ** 36 + 77
    taps = synth.random_taps(T, 31 + D)
    n_in = (n_out - 1) * D + T
    iq = _iq_int8(n_in, 400 + D)
    x = _conv(iq)
    dt, di = torch.from_numpy(taps).to(cuda_device), torch.from_numpy(iq).to(cuda_device)
    dy = torch.full((n_out + 2,), 5.0, dtype=torch.complex64, device=cuda_device)
    g.set_kernel_variant(-1)
    if nco:
        g.gsdrAdjustFrequencyFirFCInt8(fs, f, first, D, dt, T, di, dy, n_out, 0, None)
    else:
        g.gsdrFirFCInt8(D, dt, T, di, dy, n_out, 0, None)
    torch.cuda.synchronize()
    y = dy.cpu().numpy()
    assert (y[n_out:] == 5.0).all()
    n_chk = min(n_out, 3000)
    if nco:
        want = oracle.adjust_frequency_fir_fc(oracle.NCO_EXACT, fs, f, first, D, taps, x, n_chk, f64=True)
        tail0 = n_out - 400
        tail = oracle.adjust_frequency_fir_fc(oracle.NCO_EXACT, fs, f, first + tail0 * D, D, taps, x[tail0 * D:], 400, f64=True)
    else:
        want = oracle.fir("fc", D, taps, x, n_chk, f64=True)
        tail0 = n_out - 400
        tail = oracle.fir("fc", D, taps, x[tail0 * D:], 400, f64=True)
    tol = _tol(taps, x)
    assert np.abs(y[:n_chk] - want).max() <= tol
    assert np.abs(y[tail0:n_out] - tail).max() <= tol
    # the float path on the converted samples agrees too (same kernels downstream of the conversion)
    dx = torch.from_numpy(x).to(cuda_device)
    dz = torch.zeros(n_out, dtype=torch.complex64, device=cuda_device)
    if nco:
        g.gsdrAdjustFrequencyFirFC(fs, f, first, D, dt, T, dx, dz, n_out, 0, None)
    else:
        g.gsdrFirFC(D, dt, T, dx, dz, n_out, 0, None)
    torch.cuda.synchronize()
    assert float((dz - dy[:n_out]).abs().max()) <= 2 * tol


def test_int8_unaligned_input_and_forced_direct(cuda_device):
    D, T, n_out = 8, 255, 3000
    taps = synth.random_taps(T, 5)
    n_in = (n_out - 1) * D + T
    iq = _iq_int8(n_in + 1, 9)
    dt, di = torch.from_numpy(taps).to(cuda_device), torch.from_numpy(iq).to(cuda_device)
    want = oracle.fir("fc", D, taps, _conv(iq[2:]), n_out, f64=True)
    tol = _tol(taps, _conv(iq))
    for forced, inp in ((-1, di[2:]), (-2, di[2:])):  # 2-byte aligned input -> direct kernel; forced direct
        g.set_kernel_variant(forced)
        dy = torch.zeros(n_out, dtype=torch.complex64, device=cuda_device)
        g.gsdrFirFCInt8(D, dt, T, inp, dy, n_out, 0, None)
        torch.cuda.synchronize()
        assert np.abs(dy.cpu().numpy() - want).max() <= tol
