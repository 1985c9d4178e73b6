"""Parity tests proper: the sm_100a kernels, called through the C ABI, against the oracle (and against the
reference's own CUDA kernels from oracle/_ref) on the same inputs.

Tolerance (BASELINE.json north_star): max|err| <= 1e-5 * sum|h| * max|x| for FP32 results, because the
accumulation order differs from the reference's single ascending chain.  The direct kernel (CC/CF and shapes the
polyphase kernel cannot hold) keeps the reference's order and must match it bit for bit.
"""
import numpy as np
import pytest
import torch

import gsdr_b200 as g
from gsdr_b200 import synth
from oracle import oracle, ref_cuda

pytestmark = pytest.mark.gpu

KIND_FN = {"ff": g.gsdrFirFF, "fc": g.gsdrFirFC, "cc": g.gsdrFirCC, "cf": g.gsdrFirCF}
EDGE_SHAPES = [(1, 1), (2, 1), (1, 2), (16, 8), (8, 16), (31, 15), (32, 16), (33, 17)]  # ref: tests/test_fir.cpp:261-263


def _rand(cplx, n, seed):
    rng = np.random.default_rng(seed)
    if cplx:
        return (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
    return rng.uniform(-1, 1, n).astype(np.float32)


def _tol(taps, x):
    return 1e-5 * float(np.abs(taps).sum()) * float(np.abs(x).max())


def _run(kind, D, taps, x, n_out, dev, stream=None):
    dt = torch.from_numpy(taps).to(dev)
    dx = torch.from_numpy(x).to(dev)
    out_dtype = torch.float32 if kind == "ff" else torch.complex64
    # poison the output and keep guard elements around it to catch out-of-range stores
    dy = torch.full((n_out + 16,), float("nan"), dtype=out_dtype, device=dev)
    KIND_FN[kind](D, dt, taps.shape[0], dx, dy[8:], n_out, 0, stream)
    torch.cuda.synchronize()
    full = dy.cpu().numpy()
    assert np.isnan(full[:8].real).all() and np.isnan(full[8 + n_out:].real).all(), "store outside [0, numOutputs)"
    return full[8:8 + n_out]


@pytest.mark.parametrize("kind", ["ff", "fc", "cc", "cf"])
@pytest.mark.parametrize("T,n_out", EDGE_SHAPES)
@pytest.mark.parametrize("D", [1, 2, 5])
def test_edge_shapes(kind, T, n_out, D, cuda_device):
    taps = _rand(kind[0] == "c", T, 100 + T)
    x = _rand(kind[1] == "c", (n_out - 1) * D + T, 200 + n_out)
    y = _run(kind, D, taps, x, n_out, cuda_device)
    want = oracle.fir(kind, D, taps, x, n_out)
    assert np.abs(y - want).max() <= _tol(taps, x)


@pytest.mark.parametrize("kind", ["ff", "fc"])
@pytest.mark.parametrize(
    "D,T,n_in",
    [
        (1, 63, 1 << 20),       # BASELINE config 1 (FF) at full size
        (8, 255, 1 << 20),      # config 2 shape, 1M samples
        (32, 1023, 1 << 20),    # config 3 shape
        (4, 127, 1 << 18),      # config 4 per-channel shape
        (10, 255, 300_007),     # config 5 stage 1 (decimation not a power of two, ragged length)
        (5, 63, 100_003),       # config 5 stage 3
        (3, 1000, 50_000),      # many taps, odd decimation
        (7, 5, 12_345),         # fewer taps than the decimation
        (1, 1, 4097),
        (64, 2047, 1 << 19),    # large D*T
    ],
)
def test_polyphase_against_oracle(kind, D, T, n_in, cuda_device):
    taps = synth.lowpass_taps(T, D) if T > 8 else _rand(False, T, 5)
    x = synth.tone_plus_noise(0, n_in, seed=31, real=(kind == "ff"))
    n_out = g.fir_num_outputs(n_in, T, D)
    y = _run(kind, D, taps, x, n_out, cuda_device)
    want = oracle.fir(kind, D, taps, x, n_out, threads=8)
    err = np.abs(y - want).max()
    assert err <= _tol(taps, x), f"max|err| {err}"
    truth = oracle.fir(kind, D, taps, x, n_out, f64=True) if n_out * T < 5e8 else None
    if truth is not None:
        assert np.abs(y.astype(truth.dtype) - truth).max() <= _tol(taps, x)


# polyphase (cp.async) | TMA-fed | warp-specialised fused NCO | real-input | complex-tap | real x complex-tap variants
N_POLY, N_TMA, N_SPEC, N_REAL, N_CC, N_CF, N_ALL = 12, 24, 30, 40, 44, 48, 51  # [N_CF, N_ALL): wide-row kernel (D = 32)


@pytest.mark.parametrize("variant", [-2] + list(range(N_CF)))
@pytest.mark.parametrize("kind", ["ff", "fc"])
def test_every_kernel_variant(kind, variant, cuda_device):
    assert (g.num_polyphase_variants(), g.num_kernel_variants()) == (N_POLY, N_ALL)
    D, T, n_in = 8, 255, 200_000
    taps = synth.random_taps(T, 77)  # asymmetric: catches tap-order bugs
    x = synth.tone_plus_noise(5, n_in, seed=32, real=(kind == "ff"))
    n_out = g.fir_num_outputs(n_in, T, D) - 3  # ragged: last tile partly filled, input longer than needed
    g.set_kernel_variant(variant)
    info = g.describe_kernel(0 if kind == "fc" else 1, D, T, n_out)
    wrong_type = (N_POLY <= variant < N_SPEC and kind == "ff") or (variant >= N_SPEC and kind == "fc")
    if variant >= N_POLY and kind == "fc" and info.variant == -1:
        assert info.phaseGroups <= 1  # e.g. 8 branch groups requested but D = 8 has only 4 branch pairs
    elif variant >= N_SPEC and kind == "ff" and info.variant == -1:
        pass  # a real-input variant whose window does not fit this shape (128-byte rows x 1024 outputs)
    else:
        assert info.variant == (variant if variant >= 0 and not wrong_type else -1)
    y = _run(kind, D, taps, x, n_out, cuda_device)
    want = oracle.fir(kind, D, taps, x, n_out)
    if info.variant == -1:
        assert y.tobytes() == want.tobytes(), "direct kernel must reproduce the reference's accumulation order"
    else:
        assert np.abs(y - want).max() <= _tol(taps, x)


@pytest.mark.parametrize("variant", list(range(N_SPEC, N_REAL)))
@pytest.mark.parametrize("D,T,n_out", [(1, 63, 70_001), (1, 7, 5000), (2, 100, 40_000), (3, 17, 33_333), (4, 127, 30_001),
                                       (5, 63, 50_001), (5, 200, 20_000), (7, 29, 9_999), (8, 255, 20_000),
                                       (16, 1023, 9_000), (64, 300, 2_001)])
def test_real_input_kernel(variant, D, T, n_out, cuda_device):
    """gsdrFirFF on the output-pair kernel: pairs of consecutive outputs share an FFMA2, producer warps rearrange the
    bulk-copied raw window.  Odd output counts (half-filled last pair), one tap block (T <= 16 D) and several,
    interior tiles (bulk copy) and the last ones (bounds-checked loads) all occur."""
    taps = synth.random_taps(T, 9 + D)
    x = synth.tone_plus_noise(0, (n_out - 1) * D + T, seed=90 + D, real=True)
    g.set_kernel_variant(variant)
    info = g.describe_kernel(1, D, T, n_out)
    if info.variant == -1:
        pytest.skip("variant does not fit this shape")
    assert info.variant == variant
    dt, dx = torch.from_numpy(taps).to(cuda_device), torch.from_numpy(x).to(cuda_device)
    dy = torch.full((n_out + 8,), 7.0, dtype=torch.float32, device=cuda_device)
    g.gsdrFirFF(D, dt, T, dx, dy, n_out, 0, None)
    torch.cuda.synchronize()
    y = dy.cpu().numpy()
    assert (y[n_out:] == 7.0).all(), "wrote past the last output"
    want = oracle.fir("ff", D, taps, x, n_out, threads=8)
    assert np.abs(y[:n_out] - want).max() <= _tol(taps, x)
    # unaligned output (4-byte aligned only): scalar store path
    dy2 = torch.zeros(n_out + 1, dtype=torch.float32, device=cuda_device)
    g.gsdrFirFF(D, dt, T, dx, dy2[1:], n_out, 0, None)
    torch.cuda.synchronize()
    assert dy2[1:].cpu().numpy().tobytes() == y[:n_out].tobytes()


@pytest.mark.parametrize("variant", [-1] + list(range(N_REAL, N_CC)))
@pytest.mark.parametrize("D,T,n_out", [(8, 255, 40_001), (2, 33, 10_000), (4, 127, 30_000), (10, 100, 9_999),
                                       (16, 500, 7_000), (32, 1023, 5_000), (6, 7, 3_000)])
def test_complex_tap_kernel(variant, D, T, n_out, cuda_device):
    """gsdrFirCC as two real-tap filters on the same windows (real and imaginary tap planes, the second published
    multiplied by j).  Accumulation order differs from the reference's: tolerance, not bits."""
    taps = synth.random_taps(T, 19 + D, complex_taps=True)
    x = synth.tone_plus_noise(0, (n_out - 1) * D + T, seed=110 + D)
    g.set_kernel_variant(variant)
    info = g.describe_kernel(2, D, T, n_out)
    if variant >= 0 and info.variant == -1:
        pytest.skip("variant does not fit this shape")
    assert N_REAL <= info.variant < N_CC and (variant < 0 or info.variant == variant)
    y = _run("cc", D, taps, x, n_out, cuda_device)
    want = oracle.fir("cc", D, taps, x, n_out, threads=8)
    assert np.abs(y - want).max() <= _tol(taps, x)


@pytest.mark.parametrize("variant", [-1] + list(range(N_CC, N_CF)))
@pytest.mark.parametrize("D,T,n_out", [(1, 63, 50_001), (5, 63, 30_001), (2, 100, 20_000), (4, 127, 9_999),
                                       (8, 255, 8_000), (3, 7, 5_000), (16, 600, 3_001)])
def test_real_input_complex_tap_kernel(variant, D, T, n_out, cuda_device):
    """gsdrFirCF: the output-pair kernel of gsdrFirFF with a real and an imaginary tap plane over the same window."""
    taps = synth.random_taps(T, 29 + D, complex_taps=True)
    x = synth.tone_plus_noise(0, (n_out - 1) * D + T, seed=130 + D, real=True)
    g.set_kernel_variant(variant)
    info = g.describe_kernel(3, D, T, n_out)
    if variant >= 0 and info.variant == -1:
        pytest.skip("variant does not fit this shape")
    assert N_CC <= info.variant < N_CF and (variant < 0 or info.variant == variant)
    dt, dx = torch.from_numpy(taps).to(cuda_device), torch.from_numpy(x).to(cuda_device)
    dy = torch.full((n_out + 4,), 7.0, dtype=torch.complex64, device=cuda_device)
    g.gsdrFirCF(D, dt, T, dx, dy, n_out, 0, None)
    torch.cuda.synchronize()
    y = dy.cpu().numpy()
    assert (y[n_out:] == 7.0).all(), "wrote past the last output"
    want = oracle.fir("cf", D, taps, x, n_out, threads=8)
    assert np.abs(y[:n_out] - want).max() <= _tol(taps, x)
    dy2 = torch.zeros(n_out + 1, dtype=torch.complex64, device=cuda_device)
    g.gsdrFirCF(D, dt, T, dx, dy2[1:], n_out, 0, None)  # 8-byte aligned output: scalar stores
    torch.cuda.synchronize()
    assert dy2[1:].cpu().numpy().tobytes() == y[:n_out].tobytes()


def test_complex_taps_batched_and_unaligned(cuda_device):
    D, T, n_out = 8, 255, 6000
    taps = synth.random_taps(T, 23, complex_taps=True)
    x = synth.tone_plus_noise(0, (n_out - 1) * D + T + 1, seed=120)
    dt, dx = torch.from_numpy(taps).to(cuda_device), torch.from_numpy(x).to(cuda_device)
    dy = torch.zeros(n_out, dtype=torch.complex64, device=cuda_device)
    g.gsdrFirCC(D, dt, T, dx[1:], dy, n_out, 0, None)  # 8-byte aligned input: direct kernel, the reference's bits
    torch.cuda.synchronize()
    assert dy.cpu().numpy().tobytes() == oracle.fir("cc", D, taps, x[1:], n_out).tobytes()


def test_real_input_batched_and_unaligned(cuda_device):
    """Channel batches go through the same kernel (channel = part of the tile index); inputs that are not 16-byte
    aligned fall back to the cp.async kernel and still match."""
    D, T, n_out, C = 5, 63, 12_345, 7
    taps = synth.random_taps(T, 3)
    n_in = (n_out - 1) * D + T
    stride_in = (n_in + 3) // 4 * 4
    xs = np.stack([np.pad(synth.tone_plus_noise(c, n_in, seed=70 + c, real=True), (0, stride_in - n_in)) for c in range(C)])
    dt, dx = torch.from_numpy(taps).to(cuda_device), torch.from_numpy(xs).to(cuda_device)
    dy = torch.zeros((C, n_out + 3), dtype=torch.float32, device=cuda_device)
    g.gsdrFirFFBatched(D, dt, T, 0, dx, stride_in, dy, n_out + 3, n_out, C, 0, None)
    torch.cuda.synchronize()
    for c in range(C):
        want = oracle.fir("ff", D, taps, xs[c, :n_in], n_out)
        assert np.abs(dy[c, :n_out].cpu().numpy() - want).max() <= _tol(taps, xs[c])
    one = torch.zeros(n_out, dtype=torch.float32, device=cuda_device)
    g.gsdrFirFF(D, dt, T, dx[3], one, n_out, 0, None)
    torch.cuda.synchronize()
    assert one.cpu().numpy().tobytes() == dy[3, :n_out].cpu().numpy().tobytes(), "batched == single call, bit for bit"
    xo = torch.from_numpy(np.concatenate([np.zeros(1, np.float32), xs[0]])).to(cuda_device)
    g.gsdrFirFF(D, dt, T, xo[1:], one, n_out, 0, None)  # 4-byte aligned input
    torch.cuda.synchronize()
    want = oracle.fir("ff", D, taps, xs[0, :n_in], n_out)
    assert np.abs(one.cpu().numpy() - want).max() <= _tol(taps, xs[0])


@pytest.mark.parametrize("variant", [-1] + list(range(N_CF, N_ALL)))
@pytest.mark.parametrize("T,n_in", [(1023, 700_001), (1023, 155_000), (2000, 400_000), (33, 300_000), (1, 200_003)])
def test_wide_row_kernel_d32(variant, T, n_in, cuda_device):
    """firTmaWideKernel (segment-pipelined, 512-output tiles): plain FIR and the fused exact NCO, interior tiles (TMA)
    and the last tile (cp.async, zero fill), against the oracle; variant -1: the automatic choice takes it for
    captures of at least 296 tiles."""
    D = 32
    taps = synth.random_taps(T, 900 + T)
    x = synth.tone_plus_noise(0, n_in, seed=61)
    n_out = g.fir_num_outputs(n_in, T, D) - 5
    g.set_kernel_variant(variant)
    dt, dx = torch.from_numpy(taps).to(cuda_device), torch.from_numpy(x).to(cuda_device)
    info = g.describe_kernel(0, D, T, n_out)
    if variant >= 0:
        assert info.variant == variant and info.outputsPerBlock == 512
    y = _run("fc", D, taps, x, n_out, cuda_device)
    want = oracle.fir("fc", D, taps, x, n_out, threads=8)
    assert np.abs(y - want).max() <= _tol(taps, x)
    dz = torch.full((n_out + 8,), float("nan"), dtype=torch.complex64, device=cuda_device)
    first = 2 ** 36 + 11
    g.gsdrAdjustFrequencyFirFC(1.0e6, 123456.0, first, D, dt, T, dx, dz[4:], n_out, 0, None)
    torch.cuda.synchronize()
    assert torch.isnan(dz[:4].real).all() and torch.isnan(dz[4 + n_out:].real).all()
    for o0, n_chk in ((0, 1500), (n_out // 2 - 7, 1500), (n_out - 700, 700)):
        wz = oracle.adjust_frequency_fir_fc(oracle.NCO_EXACT, 1.0e6, 123456.0, first + o0 * D, D, taps, x[o0 * D:], n_chk,
                                            f64=True)
        assert np.abs(dz[4 + o0: 4 + o0 + n_chk].cpu().numpy() - wz).max() <= _tol(taps, x)


@pytest.mark.parametrize("variant", list(range(N_POLY, N_SPEC)))
@pytest.mark.timeout(120)
@pytest.mark.parametrize("D,T", [(2, 33), (4, 127), (6, 100), (8, 255), (10, 255), (14, 29), (16, 500), (32, 1023), (48, 700), (64, 129)])
def test_tma_kernel_decimations_and_swizzle_modes(variant, D, T, cuda_device):
    """Rows of 16..128 bytes: no swizzle for an odd chunk count (D = 2, 6, 10, 14), 32/64/128-byte swizzle for
    D = 4, 8, 16.  Sizes chosen so that interior tiles (TMA) and the last tile (cp.async, zero fill) both occur."""
    n_in = 150_001
    taps = synth.random_taps(T, 5 + D)
    x = synth.tone_plus_noise(0, n_in, seed=60 + D)
    n_out = g.fir_num_outputs(n_in, T, D)
    g.set_kernel_variant(variant)
    info = g.describe_kernel(4 if variant >= N_TMA else 0, D, T, n_out)
    if info.variant == -1:
        pytest.skip("variant does not fit this shape")
    if variant < N_TMA:
        y = _run("fc", D, taps, x, n_out, cuda_device)
        want = oracle.fir("fc", D, taps, x, n_out, threads=8)
        assert np.abs(y - want).max() <= _tol(taps, x)
    # fused NCO on the same kernel
    dt, dx = torch.from_numpy(taps).to(cuda_device), torch.from_numpy(x).to(cuda_device)
    dz = torch.zeros(n_out, dtype=torch.complex64, device=cuda_device)
    g.gsdrAdjustFrequencyFirFC(1.0e6, 123456.0, 2 ** 35 + 9, D, dt, T, dx, dz, n_out, 0, None)
    torch.cuda.synchronize()
    n_chk = min(n_out, 3000)
    wz = oracle.adjust_frequency_fir_fc(oracle.NCO_EXACT, 1.0e6, 123456.0, 2 ** 35 + 9, D, taps, x, n_chk, f64=True)
    assert np.abs(dz[:n_chk].cpu().numpy() - wz).max() <= _tol(taps, x)
    tail = oracle.adjust_frequency_fir_fc(oracle.NCO_EXACT, 1.0e6, 123456.0, 2 ** 35 + 9 + (n_out - 500) * D, D, taps,
                                          x[(n_out - 500) * D:], 500, f64=True)
    assert np.abs(dz[n_out - 500:].cpu().numpy() - tail).max() <= _tol(taps, x)


@pytest.mark.parametrize("kind", ["cc", "cf", "fc", "ff"])
def test_direct_kernel_is_bit_exact_to_the_reference_order(kind, cuda_device):
    g.set_kernel_variant(-2)
    D, T, n_out = 6, 300, 5000
    taps = _rand(kind[0] == "c", T, 41)
    x = _rand(kind[1] == "c", (n_out - 1) * D + T, 42)
    y = _run(kind, D, taps, x, n_out, cuda_device)
    assert y.tobytes() == oracle.fir(kind, D, taps, x, n_out).tobytes()


def test_unaligned_pointers(cuda_device):
    """cuComplex is only 8-byte aligned by the ABI; shard offsets are arbitrary sample offsets."""
    D, T, n_out = 8, 255, 3000
    taps = synth.lowpass_taps(T, D)
    x = synth.tone_plus_noise(0, (n_out - 1) * D + T + 3, seed=33)
    dt = torch.from_numpy(np.concatenate([np.zeros(1, np.float32), taps])).to(cuda_device)[1:]  # 4-byte aligned taps
    dx = torch.from_numpy(x).to(cuda_device)
    dy = torch.zeros(n_out + 1, dtype=torch.complex64, device=cuda_device)
    for off in (1, 3):
        g.gsdrFirFC(D, dt, T, dx[off:], dy[1:], n_out, 0, None)
        torch.cuda.synchronize()
        want = oracle.fir("fc", D, taps, x[off:], n_out)
        assert np.abs(dy[1:].cpu().numpy() - want).max() <= _tol(taps, x)


def test_impulse_and_dc_known_answers(cuda_device):
    D, T, n_out = 8, 255, 2048
    taps = synth.random_taps(T, 3)
    n_in = (n_out - 1) * D + T
    for k in (0, 254, 255, 8 * 1024 + 3, n_in - 1):
        x = np.zeros(n_in, dtype=np.complex64)
        x[k] = 1.0 + 2.0j
        y = _run("fc", D, taps, x, n_out, cuda_device)
        want = np.zeros(n_out, dtype=np.complex64)
        for n in range(n_out):
            i = k - n * D
            if 0 <= i < T:
                want[n] = taps[i] * x[k]
        assert np.array_equal(y, want), f"impulse at {k}"
    y = _run("fc", D, taps, np.ones(n_in, dtype=np.complex64), n_out, cuda_device)
    assert np.abs(y - taps.astype(np.float64).sum()).max() <= _tol(taps, np.ones(1))


def test_zero_taps_zero_outputs_and_bad_decimation(cuda_device):
    dx = torch.ones(64, dtype=torch.complex64, device=cuda_device)
    dy = torch.full((8,), 7.0, dtype=torch.complex64, device=cuda_device)
    g.gsdrFirFC(2, None, 0, dx, dy, 8, 0, None)  # ref: tests/test_fir.cpp:249-257 — zero taps writes zeros
    torch.cuda.synchronize()
    assert (dy == 0).all()
    dy.fill_(7.0)
    g.gsdrFirFC(2, dx, 3, dx, dy, 0, 0, None)  # zero outputs: success, nothing written
    torch.cuda.synchronize()
    assert (dy == 7.0).all()
    with pytest.raises(g.CudaError) as e:
        g.gsdrFirFC(0, dx, 3, dx, dy, 4, 0, None)
    assert e.value.code == 1  # cudaErrorInvalidValue


def test_in_stream_semantics_and_graph_capture(cuda_device):
    """Work goes to the given stream only and the call is capturable (no sync, no allocation)."""
    D, T, n_in = 8, 255, 1 << 16
    taps = synth.lowpass_taps(T, D)
    x = synth.tone_plus_noise(0, n_in, seed=34)
    n_out = g.fir_num_outputs(n_in, T, D)
    dt, dx = torch.from_numpy(taps).to(cuda_device), torch.from_numpy(x).to(cuda_device)
    dy = torch.zeros(n_out, dtype=torch.complex64, device=cuda_device)
    s = torch.cuda.Stream()
    g.gsdrFirFC(D, dt, T, dx, dy, n_out, 0, s)  # warm-up outside capture (sets the smem attribute)
    s.synchronize()
    dy.zero_()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        g.gsdrFirFC(D, dt, T, dx, dy, n_out, 0, torch.cuda.current_stream())
    assert (dy == 0).all(), "capture must not execute"
    graph.replay()
    torch.cuda.synchronize()
    assert np.abs(dy.cpu().numpy() - oracle.fir("fc", D, taps, x, n_out)).max() <= _tol(taps, x)


@pytest.mark.parametrize("kind", ["ff", "cc", "cf", "nco", "i8", "stream"])
def test_every_family_member_is_capturable(kind, cuda_device):
    """The same in-stream contract for the other entry points: nothing runs during capture, the replay gives the
    result of a direct call bit for bit (including the block-streaming helper's copies and two launches)."""
    D, T, n_out = 8, 255, 9000
    n_in = (n_out - 1) * D + T
    ctaps = kind in ("cc", "cf")
    taps = synth.random_taps(T, 51, complex_taps=ctaps)
    x = synth.tone_plus_noise(0, n_in, seed=44, real=kind in ("ff", "cf"))
    dt, dx = torch.from_numpy(taps).to(cuda_device), torch.from_numpy(x).to(cuda_device)
    if kind == "i8":
        dx = torch.view_as_real(dx).mul(100.0).round().to(torch.int8).reshape(-1).contiguous()
    out_dtype = torch.float32 if kind == "ff" else torch.complex64
    fs, f, first = 1.0e6, 12345.0, 777

    st = g.FirStream(g.FirStream.FC, D, dt, T, 0.0, 0.0, 0, 0) if kind == "stream" else None

    def call(out, stream):
        if kind == "nco":
            g.gsdrAdjustFrequencyFirFC(fs, f, first, D, dt, T, dx, out, n_out, 0, stream)
        elif kind == "i8":
            g.gsdrFirFCInt8(D, dt, T, dx, out, n_out, 0, stream)
        elif kind == "stream":
            st.reset()
            half = (n_in // 2) & ~1
            a = st.push(dx[:half], half, out, stream)
            b = st.push(dx[half:], n_in - half, out[a:], stream)
            assert a + b == n_out
        else:
            {"ff": g.gsdrFirFF, "cc": g.gsdrFirCC, "cf": g.gsdrFirCF}[kind](D, dt, T, dx, out, n_out, 0, stream)

    s = torch.cuda.Stream()
    want = torch.zeros(n_out, dtype=out_dtype, device=cuda_device)
    call(want, s)  # direct call (also warms up the shared-memory attributes)
    s.synchronize()
    assert float(want.abs().max()) > 0
    got = torch.zeros(n_out, dtype=out_dtype, device=cuda_device)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        call(got, torch.cuda.current_stream())
    assert (got == 0).all(), "capture must not execute"
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    if st is not None:
        st.close()


def test_current_device_is_preserved(cuda_device):
    before = torch.cuda.current_device()
    dx = torch.ones(300, dtype=torch.float32, device=cuda_device)
    dy = torch.zeros(10, dtype=torch.float32, device=cuda_device)
    g.gsdrFirFF(3, dx, 5, dx, dy, 10, 0, None)
    torch.cuda.synchronize()
    assert torch.cuda.current_device() == before
    assert (dy == 5.0).all()
    with pytest.raises(g.CudaError):
        g.gsdrFirFF(3, dx, 5, dx, dy, 10, 99, None)  # no such device: error code, no crash
    assert torch.cuda.current_device() == before


# ---- batching and sharding --------------------------------------------------------------------------------

@pytest.mark.parametrize("kind", ["fc", "ff"])
@pytest.mark.parametrize("shared_taps", [True, False])
def test_batched_channels_equal_single_calls_bit_exact(kind, shared_taps, cuda_device):
    D, T, C, n_in = 4, 127, 9, 20_000
    n_out = g.fir_num_outputs(n_in, T, D)
    cplx = kind == "fc"
    x = np.stack([synth.tone_plus_noise(c * 1000, n_in, seed=50 + c, real=not cplx) for c in range(C)])
    taps = np.stack([synth.random_taps(T, 60 + (0 if shared_taps else c)) for c in range(C)])
    dx, dt = torch.from_numpy(x).to(cuda_device), torch.from_numpy(taps).to(cuda_device)
    out_dtype = torch.complex64 if cplx else torch.float32
    dyb = torch.zeros((C, n_out + 5), dtype=out_dtype, device=cuda_device)
    fnb = g.gsdrFirFCBatched if cplx else g.gsdrFirFFBatched
    fnb(D, dt, T, 0 if shared_taps else T, dx, n_in, dyb, n_out + 5, n_out, C, 0, None)
    dys = torch.zeros((C, n_out), dtype=out_dtype, device=cuda_device)
    for c in range(C):
        KIND_FN[kind](D, dt[c], T, dx[c], dys[c], n_out, 0, None)
    torch.cuda.synchronize()
    assert torch.equal(dyb[:, :n_out], dys)
    assert (dyb[:, n_out:] == 0).all()
    want = oracle.fir(kind, D, taps[3], x[3], n_out)
    assert np.abs(dys[3].cpu().numpy() - want).max() <= _tol(taps[3], x[3])


@pytest.mark.parametrize("shards,n_in,tensor_cores", [(2, 1 << 20, False), (8, 1 << 20, False), (2, 1 << 21, True),
                                                       (8, 1 << 23, True), (3, 5_000_011, True)])
def test_time_shards_equal_unsharded_bit_exact(shards, n_in, tensor_cores, cuda_device):
    """BASELINE requirement: shard-vs-unsharded results identical (decimation phase and offsets bit-exact).  Shards of
    at least 65536 outputs sit on the tensor-core kernel's tile grid and take the same kernel as the whole call;
    smaller shards of a call that large are only bit-identical with gsdrB200SetFirTensorCores(0) (include/gsdr/b200.h)."""
    D, T = 8, 255
    g.set_fir_tensor_cores(tensor_cores)
    try:
        _time_shards_case(shards, n_in, D, T, cuda_device)
    finally:
        g.set_fir_tensor_cores(True)


def _time_shards_case(shards, n_in, D, T, cuda_device):
    fs, f, first = 2.4e6, 29520.0, 987_654_321
    taps = synth.lowpass_taps(T, D)
    dx = synth.tone_plus_noise(0, n_in, seed=35, device=cuda_device)
    dt = torch.from_numpy(taps).to(cuda_device)
    n_out = g.fir_num_outputs(n_in, T, D)
    whole = torch.zeros(n_out, dtype=torch.complex64, device=cuda_device)
    whole_nco = torch.zeros_like(whole)
    g.gsdrFirFC(D, dt, T, dx, whole, n_out, 0, None)
    g.gsdrAdjustFrequencyFirFC(fs, f, first, D, dt, T, dx, whole_nco, n_out, 0, None)
    parts = torch.zeros_like(whole)
    parts_nco = torch.zeros_like(whole)
    for s in range(shards):
        sh = g.shard_plan_time(n_out, D, T, first, shards, s)
        xin = dx[sh.firstInput:sh.firstInput + sh.numInputs].clone()  # the shard's own resident copy (block + halo)
        g.gsdrFirFC(D, dt, T, xin, parts[sh.firstOutput:], sh.numOutputs, 0, None)
        g.gsdrAdjustFrequencyFirFC(fs, f, sh.firstSampleIndex, D, dt, T, xin, parts_nco[sh.firstOutput:],
                                   sh.numOutputs, 0, None)
    torch.cuda.synchronize()
    assert torch.equal(parts, whole)
    assert torch.equal(parts_nco, whole_nco)


# ---- fused NCO --------------------------------------------------------------------------------------------

@pytest.mark.parametrize("D,T,n_in", [(8, 255, 1 << 18), (32, 1023, 1 << 19), (10, 255, 100_003), (1, 31, 5000)])
@pytest.mark.parametrize("first", [0, 77, 2 ** 40 + 12345])
def test_nco_exact_against_oracle(D, T, n_in, first, cuda_device):
    fs, f = 2.4e6, -310e3
    taps = synth.lowpass_taps(T, D)
    x = synth.tone_plus_noise(first, n_in, seed=36, tone_cycles_per_sample=310e3 / 2.4e6)
    n_out = g.fir_num_outputs(n_in, T, D)
    dt, dx = torch.from_numpy(taps).to(cuda_device), torch.from_numpy(x).to(cuda_device)
    dy = torch.zeros(n_out, dtype=torch.complex64, device=cuda_device)
    g.gsdrAdjustFrequencyFirFC(fs, f, first, D, dt, T, dx, dy, n_out, 0, None)
    torch.cuda.synchronize()
    y = dy.cpu().numpy()
    n_chk = min(n_out, 4000)
    want = oracle.adjust_frequency_fir_fc(oracle.NCO_EXACT, fs, f, first, D, taps, x, n_chk, f64=True)
    assert np.abs(y[:n_chk] - want).max() <= _tol(taps, x)
    # the +310 kHz tone lands at DC: |y| ~ amp * sum(h) = 0.5 everywhere, including the far end of the capture
    # (the wider the filter, i.e. the smaller D, the more of the sigma = 0.1 noise rides on it)
    mag = np.abs(y[T // D + 1:])
    assert abs(float(mag.mean()) - 0.5) < 0.02 and float(mag.min()) > 0.25


@pytest.mark.parametrize("first", [0, 5_000_003])
def test_nco_literal_against_oracle(first, cuda_device):
    fs, f, D, T, n_in = 2.4e6, 1.0e5, 8, 63, 50_000
    taps = synth.random_taps(T, 8)
    x = synth.tone_plus_noise(0, n_in, seed=37)
    n_out = g.fir_num_outputs(n_in, T, D)
    dt, dx = torch.from_numpy(taps).to(cuda_device), torch.from_numpy(x).to(cuda_device)
    dy = torch.zeros(n_out, dtype=torch.complex64, device=cuda_device)
    g.gsdrAdjustFrequencyFirFCLiteral(fs, f, first, D, dt, T, dx, dy, n_out, 0, None)
    torch.cuda.synchronize()
    want = oracle.adjust_frequency_fir_fc(oracle.NCO_LITERAL, fs, f, first, D, taps, x, n_out)
    assert np.abs(dy.cpu().numpy() - want).max() <= _tol(taps, x)


@pytest.mark.parametrize("literal", [False, True])
@pytest.mark.parametrize("D,T,n_out,forced", [(100, 400, 3000, -1), (4096, 70, 500, -1), (131, 5000, 700, -1),
                                              (8, 255, 5000, -2)])
def test_nco_direct_fallback(D, T, n_out, forced, literal, cuda_device):
    """Shapes whose window fits no staged kernel (large decimations that are not a multiple of 16, very long tap
    sets) take the direct NCO kernel — one phasor per tap per output — instead of failing."""
    fs, f, first = 2.4e6, 1.0e5, 2 ** 34 + 5 if not literal else 1_000_003
    taps = synth.random_taps(T, 21)
    x = synth.tone_plus_noise(0, (n_out - 1) * D + T, seed=38)
    dt, dx = torch.from_numpy(taps).to(cuda_device), torch.from_numpy(x).to(cuda_device)
    dy = torch.zeros(n_out, dtype=torch.complex64, device=cuda_device)
    g.set_kernel_variant(forced)
    assert g.describe_kernel(4, D, T, n_out).variant == -1
    fn = g.gsdrAdjustFrequencyFirFCLiteral if literal else g.gsdrAdjustFrequencyFirFC
    fn(fs, f, first, D, dt, T, dx, dy, n_out, 0, None)
    torch.cuda.synchronize()
    mode = oracle.NCO_LITERAL if literal else oracle.NCO_EXACT
    n_chk = min(n_out, 600)
    want = oracle.adjust_frequency_fir_fc(mode, fs, f, first, D, taps, x, n_chk, f64=not literal)
    assert np.abs(dy[:n_chk].cpu().numpy() - want).max() <= _tol(taps, x)


# ---- against the reference's own CUDA kernels (oracle/_ref) ------------------------------------------------

needs_ref = pytest.mark.skipif(not ref_cuda.available(), reason="oracle/_ref/libgsdr_ref.so not built")


@needs_ref
@pytest.mark.parametrize("kind", ["ff", "fc", "cc", "cf"])
@pytest.mark.parametrize("D,T,n_in", [(1, 63, 1 << 20), (8, 255, 1 << 22), (10, 255, 1_000_003), (32, 1023, 1 << 21)])
def test_against_reference_cuda_kernels(kind, D, T, n_in, cuda_device):
    taps = synth.random_taps(T, 70, complex_taps=(kind[0] == "c"))
    taps = (taps / np.abs(taps).sum()).astype(taps.dtype)
    dx = synth.tone_plus_noise(0, n_in, seed=38, device=cuda_device, real=(kind[1] == "f"))
    dt = torch.from_numpy(taps).to(cuda_device)
    n_out = g.fir_num_outputs(n_in, T, D)
    out_dtype = torch.float32 if kind == "ff" else torch.complex64
    ours = torch.zeros(n_out, dtype=out_dtype, device=cuda_device)
    ref = torch.zeros(n_out, dtype=out_dtype, device=cuda_device)
    KIND_FN[kind](D, dt, T, dx, ours, n_out, 0, None)
    ref_cuda.fir(kind, D, dt, T, dx, ref, n_out)
    torch.cuda.synchronize()
    err = float((ours - ref).abs().max())
    tol = 1e-5 * float(np.abs(taps).sum()) * float(dx.abs().max())
    assert err <= tol, f"max|err| vs reference CUDA {err} > {tol}"
    # and the oracle reproduces the reference's bits on a prefix
    n_chk = min(n_out, 2000)
    xh = dx[: (n_chk - 1) * D + T].cpu().numpy()
    assert oracle.fir(kind, D, taps, xh, n_chk).tobytes() == ref[:n_chk].cpu().numpy().tobytes()


@needs_ref
def test_nco_literal_against_patched_reference_kernel(cuda_device):
    fs, f, first, D, T, n_in = 2.4e6, 1.0e5, 5_000_003, 8, 255, 1 << 18
    taps = synth.lowpass_taps(T, D)
    dx = synth.tone_plus_noise(0, n_in, seed=39, device=cuda_device)
    dt = torch.from_numpy(taps).to(cuda_device)
    n_out = g.fir_num_outputs(n_in, T, D)
    ours = torch.zeros(n_out, dtype=torch.complex64, device=cuda_device)
    ref = torch.zeros_like(ours)
    g.gsdrAdjustFrequencyFirFCLiteral(fs, f, first, D, dt, T, dx, ours, n_out, 0, None)
    ref_cuda.adjust_frequency_fir_fc(fs, f, first, D, dt, T, dx, ref, n_out)
    torch.cuda.synchronize()
    tol = 1e-5 * float(np.abs(taps).sum()) * float(dx.abs().max())
    assert float((ours - ref).abs().max()) <= tol


# ---- host-buffer pipeline ---------------------------------------------------------------------------------

@pytest.fixture
def ffma2_only():
    """Chunks / blocks far smaller than 65536 outputs against one large call: bit-identical on the FFMA2 kernels
    (the tensor-core kernel would take the large call only; tests/test_tc_gpu.py covers it with chunks on its grid)."""
    g.set_fir_tensor_cores(False)
    yield
    g.set_fir_tensor_cores(True)


def test_host_pipeline_equals_device_call_bit_exact(cuda_device, ffma2_only):
    D, T, n_in = 8, 255, 3_000_017
    taps = synth.lowpass_taps(T, D)
    x = synth.tone_plus_noise(0, n_in, seed=40)
    n_out = g.fir_num_outputs(n_in, T, D)
    xin = torch.from_numpy(x).pin_memory()
    yout = torch.zeros(n_out, dtype=torch.complex64).pin_memory()
    pipe = g.HostPipeline(0, chunkInputBytes=1 << 20, numBuffers=3)  # many chunks: exercises the block seams
    pipe.gsdrFirFCHost(D, taps, T, xin, yout, n_out)
    dy = torch.zeros(n_out, dtype=torch.complex64, device=cuda_device)
    g.gsdrFirFC(D, torch.from_numpy(taps).to(cuda_device), T, xin.to(cuda_device), dy, n_out, 0, None)
    torch.cuda.synchronize()
    assert torch.equal(yout, dy.cpu())
    z = torch.zeros(n_out, dtype=torch.complex64).pin_memory()
    pipe.gsdrAdjustFrequencyFirFCHost(2.4e6, 29520.0, 11, D, taps, T, xin, z, n_out)
    dz = torch.zeros_like(dy)
    g.gsdrAdjustFrequencyFirFC(2.4e6, 29520.0, 11, D, torch.from_numpy(taps).to(cuda_device), T, xin.to(cuda_device),
                               dz, n_out, 0, None)
    torch.cuda.synchronize()
    assert torch.equal(z, dz.cpu())
    pipe.close()


# ---- BASELINE full size, size-independent properties -------------------------------------------------------

def test_config2_full_size_properties(cuda_device):
    """64M cuComplex samples, 255 taps, decimate by 8: too big for the scalar oracle in seconds, so check
    (a) a prefix and a suffix against the oracle, (b) linearity, (c) DC gain, (d) two-shard bit-exactness."""
    D, T, n_in = 8, 255, 1 << 26
    taps = synth.lowpass_taps(T, D)
    dt = torch.from_numpy(taps).to(cuda_device)
    dx = synth.tone_plus_noise(0, n_in, seed=0x5EED0002, device=cuda_device)
    n_out = g.fir_num_outputs(n_in, T, D)
    assert n_out == 8_388_577
    dy = torch.zeros(n_out, dtype=torch.complex64, device=cuda_device)
    g.gsdrFirFC(D, dt, T, dx, dy, n_out, 0, None)
    torch.cuda.synchronize()
    tol = 1e-5 * float(np.abs(taps).sum()) * float(dx.abs().max())
    k = 4096
    head = oracle.fir("fc", D, taps, dx[: (k - 1) * D + T].cpu().numpy(), k)
    assert np.abs(dy[:k].cpu().numpy() - head).max() <= tol
    tail_in = dx[(n_out - k) * D:].cpu().numpy()
    tail = oracle.fir("fc", D, taps, tail_in, k)
    assert np.abs(dy[n_out - k:].cpu().numpy() - tail).max() <= tol
    # linearity: F(a*x) == a*F(x) for a power of two is exact in floating point
    dy2 = torch.zeros_like(dy)
    g.gsdrFirFC(D, dt, T, dx * 4.0, dy2, n_out, 0, None)
    torch.cuda.synchronize()
    assert torch.equal(dy2, dy * 4.0)
    # two time shards reproduce the bits
    parts = torch.zeros_like(dy)
    for s in range(2):
        sh = g.shard_plan_time(n_out, D, T, 0, 2, s)
        g.gsdrFirFC(D, dt, T, dx[sh.firstInput:], parts[sh.firstOutput:], sh.numOutputs, 0, None)
    torch.cuda.synchronize()
    assert torch.equal(parts, dy)
    del dy2, parts
    # the in-band tone passes with |H| ~ 1: output power ~ amp^2
    p = float((dy[64:].abs() ** 2).mean())
    assert abs(p - 0.25) < 0.02


@pytest.mark.timeout(300)
@pytest.mark.parametrize("kind,D,T", [("fc", 8, 255), ("ff", 5, 63), ("nco", 10, 255)])
def test_sample_indices_beyond_2_to_31(kind, D, T, cuda_device):
    """A capture of more than 2^31 samples (the reference narrows indices to uint32, ref: src/fir.cu:30-32,53-58):
    windows at the start, across the 2^31 and 2^32-byte marks and at the very end are recomputed by the oracle."""
    free, _ = torch.cuda.mem_get_info()
    real = kind == "ff"
    n_in = (1 << 31) + (1 << 20) + 12345 if not real else (1 << 32) + (1 << 20) + 12345
    n_out = g.fir_num_outputs(n_in, T, D)
    need = n_in * (4 if real else 8) + n_out * (4 if real else 8) + (1 << 30)
    if free < need:
        pytest.skip(f"needs {need >> 30} GiB of device memory")
    taps = synth.random_taps(T, 61)
    gen = torch.Generator(device=cuda_device).manual_seed(1234)
    if real:
        dx = torch.empty(n_in, dtype=torch.float32, device=cuda_device).uniform_(-1, 1, generator=gen)
    else:
        dx = torch.view_as_complex(torch.empty(n_in, 2, dtype=torch.float32, device=cuda_device).uniform_(-1, 1, generator=gen))
    dt = torch.from_numpy(taps).to(cuda_device)
    dy = torch.zeros(n_out, dtype=torch.float32 if real else torch.complex64, device=cuda_device)
    fs, f, first = 2.4e6, 29520.0, 2 ** 40 + 3
    if kind == "nco":
        g.gsdrAdjustFrequencyFirFC(fs, f, first, D, dt, T, dx, dy, n_out, 0, None)
    else:
        (g.gsdrFirFF if real else g.gsdrFirFC)(D, dt, T, dx, dy, n_out, 0, None)
    torch.cuda.synchronize()
    elem = 4 if real else 8
    marks = [0, ((1 << 31) // D) - 100, ((1 << 32) // elem // D) - 100, ((1 << 31) + (1 << 19)) // D, n_out - 200]
    for o0 in marks:
        o0 = max(0, min(o0, n_out - 200))
        xs = dx[o0 * D:(o0 + 199) * D + T].cpu().numpy()
        if kind == "nco":
            want = oracle.adjust_frequency_fir_fc(oracle.NCO_EXACT, fs, f, first + o0 * D, D, taps, xs, 200, f64=True)
        else:
            want = oracle.fir(kind, D, taps, xs, 200, f64=True)
        got = dy[o0:o0 + 200].cpu().numpy()
        assert np.abs(got - want).max() <= _tol(taps, xs), f"outputs {o0}.."
    del dx, dy
    torch.cuda.empty_cache()


@pytest.mark.parametrize("D,T", [(8, 255), (10, 255), (32, 1023), (4, 127), (5, 63)])
@pytest.mark.parametrize("bad", [float("inf"), float("nan")])
def test_non_finite_sample_reach_is_pinned(D, T, bad, cuda_device):
    """Documented deviation (include/gsdr/fir.h): the staged kernels multiply zero-padded taps with real samples, so a
    non-finite sample also poisons outputs whose window ends up to 16*D samples BEFORE it (0 * Inf = NaN); the
    reference (ref: src/fir.cu:57-70) touches a sample only for the outputs whose T taps cover it.  Pinned here: every
    output the reference would poison is non-finite, nothing after the sample's last window is touched, and the extra
    reach before it is at most 16 outputs."""
    n_out = 40_000
    n_in = g.fir_num_inputs(n_out, T, D)
    taps = synth.lowpass_taps(T, D)
    x = synth.tone_plus_noise(0, n_in, seed=88)
    k = 20_011 * D + 3
    x[k] = complex(bad, 0.25)
    y = _run("fc", D, taps, x, n_out, cuda_device)
    finite = np.isfinite(y.real) & np.isfinite(y.imag)
    first = max(0, -(-(k - T + 1) // D))   # smallest n with n*D + T > k
    last = k // D                          # largest n with n*D <= k
    assert not finite[first:last + 1].any(), "every output whose taps cover the sample is non-finite"
    assert finite[last + 1:].all(), "no output that starts after the sample is touched"
    assert finite[:max(0, first - 16)].all(), "the padded taps reach at most 16 outputs further back"
