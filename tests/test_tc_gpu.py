"""The tensor-core FIR (gsdr_b200/csrc/fir_tc_kernel.cuh: banded-Toeplitz FP16 GEMM, operands scaled by powers of two
and split head + remainder, all four partial products accumulated in FP32) against the oracle: BASELINE tolerance
max|err| <= 1e-5 * sum|h| * max|x| (FP32-grade, accumulation order differs from ref: src/fir.cu:57-70), ragged ends,
batched channels, dynamic range (the per-tile scale), and time shards against the unsharded call."""
import numpy as np
import pytest
import torch

import gsdr_b200 as g
from gsdr_b200 import synth
from oracle import oracle

pytestmark = pytest.mark.gpu

TC = -4       # gsdrB200SetKernelVariant: tensor-core kernel wherever its shape rules allow (tuning build)
NO_TC = -3    # automatic choice among the FFMA2 kernels only


def _tol(taps, x):
    return 1e-5 * float(np.abs(taps).sum()) * float(np.abs(x).max())


def _run_fc(D, taps, dx, n_out, dev):
    y = torch.full((n_out + 16,), float("nan"), dtype=torch.complex64, device=dev)
    g.gsdrFirFC(D, torch.from_numpy(taps).to(dev), taps.shape[0], dx, y[8:], n_out, 0, None)
    torch.cuda.synchronize()
    out = y.cpu().numpy()
    assert np.isnan(out[:8].real).all() and np.isnan(out[8 + n_out:].real).all(), "wrote outside the output"
    return out[8:8 + n_out]


@pytest.fixture(autouse=True)
def _default_settings():
    yield
    g.set_kernel_variant(-1)
    g.set_fir_tensor_cores(True)


def test_where_the_tensor_core_kernel_is_selected(cuda_device):
    """Chosen by measurement: decimation 8 with more than 128 taps, 4 with more than 64 or 16 with more
    than 256, at least 65536 outputs per channel; gsdrB200SetFirTensorCores(0) and the tuning override move the line."""
    tc_id = g.num_kernel_variants()
    assert g.describe_kernel(0, 8, 255, 8_388_577).variant == tc_id      # BASELINE config 2
    assert g.describe_kernel(0, 8, 160, 65_536).variant == tc_id
    assert g.describe_kernel(0, 8, 129, 65_536).variant == tc_id
    assert g.describe_kernel(0, 8, 128, 8_388_577).variant != tc_id
    assert g.describe_kernel(0, 8, 255, 65_535).variant != tc_id         # below the size shards are aligned from
    assert g.describe_kernel(0, 8, 127, 8_388_577).variant != tc_id      # the FFMA2 kernel is HBM-bound there
    assert g.describe_kernel(0, 8, 265, 8_388_577).variant != tc_id      # a window must fit two segments
    assert g.describe_kernel(0, 4, 127, 1_048_545).variant == tc_id      # BASELINE config 4 (per channel)
    assert g.describe_kernel(0, 4, 65, 65_536).variant == tc_id
    assert g.describe_kernel(0, 4, 260, 65_536).variant == tc_id
    assert g.describe_kernel(0, 4, 64, 8_388_577).variant != tc_id       # the FFMA2 kernel wins up to 64 taps
    assert g.describe_kernel(0, 4, 261, 8_388_577).variant != tc_id
    assert g.describe_kernel(0, 16, 511, 8_388_577).variant == tc_id
    assert g.describe_kernel(0, 16, 257, 65_536).variant == tc_id
    assert g.describe_kernel(0, 16, 256, 8_388_577).variant != tc_id     # the FFMA2 kernel wins up to 256 taps
    assert g.describe_kernel(0, 16, 529, 8_388_577).variant != tc_id
    assert g.describe_kernel(4, 8, 255, 8_388_577).variant != tc_id      # fused NCO: FFMA2 kernels only
    assert g.set_fir_tensor_cores(False) is True
    assert g.describe_kernel(0, 8, 255, 8_388_577).variant != tc_id
    assert g.set_fir_tensor_cores(True) is False
    g.set_kernel_variant(TC)
    assert g.describe_kernel(0, 8, 255, 100_000).variant == tc_id
    assert g.describe_kernel(0, 4, 33, 100_000).variant == tc_id
    assert g.describe_kernel(0, 32, 1023, 8_388_577).variant != tc_id    # segments would not fit shared memory
    assert g.describe_kernel(0, 10, 255, 8_388_577).variant != tc_id     # k-steps of 16 would straddle tap rows
    g.set_kernel_variant(NO_TC)
    assert g.describe_kernel(0, 8, 255, 8_388_577).variant != tc_id


@pytest.mark.parametrize("D,T", [(8, 255), (4, 127), (16, 511)])
def test_default_path_on_a_large_call_and_the_switch(D, T, cuda_device):
    """The release library's own choice (no override): a call over the size line runs on the tensor cores, agrees with
    the oracle on windows and with the FFMA2 kernels everywhere; shards on the plan's 1024-output grid reproduce the
    unsharded bits; with the switch off the call is bit-identical to the FFMA2 result."""
    n_in = 20_000_003
    taps = synth.random_taps(T, 5)
    x = synth.tone_plus_noise(0, n_in, seed=310, device=cuda_device)
    n_out = g.fir_num_outputs(n_in, T, D)
    dt = torch.from_numpy(taps).to(cuda_device)
    assert g.describe_kernel(0, D, T, n_out).variant == g.num_kernel_variants()
    y = torch.zeros(n_out, dtype=torch.complex64, device=cuda_device)
    g.gsdrFirFC(D, dt, T, x, y, n_out, 0, None)
    g.set_fir_tensor_cores(False)
    y_ffma = torch.zeros_like(y)
    g.gsdrFirFC(D, dt, T, x, y_ffma, n_out, 0, None)
    g.set_fir_tensor_cores(True)
    torch.cuda.synchronize()
    tol = 1e-5 * float(np.abs(taps).sum()) * float(x.abs().max())
    assert not torch.equal(y, y_ffma)                       # it really is another kernel
    assert float((y - y_ffma).abs().max()) <= 0.2 * tol
    for o0 in (0, 2047, 1_234_567, n_out - 3000):
        xs = x[o0 * D:(o0 + 2999) * D + T].cpu().numpy()
        want = oracle.fir("fc", D, taps, xs, 3000, f64=True)
        assert np.abs(y[o0:o0 + 3000].cpu().numpy() - want).max() <= tol
    for shards in (2, 3, 7):
        parts = torch.zeros_like(y)
        for s in range(shards):
            sh = g.shard_plan_time(n_out, D, T, 0, shards, s)
            assert s == 0 or sh.firstOutput % 2048 == 0
            xs = x[sh.firstInput: sh.firstInput + sh.numInputs].clone()
            g.gsdrFirFC(D, dt, T, xs, parts[sh.firstOutput: sh.firstOutput + sh.numOutputs], sh.numOutputs, 0, None)
        torch.cuda.synchronize()
        assert torch.equal(parts, y), f"{shards} shards"
    # the host pipeline cuts its chunks on the same grid
    pipe = g.HostPipeline(0, 32 << 20, 3)
    hx = x.cpu().numpy()
    hy = np.zeros(n_out, dtype=np.complex64)
    pipe.gsdrFirFCHost(D, taps, T, hx, hy, n_out)
    pipe.close()
    assert hy.tobytes() == y.cpu().numpy().tobytes()


@pytest.mark.parametrize("D,T", [(8, 255), (4, 127), (16, 511)])
def test_non_finite_sample_reach_on_tensor_cores_is_pinned(D, T, cuda_device):
    """ref: src/fir.cu:57-70 multiplies only the taps that overlap a sample, so an Inf/NaN there reaches ceil(T/D)
    outputs.  The tensor-core kernel multiplies the zeros of the band too (0 * Inf = NaN) and scales per segment of
    256 samples: the value reaches every output of the windows (32 outputs; 64 at decimation 4) that read its segment
    and nothing else."""
    n_out = 1_200_000
    n_in = g.fir_num_inputs(n_out, T, D)
    taps = synth.lowpass_taps(T, D)
    for bad, k in ((np.inf, (n_in * 3) // 5 + 123), (np.nan, 700_001)):
        x = synth.tone_plus_noise(0, n_in, seed=311)
        x[k] = bad
        y = _run_fc(D, taps, torch.from_numpy(x).to(cuda_device), n_out, cuda_device)
        nonfinite = ~(np.isfinite(y.real) & np.isfinite(y.imag))
        idx = np.flatnonzero(nonfinite)
        S = 64 if D == 4 else 32            # outputs per window
        seg = k // (S * D)                  # segment of S*D samples; window w reads segments w and w+1
        lo, hi = max(0, seg - 1) * S, (seg + 1) * S
        assert idx.min() >= lo and idx.max() < hi, (idx.min(), idx.max(), lo, hi)
        ref_lo, ref_hi = (k - T) // D + 1, k // D       # what the reference would touch
        assert nonfinite[max(ref_lo, 0):ref_hi + 1].all()
        clean = np.ones(n_out, dtype=bool)
        clean[lo:hi] = False
        xz = x.copy()
        xz[k] = 0
        want = _run_fc(D, taps, torch.from_numpy(xz).to(cuda_device), n_out, cuda_device)
        assert np.array_equal(y[clean], want[clean])


@pytest.mark.parametrize("D,T,n_out", [
    (8, 255, 200_000), (8, 255, 65_536), (8, 255, 66_561), (8, 264, 70_001), (8, 193, 80_003), (8, 1, 70_000),
    (4, 127, 150_001), (4, 132, 66_000), (4, 97, 99_999), (16, 511, 70_000), (16, 528, 65_537), (16, 400, 66_666)])
def test_tensor_core_fir_against_oracle(D, T, n_out, cuda_device):
    taps = synth.random_taps(T, 100 + D)           # asymmetric: catches tap-order and band-offset mistakes
    n_in = g.fir_num_inputs(n_out, T, D)
    x = synth.tone_plus_noise(0, n_in, seed=200 + T)
    g.set_kernel_variant(TC)
    assert g.describe_kernel(0, D, T, n_out).variant == g.num_kernel_variants()
    y = _run_fc(D, taps, torch.from_numpy(x).to(cuda_device), n_out, cuda_device)
    want = oracle.fir("fc", D, taps, x, n_out, f64=True)
    err = np.abs(y - want).max()
    assert err <= _tol(taps, x), f"max|err| {err} > {_tol(taps, x)}"
    # FP32-grade, not TF32-grade: the split must leave an error comparable to the FFMA2 kernel's
    g.set_kernel_variant(NO_TC)
    y2 = _run_fc(D, taps, torch.from_numpy(x).to(cuda_device), n_out, cuda_device)
    err2 = np.abs(y2 - want).max()
    assert err <= max(8.0 * err2, 0.2 * _tol(taps, x))


def test_impulse_and_dc_known_answers_on_tensor_cores(cuda_device):
    D, T, n_out = 8, 255, 70_000
    n_in = g.fir_num_inputs(n_out, T, D)
    taps = (np.arange(T, dtype=np.float32) + 1.0) / 64.0
    g.set_kernel_variant(TC)
    for k in (0, 7, 8, 254, 255, 256, 100_003, n_in - 1):
        x = np.zeros(n_in, dtype=np.complex64)
        x[k] = 3.0 - 2.0j
        y = _run_fc(D, taps, torch.from_numpy(x).to(cuda_device), n_out, cuda_device)
        want = np.zeros(n_out, dtype=np.complex64)
        for n in range(max(0, (k - T) // D), min(n_out, k // D + 1)):
            i = k - n * D
            if 0 <= i < T:
                want[n] = taps[i] * x[k]
        assert np.array_equal(y, want), f"impulse at {k}"   # exactly representable products: the split loses nothing
    ones = np.ones(n_in, dtype=np.complex64) * (1.0 + 0.5j)
    y = _run_fc(D, taps, torch.from_numpy(ones).to(cuda_device), n_out, cuda_device)
    assert np.abs(y - taps.astype(np.float64).sum() * (1.0 + 0.5j)).max() <= _tol(taps, ones)


def test_batched_channels_on_tensor_cores_equal_single_calls_bit_exact(cuda_device):
    D, T, C, n_in = 4, 127, 6, 300_000
    n_out = g.fir_num_outputs(n_in, T, D)
    taps = synth.lowpass_taps(T, D)
    xs = synth.tone_plus_noise(0, C * n_in, seed=301, device=cuda_device).view(C, n_in)
    dt = torch.from_numpy(taps).to(cuda_device)
    g.set_kernel_variant(TC)
    yb = torch.zeros((C, n_out + 3), dtype=torch.complex64, device=cuda_device)
    g.gsdrFirFCBatched(D, dt, T, 0, xs, n_in, yb, n_out + 3, n_out, C, 0, None)
    for c in range(C):
        one = torch.zeros(n_out, dtype=torch.complex64, device=cuda_device)
        g.gsdrFirFC(D, dt, T, xs[c], one, n_out, 0, None)
        torch.cuda.synchronize()
        assert torch.equal(one, yb[c, :n_out])
    assert float(yb[:, n_out:].abs().max()) == 0.0
    want = oracle.fir("fc", D, taps, xs[C - 1].cpu().numpy(), n_out, f64=True)
    assert np.abs(yb[C - 1, :n_out].cpu().numpy() - want).max() <= _tol(taps, xs[C - 1].cpu().numpy())


@pytest.mark.parametrize("shards", [2, 3, 8])
def test_tensor_core_time_shards_reproduce_the_unsharded_call(shards, cuda_device):
    """Shards start at arbitrary output indices, so their tiles (scale factor, k-step alignment) differ from the
    unsharded call's: the results agree to a tenth of the oracle tolerance, not bit for bit (the FFMA2 kernels are
    the bit-reproducible ones)."""
    D, T, n_in = 8, 255, 3_000_017
    taps = synth.lowpass_taps(T, D)
    x = synth.tone_plus_noise(0, n_in, seed=302, device=cuda_device)
    n_out = g.fir_num_outputs(n_in, T, D)
    dt = torch.from_numpy(taps).to(cuda_device)
    g.set_kernel_variant(TC)
    whole = torch.zeros(n_out, dtype=torch.complex64, device=cuda_device)
    g.gsdrFirFC(D, dt, T, x, whole, n_out, 0, None)
    parts = torch.zeros_like(whole)
    for s in range(shards):
        sh = g.shard_plan_time(n_out, D, T, 0, shards, s)
        xs = x[sh.firstInput: sh.firstInput + sh.numInputs].clone()   # a fresh, 16-byte aligned shard buffer
        g.gsdrFirFC(D, dt, T, xs, parts[sh.firstOutput: sh.firstOutput + sh.numOutputs], sh.numOutputs, 0, None)
    torch.cuda.synchronize()
    tol = 1e-6 * float(np.abs(taps).sum()) * float(x.abs().max())
    assert float((whole - parts).abs().max()) <= tol


@pytest.mark.parametrize("scale", [1e-30, 3e-9, 1.0, 7e4, 1e12, 2e30])
def test_tensor_core_dynamic_range(scale, cuda_device):
    """FP16 operands only work because every tile is scaled into [0.5, 1) first: amplitudes far outside FP16's range,
    a 2^40 step in level between neighbouring tiles, and taps scaled the other way."""
    D, T, n_out = 8, 255, 150_000
    n_in = g.fir_num_inputs(n_out, T, D)
    taps = (synth.random_taps(T, 7) / np.float32(scale)).astype(np.float32) if scale < 1e20 else synth.random_taps(T, 7)
    x = synth.tone_plus_noise(0, n_in, seed=304).astype(np.complex64) * np.float32(scale)
    x[n_in // 2:] *= np.float32(2.0 ** -40 if scale > 1 else 2.0 ** 40)   # later tiles: a very different level
    g.set_kernel_variant(TC)
    y = _run_fc(D, taps, torch.from_numpy(x).to(cuda_device), n_out, cuda_device)
    want = oracle.fir("fc", D, taps, x, n_out, f64=True)
    # judged per tile-sized block against that block's own level: the scale is per tile
    h1 = float(np.abs(taps).sum())
    for lo in range(0, n_out, 1024):
        hi = min(n_out, lo + 1024)
        seg = x[lo * D: (hi - 1) * D + T]
        m = float(max(np.abs(seg.real).max(), np.abs(seg.imag).max()))
        assert np.abs(y[lo:hi] - want[lo:hi]).max() <= 1e-5 * h1 * m, f"outputs {lo}..{hi}"


def test_tensor_core_all_zero_and_subnormal_tiles(cuda_device):
    D, T, n_out = 8, 255, 70_000
    n_in = g.fir_num_inputs(n_out, T, D)
    taps = synth.random_taps(T, 9)
    x = np.zeros(n_in, dtype=np.complex64)
    x[200_000:200_100] = np.float32(1e-41) * (1 + 1j)      # subnormal samples in an otherwise silent capture
    g.set_kernel_variant(TC)
    y = _run_fc(D, taps, torch.from_numpy(x).to(cuda_device), n_out, cuda_device)
    want = oracle.fir("fc", D, taps, x, n_out, f64=True)
    assert np.isfinite(y.real).all() and np.isfinite(y.imag).all()
    assert np.abs(y - want).max() <= max(1e-5 * float(np.abs(taps).sum()) * 1.5e-41, 2e-45)
    assert not y[:20_000].any()


def test_unaligned_input_falls_back_and_still_matches(cuda_device):
    g.set_kernel_variant(TC)
    D, T, n_out = 8, 255, 100_000
    n_in = g.fir_num_inputs(n_out, T, D)
    taps = synth.lowpass_taps(T, D)
    x = synth.tone_plus_noise(0, n_in + 1, seed=303)
    dx = torch.from_numpy(x).to(cuda_device)[1:]   # 8-byte aligned only: bulk copies are not possible
    y = _run_fc(D, taps, dx, n_out, cuda_device)
    want = oracle.fir("fc", D, taps, x[1:], n_out, f64=True)
    assert np.abs(y - want).max() <= _tol(taps, x)


def test_tensor_core_kernel_is_deterministic_run_to_run(cuda_device):
    """Persistent CTAs, mbarrier hand-overs between three kinds of warps, in-place conversion of the staged samples: a
    missing ordering would show up as run-to-run differences.  40 launches of a ragged shape, identical bits."""
    D, T, n_out = 8, 255, 700_001
    n_in = g.fir_num_inputs(n_out, T, D)
    taps = torch.from_numpy(synth.random_taps(T, 21)).to(cuda_device)
    x = synth.tone_plus_noise(0, n_in, seed=320, device=cuda_device)
    g.set_kernel_variant(TC)
    first = torch.zeros(n_out, dtype=torch.complex64, device=cuda_device)
    g.gsdrFirFC(D, taps, T, x, first, n_out, 0, None)
    for _ in range(40):
        y = torch.full((n_out,), float("nan"), dtype=torch.complex64, device=cuda_device)
        g.gsdrFirFC(D, taps, T, x, y, n_out, 0, None)
        torch.cuda.synchronize()
        assert torch.equal(y, first)


def test_tensor_core_random_shapes_against_the_ffma2_kernels(cuda_device):
    """Random (decimation, taps, outputs, channels): the two kernel families agree within a fifth of the oracle
    tolerance everywhere, and nothing is written outside the outputs.  GSDR_TC_FUZZ_CASES widens it."""
    import os

    rng = np.random.default_rng(int(os.environ.get("GSDR_TC_FUZZ_SEED", "7")))
    for case in range(int(os.environ.get("GSDR_TC_FUZZ_CASES", "24"))):
        D = int(rng.choice([4, 8, 8, 8, 16]))
        T = int(rng.integers(1, 33 * D + 1))
        n_out = int(rng.integers(1, 200_000))
        C = int(rng.choice([1, 1, 1, 3]))
        n_in = g.fir_num_inputs(n_out, T, D)
        n_in += n_in & 1            # even channel stride: the batched bulk-copy path needs 16-byte aligned channels
        taps = synth.random_taps(T, 1000 + case)
        xs = synth.tone_plus_noise(0, C * n_in, seed=2000 + case, device=cuda_device).view(C, n_in)
        dt = torch.from_numpy(taps).to(cuda_device)
        outs = {}
        for name, v in (("tc", TC), ("ffma2", NO_TC)):
            g.set_kernel_variant(v)
            if name == "tc":
                assert g.describe_kernel(0, D, T, n_out).variant == g.num_kernel_variants(), (D, T, n_out)
            y = torch.full((C, n_out + 9), float("nan"), dtype=torch.complex64, device=cuda_device)
            g.gsdrFirFCBatched(D, dt, T, 0, xs, n_in, y, n_out + 9, n_out, C, 0, None)
            torch.cuda.synchronize()
            assert bool(torch.isnan(y[:, n_out:].real).all()), f"case {case}: wrote past the outputs"
            outs[name] = y[:, :n_out]
        tol = 1e-5 * float(np.abs(taps).sum()) * float(xs.abs().max())
        err = float((outs["tc"] - outs["ffma2"]).abs().max())
        assert err <= 0.2 * tol + 1e-30, f"case {case}: D={D} T={T} n_out={n_out} C={C}: {err} > {0.2 * tol}"


def test_tensor_core_call_in_a_cuda_graph_and_on_two_streams(cuda_device):
    """include/gsdr/fir.h: work is enqueued on cudaStream only and can be captured into a CUDA graph.  Two streams
    running the kernel at the same time share the SM's tensor memory (every CTA allocates 128 of its 512 columns and
    waits for them if another kernel's CTAs hold them): same bits as the calls run one after the other."""
    D, T, n_out = 8, 255, 300_000
    n_in = g.fir_num_inputs(n_out, T, D)
    taps = torch.from_numpy(synth.random_taps(T, 31)).to(cuda_device)
    xs = [synth.tone_plus_noise(0, n_in, seed=330 + i, device=cuda_device) for i in range(2)]
    want = [torch.zeros(n_out, dtype=torch.complex64, device=cuda_device) for _ in range(2)]
    for i in range(2):
        g.gsdrFirFC(D, taps, T, xs[i], want[i], n_out, 0, None)
    torch.cuda.synchronize()
    assert g.describe_kernel(0, D, T, n_out).variant == g.num_kernel_variants()   # the release library's own choice
    # two streams at once, several rounds
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    got = [torch.zeros_like(w) for w in want]
    for _ in range(10):
        for i in range(2):
            g.gsdrFirFC(D, taps, T, xs[i], got[i], n_out, 0, streams[i])
    for s in streams:
        s.synchronize()
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
    # captured into a graph and replayed
    out = torch.zeros_like(want[0])
    cap = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(cap):
        with torch.cuda.graph(graph, stream=cap):
            g.gsdrFirFC(D, taps, T, xs[0], out, n_out, 0, cap)
            g.gsdrFirFC(D, taps, T, xs[0], out, n_out, 0, cap)
    out.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, want[0])


def test_tensor_core_components_keep_their_own_precision(cuda_device):
    """re and im are scaled separately: an imaginary part seven orders of magnitude below the real part (and the other
    way round) comes out with the tolerance of ITS OWN level, as on the FP32 pipes."""
    D, T, n_out = 8, 255, 150_000
    n_in = g.fir_num_inputs(n_out, T, D)
    taps = synth.random_taps(T, 41)
    h1 = float(np.abs(taps).sum())
    g.set_kernel_variant(TC)
    for small_im in (True, False):
        x = synth.tone_plus_noise(0, n_in, seed=340)
        x = (x.real + 1j * x.imag * np.float32(1e-7)) if small_im else (x.real * np.float32(1e-7) + 1j * x.imag)
        x = x.astype(np.complex64)
        y = _run_fc(D, taps, torch.from_numpy(x).to(cuda_device), n_out, cuda_device)
        want = oracle.fir("fc", D, taps, x, n_out, f64=True)
        assert np.abs(y.real - want.real).max() <= 1e-5 * h1 * float(np.abs(x.real).max())
        assert np.abs(y.imag - want.imag).max() <= 1e-5 * h1 * float(np.abs(x.imag).max())
