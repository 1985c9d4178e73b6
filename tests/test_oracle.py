"""CPU tests of the oracle itself (known-answer tests the reference's suite lacks) and of the
oracle against the golden vectors captured from the reference's own CUDA kernels (tests/golden/)."""
import math
from pathlib import Path

import numpy as np
import pytest

from gsdr_b200 import synth
from oracle import oracle

GOLDEN = Path(__file__).resolve().parent / "golden"

# (tapCount, numOutputs) edge list of the reference's EdgeCasesTest (ref: tests/test_fir.cpp:261-263)
EDGE_SHAPES = [(1, 1), (2, 1), (1, 2), (16, 8), (8, 16), (31, 15), (32, 16), (33, 17)]
KINDS = ["ff", "fc", "cc", "cf"]


def _rand(kind_is_complex, n, seed):
    rng = np.random.default_rng(seed)
    if kind_is_complex:
        return (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
    return rng.uniform(-1, 1, n).astype(np.float32)


def _fma32(a, b, c):
    """Correctly rounded float32 fma via exact rational arithmetic (math.fma needs Python 3.13)."""
    from fractions import Fraction

    exact = Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c))
    cand = np.float32(float(exact))
    best = None
    for v in (np.nextafter(cand, np.float32(-np.inf)), cand, np.nextafter(cand, np.float32(np.inf))):
        d = abs(Fraction(float(v)) - exact)
        even = (int(np.float32(v).view(np.uint32)) & 1) == 0
        if best is None or d < best[0] or (d == best[0] and even):
            best = (d, v)
    return np.float32(best[1])


def _taps_complex(kind):
    return kind[0] == "c"


def _in_complex(kind):
    return kind[1] == "c"


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("D", [1, 2, 3, 8])
def test_impulse_response(kind, D):
    """Impulse at in[k] => out[n] = taps[k - n*D] where 0 <= k - n*D < T, else 0 (correlation semantics,
    ref: src/fir.cu:57-70 — NOT the out[i] == taps[i] the reference's ImpulseResponseTest assumes)."""
    T, n_out = 17, 9
    n_in = (n_out - 1) * D + T
    taps = _rand(_taps_complex(kind), T, 1)
    for k in [0, 1, T - 1, T, n_in - 1]:
        x = np.zeros(n_in, dtype=np.complex64 if _in_complex(kind) else np.float32)
        x[k] = 1.0
        y = oracle.fir(kind, D, taps, x, n_out)
        for n in range(n_out):
            i = k - n * D
            want = taps[i] if 0 <= i < T else 0.0
            assert y[n] == np.asarray(want, dtype=y.dtype), (k, n)


@pytest.mark.parametrize("kind", KINDS)
def test_dc_gain(kind):
    T, D, n_out = 63, 4, 50
    taps = _rand(_taps_complex(kind), T, 2)
    x = np.ones((n_out - 1) * D + T, dtype=np.complex64 if _in_complex(kind) else np.float32)
    y = oracle.fir(kind, D, taps, x, n_out, f64=True)
    assert np.allclose(y, taps.astype(np.complex128).sum(), rtol=0, atol=1e-12)


def test_tone_response_fc():
    """A complex tone at w comes out scaled by H(e^{jw}) = sum_i h[i] e^{jwi} and decimated."""
    T, D, n_out = 255, 8, 64
    h = synth.lowpass_taps(T, D)
    w = 2 * math.pi * 0.0123
    n = np.arange((n_out - 1) * D + T)
    x = np.exp(1j * w * n).astype(np.complex64)
    H = np.sum(h.astype(np.float64) * np.exp(1j * w * np.arange(T)))
    want = H * np.exp(1j * w * D * np.arange(n_out))
    y = oracle.fir("fc", D, h, x, n_out, f64=True)
    assert np.abs(y - want).max() < 5e-6  # limited by the float32 rounding of x


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("T,n_out", EDGE_SHAPES)
@pytest.mark.parametrize("D", [1, 5])
def test_edge_shapes_float_vs_truth(kind, T, n_out, D):
    taps = _rand(_taps_complex(kind), T, 10 + T)
    x = _rand(_in_complex(kind), (n_out - 1) * D + T, 20 + n_out)
    y = oracle.fir(kind, D, taps, x, n_out)
    truth = oracle.fir(kind, D, taps, x, n_out, f64=True)
    second = oracle.fir_numpy_f64(D, taps, x, n_out)
    assert np.abs(truth - second).max() <= 1e-12
    tol = 1e-5 * float(np.abs(taps).sum()) * float(np.abs(x).max())
    assert np.abs(y.astype(truth.dtype) - truth).max() <= tol


def test_zero_taps_and_zero_outputs():
    """ref: tests/test_fir.cpp:249-257 (ZeroTapsTest): zero taps writes zeros; zero outputs writes nothing."""
    x = np.ones(16, dtype=np.complex64)
    assert np.all(oracle.fir("fc", 2, np.zeros(0, np.float32), x, 4) == 0)
    assert oracle.fir("fc", 2, np.ones(3, np.float32), x, 0).shape == (0,)


def test_order_faithful_chain_is_sequential_fmaf():
    """The float oracle is exactly the ascending single-accumulator fmaf chain (ref: src/fir.cu:67-70)."""
    T, D, n_out = 40, 3, 7
    h = _rand(False, T, 5)
    x = _rand(False, (n_out - 1) * D + T, 6)
    y = oracle.fir("ff", D, h, x, n_out)
    for n in range(n_out):
        acc = np.float32(0)
        for i in range(T):
            acc = _fma32(x[n * D + i], h[i], acc)
        assert y[n] == acc


def test_threaded_matches_scalar():
    T, D = 255, 8
    h = synth.lowpass_taps(T, D)
    x = synth.tone_plus_noise(0, 1 << 16, seed=3)
    a = oracle.fir("fc", D, h, x)
    b = oracle.fir("fc", D, h, x, threads=5)
    assert a.tobytes() == b.tobytes()


def test_num_outputs_formula():
    for n_in in range(0, 70):
        for T in range(1, 9):
            for D in range(1, 6):
                n = oracle.num_outputs(n_in, T, D)
                if n:
                    assert (n - 1) * D + T <= n_in < n * D + T
                else:
                    assert n_in < T


# ---- NCO ---------------------------------------------------------------------------------------------------

def test_nco_literal_phase_matches_hand_arithmetic():
    """ref: src/adjustFrequency.cu:23,35-43 evaluated step by step with numpy float32 scalars."""
    fs, f = np.float32(2.4e6), np.float32(1.0e5)
    for idx in [0, 1, 7, 23, 2_399_999, 2_400_000, 2_400_001, 16_777_217, 0xFFFFFFFF]:
        period = np.float32(1.0) / f
        t = np.float32(np.fmod(np.float32(idx), fs)) / fs
        u = np.float32(np.fmod(t, period))
        want = np.float32(u * np.float32(2.0))
        assert oracle.nco_literal_theta_div_pi(float(f), idx, float(fs)) == want


def test_nco_literal_first_index_reduction():
    """ref: src/fm.cu:202."""
    assert oracle.reduce_first_sample_index(5_000_000, 2.4e6) == 200_000
    assert oracle.reduce_first_sample_index(123, 2.4e6) == 123


def test_nco_exact_phase_step_and_shard_continuity():
    step = oracle.nco_exact_phase_step(100e3, 2.4e6)
    assert step == round((100e3 / 2.4e6) * 2 ** 64) & (2 ** 64 - 1)
    neg = oracle.nco_exact_phase_step(-100e3, 2.4e6)
    assert (step + neg) % 2 ** 64 == 0
    # phase(n0 + k) == phase(n0) + k*step  (mod 2^64): what makes time shards exact
    for n0 in [0, 12345, 2 ** 40 + 17]:
        for k in [0, 1, 1000, 2 ** 33]:
            a = ((n0 + k) * step) % 2 ** 64
            b = ((n0 * step) % 2 ** 64 + k * step) % 2 ** 64
            assert a == b
            q = (a >> 32)
            q = q - 2 ** 32 if q >= 2 ** 31 else q
            assert oracle.nco_exact_phase_q31(step, n0 + k) == q


def test_nco_exact_mix_then_fir_equals_fir_of_mixed():
    fs, f, first, D, T = 2.4e6, -310e3, 987654321, 4, 63
    h = synth.lowpass_taps(T, D)
    x = synth.tone_plus_noise(0, 4096, seed=9, tone_cycles_per_sample=310e3 / 2.4e6)
    y = oracle.adjust_frequency_fir_fc(oracle.NCO_EXACT, fs, f, first, D, h, x, f64=True)
    step = oracle.nco_exact_phase_step(f, fs)
    n = np.arange(x.shape[0], dtype=object) + first
    v = np.array([np.float32(oracle.nco_exact_phase_q31(step, int(k))) * np.float32(2.0 ** -31) for k in n], dtype=np.float64)
    mixed = x.astype(np.complex128) * np.exp(1j * math.pi * v)
    want = oracle.fir_numpy_f64(D, h, mixed)
    assert np.abs(y - want).max() < 1e-12
    # the tone at +310 kHz is brought to DC: output magnitude ~ amp * sum(h) = 0.5
    assert abs(np.abs(y[8:]).mean() - 0.5) < 0.02
    yf = oracle.adjust_frequency_fir_fc(oracle.NCO_EXACT, fs, f, first, D, h, x)
    tol = 1e-5 * float(np.abs(h).sum()) * float(np.abs(x).max())
    assert np.abs(yf - y).max() <= tol


def test_quad_fm_demod_of_tone_is_constant():
    w = 0.3
    x = np.exp(1j * w * np.arange(100)).astype(np.complex64)
    y = oracle.quad_fm_demod(x, gain=2.0)
    assert np.allclose(y, 2.0 * w, atol=1e-5)


# ---- golden vectors captured from the reference's own CUDA kernels -----------------------------------------

def _golden_files():
    return sorted(GOLDEN.glob("ref_*.npz"))


@pytest.mark.parametrize("path", _golden_files(), ids=lambda p: p.stem)
def test_oracle_reproduces_reference_cuda_bits(path):
    """tests/golden/ref_*.npz hold inputs and the outputs of the reference's kernels (oracle/_ref, built from
    /root/reference/src/fir.cu for sm_100) run on a B200 by tests/golden/make_golden.py.  FF/FC/CF/CC are
    deterministic fixed-order FMA chains, so the oracle must match bit for bit."""
    z = np.load(path)
    kind = str(z["kind"])
    if kind in KINDS:
        y = oracle.fir(kind, int(z["decimation"]), z["taps"], z["input"], int(z["num_outputs"]))
        assert y.tobytes() == z["output"].tobytes(), f"{path.name}: oracle differs from the reference's CUDA output"
    elif kind == "adjust_literal":
        y = oracle.adjust_frequency_fir_fc(oracle.NCO_LITERAL, float(z["sample_rate"]), float(z["frequency_shift"]),
                                           int(z["first_sample_index"]), int(z["decimation"]), z["taps"], z["input"],
                                           int(z["num_outputs"]))
        tol = 1e-5 * float(np.abs(z["taps"]).sum()) * float(np.abs(z["input"]).max())
        # sincospif differs by <= 2 ulp between CUDA and libm, so this one is tolerance-, not bit-, pinned
        assert np.abs(y - z["output"]).max() <= tol
    elif kind == "quad_fm":
        y = oracle.quad_fm_demod(z["input"], float(z["gain"]), int(z["num_outputs"]))
        # ref: src/quad_demod.cu:23-37 — atan2f differs by a few ulp between CUDA and libm
        assert np.abs(y - z["output"]).max() <= float(z["gain"]) * 4e-7 * math.pi
    else:
        pytest.fail(f"unknown golden kind {kind}")


def test_golden_vectors_present():
    assert _golden_files(), "tests/golden/ref_*.npz missing — run tests/golden/make_golden.py on a GPU box"
