"""Deterministic synthetic signals and filter taps for tests and benchmarks.

Signals are "complex tone + uniform noise" generated from a counter-based hash of the GLOBAL sample index, so any
shard of a capture regenerates exactly the bits the unsharded capture has.  Works on numpy arrays
(host) and torch tensors (device) with identical integer arithmetic; the sin/cos of the tone differ between CPU
and GPU in the last ulp, so parity tests copy the device buffer back rather than regenerating it.
"""
from __future__ import annotations

import math

import numpy as np

_M64 = (1 << 64) - 1


def _s64(v: int) -> int:
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


_C1 = _s64(0x9E3779B97F4A7C15)
_C2 = _s64(0xBF58476D1CE4E5B9)
_C3 = _s64(0x94D049BB133111EB)


def _lsr(z, k):
    """Logical shift right of an int64 array/tensor."""
    return (z >> k) & ((1 << (64 - k)) - 1)


def hash_uniform(idx, seed: int):
    """splitmix64 of (idx + seed*golden) -> float32 uniform in [-1, 1).  idx: int64 numpy array or torch tensor."""
    is_np = isinstance(idx, np.ndarray)
    if is_np:
        with np.errstate(over="ignore"):
            z = idx.astype(np.int64) + np.int64(_s64(seed * 0x9E3779B97F4A7C15))
            z = z + np.int64(_C1)
            z = (z ^ _lsr(z, 30)) * np.int64(_C2)
            z = (z ^ _lsr(z, 27)) * np.int64(_C3)
            z = z ^ _lsr(z, 31)
            top = _lsr(z, 40).astype(np.float32)
        return top * np.float32(1.0 / (1 << 23)) - np.float32(1.0)
    import torch

    z = idx.to(torch.int64) + _s64(seed * 0x9E3779B97F4A7C15)
    z = z + _C1
    z = (z ^ _lsr(z, 30)) * _C2
    z = (z ^ _lsr(z, 27)) * _C3
    z = z ^ _lsr(z, 31)
    top = _lsr(z, 40).to(torch.float32)
    return top * (1.0 / (1 << 23)) - 1.0


def tone_plus_noise(first: int, count: int, seed: int, tone_cycles_per_sample: float = 0.0123, amp: float = 0.5,
                    sigma: float = 0.1, device=None, real: bool = False, chunk: int = 1 << 24):
    """x[n] = amp*exp(j*2*pi*f0*n) + sigma*(u1[n] + j*u2[n]) for n in [first, first+count).

    device=None -> numpy complex64 (float32 if real); otherwise a torch tensor on `device`.
    The tone phase uses a 32-bit integer accumulator so it is exact for any n.
    """
    step = int(round((tone_cycles_per_sample % 1.0) * (1 << 32)))
    if device is None:
        n = np.arange(first, first + count, dtype=np.int64)
        ph = ((n * step) & 0xFFFFFFFF).astype(np.float64) * (2.0 * math.pi / (1 << 32))
        re = (amp * np.cos(ph)).astype(np.float32) + np.float32(sigma) * hash_uniform(2 * n, seed)
        if real:
            return re.astype(np.float32)
        im = (amp * np.sin(ph)).astype(np.float32) + np.float32(sigma) * hash_uniform(2 * n + 1, seed)
        return (re + 1j * im).astype(np.complex64)
    import torch

    out = torch.empty(count, dtype=torch.float32 if real else torch.complex64, device=device)
    view = out if real else torch.view_as_real(out)
    for c0 in range(0, count, chunk):
        c1 = min(count, c0 + chunk)
        n = torch.arange(first + c0, first + c1, dtype=torch.int64, device=device)
        ph = ((n * step) & 0xFFFFFFFF).to(torch.float64) * (2.0 * math.pi / (1 << 32))
        re = (amp * torch.cos(ph)).to(torch.float32) + sigma * hash_uniform(2 * n, seed)
        if real:
            view[c0:c1] = re
        else:
            im = (amp * torch.sin(ph)).to(torch.float32) + sigma * hash_uniform(2 * n + 1, seed)
            view[c0:c1, 0] = re
            view[c0:c1, 1] = im
    return out


def lowpass_taps(tap_count: int, decimation: int, cutoff: float = 0.4) -> np.ndarray:
    """Hamming-windowed sinc low-pass, fc = cutoff/decimation cycles/sample, sum(h) = 1, symmetric, float32."""
    if tap_count <= 0:
        return np.zeros(0, dtype=np.float32)
    fc = cutoff / max(decimation, 1)
    k = np.arange(tap_count, dtype=np.float64)
    mid = (tap_count - 1) / 2.0
    h = 2.0 * fc * np.sinc(2.0 * fc * (k - mid))
    if tap_count > 1:
        h *= 0.54 - 0.46 * np.cos(2.0 * math.pi * k / (tap_count - 1))
    h /= h.sum()
    return h.astype(np.float32)


def random_taps(tap_count: int, seed: int, complex_taps: bool = False) -> np.ndarray:
    """Asymmetric taps (catches tap-order bugs)."""
    i = np.arange(tap_count, dtype=np.int64)
    re = hash_uniform(2 * i, seed ^ 0x7A95)
    if not complex_taps:
        return re.astype(np.float32)
    im = hash_uniform(2 * i + 1, seed ^ 0x7A95)
    return (re + 1j * im).astype(np.complex64)
