"""TEST INFRASTRUCTURE ONLY — numpy/ctypes front end of oracle/libgsdr_oracle.so (see oracle/gsdr_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl cpu legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libgsdr_oracle.so"

NCO_LITERAL = 0
NCO_EXACT = 1


def build(force: bool = False) -> Path:
    src = [HERE / "gsdr_oracle.c", HERE / "gsdr_oracle.h"]
    if force or not LIB_PATH.exists() or any(s.stat().st_mtime > LIB_PATH.stat().st_mtime for s in src):
        subprocess.run(["make", "-C", str(HERE), "-B"], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(str(LIB_PATH))
        sz, vp, f32 = C.c_size_t, C.c_void_p, C.c_float
        for name in ("ff", "fc", "cc", "cf"):
            getattr(l, f"gsdr_oracle_fir_{name}").argtypes = [sz, vp, sz, vp, vp, sz]
            getattr(l, f"gsdr_oracle_fir_{name}").restype = None
            getattr(l, f"gsdr_oracle_fir_{name}_f64").argtypes = [sz, vp, sz, vp, vp, sz]
            getattr(l, f"gsdr_oracle_fir_{name}_f64").restype = None
        for name in ("fc", "ff"):
            getattr(l, f"gsdr_oracle_fir_{name}_mt").argtypes = [sz, vp, sz, vp, vp, sz, C.c_int]
            getattr(l, f"gsdr_oracle_fir_{name}_mt").restype = None
        l.gsdr_oracle_reduce_first_sample_index.argtypes = [sz, f32]
        l.gsdr_oracle_reduce_first_sample_index.restype = C.c_uint32
        l.gsdr_oracle_nco_literal_theta_div_pi.argtypes = [f32, C.c_uint32, f32]
        l.gsdr_oracle_nco_literal_theta_div_pi.restype = f32
        l.gsdr_oracle_nco_exact_phase_step.argtypes = [f32, f32]
        l.gsdr_oracle_nco_exact_phase_step.restype = C.c_uint64
        l.gsdr_oracle_nco_exact_phase_q31.argtypes = [C.c_uint64, C.c_uint64]
        l.gsdr_oracle_nco_exact_phase_q31.restype = C.c_int32
        for name in ("gsdr_oracle_adjust_frequency_fir_fc", "gsdr_oracle_adjust_frequency_fir_fc_f64"):
            getattr(l, name).argtypes = [C.c_int, f32, f32, sz, sz, vp, sz, vp, vp, sz]
            getattr(l, name).restype = None
        l.gsdr_oracle_quad_fm_demod.argtypes = [vp, vp, f32, sz]
        l.gsdr_oracle_quad_fm_demod.restype = None
        _lib = l
    return _lib


_IN = {"ff": np.float32, "fc": np.complex64, "cc": np.complex64, "cf": np.float32}
_TAP = {"ff": np.float32, "fc": np.float32, "cc": np.complex64, "cf": np.complex64}
_OUT = {"ff": np.float32, "fc": np.complex64, "cc": np.complex64, "cf": np.complex64}


def _p(a: np.ndarray) -> int:
    return a.ctypes.data


def num_outputs(num_inputs: int, tap_count: int, decimation: int) -> int:
    """N_out = floor((N_in - T)/D) + 1 (ref: the caller contract implied by src/fir.cu:57-70)."""
    if decimation <= 0 or tap_count <= 0 or num_inputs < tap_count:
        return 0
    return (num_inputs - tap_count) // decimation + 1


def fir(kind: str, decimation: int, taps, x, num_out: int | None = None, f64: bool = False, threads: int = 1):
    """kind in {'ff','fc','cc','cf'} (<TapType><InputType>, ref: include/gsdr/fir.h:30-68)."""
    taps = np.ascontiguousarray(taps, dtype=_TAP[kind])
    x = np.ascontiguousarray(x, dtype=_IN[kind])
    T = taps.shape[0]
    if num_out is None:
        num_out = num_outputs(x.shape[0], T, decimation)
    if num_out > 0 and T > 0:
        assert (num_out - 1) * decimation + T <= x.shape[0], "input too short for numOutputs"
    if f64:
        out = np.zeros(num_out, dtype=np.float64 if kind == "ff" else np.complex128)
        getattr(lib(), f"gsdr_oracle_fir_{kind}_f64")(decimation, _p(taps), T, _p(x), _p(out), num_out)
        return out
    out = np.zeros(num_out, dtype=_OUT[kind])
    if threads > 1 and kind in ("fc", "ff"):
        getattr(lib(), f"gsdr_oracle_fir_{kind}_mt")(decimation, _p(taps), T, _p(x), _p(out), num_out, threads)
    else:
        getattr(lib(), f"gsdr_oracle_fir_{kind}")(decimation, _p(taps), T, _p(x), _p(out), num_out)
    return out


def adjust_frequency_fir_fc(mode: int, sample_rate: float, freq_shift: float, first_sample_index: int,
                            decimation: int, taps, x, num_out: int | None = None, f64: bool = False):
    taps = np.ascontiguousarray(taps, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.complex64)
    T = taps.shape[0]
    if num_out is None:
        num_out = num_outputs(x.shape[0], T, decimation)
    out = np.zeros(num_out, dtype=np.complex128 if f64 else np.complex64)
    fn = lib().gsdr_oracle_adjust_frequency_fir_fc_f64 if f64 else lib().gsdr_oracle_adjust_frequency_fir_fc
    fn(mode, sample_rate, freq_shift, first_sample_index, decimation, _p(taps), T, _p(x), _p(out), num_out)
    return out


def nco_exact_phase_step(freq_shift: float, sample_rate: float) -> int:
    return int(lib().gsdr_oracle_nco_exact_phase_step(freq_shift, sample_rate))


def nco_exact_phase_q31(step: int, index: int) -> int:
    return int(lib().gsdr_oracle_nco_exact_phase_q31(step, index))


def nco_literal_theta_div_pi(freq_shift: float, index: int, sample_rate: float) -> float:
    return float(lib().gsdr_oracle_nco_literal_theta_div_pi(freq_shift, index & 0xFFFFFFFF, sample_rate))


def reduce_first_sample_index(first: int, sample_rate: float) -> int:
    return int(lib().gsdr_oracle_reduce_first_sample_index(first, sample_rate))


def quad_fm_demod(x, gain: float, num_out: int | None = None):
    x = np.ascontiguousarray(x, dtype=np.complex64)
    if num_out is None:
        num_out = max(x.shape[0] - 1, 0)
    out = np.zeros(num_out, dtype=np.float32)
    lib().gsdr_oracle_quad_fm_demod(_p(x), _p(out), gain, num_out)
    return out


def quad_am_demod(x):
    """2 * saturate(|x|) - 1 per sample in float32 (ref: src/quad_demod.cu:46-49, src/am.cu:49)."""
    x = np.ascontiguousarray(x, dtype=np.complex64)
    mag = np.hypot(x.real.astype(np.float32), x.imag.astype(np.float32)).astype(np.float32)
    return (np.float32(2.0) * np.clip(mag, np.float32(0.0), np.float32(1.0)) - np.float32(1.0)).astype(np.float32)


# ---- pure-numpy second opinion (tiny cases only) -----------------------------------------------------------

def fir_numpy_f64(decimation: int, taps, x, num_out: int | None = None):
    """Direct evaluation of out[n] = sum_i x[n*D+i]*taps[i] in float64/complex128 with numpy (small sizes)."""
    taps = np.asarray(taps)
    x = np.asarray(x)
    T = taps.shape[0]
    if num_out is None:
        num_out = num_outputs(x.shape[0], T, decimation)
    cplx = np.iscomplexobj(taps) or np.iscomplexobj(x)
    out = np.zeros(num_out, dtype=np.complex128 if cplx else np.float64)
    td = taps.astype(np.complex128 if np.iscomplexobj(taps) else np.float64)
    xd = x.astype(np.complex128 if np.iscomplexobj(x) else np.float64)
    for n in range(num_out):
        out[n] = np.dot(xd[n * decimation:n * decimation + T], td)
    return out
