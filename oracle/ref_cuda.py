"""TEST INFRASTRUCTURE ONLY — ctypes binding of oracle/_ref/libgsdr_ref.so, the reference's own CUDA kernels
compiled for sm_100 by oracle/build_ref.sh (ref: src/fir.cu, src/quad_demod.cu, src/adjustFrequency.cu).
Needs a GPU to run; used by the `-m gpu` parity tests, tests/golden/make_golden.py and `bench.py --impl reference`.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "_ref" / "libgsdr_ref.so"

_lib = None


def available() -> bool:
    return LIB_PATH.exists()


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        l = C.CDLL(str(LIB_PATH))  # RTLD_LOCAL: its gsdrFir* do not clash with the product library's
        sz, vp = C.c_size_t, C.c_void_p
        for n in ("gsdrFirFC", "gsdrFirFF", "gsdrFirCC", "gsdrFirCF"):
            getattr(l, n).argtypes = [sz, vp, sz, vp, vp, sz, C.c_int32, vp]
            getattr(l, n).restype = C.c_int
        l.refAdjustFrequencyFirFC.argtypes = [C.c_float, C.c_float, sz, sz, vp, sz, vp, vp, sz, C.c_int32, vp]
        l.refAdjustFrequencyFirFC.restype = C.c_int
        l.gsdrQuadFmDemod.argtypes = [vp, vp, C.c_float, sz, C.c_int32, vp]
        l.gsdrQuadFmDemod.restype = C.c_int
        _lib = l
    return _lib


def fir(kind: str, decimation, taps, tapCount, input, output, numOutputs, device=0, stream=0) -> None:
    fn = getattr(lib(), "gsdrFir" + kind.upper())
    rc = fn(decimation, taps.data_ptr(), tapCount, input.data_ptr(), output.data_ptr(), numOutputs, device, stream)
    if rc:
        raise RuntimeError(f"reference gsdrFir{kind.upper()} returned cudaError_t {rc}")


def adjust_frequency_fir_fc(sampleRate, frequencyShift, firstSampleIndex, decimation, taps, tapCount, input, output,
                            numOutputs, device=0, stream=0) -> None:
    rc = lib().refAdjustFrequencyFirFC(sampleRate, frequencyShift, firstSampleIndex, decimation, taps.data_ptr(),
                                       tapCount, input.data_ptr(), output.data_ptr(), numOutputs, device, stream)
    if rc:
        raise RuntimeError(f"reference adjustFrequency harness returned cudaError_t {rc}")
