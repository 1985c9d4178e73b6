"""ctypes binding of libgsdr_b200.so — the C ABI declared in include/gsdr/*.h.

There is deliberately no fallback: if the shared library is missing or a symbol cannot be resolved, importing
this module raises.  (Use `python gsdr_b200/build.py` or `__graft_entry__.build()` to compile it.)
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

# GSDR_B200_LIB: development hook to load an experimental build of the same library
LIB_PATH = Path(os.environ.get("GSDR_B200_LIB") or Path(__file__).resolve().parent / "csrc" / "libgsdr_b200.so")
# The tuning build (-DGSDR_B200_TUNING): same code plus gsdrB200SetKernelVariant / gsdrB200SetDebugFlags.  Loaded on
# demand by api.set_kernel_variant / api.set_debug_flags (variant-coverage tests, tools/sweep.py) — never by default.
TUNING_LIB_PATH = Path(os.environ.get("GSDR_B200_TUNING_LIB") or Path(__file__).resolve().parent / "csrc" / "libgsdr_b200_tuning.so")

c_size_t = C.c_size_t
c_void_p = C.c_void_p
c_int32 = C.c_int32
c_float = C.c_float
cudaError_t = C.c_int


class Shard(C.Structure):
    """gsdrShard of include/gsdr/b200.h."""

    _fields_ = [
        ("firstOutput", C.c_uint64),
        ("numOutputs", C.c_uint64),
        ("firstInput", C.c_uint64),
        ("numInputs", C.c_uint64),
        ("firstSampleIndex", C.c_uint64),
    ]


class StreamPlan(C.Structure):
    """gsdrStreamPlan of include/gsdr/stream.h."""

    _fields_ = [(n, C.c_uint64) for n in ("numOutputs", "skippedInputs", "headOutputs", "headNewInputs", "bodyOutputs",
                                         "bodyOffset", "carryLength", "newCarryLength", "newNextStart")]


class KernelInfo(C.Structure):
    """gsdrB200KernelInfo of include/gsdr/b200.h."""

    _fields_ = [
        ("variant", C.c_int),
        ("outputsPerThread", C.c_int),
        ("threadsPerBlock", C.c_int),
        ("phaseGroups", C.c_int),
        ("windowBuffers", C.c_int),
        ("smCount", C.c_int),
        ("outputsPerBlock", c_size_t),
        ("sharedBytesPerBlock", c_size_t),
        ("numBlocks", c_size_t),
    ]


_FIR_ARGS = [c_size_t, c_void_p, c_size_t, c_void_p, c_void_p, c_size_t, c_int32, c_void_p]
_NCO_ARGS = [c_float, c_float, c_size_t] + _FIR_ARGS
_BATCH_ARGS = [c_size_t, c_void_p, c_size_t, c_size_t, c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t,
               c_int32, c_void_p]

# name -> (restype, argtypes); every symbol the headers declare is listed here and checked at import.
SIGNATURES = {
    # include/gsdr/fir.h
    "gsdrFirFC": (cudaError_t, _FIR_ARGS),
    "gsdrFirFF": (cudaError_t, _FIR_ARGS),
    "gsdrFirCC": (cudaError_t, _FIR_ARGS),
    "gsdrFirCF": (cudaError_t, _FIR_ARGS),
    # include/gsdr/adjust_frequency.h
    "gsdrAdjustFrequencyFirFC": (cudaError_t, _NCO_ARGS),
    "gsdrAdjustFrequencyFirFCLiteral": (cudaError_t, _NCO_ARGS),
    "gsdrNcoPhaseStep": (C.c_uint64, [c_float, c_float]),
    "gsdrChannelizeFC": (cudaError_t, [c_float, C.POINTER(c_float), c_size_t, c_size_t, c_size_t, c_void_p, c_size_t,
                                       c_void_p, c_void_p, c_size_t, c_size_t, c_int32, c_void_p]),
    # include/gsdr/quad_demod.h, include/gsdr/fm.h
    "gsdrQuadFmDemod": (cudaError_t, [c_void_p, c_void_p, c_float, c_size_t, c_int32, c_void_p]),
    "gsdrQuadAmDemod": (cudaError_t, [c_void_p, c_void_p, c_size_t, c_int32, c_void_p]),
    "gsdrFmDemod": (cudaError_t, [c_float, c_float, c_float, c_float, C.c_uint32, c_size_t, c_void_p, c_size_t, c_void_p,
                                  c_void_p, c_size_t, c_int32, c_void_p]),
    "gsdrAmDemod": (cudaError_t, [c_float, c_float, c_float, C.c_uint32, c_size_t, c_void_p, c_size_t, c_void_p, c_void_p,
                                  c_size_t, c_int32, c_void_p]),
    "gsdrFmDemodFused": (cudaError_t, [c_float, c_float, c_float, c_float, C.c_uint32, c_size_t, c_void_p, c_size_t,
                                       c_void_p, c_void_p, c_size_t, c_int32, c_void_p]),
    "gsdrFmDemodWorkspaceBytes": (c_size_t, [c_size_t]),
    "gsdrFmDemodWorkspace": (cudaError_t, [c_float, c_float, c_float, c_float, C.c_uint32, c_size_t, c_void_p, c_size_t,
                                           c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_int32, c_void_p]),
    "gsdrB200ReleaseScratch": (cudaError_t, [c_int32]),
    # include/gsdr/b200.h
    "gsdrFirNumOutputs": (c_size_t, [c_size_t, c_size_t, c_size_t]),
    "gsdrFirNumInputs": (c_size_t, [c_size_t, c_size_t, c_size_t]),
    "gsdrFirFCBatched": (cudaError_t, _BATCH_ARGS),
    "gsdrFirFFBatched": (cudaError_t, _BATCH_ARGS),
    "gsdrShardPlanTime": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32,
                                    C.POINTER(Shard)]),
    "gsdrShardPlanChannels": (C.c_int, [C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64),
                                        C.POINTER(C.c_uint64)]),
    "gsdrHostPipelineCreate": (cudaError_t, [c_int32, c_size_t, C.c_int, C.POINTER(c_void_p)]),
    "gsdrHostPipelineDestroy": (None, [c_void_p]),
    "gsdrFirFCHost": (cudaError_t, [c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_void_p, c_size_t]),
    "gsdrFirFFHost": (cudaError_t, [c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_void_p, c_size_t]),
    "gsdrAdjustFrequencyFirFCHost": (cudaError_t, [c_void_p, c_float, c_float, c_size_t, c_size_t, c_void_p, c_size_t,
                                                   c_void_p, c_void_p, c_size_t]),
    "gsdrFirFCMultiGpuHost": (cudaError_t, [C.POINTER(c_void_p), C.c_int, c_size_t, c_void_p, c_size_t, c_void_p,
                                            c_void_p, c_size_t]),
    "gsdrFirFCInt8Host": (cudaError_t, [c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_void_p, c_size_t]),
    "gsdrAdjustFrequencyFirFCInt8Host": (cudaError_t, [c_void_p, c_float, c_float, c_size_t, c_size_t, c_void_p,
                                                       c_size_t, c_void_p, c_void_p, c_size_t]),
    "gsdrAdjustFrequencyFirFCMultiGpuHost": (cudaError_t, [C.POINTER(c_void_p), C.c_int, c_float, c_float, c_size_t,
                                                           c_size_t, c_void_p, c_size_t, c_void_p, c_void_p, c_size_t]),
    "gsdrFirFCChannelsMultiGpuHost": (cudaError_t, [C.POINTER(c_void_p), C.c_int, c_size_t, c_void_p, c_size_t, c_void_p,
                                                    c_size_t, c_void_p, c_size_t, c_size_t, c_size_t]),
    "gsdrMultiGpuCreate": (cudaError_t, [C.POINTER(c_int32), C.c_int, C.POINTER(c_void_p)]),
    "gsdrMultiGpuDestroy": (None, [c_void_p]),
    "gsdrMultiGpuPeerOk": (C.c_int, [c_void_p, C.c_int]),
    "gsdrFirFCMultiGpu": (cudaError_t, [c_void_p, c_float, c_float, c_size_t, c_size_t, C.POINTER(c_void_p), c_size_t,
                                        C.POINTER(c_void_p), C.POINTER(c_void_p), c_void_p, c_size_t, C.c_int,
                                        C.POINTER(c_float)]),
    "gsdrMultiGpuGather": (cudaError_t, [c_void_p, c_size_t, c_size_t, C.POINTER(c_void_p), c_void_p, c_size_t,
                                         C.POINTER(c_float)]),
    "gsdrSharedBufferCreate": (cudaError_t, [c_size_t, c_int32, C.POINTER(c_void_p), C.c_char_p]),
    "gsdrSharedBufferOpen": (cudaError_t, [C.c_char_p, c_int32, C.POINTER(c_void_p)]),
    "gsdrSharedBufferClose": (cudaError_t, [c_void_p, c_int32]),
    "gsdrSharedBufferDestroy": (cudaError_t, [c_void_p, c_int32]),
    # include/gsdr/conversion.h
    "gsdrInt8ToNormFloat": (cudaError_t, [c_void_p, c_void_p, c_size_t, c_int32, c_void_p]),
    "gsdrFirFCInt8": (cudaError_t, _FIR_ARGS),
    "gsdrAdjustFrequencyFirFCInt8": (cudaError_t, _NCO_ARGS),
    # include/gsdr/stream.h
    "gsdrFirStreamPlan": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32,
                                    C.POINTER(StreamPlan)]),
    "gsdrFirStreamCreate": (cudaError_t, [C.POINTER(c_void_p), C.c_int, c_size_t, c_void_p, c_size_t, c_float, c_float,
                                          c_size_t, c_int32]),
    "gsdrFirStreamDestroy": (None, [c_void_p]),
    "gsdrFirStreamReset": (None, [c_void_p]),
    "gsdrFirStreamNumOutputs": (c_size_t, [c_void_p, c_size_t]),
    "gsdrFirStreamPush": (cudaError_t, [c_void_p, c_void_p, c_size_t, c_void_p, C.POINTER(c_size_t), c_void_p]),
    "gsdrB200DescribeKernel": (C.c_int, [C.c_int, c_size_t, c_size_t, c_size_t, c_int32, C.POINTER(KernelInfo)]),
    "gsdrB200NumKernelVariants": (C.c_int, []),
    "gsdrB200NumPolyphaseVariants": (C.c_int, []),
    "gsdrB200HasTuningHooks": (C.c_int, []),
    "gsdrB200SetFirTensorCores": (C.c_int, [C.c_int]),
}

# exported by the tuning build only (include/gsdr/b200.h, #ifdef GSDR_B200_TUNING)
TUNING_SIGNATURES = {
    "gsdrB200SetKernelVariant": (C.c_int, [C.c_int]),
    "gsdrB200SetDebugFlags": (C.c_int, [C.c_int]),
}


def _load(path: Path, tuning: bool) -> C.CDLL:
    if not path.exists():
        raise ImportError(
            f"{path} is missing: the CUDA library has not been built (run `python gsdr_b200/build.py`). "
            "gsdr_b200 has no CPU fallback.")
    lib = C.CDLL(str(path))
    sigs = dict(SIGNATURES)
    if tuning:
        sigs.update(TUNING_SIGNATURES)
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load(LIB_PATH, tuning=False)  # the release library: what every call goes through by default
_tuning = None
_active = lib


def tuning_lib() -> C.CDLL:
    global _tuning
    if _tuning is None:
        _tuning = _load(TUNING_LIB_PATH, tuning=True)
    return _tuning


def active() -> C.CDLL:
    """The library API calls go through: the release build unless a test / sweep has selected the tuning build."""
    return _active


def use_tuning(on: bool) -> None:
    global _active
    _active = tuning_lib() if on else lib
