"""Python mirror of the gsdr C ABI (include/gsdr/fir.h, adjust_frequency.h, b200.h).

Function names, argument order and meaning are the reference's (ref: include/gsdr/fir.h:30-68):
    gsdrFirFC(decimation, taps, tapCount, input, output, numOutputs, cudaDevice, cudaStream)
`taps`, `input`, `output` may be torch CUDA tensors (complex64 or float32) or raw integer device addresses;
`cudaStream` may be None (the NULL stream), a torch.cuda.Stream or a raw handle.  A non-zero cudaError_t raises
CudaError — there is no CPU path behind any of these.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

from ._lib import KernelInfo, Shard, StreamPlan, lib


class CudaError(RuntimeError):
    def __init__(self, code: int, where: str):
        super().__init__(f"{where} failed with cudaError_t {code}")
        self.code = code


def _ptr(x) -> int:
    if x is None:
        return 0
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return int(x.data_ptr())
    if hasattr(x, "ctypes"):  # numpy (host pointers, for the *Host entry points)
        return int(x.ctypes.data)
    raise TypeError(f"cannot take an address of {type(x)!r}")


def _stream(s) -> int:
    if s is None:
        return 0
    if isinstance(s, int):
        return s
    if hasattr(s, "cuda_stream"):
        return int(s.cuda_stream)
    raise TypeError(f"not a CUDA stream: {type(s)!r}")


def _check(code: int, where: str) -> None:
    if code != 0:
        raise CudaError(code, where)


def _fir(name):
    fn = getattr(lib, name)

    def call(decimation, taps, tapCount, input, output, numOutputs, cudaDevice=0, cudaStream=None):
        _check(fn(decimation, _ptr(taps), tapCount, _ptr(input), _ptr(output), numOutputs, cudaDevice,
                  _stream(cudaStream)), name)

    call.__name__ = name
    call.__doc__ = f"{name}(decimation, taps, tapCount, input, output, numOutputs, cudaDevice, cudaStream)"
    return call


gsdrFirFC = _fir("gsdrFirFC")
gsdrFirFF = _fir("gsdrFirFF")
gsdrFirCC = _fir("gsdrFirCC")
gsdrFirCF = _fir("gsdrFirCF")


def _nco(name):
    fn = getattr(lib, name)

    def call(sampleRate, frequencyShift, firstSampleIndex, decimation, taps, tapCount, input, output, numOutputs,
             cudaDevice=0, cudaStream=None):
        _check(fn(sampleRate, frequencyShift, firstSampleIndex, decimation, _ptr(taps), tapCount, _ptr(input),
                  _ptr(output), numOutputs, cudaDevice, _stream(cudaStream)), name)

    call.__name__ = name
    return call


gsdrAdjustFrequencyFirFC = _nco("gsdrAdjustFrequencyFirFC")
gsdrAdjustFrequencyFirFCLiteral = _nco("gsdrAdjustFrequencyFirFCLiteral")


gsdrFirFCInt8 = _fir("gsdrFirFCInt8")
gsdrAdjustFrequencyFirFCInt8 = _nco("gsdrAdjustFrequencyFirFCInt8")


def gsdrInt8ToNormFloat(input, output, numElements, cudaDevice=0, cudaStream=None):
    _check(lib.gsdrInt8ToNormFloat(_ptr(input), _ptr(output), numElements, cudaDevice, _stream(cudaStream)),
           "gsdrInt8ToNormFloat")


def gsdrQuadFmDemod(input, output, gain, numOutputElements, cudaDevice=0, cudaStream=None):
    _check(lib.gsdrQuadFmDemod(_ptr(input), _ptr(output), gain, numOutputElements, cudaDevice, _stream(cudaStream)),
           "gsdrQuadFmDemod")


def gsdrQuadAmDemod(input, output, numOutputElements, cudaDevice=0, cudaStream=None):
    _check(lib.gsdrQuadAmDemod(_ptr(input), _ptr(output), numOutputElements, cudaDevice, _stream(cudaStream)),
           "gsdrQuadAmDemod")


def gsdrFmDemod(rfSampleRate, tuningFrequency, channelFrequency, frequencyDeviation, decimation, firstSampleIndex,
                lowPassTaps, numLowPassTaps, input, output, numOutputs, cudaDevice=0, cudaStream=None):
    _check(lib.gsdrFmDemod(rfSampleRate, tuningFrequency, channelFrequency, frequencyDeviation, decimation,
                           firstSampleIndex, _ptr(lowPassTaps), numLowPassTaps, _ptr(input), _ptr(output), numOutputs,
                           cudaDevice, _stream(cudaStream)), "gsdrFmDemod")


def _batched(name):
    fn = getattr(lib, name)

    def call(decimation, taps, tapCount, tapStride, input, inputStride, output, outputStride, numOutputs,
             numChannels, cudaDevice=0, cudaStream=None):
        _check(fn(decimation, _ptr(taps), tapCount, tapStride, _ptr(input), inputStride, _ptr(output), outputStride,
                  numOutputs, numChannels, cudaDevice, _stream(cudaStream)), name)

    call.__name__ = name
    return call


gsdrFirFCBatched = _batched("gsdrFirFCBatched")
gsdrFirFFBatched = _batched("gsdrFirFFBatched")


def nco_phase_step(frequencyShift: float, sampleRate: float) -> int:
    return int(lib.gsdrNcoPhaseStep(frequencyShift, sampleRate))


def fir_num_outputs(numInputs: int, tapCount: int, decimation: int) -> int:
    return int(lib.gsdrFirNumOutputs(numInputs, tapCount, decimation))


def fir_num_inputs(numOutputs: int, tapCount: int, decimation: int) -> int:
    return int(lib.gsdrFirNumInputs(numOutputs, tapCount, decimation))


def shard_plan_time(numOutputs, decimation, tapCount, firstSampleIndex, numShards, shardIndex) -> Shard:
    s = Shard()
    if lib.gsdrShardPlanTime(numOutputs, decimation, tapCount, firstSampleIndex, numShards, shardIndex, C.byref(s)):
        raise ValueError("gsdrShardPlanTime: invalid arguments")
    return s


def shard_plan_channels(numChannels, numShards, shardIndex):
    a, n = C.c_uint64(), C.c_uint64()
    if lib.gsdrShardPlanChannels(numChannels, numShards, shardIndex, C.byref(a), C.byref(n)):
        raise ValueError("gsdrShardPlanChannels: invalid arguments")
    return int(a.value), int(n.value)


def describe_kernel(firType: int, decimation: int, tapCount: int, numOutputs: int, cudaDevice: int = 0) -> KernelInfo:
    info = KernelInfo()
    if lib.gsdrB200DescribeKernel(firType, decimation, tapCount, numOutputs, cudaDevice, C.byref(info)):
        raise RuntimeError("gsdrB200DescribeKernel failed")
    return info


def set_kernel_variant(variant: int) -> None:
    if lib.gsdrB200SetKernelVariant(variant):
        raise ValueError(f"no kernel variant {variant}")


def num_kernel_variants() -> int:
    return int(lib.gsdrB200NumKernelVariants())


def num_polyphase_variants() -> int:
    return int(lib.gsdrB200NumPolyphaseVariants())


class HostPipeline:
    """gsdrHostPipeline: host buffers in, host buffers out, H2D / kernel / D2H overlapped chunk by chunk."""

    def __init__(self, cudaDevice: int = 0, chunkInputBytes: int = 32 << 20, numBuffers: int = 3):
        h = C.c_void_p()
        _check(lib.gsdrHostPipelineCreate(cudaDevice, chunkInputBytes, numBuffers, C.byref(h)),
               "gsdrHostPipelineCreate")
        self._h: Optional[C.c_void_p] = h

    def close(self) -> None:
        if self._h is not None:
            lib.gsdrHostPipelineDestroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    @property
    def handle(self):
        return self._h

    def gsdrFirFCHost(self, decimation, taps, tapCount, input, output, numOutputs):
        _check(lib.gsdrFirFCHost(self._h, decimation, _ptr(taps), tapCount, _ptr(input), _ptr(output), numOutputs),
               "gsdrFirFCHost")

    def gsdrFirFFHost(self, decimation, taps, tapCount, input, output, numOutputs):
        _check(lib.gsdrFirFFHost(self._h, decimation, _ptr(taps), tapCount, _ptr(input), _ptr(output), numOutputs),
               "gsdrFirFFHost")

    def gsdrAdjustFrequencyFirFCHost(self, sampleRate, frequencyShift, firstSampleIndex, decimation, taps, tapCount,
                                     input, output, numOutputs):
        _check(lib.gsdrAdjustFrequencyFirFCHost(self._h, sampleRate, frequencyShift, firstSampleIndex, decimation,
                                                _ptr(taps), tapCount, _ptr(input), _ptr(output), numOutputs),
               "gsdrAdjustFrequencyFirFCHost")


def stream_plan(decimation, tapCount, totalInputs, nextStart, numInputs, align=2) -> StreamPlan:
    p = StreamPlan()
    if lib.gsdrFirStreamPlan(decimation, tapCount, totalInputs, nextStart, numInputs, align, C.byref(p)) != 0:
        raise ValueError("gsdrFirStreamPlan: invalid arguments")
    return p


class FirStream:
    """gsdrFirStream: feed blocks of any length; the concatenated outputs equal one call over the whole input."""

    FC, FF, FC_NCO, FC_INT8, FC_NCO_INT8 = 0, 1, 4, 5, 6

    def __init__(self, firType, decimation, taps, tapCount, sampleRate=0.0, frequencyShift=0.0, firstSampleIndex=0,
                 cudaDevice=0):
        h = C.c_void_p()
        _check(lib.gsdrFirStreamCreate(C.byref(h), firType, decimation, _ptr(taps), tapCount, sampleRate,
                                       frequencyShift, firstSampleIndex, cudaDevice), "gsdrFirStreamCreate")
        self._h: Optional[C.c_void_p] = h

    def close(self) -> None:
        if self._h is not None:
            lib.gsdrFirStreamDestroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def reset(self) -> None:
        lib.gsdrFirStreamReset(self._h)

    def num_outputs(self, numInputs: int) -> int:
        return int(lib.gsdrFirStreamNumOutputs(self._h, numInputs))

    def push(self, input, numInputs, output, cudaStream=None) -> int:
        n = C.c_size_t(0)
        _check(lib.gsdrFirStreamPush(self._h, _ptr(input) if numInputs else None, numInputs,
                                     _ptr(output) if output is not None else None, C.byref(n), _stream(cudaStream)),
               "gsdrFirStreamPush")
        return int(n.value)
