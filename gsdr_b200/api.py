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

from . import _lib
from ._lib import KernelInfo, Shard, StreamPlan


def L():
    """The loaded library calls go through (release build; the tuning build while a variant is forced)."""
    return _lib.active()


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
    def call(decimation, taps, tapCount, input, output, numOutputs, cudaDevice=0, cudaStream=None):
        _check(getattr(L(), name)(decimation, _ptr(taps), tapCount, _ptr(input), _ptr(output), numOutputs, cudaDevice,
                  _stream(cudaStream)), name)

    call.__name__ = name
    call.__doc__ = f"{name}(decimation, taps, tapCount, input, output, numOutputs, cudaDevice, cudaStream)"
    return call


gsdrFirFC = _fir("gsdrFirFC")
gsdrFirFF = _fir("gsdrFirFF")
gsdrFirCC = _fir("gsdrFirCC")
gsdrFirCF = _fir("gsdrFirCF")


def _nco(name):
    def call(sampleRate, frequencyShift, firstSampleIndex, decimation, taps, tapCount, input, output, numOutputs,
             cudaDevice=0, cudaStream=None):
        _check(getattr(L(), name)(sampleRate, frequencyShift, firstSampleIndex, decimation, _ptr(taps), tapCount, _ptr(input),
                  _ptr(output), numOutputs, cudaDevice, _stream(cudaStream)), name)

    call.__name__ = name
    return call


gsdrAdjustFrequencyFirFC = _nco("gsdrAdjustFrequencyFirFC")
gsdrAdjustFrequencyFirFCLiteral = _nco("gsdrAdjustFrequencyFirFCLiteral")


def gsdrChannelizeFC(sampleRate, frequencyShifts, firstSampleIndex, decimation, taps, tapCount, input, output,
                     outputStride, numOutputs, cudaDevice=0, cudaStream=None):
    """One input, len(frequencyShifts) channels: output[k * outputStride + n] (frequencyShifts: host floats)."""
    shifts = (C.c_float * len(frequencyShifts))(*[float(f) for f in frequencyShifts])
    _check(L().gsdrChannelizeFC(sampleRate, shifts, len(frequencyShifts), firstSampleIndex, decimation, _ptr(taps),
                                tapCount, _ptr(input), _ptr(output), outputStride, numOutputs, cudaDevice,
                                _stream(cudaStream)), "gsdrChannelizeFC")


gsdrFirFCInt8 = _fir("gsdrFirFCInt8")
gsdrAdjustFrequencyFirFCInt8 = _nco("gsdrAdjustFrequencyFirFCInt8")


def gsdrInt8ToNormFloat(input, output, numElements, cudaDevice=0, cudaStream=None):
    _check(L().gsdrInt8ToNormFloat(_ptr(input), _ptr(output), numElements, cudaDevice, _stream(cudaStream)),
           "gsdrInt8ToNormFloat")


def gsdrQuadFmDemod(input, output, gain, numOutputElements, cudaDevice=0, cudaStream=None):
    _check(L().gsdrQuadFmDemod(_ptr(input), _ptr(output), gain, numOutputElements, cudaDevice, _stream(cudaStream)),
           "gsdrQuadFmDemod")


def gsdrQuadAmDemod(input, output, numOutputElements, cudaDevice=0, cudaStream=None):
    _check(L().gsdrQuadAmDemod(_ptr(input), _ptr(output), numOutputElements, cudaDevice, _stream(cudaStream)),
           "gsdrQuadAmDemod")


def gsdrFmDemod(rfSampleRate, tuningFrequency, channelFrequency, frequencyDeviation, decimation, firstSampleIndex,
                lowPassTaps, numLowPassTaps, input, output, numOutputs, cudaDevice=0, cudaStream=None):
    _check(L().gsdrFmDemod(rfSampleRate, tuningFrequency, channelFrequency, frequencyDeviation, decimation,
                           firstSampleIndex, _ptr(lowPassTaps), numLowPassTaps, _ptr(input), _ptr(output), numOutputs,
                           cudaDevice, _stream(cudaStream)), "gsdrFmDemod")


def gsdrAmDemod(rfSampleRate, tuningFrequency, channelFrequency, decimation, firstSampleIndex, lowPassTaps,
                numLowPassTaps, input, output, numElements, cudaDevice=0, cudaStream=None):
    _check(L().gsdrAmDemod(rfSampleRate, tuningFrequency, channelFrequency, decimation, firstSampleIndex,
                           _ptr(lowPassTaps), numLowPassTaps, _ptr(input), _ptr(output), numElements, cudaDevice,
                           _stream(cudaStream)), "gsdrAmDemod")


def gsdrFmDemodFused(rfSampleRate, tuningFrequency, channelFrequency, frequencyDeviation, decimation, firstSampleIndex,
                     lowPassTaps, numLowPassTaps, input, output, numOutputs, cudaDevice=0, cudaStream=None):
    _check(L().gsdrFmDemodFused(rfSampleRate, tuningFrequency, channelFrequency, frequencyDeviation, decimation,
                                firstSampleIndex, _ptr(lowPassTaps), numLowPassTaps, _ptr(input), _ptr(output),
                                numOutputs, cudaDevice, _stream(cudaStream)), "gsdrFmDemodFused")


def gsdrFmDemodWorkspace(rfSampleRate, tuningFrequency, channelFrequency, frequencyDeviation, decimation,
                         firstSampleIndex, lowPassTaps, numLowPassTaps, input, output, numOutputs, workspace,
                         workspaceBytes, cudaDevice=0, cudaStream=None):
    _check(L().gsdrFmDemodWorkspace(rfSampleRate, tuningFrequency, channelFrequency, frequencyDeviation, decimation,
                                    firstSampleIndex, _ptr(lowPassTaps), numLowPassTaps, _ptr(input), _ptr(output),
                                    numOutputs, _ptr(workspace), workspaceBytes, cudaDevice, _stream(cudaStream)),
           "gsdrFmDemodWorkspace")


def fm_demod_workspace_bytes(numOutputs: int) -> int:
    return int(L().gsdrFmDemodWorkspaceBytes(numOutputs))


def release_scratch(cudaDevice: int = 0) -> None:
    _check(L().gsdrB200ReleaseScratch(cudaDevice), "gsdrB200ReleaseScratch")


def _batched(name):
    def call(decimation, taps, tapCount, tapStride, input, inputStride, output, outputStride, numOutputs,
             numChannels, cudaDevice=0, cudaStream=None):
        _check(getattr(L(), name)(decimation, _ptr(taps), tapCount, tapStride, _ptr(input), inputStride, _ptr(output), outputStride,
                  numOutputs, numChannels, cudaDevice, _stream(cudaStream)), name)

    call.__name__ = name
    return call


gsdrFirFCBatched = _batched("gsdrFirFCBatched")
gsdrFirFFBatched = _batched("gsdrFirFFBatched")


def nco_phase_step(frequencyShift: float, sampleRate: float) -> int:
    return int(L().gsdrNcoPhaseStep(frequencyShift, sampleRate))


def fir_num_outputs(numInputs: int, tapCount: int, decimation: int) -> int:
    return int(L().gsdrFirNumOutputs(numInputs, tapCount, decimation))


def fir_num_inputs(numOutputs: int, tapCount: int, decimation: int) -> int:
    return int(L().gsdrFirNumInputs(numOutputs, tapCount, decimation))


def shard_plan_time(numOutputs, decimation, tapCount, firstSampleIndex, numShards, shardIndex) -> Shard:
    s = Shard()
    if L().gsdrShardPlanTime(numOutputs, decimation, tapCount, firstSampleIndex, numShards, shardIndex, C.byref(s)):
        raise ValueError("gsdrShardPlanTime: invalid arguments")
    return s


def shard_plan_channels(numChannels, numShards, shardIndex):
    a, n = C.c_uint64(), C.c_uint64()
    if L().gsdrShardPlanChannels(numChannels, numShards, shardIndex, C.byref(a), C.byref(n)):
        raise ValueError("gsdrShardPlanChannels: invalid arguments")
    return int(a.value), int(n.value)


def describe_kernel(firType: int, decimation: int, tapCount: int, numOutputs: int, cudaDevice: int = 0) -> KernelInfo:
    info = KernelInfo()
    if L().gsdrB200DescribeKernel(firType, decimation, tapCount, numOutputs, cudaDevice, C.byref(info)):
        raise RuntimeError("gsdrB200DescribeKernel failed")
    return info


def set_kernel_variant(variant: int) -> None:
    """Forces a kernel variant (tests, tools/sweep.py).  The release library has no such hook: any value other than
    -1 switches this process's calls to the tuning build (libgsdr_b200_tuning.so); -1 switches back."""
    if variant == -1:
        if _lib._tuning is not None:
            _lib.tuning_lib().gsdrB200SetKernelVariant(-1)
        _lib.use_tuning(False)
        return
    _lib.use_tuning(True)
    if _lib.tuning_lib().gsdrB200SetKernelVariant(variant):
        _lib.use_tuning(False)
        raise ValueError(f"no kernel variant {variant}")


def set_debug_flags(flags: int) -> None:
    """Work-skipping measurement flags of the tuning build (results are wrong while set); 0 restores normal operation
    (the active library stays the tuning build until set_kernel_variant(-1))."""
    if flags:
        _lib.use_tuning(True)
    if _lib._tuning is not None:
        _lib.tuning_lib().gsdrB200SetDebugFlags(flags)


def set_fir_tensor_cores(enable: bool) -> bool:
    """gsdrB200SetFirTensorCores: False keeps every FIR call on the FFMA2 kernels.  Applies to both loaded builds;
    returns the previous setting."""
    prev = bool(L().gsdrB200SetFirTensorCores(1 if enable else 0))
    for other in (_lib.lib, _lib._tuning):
        if other is not None and other is not L():
            other.gsdrB200SetFirTensorCores(1 if enable else 0)
    return prev


def has_tuning_hooks() -> bool:
    return bool(L().gsdrB200HasTuningHooks())


def library_path() -> str:
    return str(_lib.TUNING_LIB_PATH if _lib.active() is not _lib.lib else _lib.LIB_PATH)


def fm_chain_launches() -> int:
    """Kernel launches of one FM receive chain step (gsdrFmDemod: fused mix + FIR, quad demod; then gsdrFirFF)."""
    return 3


def num_kernel_variants() -> int:
    return int(L().gsdrB200NumKernelVariants())


def num_polyphase_variants() -> int:
    return int(L().gsdrB200NumPolyphaseVariants())


class HostPipeline:
    """gsdrHostPipeline: host buffers in, host buffers out, H2D / kernel / D2H overlapped chunk by chunk."""

    def __init__(self, cudaDevice: int = 0, chunkInputBytes: int = 32 << 20, numBuffers: int = 3):
        h = C.c_void_p()
        _check(L().gsdrHostPipelineCreate(cudaDevice, chunkInputBytes, numBuffers, C.byref(h)),
               "gsdrHostPipelineCreate")
        self._h: Optional[C.c_void_p] = h

    def close(self) -> None:
        if self._h is not None:
            L().gsdrHostPipelineDestroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    @property
    def handle(self):
        return self._h

    def gsdrFirFCHost(self, decimation, taps, tapCount, input, output, numOutputs):
        _check(L().gsdrFirFCHost(self._h, decimation, _ptr(taps), tapCount, _ptr(input), _ptr(output), numOutputs),
               "gsdrFirFCHost")

    def gsdrFirFFHost(self, decimation, taps, tapCount, input, output, numOutputs):
        _check(L().gsdrFirFFHost(self._h, decimation, _ptr(taps), tapCount, _ptr(input), _ptr(output), numOutputs),
               "gsdrFirFFHost")

    def gsdrFirFCInt8Host(self, decimation, taps, tapCount, input, output, numOutputs):
        _check(L().gsdrFirFCInt8Host(self._h, decimation, _ptr(taps), tapCount, _ptr(input), _ptr(output), numOutputs),
               "gsdrFirFCInt8Host")

    def gsdrAdjustFrequencyFirFCInt8Host(self, sampleRate, frequencyShift, firstSampleIndex, decimation, taps, tapCount,
                                         input, output, numOutputs):
        _check(L().gsdrAdjustFrequencyFirFCInt8Host(self._h, sampleRate, frequencyShift, firstSampleIndex, decimation,
                                                    _ptr(taps), tapCount, _ptr(input), _ptr(output), numOutputs),
               "gsdrAdjustFrequencyFirFCInt8Host")

    def gsdrAdjustFrequencyFirFCHost(self, sampleRate, frequencyShift, firstSampleIndex, decimation, taps, tapCount,
                                     input, output, numOutputs):
        _check(L().gsdrAdjustFrequencyFirFCHost(self._h, sampleRate, frequencyShift, firstSampleIndex, decimation,
                                                _ptr(taps), tapCount, _ptr(input), _ptr(output), numOutputs),
               "gsdrAdjustFrequencyFirFCHost")


def stream_plan(decimation, tapCount, totalInputs, nextStart, numInputs, align=2) -> StreamPlan:
    p = StreamPlan()
    if L().gsdrFirStreamPlan(decimation, tapCount, totalInputs, nextStart, numInputs, align, C.byref(p)) != 0:
        raise ValueError("gsdrFirStreamPlan: invalid arguments")
    return p


class FirStream:
    """gsdrFirStream: feed blocks of any length; the concatenated outputs equal one call over the whole input."""

    FC, FF, FC_NCO, FC_INT8, FC_NCO_INT8 = 0, 1, 4, 5, 6

    def __init__(self, firType, decimation, taps, tapCount, sampleRate=0.0, frequencyShift=0.0, firstSampleIndex=0,
                 cudaDevice=0):
        h = C.c_void_p()
        _check(L().gsdrFirStreamCreate(C.byref(h), firType, decimation, _ptr(taps), tapCount, sampleRate,
                                       frequencyShift, firstSampleIndex, cudaDevice), "gsdrFirStreamCreate")
        self._h: Optional[C.c_void_p] = h

    def close(self) -> None:
        if self._h is not None:
            L().gsdrFirStreamDestroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def reset(self) -> None:
        L().gsdrFirStreamReset(self._h)

    def num_outputs(self, numInputs: int) -> int:
        return int(L().gsdrFirStreamNumOutputs(self._h, numInputs))

    def push(self, input, numInputs, output, cudaStream=None) -> int:
        n = C.c_size_t(0)
        _check(L().gsdrFirStreamPush(self._h, _ptr(input) if numInputs else None, numInputs,
                                     _ptr(output) if output is not None else None, C.byref(n), _stream(cudaStream)),
               "gsdrFirStreamPush")
        return int(n.value)


def _handles(pipelines):
    arr = (C.c_void_p * len(pipelines))(*[p.handle for p in pipelines])
    return arr


def gsdrFirFCMultiGpuHost(pipelines, decimation, taps, tapCount, input, output, numOutputs):
    """Host buffers in and out, the capture time-sharded over one HostPipeline per device."""
    _check(L().gsdrFirFCMultiGpuHost(_handles(pipelines), len(pipelines), decimation, _ptr(taps), tapCount, _ptr(input),
                                     _ptr(output), numOutputs), "gsdrFirFCMultiGpuHost")


def gsdrAdjustFrequencyFirFCMultiGpuHost(pipelines, sampleRate, frequencyShift, firstSampleIndex, decimation, taps,
                                         tapCount, input, output, numOutputs):
    _check(L().gsdrAdjustFrequencyFirFCMultiGpuHost(_handles(pipelines), len(pipelines), sampleRate, frequencyShift,
                                                    firstSampleIndex, decimation, _ptr(taps), tapCount, _ptr(input),
                                                    _ptr(output), numOutputs), "gsdrAdjustFrequencyFirFCMultiGpuHost")


def gsdrFirFCChannelsMultiGpuHost(pipelines, decimation, taps, tapCount, input, inputStride, output, outputStride,
                                  numOutputs, numChannels):
    _check(L().gsdrFirFCChannelsMultiGpuHost(_handles(pipelines), len(pipelines), decimation, _ptr(taps), tapCount,
                                             _ptr(input), inputStride, _ptr(output), outputStride, numOutputs,
                                             numChannels), "gsdrFirFCChannelsMultiGpuHost")


class MultiGpu:
    """gsdrMultiGpu: device-resident shards, one persistent host thread + stream per device, fused gather."""

    def __init__(self, devices):
        self.devices = list(devices)
        arr = (C.c_int32 * len(self.devices))(*self.devices)
        h = C.c_void_p()
        _check(L().gsdrMultiGpuCreate(arr, len(self.devices), C.byref(h)), "gsdrMultiGpuCreate")
        self._h = h

    def close(self):
        if self._h is not None:
            L().gsdrMultiGpuDestroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def peer_ok(self, g: int) -> bool:
        return bool(L().gsdrMultiGpuPeerOk(self._h, g))

    @staticmethod
    def _ptrs(seq):
        return (C.c_void_p * len(seq))(*[_ptr(x) for x in seq])

    def gsdrFirFCMultiGpu(self, sampleRate, frequencyShift, firstSampleIndex, decimation, taps, tapCount, inputs, outputs,
                          gatherOutput, numOutputs, repeats=1) -> float:
        """Returns the longest per-device time (ms per launch, CUDA events)."""
        ms = C.c_float(0.0)
        _check(L().gsdrFirFCMultiGpu(self._h, sampleRate, frequencyShift, firstSampleIndex, decimation, self._ptrs(taps),
                                     tapCount, self._ptrs(inputs), self._ptrs(outputs) if outputs is not None else None,
                                     _ptr(gatherOutput), numOutputs, repeats, C.byref(ms)), "gsdrFirFCMultiGpu")
        return float(ms.value)

    def gsdrMultiGpuGather(self, decimation, tapCount, outputs, dst, numOutputs) -> float:
        ms = C.c_float(0.0)
        _check(L().gsdrMultiGpuGather(self._h, decimation, tapCount, self._ptrs(outputs), _ptr(dst), numOutputs,
                                      C.byref(ms)), "gsdrMultiGpuGather")
        return float(ms.value)


def shared_buffer_create(nbytes: int, cudaDevice: int):
    """-> (device pointer, 64-byte handle) of a buffer other processes can map with shared_buffer_open."""
    p = C.c_void_p()
    handle = C.create_string_buffer(64)
    _check(L().gsdrSharedBufferCreate(nbytes, cudaDevice, C.byref(p), handle), "gsdrSharedBufferCreate")
    return int(p.value), handle.raw


def shared_buffer_open(handle: bytes, cudaDevice: int) -> int:
    p = C.c_void_p()
    _check(L().gsdrSharedBufferOpen(handle, cudaDevice, C.byref(p)), "gsdrSharedBufferOpen")
    return int(p.value)


def shared_buffer_close(ptr: int, cudaDevice: int) -> None:
    _check(L().gsdrSharedBufferClose(ptr, cudaDevice), "gsdrSharedBufferClose")


def shared_buffer_destroy(ptr: int, cudaDevice: int) -> None:
    _check(L().gsdrSharedBufferDestroy(ptr, cudaDevice), "gsdrSharedBufferDestroy")
