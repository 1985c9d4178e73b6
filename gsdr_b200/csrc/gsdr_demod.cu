// gsdr_demod.cu — quadrature demodulators and the FM receive stage (<gsdr/quad_demod.h>, <gsdr/fm.h>).
// Replaces ref: src/quad_demod.cu:23-74 and ref: src/fm.cu:21-69,181-218.  HBM-bound elementwise kernels:
// 8 bytes read + 4 bytes written per output, four outputs per thread with 16-byte accesses when aligned.
#include <gsdr/am.h>
#include <gsdr/fm.h>
#include <gsdr/quad_demod.h>

#include <cmath>
#include <cstdint>
#include <mutex>

#include "launch.h"

namespace gsdr_b200 {

constexpr int kDemodThreads = 256;
constexpr int kDemodPerThread = 4;

// m = next * conj(cur) with the expression shape nvcc gives the reference's cuCmulf(next, cuConjf(cur)):
// FMUL, FMUL, FFMA, FFMA (ref: src/quad_demod.cu:30; checked against the sm_100 SASS of the compiled reference),
// then gain * atan2f(Im, Re) (ref: src/quad_demod.cu:31).  Same libdevice atan2f => same bits as the reference.
__device__ __forceinline__ float quadFm(float2 cur, float2 next, float gain) {
  const float a = __fmul_rn(next.x, cur.y);
  const float b = __fmul_rn(next.y, cur.y);
  const float im = __fmaf_rn(next.y, cur.x, -a);
  const float re = __fmaf_rn(next.x, cur.x, b);
  return __fmul_rn(gain, atan2f(im, re));
}

// ref: src/quad_demod.cu:46-49 — scalbnf(__saturatef(hypotf(x, y)), 1) - 1
__device__ __forceinline__ float quadAm(float2 v) { return __fadd_rn(scalbnf(__saturatef(hypotf(v.x, v.y)), 1), -1.0f); }

__global__ void __launch_bounds__(kDemodThreads) quadFmDemodKernel(const float2* __restrict__ in, float* __restrict__ out,
                                                                  float gain, unsigned long long n, int aligned) {
  const unsigned long long i0 = ((unsigned long long)blockIdx.x * kDemodThreads + threadIdx.x) * kDemodPerThread;
  if (i0 >= n) return;
  if (aligned && i0 + kDemodPerThread <= n) {
    const float4 p0 = __ldg(reinterpret_cast<const float4*>(in + i0));
    const float4 p1 = __ldg(reinterpret_cast<const float4*>(in + i0 + 2));
    const float2 p2 = __ldg(in + i0 + 4);
    float4 r;
    r.x = quadFm(make_float2(p0.x, p0.y), make_float2(p0.z, p0.w), gain);
    r.y = quadFm(make_float2(p0.z, p0.w), make_float2(p1.x, p1.y), gain);
    r.z = quadFm(make_float2(p1.x, p1.y), make_float2(p1.z, p1.w), gain);
    r.w = quadFm(make_float2(p1.z, p1.w), p2, gain);
    *reinterpret_cast<float4*>(out + i0) = r;
  } else {
    float2 cur = __ldg(in + i0);
    for (int k = 0; k < kDemodPerThread && i0 + k < n; k++) {
      const float2 next = __ldg(in + i0 + k + 1);
      out[i0 + k] = quadFm(cur, next, gain);
      cur = next;
    }
  }
}

__global__ void __launch_bounds__(kDemodThreads) quadAmDemodKernel(const float2* __restrict__ in, float* __restrict__ out,
                                                                  unsigned long long n, int aligned) {
  const unsigned long long i0 = ((unsigned long long)blockIdx.x * kDemodThreads + threadIdx.x) * kDemodPerThread;
  if (i0 >= n) return;
  if (aligned && i0 + kDemodPerThread <= n) {
    const float4 p0 = __ldg(reinterpret_cast<const float4*>(in + i0));
    const float4 p1 = __ldg(reinterpret_cast<const float4*>(in + i0 + 2));
    *reinterpret_cast<float4*>(out + i0) =
        make_float4(quadAm(make_float2(p0.x, p0.y)), quadAm(make_float2(p0.z, p0.w)), quadAm(make_float2(p1.x, p1.y)),
                    quadAm(make_float2(p1.z, p1.w)));
  } else {
    for (int k = 0; k < kDemodPerThread && i0 + k < n; k++) out[i0 + k] = quadAm(__ldg(in + i0 + k));
  }
}

static cudaError_t enqueueQuadFm(const cuComplex* in, float* out, float gain, size_t n, cudaStream_t stream) noexcept {
  if (n == 0) return cudaSuccess;
  const unsigned long long per = (unsigned long long)kDemodThreads * kDemodPerThread;
  const unsigned long long blocks = (n + per - 1) / per;
  if (blocks > 0x7fffffffull) return cudaErrorInvalidValue;
  const int aligned = ((uintptr_t)in % 16 == 0 && (uintptr_t)out % 16 == 0) ? 1 : 0;
  quadFmDemodKernel<<<(unsigned)blocks, kDemodThreads, 0, stream>>>((const float2*)in, out, gain, n, aligned);
  return cudaPeekAtLastError();
}

// Library-private stream-ordered memory pool per device for gsdrFmDemod's low-pass scratch (the workspace variant
// gsdrFmDemodWorkspace allocates nothing).  The pool keeps freed blocks so that back-to-back calls reuse them instead
// of going back to the driver at every synchronisation (the default pool's release threshold of 0 made every call
// re-allocate: 4.3 ms instead of 1.5 ms per call in bench.py --workload cfg5); gsdrB200ReleaseScratch() hands the
// memory back.  The handle is cached only once the pool is fully configured.  The device's default pool is untouched.
static std::mutex gPoolMutex;
static cudaMemPool_t gPools[64] = {};

static cudaError_t scratchPool(int dev, cudaMemPool_t* pool) noexcept {
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  std::lock_guard<std::mutex> lock(gPoolMutex);
  if (!gPools[dev]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t fresh = nullptr;
    cudaError_t st = cudaMemPoolCreate(&fresh, &props);
    if (st != cudaSuccess) return st;
    uint64_t threshold = UINT64_MAX;
    st = cudaMemPoolSetAttribute(fresh, cudaMemPoolAttrReleaseThreshold, &threshold);
    if (st != cudaSuccess) {
      cudaMemPoolDestroy(fresh);
      return st;
    }
    gPools[dev] = fresh;
  }
  *pool = gPools[dev];
  return cudaSuccess;
}

// The stage as ONE kernel: the quadrature demodulator runs in the FIR's store path (no low-pass values in HBM, no
// scratch).  cudaErrorNotSupported when the shape's kernel has no such stage (nothing has been enqueued then).
static cudaError_t fmDemodFused(float rfSampleRate, float tuningFrequency, float channelFrequency,
                                float frequencyDeviation, uint32_t decimation, size_t firstSampleIndex,
                                const float* lowPassTaps, size_t numLowPassTaps, const cuComplex* input, float* output,
                                size_t numOutputs, cudaStream_t stream) noexcept {
  FirCall c;
  c.type = kFirFC;
  c.nco = kNcoExact;
  c.decimation = decimation;
  c.taps = lowPassTaps;
  c.tapCount = numLowPassTaps;
  c.input = input;
  c.output = output;
  c.numOutputs = numOutputs;
  c.sampleRate = rfSampleRate;
  c.frequencyShift = tuningFrequency - channelFrequency;  // ref: src/fm.cu:204
  c.firstSampleIndex = firstSampleIndex;
  c.epilogue = kFirEpiFmDemod;
  c.epilogueGain = rfSampleRate / (2.0f * 3.14159265358979323846f * frequencyDeviation);  // ref: src/fm.cu:203
  return enqueueFir(c, stream);
}

static cudaError_t fmDemodStage(float rfSampleRate, float tuningFrequency, float channelFrequency,
                                float frequencyDeviation, uint32_t decimation, size_t firstSampleIndex,
                                const float* lowPassTaps, size_t numLowPassTaps, const cuComplex* input, float* output,
                                size_t numOutputs, void* lowPassed, cudaStream_t stream) noexcept {
  FirCall c;
  c.type = kFirFC;
  c.nco = kNcoExact;
  c.decimation = decimation;
  c.taps = lowPassTaps;
  c.tapCount = numLowPassTaps;
  c.input = input;
  c.output = lowPassed;
  c.numOutputs = numOutputs + 1;
  c.sampleRate = rfSampleRate;
  c.frequencyShift = tuningFrequency - channelFrequency;  // ref: src/fm.cu:204
  c.firstSampleIndex = firstSampleIndex;
  cudaError_t st = enqueueFir(c, stream);
  if (st == cudaSuccess) {
    const float gain = rfSampleRate / (2.0f * 3.14159265358979323846f * frequencyDeviation);  // ref: src/fm.cu:203
    st = enqueueQuadFm((const cuComplex*)lowPassed, output, gain, numOutputs, stream);
  }
  return st;
}

}  // namespace gsdr_b200

using namespace gsdr_b200;

GSDR_C_LINKAGE cudaError_t gsdrQuadFmDemod(const cuComplex* input, float* output, float gain, size_t numOutputElements,
                                           int32_t cudaDevice, cudaStream_t cudaStream) GSDR_NO_EXCEPT {
  DeviceScope scope(cudaDevice);
  if (scope.status() != cudaSuccess) return scope.status();
  return enqueueQuadFm(input, output, gain, numOutputElements, cudaStream);
}

GSDR_C_LINKAGE cudaError_t gsdrQuadAmDemod(const cuComplex* input, float* output, size_t numOutputElements,
                                           int32_t cudaDevice, cudaStream_t cudaStream) GSDR_NO_EXCEPT {
  DeviceScope scope(cudaDevice);
  if (scope.status() != cudaSuccess) return scope.status();
  if (numOutputElements == 0) return cudaSuccess;
  const unsigned long long per = (unsigned long long)kDemodThreads * kDemodPerThread;
  const unsigned long long blocks = (numOutputElements + per - 1) / per;
  if (blocks > 0x7fffffffull) return cudaErrorInvalidValue;
  const int aligned = ((uintptr_t)input % 16 == 0 && (uintptr_t)output % 16 == 0) ? 1 : 0;
  quadAmDemodKernel<<<(unsigned)blocks, kDemodThreads, 0, cudaStream>>>((const float2*)input, output,
                                                                        numOutputElements, aligned);
  return cudaPeekAtLastError();
}

GSDR_C_LINKAGE cudaError_t gsdrFmDemod(float rfSampleRate, float tuningFrequency, float channelFrequency,
                                       float frequencyDeviation, uint32_t decimation, size_t firstSampleIndex,
                                       const float* lowPassTaps, size_t numLowPassTaps, const cuComplex* input,
                                       float* output, size_t numOutputs, int32_t cudaDevice,
                                       cudaStream_t cudaStream) GSDR_NO_EXCEPT {
  DeviceScope scope(cudaDevice);
  if (scope.status() != cudaSuccess) return scope.status();
  if (numOutputs == 0) return cudaSuccess;
  if (decimation == 0) return cudaErrorInvalidValue;
  // Two kernels through pool scratch: measured FASTER than the fused stage (BASELINE config 5 chain 1.026 vs 1.072 ms:
  // atan2f costs the issue-bound FIR kernel more than the separate HBM-bound demodulator launch costs).  The fused
  // form is gsdrFmDemodFused.
  void* lowPassed = nullptr;
  cudaMemPool_t pool = nullptr;
  cudaError_t st = scratchPool(cudaDevice, &pool);
  if (st != cudaSuccess) return st;
  st = cudaMallocFromPoolAsync(&lowPassed, (numOutputs + 1) * sizeof(cuComplex), pool, cudaStream);
  if (st != cudaSuccess) return st;
  st = fmDemodStage(rfSampleRate, tuningFrequency, channelFrequency, frequencyDeviation, decimation, firstSampleIndex,
                    lowPassTaps, numLowPassTaps, input, output, numOutputs, lowPassed, cudaStream);
  const cudaError_t fr = cudaFreeAsync(lowPassed, cudaStream);
  return st != cudaSuccess ? st : fr;
}

GSDR_C_LINKAGE cudaError_t gsdrFmDemodFused(float rfSampleRate, float tuningFrequency, float channelFrequency,
                                            float frequencyDeviation, uint32_t decimation, size_t firstSampleIndex,
                                            const float* lowPassTaps, size_t numLowPassTaps, const cuComplex* input,
                                            float* output, size_t numOutputs, int32_t cudaDevice,
                                            cudaStream_t cudaStream) GSDR_NO_EXCEPT {
  DeviceScope scope(cudaDevice);
  if (scope.status() != cudaSuccess) return scope.status();
  if (numOutputs == 0) return cudaSuccess;
  if (decimation == 0) return cudaErrorInvalidValue;
  return fmDemodFused(rfSampleRate, tuningFrequency, channelFrequency, frequencyDeviation, decimation, firstSampleIndex,
                      lowPassTaps, numLowPassTaps, input, output, numOutputs, cudaStream);
}

GSDR_C_LINKAGE size_t gsdrFmDemodWorkspaceBytes(size_t numOutputs) GSDR_NO_EXCEPT {
  return (numOutputs + 1) * sizeof(cuComplex);
}

GSDR_C_LINKAGE cudaError_t gsdrFmDemodWorkspace(float rfSampleRate, float tuningFrequency, float channelFrequency,
                                                float frequencyDeviation, uint32_t decimation, size_t firstSampleIndex,
                                                const float* lowPassTaps, size_t numLowPassTaps, const cuComplex* input,
                                                float* output, size_t numOutputs, void* workspace,
                                                size_t workspaceBytes, int32_t cudaDevice,
                                                cudaStream_t cudaStream) GSDR_NO_EXCEPT {
  DeviceScope scope(cudaDevice);
  if (scope.status() != cudaSuccess) return scope.status();
  if (numOutputs == 0) return cudaSuccess;
  if (decimation == 0 || !workspace || workspaceBytes < gsdrFmDemodWorkspaceBytes(numOutputs) ||
      (uintptr_t)workspace % 16 != 0)
    return cudaErrorInvalidValue;
  return fmDemodStage(rfSampleRate, tuningFrequency, channelFrequency, frequencyDeviation, decimation, firstSampleIndex,
                      lowPassTaps, numLowPassTaps, input, output, numOutputs, workspace, cudaStream);
}

GSDR_C_LINKAGE cudaError_t gsdrB200ReleaseScratch(int32_t cudaDevice) GSDR_NO_EXCEPT {
  if (cudaDevice < 0 || cudaDevice >= 64) return cudaErrorInvalidDevice;
  std::lock_guard<std::mutex> lock(gPoolMutex);
  if (!gPools[cudaDevice]) return cudaSuccess;
  return cudaMemPoolTrimTo(gPools[cudaDevice], 0);  // blocks in use by enqueued work stay; everything else goes back
}

// ---- <gsdr/am.h> ------------------------------------------------------------------------------------------------

GSDR_C_LINKAGE cudaError_t gsdrAmDemod(float rfSampleRate, float tuningFrequency, float channelFrequency,
                                       uint32_t decimation, size_t firstSampleIndex, const float* lowPassTaps,
                                       size_t numLowPassTaps, const cuComplex* input, float* output, size_t numElements,
                                       int32_t cudaDevice, cudaStream_t cudaStream) GSDR_NO_EXCEPT {
  DeviceScope scope(cudaDevice);
  if (scope.status() != cudaSuccess) return scope.status();
  if (numElements == 0) return cudaSuccess;
  if (decimation == 0) return cudaErrorInvalidValue;
  FirCall c;
  c.type = kFirFC;
  c.nco = kNcoExact;
  c.decimation = decimation;
  c.taps = lowPassTaps;
  c.tapCount = numLowPassTaps;
  c.input = input;
  c.output = output;
  c.numOutputs = numElements;
  c.sampleRate = rfSampleRate;
  c.frequencyShift = tuningFrequency - channelFrequency;  // ref: src/am.cu:68
  c.firstSampleIndex = firstSampleIndex;
  c.epilogue = kFirEpiAmEnvelope;  // 2 * sat(|y|) - 1 in the FIR's store path (ref: src/am.cu:49)
  cudaError_t st = enqueueFir(c, cudaStream);
  if (st != cudaErrorNotSupported) return st;
  // shapes without the fused stage: low-pass values through pool scratch, then the envelope kernel
  void* lowPassed = nullptr;
  cudaMemPool_t pool = nullptr;
  st = scratchPool(cudaDevice, &pool);
  if (st != cudaSuccess) return st;
  st = cudaMallocFromPoolAsync(&lowPassed, numElements * sizeof(cuComplex), pool, cudaStream);
  if (st != cudaSuccess) return st;
  c.epilogue = kFirEpiNone;
  c.output = lowPassed;
  st = enqueueFir(c, cudaStream);
  if (st == cudaSuccess) {
    const unsigned long long per = (unsigned long long)kDemodThreads * kDemodPerThread;
    const unsigned long long blocks = (numElements + per - 1) / per;
    if (blocks > 0x7fffffffull) {
      st = cudaErrorInvalidValue;
    } else {
      const int aligned = ((uintptr_t)output % 16 == 0) ? 1 : 0;
      quadAmDemodKernel<<<(unsigned)blocks, kDemodThreads, 0, cudaStream>>>((const float2*)lowPassed, output, numElements,
                                                                            aligned);
      st = cudaPeekAtLastError();
    }
  }
  const cudaError_t fr = cudaFreeAsync(lowPassed, cudaStream);
  return st != cudaSuccess ? st : fr;
}
