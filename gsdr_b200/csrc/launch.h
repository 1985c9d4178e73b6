// launch.h — internal host-side launch interface shared by the C-ABI translation units.
#pragma once

#include <cuComplex.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace gsdr_b200 {

enum FirType : int { kFirFC = 0, kFirFF = 1, kFirCC = 2, kFirCF = 3 };
enum NcoMode : int { kNcoNone = 0, kNcoExact = 1, kNcoLiteral = 2 };
// Output stage fused into the FIR's store path (FC only): the complex outputs themselves, their AM envelope
// 2 * sat(|y|) - 1 (ref: src/am.cu:49), or the FM quadrature demodulation gain * arg(y[n+1] * conj(y[n]))
// (ref: src/quad_demod.cu:23-37).  With kFirEpiFmDemod, numOutputs counts demodulated values: the FIR produces one more.
enum FirEpilogue : int { kFirEpiNone = 0, kFirEpiAmEnvelope = 1, kFirEpiFmDemod = 2 };

struct FirCall {
  FirType type = kFirFC;
  NcoMode nco = kNcoNone;
  size_t decimation = 1;
  const void* taps = nullptr;
  size_t tapCount = 0;
  const void* input = nullptr;
  void* output = nullptr;
  size_t numOutputs = 0;
  // batching: numChannels independent filters in one launch; strides in elements (tapStride 0 = shared taps)
  size_t numChannels = 1;
  size_t inputStride = 0, outputStride = 0, tapStride = 0;
  // NCO
  float sampleRate = 0.0f, frequencyShift = 0.0f;
  size_t firstSampleIndex = 0;
  // fused output stage: `output` is then a float array (strides in floats)
  FirEpilogue epilogue = kFirEpiNone;
  float epilogueGain = 0.0f;
};

// Enqueue on `stream` of the CURRENT device (callers switch devices). No sync, no allocation.
// A call with a fused output stage returns cudaErrorNotSupported (nothing enqueued) when the kernel that fits its shape
// has no such stage; the caller then runs the stage as a separate kernel.
cudaError_t enqueueFir(const FirCall& call, cudaStream_t stream) noexcept;

// int8 IQ input (<gsdr/conversion.h>): conversion fused into the staging, optional exact NCO.  Same conventions.
cudaError_t enqueueFirInt8(bool nco, float sampleRate, float frequencyShift, size_t firstSampleIndex, size_t decimation,
                           const float* taps, size_t tapCount, const signed char* input, float2* output,
                           size_t numOutputs, cudaStream_t stream) noexcept;

// Saves the current device, switches to `device`, restores on destruction (ref: the SIMPLE_CUDA_FNC_START/END
// pair at src/cuComplexOperatorOverloads.cuh:74-93).
class DeviceScope {
 public:
  explicit DeviceScope(int32_t device) noexcept;
  ~DeviceScope() noexcept;
  cudaError_t status() const noexcept { return status_; }

 private:
  int previous_ = -1;
  bool switched_ = false;
  cudaError_t status_ = cudaSuccess;
};

uint64_t ncoPhaseStep(float frequencyShift, float sampleRate) noexcept;

}  // namespace gsdr_b200
