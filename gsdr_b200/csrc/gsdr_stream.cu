// gsdr_stream.cu — block-streaming state over the stateless FIR entry points (<gsdr/stream.h>).
// New code: the reference leaves the overlap of tapCount - 1 samples and the running firstSampleIndex to every
// caller (ref: include/gsdr/fm.h:26,34; src/fm.cu:202).  Pure host bookkeeping + small device-to-device copies;
// the filtering itself is enqueueFir() — the same kernels as gsdrFirFC / gsdrFirFF / gsdrAdjustFrequencyFirFC.
#include <gsdr/stream.h>

#include <algorithm>
#include <new>

#include "launch.h"

using namespace gsdr_b200;

struct gsdrFirStream {
  int firType = GSDR_STREAM_FIR_FC;
  size_t decimation = 1, tapCount = 0;
  float sampleRate = 0.0f, frequencyShift = 0.0f;
  size_t firstSampleIndex = 0;
  int32_t device = 0;
  size_t elemBytes = 8;   // input element size (cuComplex, float, or 2 for int8 IQ)
  size_t outBytes = 8;    // output element size
  uint32_t align = 2;     // input samples per 16 bytes
  float* taps = nullptr;  // device copy
  unsigned char* buffer[2] = {nullptr, nullptr};  // carry / staging, ping-pong
  int current = 0;
  uint64_t totalInputs = 0;  // samples pushed so far
  uint64_t nextStart = 0;    // absolute index of the first sample of the next output's window
};

GSDR_C_LINKAGE int gsdrFirStreamPlan(uint64_t decimation, uint64_t tapCount, uint64_t totalInputs, uint64_t nextStart,
                                     uint64_t numInputs, uint32_t align, gsdrStreamPlan* plan) GSDR_NO_EXCEPT {
  if (!plan || decimation == 0 || tapCount == 0 || align == 0) return -1;
  gsdrStreamPlan p{};
  const uint64_t D = decimation, T = tapCount;
  const uint64_t carry = totalInputs > nextStart ? totalInputs - nextStart : 0;
  const uint64_t skip = nextStart > totalInputs ? nextStart - totalInputs : 0;
  const uint64_t skipped = std::min(skip, numInputs);
  const uint64_t fresh = numInputs - skipped;  // block samples that take part
  const uint64_t avail = carry + fresh;
  const uint64_t nOut = avail >= T ? (avail - T) / D + 1 : 0;
  uint64_t head = 0;
  if (nOut > 0) {
    if (carry > 0) head = std::min(nOut, (carry + D - 1) / D);  // windows that start inside the carried samples
    // a few more head outputs when that puts the first body window on a 16-byte boundary of the block
    for (uint32_t extra = 0; extra < align; extra++) {
      const uint64_t h = head + extra;
      if (h > nOut) break;
      if (h == nOut || (skipped + h * D - carry) % align == 0) {
        head = h;
        break;
      }
    }
  }
  const uint64_t headSpan = head > 0 ? (head - 1) * D + T : 0;  // staging samples the head outputs read
  p.numOutputs = nOut;
  p.skippedInputs = skipped;
  p.headOutputs = head;
  p.headNewInputs = headSpan > carry ? headSpan - carry : 0;
  p.bodyOutputs = nOut - head;
  p.bodyOffset = p.bodyOutputs > 0 ? skipped + head * D - carry : 0;
  p.carryLength = carry;
  p.newNextStart = nextStart + nOut * D;
  const uint64_t newTotal = totalInputs + numInputs;
  p.newCarryLength = newTotal > p.newNextStart ? newTotal - p.newNextStart : 0;
  *plan = p;
  return 0;
}

GSDR_C_LINKAGE cudaError_t gsdrFirStreamCreate(gsdrFirStream** stream, int firType, size_t decimation,
                                               const float* taps, size_t tapCount, float sampleRate,
                                               float frequencyShift, size_t firstSampleIndex,
                                               int32_t cudaDevice) GSDR_NO_EXCEPT {
  if (!stream) return cudaErrorInvalidValue;
  *stream = nullptr;
  if (decimation == 0 || tapCount == 0 || !taps) return cudaErrorInvalidValue;
  if (firType != GSDR_STREAM_FIR_FC && firType != GSDR_STREAM_FIR_FF && firType != GSDR_STREAM_FIR_FC_NCO &&
      firType != GSDR_STREAM_FIR_FC_INT8 && firType != GSDR_STREAM_FIR_FC_NCO_INT8) {
    return cudaErrorInvalidValue;
  }
  DeviceScope scope(cudaDevice);
  if (scope.status() != cudaSuccess) return scope.status();
  gsdrFirStream* s = new (std::nothrow) gsdrFirStream;
  if (!s) return cudaErrorMemoryAllocation;
  s->firType = firType;
  s->decimation = decimation;
  s->tapCount = tapCount;
  s->sampleRate = sampleRate;
  s->frequencyShift = frequencyShift;
  s->firstSampleIndex = firstSampleIndex;
  s->device = cudaDevice;
  const bool int8In = firType == GSDR_STREAM_FIR_FC_INT8 || firType == GSDR_STREAM_FIR_FC_NCO_INT8;
  s->elemBytes = firType == GSDR_STREAM_FIR_FF ? 4 : (int8In ? 2 : 8);
  s->outBytes = firType == GSDR_STREAM_FIR_FF ? 4 : 8;
  s->align = (uint32_t)(16 / s->elemBytes);
  // carry < tapCount; the staging span of the head outputs is below carry + (align + 1) * decimation + tapCount
  const size_t capacity = (2 * tapCount + (s->align + 2) * decimation + 8) * s->elemBytes;
  cudaError_t st = cudaMalloc((void**)&s->taps, tapCount * sizeof(float));
  if (st == cudaSuccess) st = cudaMalloc((void**)&s->buffer[0], capacity);
  if (st == cudaSuccess) st = cudaMalloc((void**)&s->buffer[1], capacity);
  if (st == cudaSuccess) st = cudaMemcpy(s->taps, taps, tapCount * sizeof(float), cudaMemcpyDeviceToDevice);
  if (st != cudaSuccess) {
    gsdrFirStreamDestroy(s);
    return st;
  }
  *stream = s;
  return cudaSuccess;
}

GSDR_C_LINKAGE void gsdrFirStreamDestroy(gsdrFirStream* s) GSDR_NO_EXCEPT {
  if (!s) return;
  DeviceScope scope(s->device);
  if (s->taps) cudaFree(s->taps);
  if (s->buffer[0]) cudaFree(s->buffer[0]);
  if (s->buffer[1]) cudaFree(s->buffer[1]);
  delete s;
}

GSDR_C_LINKAGE void gsdrFirStreamReset(gsdrFirStream* s) GSDR_NO_EXCEPT {
  if (!s) return;
  s->totalInputs = 0;
  s->nextStart = 0;
}

GSDR_C_LINKAGE size_t gsdrFirStreamNumOutputs(const gsdrFirStream* s, size_t numInputs) GSDR_NO_EXCEPT {
  if (!s) return 0;
  gsdrStreamPlan p;
  if (gsdrFirStreamPlan(s->decimation, s->tapCount, s->totalInputs, s->nextStart, numInputs, s->align, &p) != 0) return 0;
  return (size_t)p.numOutputs;
}

static cudaError_t streamFir(const gsdrFirStream* s, const void* input, void* output, size_t numOutputs,
                             uint64_t absoluteIndex, cudaStream_t stream) noexcept {
  if (s->firType == GSDR_STREAM_FIR_FC_INT8 || s->firType == GSDR_STREAM_FIR_FC_NCO_INT8) {
    return enqueueFirInt8(s->firType == GSDR_STREAM_FIR_FC_NCO_INT8, s->sampleRate, s->frequencyShift,
                          s->firstSampleIndex + (size_t)absoluteIndex, s->decimation, s->taps, s->tapCount,
                          (const signed char*)input, (float2*)output, numOutputs, stream);
  }
  FirCall c;
  c.type = s->firType == GSDR_STREAM_FIR_FF ? kFirFF : kFirFC;
  c.nco = s->firType == GSDR_STREAM_FIR_FC_NCO ? kNcoExact : kNcoNone;
  c.decimation = s->decimation;
  c.taps = s->taps;
  c.tapCount = s->tapCount;
  c.input = input;
  c.output = output;
  c.numOutputs = numOutputs;
  c.sampleRate = s->sampleRate;
  c.frequencyShift = s->frequencyShift;
  c.firstSampleIndex = s->firstSampleIndex + (size_t)absoluteIndex;
  return enqueueFir(c, stream);
}

GSDR_C_LINKAGE cudaError_t gsdrFirStreamPush(gsdrFirStream* s, const void* input, size_t numInputs, void* output,
                                             size_t* numOutputs, cudaStream_t stream) GSDR_NO_EXCEPT {
  if (numOutputs) *numOutputs = 0;
  if (!s || (numInputs > 0 && !input)) return cudaErrorInvalidValue;
  DeviceScope scope(s->device);
  if (scope.status() != cudaSuccess) return scope.status();
  gsdrStreamPlan p;
  if (gsdrFirStreamPlan(s->decimation, s->tapCount, s->totalInputs, s->nextStart, numInputs, s->align, &p) != 0) {
    return cudaErrorInvalidValue;
  }
  if (p.numOutputs > 0 && !output) return cudaErrorInvalidValue;
  const size_t eb = s->elemBytes;
  const unsigned char* fresh = (const unsigned char*)input + p.skippedInputs * eb;
  const uint64_t freshCount = numInputs - p.skippedInputs;
  unsigned char* staging = s->buffer[s->current];
  cudaError_t st = cudaSuccess;

  if (p.numOutputs == 0) {
    // not enough samples for a window yet: the block joins the carry
    if (freshCount > 0) {
      st = cudaMemcpyAsync(staging + p.carryLength * eb, fresh, freshCount * eb, cudaMemcpyDeviceToDevice, stream);
      if (st != cudaSuccess) return st;
    }
    s->totalInputs += numInputs;
    return cudaSuccess;
  }

  if (p.headOutputs > 0) {
    if (p.headNewInputs > 0) {
      st = cudaMemcpyAsync(staging + p.carryLength * eb, fresh, p.headNewInputs * eb, cudaMemcpyDeviceToDevice, stream);
      if (st != cudaSuccess) return st;
    }
    st = streamFir(s, staging, output, (size_t)p.headOutputs, s->nextStart, stream);
    if (st != cudaSuccess) return st;
  }
  if (p.bodyOutputs > 0) {
    st = streamFir(s, (const unsigned char*)input + p.bodyOffset * eb, (unsigned char*)output + p.headOutputs * s->outBytes,
                   (size_t)p.bodyOutputs, s->nextStart + p.headOutputs * s->decimation, stream);
    if (st != cudaSuccess) return st;
  }

  // the samples after the last consumed window start become the next carry, built in the other buffer
  if (p.newCarryLength > 0) {
    unsigned char* next = s->buffer[s->current ^ 1];
    const uint64_t consumed = p.numOutputs * s->decimation;  // counted from the start of the carry
    if (consumed >= p.carryLength) {
      st = cudaMemcpyAsync(next, fresh + (consumed - p.carryLength) * eb, p.newCarryLength * eb,
                           cudaMemcpyDeviceToDevice, stream);
    } else {
      const uint64_t keep = p.carryLength - consumed;  // tail of the old carry, then the whole block
      st = cudaMemcpyAsync(next, staging + consumed * eb, keep * eb, cudaMemcpyDeviceToDevice, stream);
      if (st == cudaSuccess && freshCount > 0) {
        st = cudaMemcpyAsync(next + keep * eb, fresh, freshCount * eb, cudaMemcpyDeviceToDevice, stream);
      }
    }
    if (st != cudaSuccess) return st;
  }
  s->current ^= 1;
  s->totalInputs += numInputs;
  s->nextStart = p.newNextStart;
  if (numOutputs) *numOutputs = (size_t)p.numOutputs;
  return cudaSuccess;
}
