// gsdr_host.cu — the host-side batching / sharding layer declared in <gsdr/b200.h>.
// New code (the reference has no multi-call orchestration at all: ref: src/fir.cu:73-96 is one launch per
// call); it automates the reference's documented streaming contract — caller-supplied overlap and running
// firstSampleIndex (ref: include/gsdr/fm.h:26,34).
#include <gsdr/b200.h>

#include <algorithm>
#include <new>
#include <thread>
#include <vector>

#include "launch.h"

using namespace gsdr_b200;

// ---- size arithmetic --------------------------------------------------------------------------------------

GSDR_C_LINKAGE size_t gsdrFirNumOutputs(size_t numInputs, size_t tapCount, size_t decimation) GSDR_NO_EXCEPT {
  if (decimation == 0 || tapCount == 0 || numInputs < tapCount) return 0;
  return (numInputs - tapCount) / decimation + 1;
}

GSDR_C_LINKAGE size_t gsdrFirNumInputs(size_t numOutputs, size_t tapCount, size_t decimation) GSDR_NO_EXCEPT {
  if (numOutputs == 0) return 0;
  return (numOutputs - 1) * decimation + tapCount;
}

// ---- channel batching -------------------------------------------------------------------------------------

static cudaError_t batched(FirType type, size_t decimation, const void* taps, size_t tapCount, size_t tapStride,
                           const void* input, size_t inputStride, void* output, size_t outputStride,
                           size_t numOutputs, size_t numChannels, int32_t cudaDevice, cudaStream_t stream) noexcept {
  DeviceScope scope(cudaDevice);
  if (scope.status() != cudaSuccess) return scope.status();
  FirCall c;
  c.type = type;
  c.decimation = decimation;
  c.taps = taps;
  c.tapCount = tapCount;
  c.tapStride = tapStride;
  c.input = input;
  c.inputStride = inputStride;
  c.output = output;
  c.outputStride = outputStride;
  c.numOutputs = numOutputs;
  c.numChannels = numChannels;
  return enqueueFir(c, stream);
}

GSDR_C_LINKAGE cudaError_t gsdrFirFCBatched(size_t decimation, const float* taps, size_t tapCount, size_t tapStride,
                                            const cuComplex* input, size_t inputStride, cuComplex* output,
                                            size_t outputStride, size_t numOutputs, size_t numChannels,
                                            int32_t cudaDevice, cudaStream_t cudaStream) GSDR_NO_EXCEPT {
  return batched(kFirFC, decimation, taps, tapCount, tapStride, input, inputStride, output, outputStride, numOutputs,
                 numChannels, cudaDevice, cudaStream);
}

GSDR_C_LINKAGE cudaError_t gsdrFirFFBatched(size_t decimation, const float* taps, size_t tapCount, size_t tapStride,
                                            const float* input, size_t inputStride, float* output,
                                            size_t outputStride, size_t numOutputs, size_t numChannels,
                                            int32_t cudaDevice, cudaStream_t cudaStream) GSDR_NO_EXCEPT {
  return batched(kFirFF, decimation, taps, tapCount, tapStride, input, inputStride, output, outputStride, numOutputs,
                 numChannels, cudaDevice, cudaStream);
}

// ---- shard planning ---------------------------------------------------------------------------------------

static uint64_t splitPoint(uint64_t n, uint32_t shards, uint32_t s) noexcept {
  return (uint64_t)(((unsigned __int128)n * s) / shards);
}

GSDR_C_LINKAGE int gsdrShardPlanTime(uint64_t numOutputs, uint64_t decimation, uint64_t tapCount,
                                     uint64_t firstSampleIndex, uint32_t numShards, uint32_t shardIndex,
                                     gsdrShard* shard) GSDR_NO_EXCEPT {
  if (!shard || numShards == 0 || shardIndex >= numShards || decimation == 0) return -1;
  const uint64_t a = splitPoint(numOutputs, numShards, shardIndex);
  const uint64_t b = splitPoint(numOutputs, numShards, shardIndex + 1);
  shard->firstOutput = a;
  shard->numOutputs = b - a;
  shard->firstInput = a * decimation;
  shard->numInputs = (b > a) ? (b - a - 1) * decimation + tapCount : 0;
  shard->firstSampleIndex = firstSampleIndex + a * decimation;
  return 0;
}

GSDR_C_LINKAGE int gsdrShardPlanChannels(uint64_t numChannels, uint32_t numShards, uint32_t shardIndex,
                                         uint64_t* firstChannel, uint64_t* channelCount) GSDR_NO_EXCEPT {
  if (!firstChannel || !channelCount || numShards == 0 || shardIndex >= numShards) return -1;
  const uint64_t a = splitPoint(numChannels, numShards, shardIndex);
  const uint64_t b = splitPoint(numChannels, numShards, shardIndex + 1);
  *firstChannel = a;
  *channelCount = b - a;
  return 0;
}

// ---- host-buffer pipeline ---------------------------------------------------------------------------------

struct gsdrHostPipeline {
  int32_t device = 0;
  size_t chunkInputBytes = 0;
  int numBuffers = 0;
  size_t tapsCapacityBytes = 0;
  void* dTaps = nullptr;
  cudaEvent_t tapsReady = nullptr;
  std::vector<cudaStream_t> streams;
  std::vector<void*> dIn;
  std::vector<void*> dOut;
};

static void destroyPipeline(gsdrHostPipeline* p) noexcept {
  if (!p) return;
  DeviceScope scope(p->device);
  for (cudaStream_t s : p->streams)
    if (s) cudaStreamDestroy(s);
  for (void* b : p->dIn)
    if (b) cudaFree(b);
  for (void* b : p->dOut)
    if (b) cudaFree(b);
  if (p->dTaps) cudaFree(p->dTaps);
  if (p->tapsReady) cudaEventDestroy(p->tapsReady);
  delete p;
}

GSDR_C_LINKAGE cudaError_t gsdrHostPipelineCreate(int32_t cudaDevice, size_t chunkInputBytes, int numBuffers,
                                                  gsdrHostPipeline** pipeline) GSDR_NO_EXCEPT {
  if (!pipeline || numBuffers < 2 || numBuffers > 16 || chunkInputBytes < 4096) return cudaErrorInvalidValue;
  *pipeline = nullptr;
  DeviceScope scope(cudaDevice);
  if (scope.status() != cudaSuccess) return scope.status();
  gsdrHostPipeline* p = new (std::nothrow) gsdrHostPipeline();
  if (!p) return cudaErrorMemoryAllocation;
  p->device = cudaDevice;
  p->chunkInputBytes = chunkInputBytes;
  p->numBuffers = numBuffers;
  p->tapsCapacityBytes = (size_t)8 << 20;
  cudaError_t st = cudaMalloc(&p->dTaps, p->tapsCapacityBytes);
  if (st == cudaSuccess) st = cudaEventCreateWithFlags(&p->tapsReady, cudaEventDisableTiming);
  for (int i = 0; i < numBuffers && st == cudaSuccess; i++) {
    cudaStream_t s = nullptr;
    void* a = nullptr;
    void* b = nullptr;
    st = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    p->streams.push_back(s);
    if (st == cudaSuccess) st = cudaMalloc(&a, chunkInputBytes);
    p->dIn.push_back(a);
    // an output chunk can never be larger than its input chunk (decimation >= 1, complex out of real in at most 2x)
    if (st == cudaSuccess) st = cudaMalloc(&b, chunkInputBytes * 2);
    p->dOut.push_back(b);
  }
  if (st != cudaSuccess) {
    destroyPipeline(p);
    return st;
  }
  *pipeline = p;
  return cudaSuccess;
}

GSDR_C_LINKAGE void gsdrHostPipelineDestroy(gsdrHostPipeline* pipeline) GSDR_NO_EXCEPT { destroyPipeline(pipeline); }

struct HostJob {
  FirType type;
  NcoMode nco;
  float sampleRate, frequencyShift;
  size_t firstSampleIndex;
  size_t decimation;
  const void* taps;
  size_t tapCount;
  const void* input;
  void* output;
  size_t numOutputs;
};

static size_t inElem(FirType t) noexcept { return (t == kFirFF || t == kFirCF) ? 4 : 8; }
static size_t outElem(FirType t) noexcept { return t == kFirFF ? 4 : 8; }
static size_t tapElem(FirType t) noexcept { return (t == kFirCC || t == kFirCF) ? 8 : 4; }

static cudaError_t runHostJob(gsdrHostPipeline* p, const HostJob& j) noexcept {
  if (!p) return cudaErrorInvalidValue;
  if (j.numOutputs == 0) return cudaSuccess;
  if (j.decimation == 0) return cudaErrorInvalidValue;
  const size_t ie = inElem(j.type), oe = outElem(j.type), te = tapElem(j.type);
  if (j.tapCount * te > p->tapsCapacityBytes) return cudaErrorInvalidValue;
  const size_t chunkElems = p->chunkInputBytes / ie;
  if (chunkElems < j.tapCount + j.decimation) return cudaErrorInvalidValue;  // a chunk must hold at least one window
  const size_t outsPerChunk = j.tapCount ? (chunkElems - j.tapCount) / j.decimation + 1 : chunkElems;
  DeviceScope scope(p->device);
  if (scope.status() != cudaSuccess) return scope.status();

  cudaError_t st = cudaSuccess;
  if (j.tapCount) {
    st = cudaMemcpyAsync(p->dTaps, j.taps, j.tapCount * te, cudaMemcpyHostToDevice, p->streams[0]);
    if (st == cudaSuccess) st = cudaEventRecord(p->tapsReady, p->streams[0]);
    for (int i = 1; i < p->numBuffers && st == cudaSuccess; i++) st = cudaStreamWaitEvent(p->streams[i], p->tapsReady, 0);
  }
  size_t k = 0;
  for (size_t o0 = 0; o0 < j.numOutputs && st == cudaSuccess; o0 += outsPerChunk, k++) {
    const int slot = (int)(k % (size_t)p->numBuffers);
    cudaStream_t s = p->streams[slot];
    const size_t n = std::min(outsPerChunk, j.numOutputs - o0);
    const size_t firstIn = o0 * j.decimation;
    const size_t nIn = j.tapCount ? (n - 1) * j.decimation + j.tapCount : 0;
    if (nIn) {
      st = cudaMemcpyAsync(p->dIn[slot], (const unsigned char*)j.input + firstIn * ie, nIn * ie,
                           cudaMemcpyHostToDevice, s);
      if (st != cudaSuccess) break;
    }
    FirCall c;
    c.type = j.type;
    c.nco = j.nco;
    c.decimation = j.decimation;
    c.taps = p->dTaps;
    c.tapCount = j.tapCount;
    c.input = p->dIn[slot];
    c.output = p->dOut[slot];
    c.numOutputs = n;
    c.sampleRate = j.sampleRate;
    c.frequencyShift = j.frequencyShift;
    c.firstSampleIndex = j.firstSampleIndex + firstIn;
    st = enqueueFir(c, s);
    if (st != cudaSuccess) break;
    st = cudaMemcpyAsync((unsigned char*)j.output + o0 * oe, p->dOut[slot], n * oe, cudaMemcpyDeviceToHost, s);
  }
  for (cudaStream_t s : p->streams) {
    const cudaError_t e = cudaStreamSynchronize(s);
    if (st == cudaSuccess) st = e;
  }
  return st;
}

GSDR_C_LINKAGE cudaError_t gsdrFirFCHost(gsdrHostPipeline* pipeline, size_t decimation, const float* taps,
                                         size_t tapCount, const cuComplex* input, cuComplex* output,
                                         size_t numOutputs) GSDR_NO_EXCEPT {
  return runHostJob(pipeline, HostJob{kFirFC, kNcoNone, 0.f, 0.f, 0, decimation, taps, tapCount, input, output, numOutputs});
}

GSDR_C_LINKAGE cudaError_t gsdrFirFFHost(gsdrHostPipeline* pipeline, size_t decimation, const float* taps,
                                         size_t tapCount, const float* input, float* output,
                                         size_t numOutputs) GSDR_NO_EXCEPT {
  return runHostJob(pipeline, HostJob{kFirFF, kNcoNone, 0.f, 0.f, 0, decimation, taps, tapCount, input, output, numOutputs});
}

GSDR_C_LINKAGE cudaError_t gsdrAdjustFrequencyFirFCHost(gsdrHostPipeline* pipeline, float sampleRate,
                                                        float frequencyShift, size_t firstSampleIndex,
                                                        size_t decimation, const float* taps, size_t tapCount,
                                                        const cuComplex* input, cuComplex* output,
                                                        size_t numOutputs) GSDR_NO_EXCEPT {
  return runHostJob(pipeline, HostJob{kFirFC, kNcoExact, sampleRate, frequencyShift, firstSampleIndex, decimation, taps,
                                      tapCount, input, output, numOutputs});
}

GSDR_C_LINKAGE cudaError_t gsdrFirFCMultiGpuHost(gsdrHostPipeline* const* pipelines, int numPipelines,
                                                 size_t decimation, const float* taps, size_t tapCount,
                                                 const cuComplex* input, cuComplex* output,
                                                 size_t numOutputs) GSDR_NO_EXCEPT {
  if (!pipelines || numPipelines < 1 || decimation == 0) return cudaErrorInvalidValue;
  std::vector<cudaError_t> results((size_t)numPipelines, cudaSuccess);
  try {
    std::vector<std::thread> workers;
    for (int s = 0; s < numPipelines; s++) {
      workers.emplace_back([&, s]() {
        gsdrShard sh;
        if (gsdrShardPlanTime(numOutputs, decimation, tapCount, 0, (uint32_t)numPipelines, (uint32_t)s, &sh) != 0) {
          results[(size_t)s] = cudaErrorInvalidValue;
          return;
        }
        results[(size_t)s] = runHostJob(
            pipelines[s], HostJob{kFirFC, kNcoNone, 0.f, 0.f, 0, decimation, taps, tapCount, input + sh.firstInput,
                                  output + sh.firstOutput, (size_t)sh.numOutputs});
      });
    }
    for (std::thread& t : workers) t.join();
  } catch (...) {
    return cudaErrorUnknown;
  }
  for (cudaError_t r : results)
    if (r != cudaSuccess) return r;
  return cudaSuccess;
}
