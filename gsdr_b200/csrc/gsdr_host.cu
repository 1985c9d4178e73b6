// gsdr_host.cu — the host-side batching / sharding layer declared in <gsdr/b200.h>.
// New code (the reference has no multi-call orchestration at all: ref: src/fir.cu:73-96 is one launch per
// call); it automates the reference's documented streaming contract — caller-supplied overlap and running
// firstSampleIndex (ref: include/gsdr/fm.h:26,34).
#include <gsdr/b200.h>

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
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

// Time shards of at least 64Ki outputs start on a multiple of 2048 outputs: the tensor-core FIR kernel works in tiles
// of 1024 outputs (2048 at decimation 4) counted from the first output of a call, and its rounding depends on an
// output's position in the tile — with aligned shards every output sits where it sits in the unsharded call, so the
// shards reproduce its bits.
constexpr uint64_t kShardAlign = 2048, kShardAlignFrom = 65536;
static uint64_t timeSplitPoint(uint64_t n, uint32_t shards, uint32_t s) noexcept {
  const uint64_t p = splitPoint(n, shards, s);
  if (s == 0 || s >= shards || n / shards < kShardAlignFrom) return p;
  return p - p % kShardAlign;
}

GSDR_C_LINKAGE int gsdrShardPlanTime(uint64_t numOutputs, uint64_t decimation, uint64_t tapCount,
                                     uint64_t firstSampleIndex, uint32_t numShards, uint32_t shardIndex,
                                     gsdrShard* shard) GSDR_NO_EXCEPT {
  if (!shard || numShards == 0 || shardIndex >= numShards || decimation == 0) return -1;
  const uint64_t a = timeSplitPoint(numOutputs, numShards, shardIndex);
  const uint64_t b = timeSplitPoint(numOutputs, numShards, shardIndex + 1);
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
    // an output chunk is at most 4x its input chunk (int8 I/Q in, cuComplex out, decimation 1)
    if (st == cudaSuccess) st = cudaMalloc(&b, chunkInputBytes * 4);
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
  bool int8Input = false;  // interleaved int8 I/Q (2 bytes per complex sample), FC only
};

static size_t inElem(FirType t) noexcept { return (t == kFirFF || t == kFirCF) ? 4 : 8; }
static size_t outElem(FirType t) noexcept { return t == kFirFF ? 4 : 8; }
static size_t tapElem(FirType t) noexcept { return (t == kFirCC || t == kFirCF) ? 8 : 4; }

static cudaError_t runHostJob(gsdrHostPipeline* p, const HostJob& j) noexcept {
  if (!p) return cudaErrorInvalidValue;
  if (j.numOutputs == 0) return cudaSuccess;
  if (j.decimation == 0) return cudaErrorInvalidValue;
  const size_t ie = j.int8Input ? 2 : inElem(j.type), oe = outElem(j.type), te = tapElem(j.type);
  if (j.int8Input && j.type != kFirFC) return cudaErrorInvalidValue;
  if (j.tapCount * te > p->tapsCapacityBytes) return cudaErrorInvalidValue;
  const size_t chunkElems = p->chunkInputBytes / ie;
  if (chunkElems < j.tapCount + j.decimation) return cudaErrorInvalidValue;  // a chunk must hold at least one window
  size_t outsPerChunk = j.tapCount ? (chunkElems - j.tapCount) / j.decimation + 1 : chunkElems;
  if (outsPerChunk >= kShardAlignFrom) outsPerChunk -= outsPerChunk % kShardAlign;  // see timeSplitPoint
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
    if (j.int8Input) {
      st = enqueueFirInt8(j.nco != kNcoNone, j.sampleRate, j.frequencyShift, j.firstSampleIndex + firstIn, j.decimation,
                          (const float*)p->dTaps, j.tapCount, (const signed char*)p->dIn[slot], (float2*)p->dOut[slot],
                          n, s);
    } else {
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
    }
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

GSDR_C_LINKAGE cudaError_t gsdrFirFCInt8Host(gsdrHostPipeline* pipeline, size_t decimation, const float* taps,
                                             size_t tapCount, const int8_t* input, cuComplex* output,
                                             size_t numOutputs) GSDR_NO_EXCEPT {
  HostJob j{kFirFC, kNcoNone, 0.f, 0.f, 0, decimation, taps, tapCount, input, output, numOutputs};
  j.int8Input = true;
  return runHostJob(pipeline, j);
}

GSDR_C_LINKAGE cudaError_t gsdrAdjustFrequencyFirFCInt8Host(gsdrHostPipeline* pipeline, float sampleRate,
                                                            float frequencyShift, size_t firstSampleIndex,
                                                            size_t decimation, const float* taps, size_t tapCount,
                                                            const int8_t* input, cuComplex* output,
                                                            size_t numOutputs) GSDR_NO_EXCEPT {
  HostJob j{kFirFC, kNcoExact, sampleRate, frequencyShift, firstSampleIndex, decimation, taps, tapCount, input, output,
            numOutputs};
  j.int8Input = true;
  return runHostJob(pipeline, j);
}

// ---- single-process multi-GPU over host buffers: one pipeline and one host thread per device --------------------

template <class MakeJob>
static cudaError_t runOnPipelines(gsdrHostPipeline* const* pipelines, int numPipelines, MakeJob makeJob) noexcept {
  if (!pipelines || numPipelines < 1) return cudaErrorInvalidValue;
  std::vector<cudaError_t> results((size_t)numPipelines, cudaSuccess);
  try {
    std::vector<std::thread> workers;
    for (int s = 0; s < numPipelines; s++) {
      workers.emplace_back([&, s]() { results[(size_t)s] = makeJob(s); });
    }
    for (std::thread& t : workers) t.join();
  } catch (...) {
    return cudaErrorUnknown;
  }
  for (cudaError_t r : results)
    if (r != cudaSuccess) return r;
  return cudaSuccess;
}

static cudaError_t timeShardedHost(gsdrHostPipeline* const* pipelines, int numPipelines, NcoMode nco, float sampleRate,
                                   float frequencyShift, size_t firstSampleIndex, size_t decimation, const float* taps,
                                   size_t tapCount, const cuComplex* input, cuComplex* output,
                                   size_t numOutputs) noexcept {
  if (decimation == 0) return cudaErrorInvalidValue;
  return runOnPipelines(pipelines, numPipelines, [&](int s) -> cudaError_t {
    gsdrShard sh;
    if (gsdrShardPlanTime(numOutputs, decimation, tapCount, firstSampleIndex, (uint32_t)numPipelines, (uint32_t)s,
                          &sh) != 0)
      return cudaErrorInvalidValue;
    return runHostJob(pipelines[s], HostJob{kFirFC, nco, sampleRate, frequencyShift, (size_t)sh.firstSampleIndex,
                                            decimation, taps, tapCount, input + sh.firstInput,
                                            output + sh.firstOutput, (size_t)sh.numOutputs});
  });
}

GSDR_C_LINKAGE cudaError_t gsdrFirFCMultiGpuHost(gsdrHostPipeline* const* pipelines, int numPipelines,
                                                 size_t decimation, const float* taps, size_t tapCount,
                                                 const cuComplex* input, cuComplex* output,
                                                 size_t numOutputs) GSDR_NO_EXCEPT {
  return timeShardedHost(pipelines, numPipelines, kNcoNone, 0.f, 0.f, 0, decimation, taps, tapCount, input, output,
                         numOutputs);
}

GSDR_C_LINKAGE cudaError_t gsdrAdjustFrequencyFirFCMultiGpuHost(gsdrHostPipeline* const* pipelines, int numPipelines,
                                                                float sampleRate, float frequencyShift,
                                                                size_t firstSampleIndex, size_t decimation,
                                                                const float* taps, size_t tapCount,
                                                                const cuComplex* input, cuComplex* output,
                                                                size_t numOutputs) GSDR_NO_EXCEPT {
  return timeShardedHost(pipelines, numPipelines, kNcoExact, sampleRate, frequencyShift, firstSampleIndex, decimation,
                         taps, tapCount, input, output, numOutputs);
}

GSDR_C_LINKAGE cudaError_t gsdrFirFCChannelsMultiGpuHost(gsdrHostPipeline* const* pipelines, int numPipelines,
                                                         size_t decimation, const float* taps, size_t tapCount,
                                                         const cuComplex* input, size_t inputStride, cuComplex* output,
                                                         size_t outputStride, size_t numOutputs,
                                                         size_t numChannels) GSDR_NO_EXCEPT {
  if (decimation == 0) return cudaErrorInvalidValue;
  return runOnPipelines(pipelines, numPipelines, [&](int s) -> cudaError_t {
    uint64_t c0 = 0, cn = 0;
    if (gsdrShardPlanChannels(numChannels, (uint32_t)numPipelines, (uint32_t)s, &c0, &cn) != 0)
      return cudaErrorInvalidValue;
    for (uint64_t c = c0; c < c0 + cn; c++) {
      const cudaError_t st =
          runHostJob(pipelines[s], HostJob{kFirFC, kNcoNone, 0.f, 0.f, 0, decimation, taps, tapCount,
                                           input + c * inputStride, output + c * outputStride, numOutputs});
      if (st != cudaSuccess) return st;
    }
    return cudaSuccess;
  });
}

// ---- device-resident multi-GPU executor: persistent worker per device, fused gather by peer stores --------------

struct gsdrMultiGpu {
  struct Worker {
    int32_t device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool peerOk = false;
    std::thread thread;
    cudaError_t result = cudaSuccess;
    float ms = 0.f;
  };
  std::vector<Worker> workers;
  std::mutex mu;
  std::condition_variable wake, done;
  uint64_t generation = 0;
  int pending = 0;
  bool quit = false;
  std::function<void(int)> job;  // runs on worker g's thread with devices[g] current
};

static void multiGpuWorker(gsdrMultiGpu* mg, int g) noexcept {
  cudaSetDevice(mg->workers[(size_t)g].device);
  uint64_t seen = 0;
  for (;;) {
    std::unique_lock<std::mutex> lock(mg->mu);
    mg->wake.wait(lock, [&] { return mg->quit || mg->generation != seen; });
    if (mg->quit) return;
    seen = mg->generation;
    lock.unlock();
    mg->job(g);
    lock.lock();
    if (--mg->pending == 0) mg->done.notify_all();
  }
}

static cudaError_t multiGpuRun(gsdrMultiGpu* mg, std::function<void(int)> job) noexcept {
  std::unique_lock<std::mutex> lock(mg->mu);
  mg->job = std::move(job);
  mg->pending = (int)mg->workers.size();
  mg->generation++;
  mg->wake.notify_all();
  mg->done.wait(lock, [&] { return mg->pending == 0; });
  for (const auto& w : mg->workers)
    if (w.result != cudaSuccess) return w.result;
  return cudaSuccess;
}

GSDR_C_LINKAGE void gsdrMultiGpuDestroy(gsdrMultiGpu* mg) GSDR_NO_EXCEPT {
  if (!mg) return;
  {
    std::lock_guard<std::mutex> lock(mg->mu);
    mg->quit = true;
  }
  mg->wake.notify_all();
  for (auto& w : mg->workers) {
    if (w.thread.joinable()) w.thread.join();
    DeviceScope scope(w.device);
    if (w.stream) cudaStreamDestroy(w.stream);
    if (w.e0) cudaEventDestroy(w.e0);
    if (w.e1) cudaEventDestroy(w.e1);
  }
  delete mg;
}

GSDR_C_LINKAGE cudaError_t gsdrMultiGpuCreate(const int32_t* devices, int numDevices,
                                              gsdrMultiGpu** executor) GSDR_NO_EXCEPT {
  if (!devices || !executor || numDevices < 1 || numDevices > 64) return cudaErrorInvalidValue;
  *executor = nullptr;
  gsdrMultiGpu* mg = new (std::nothrow) gsdrMultiGpu();
  if (!mg) return cudaErrorMemoryAllocation;
  cudaError_t st = cudaSuccess;
  try {
    mg->workers.resize((size_t)numDevices);
    for (int g = 0; g < numDevices && st == cudaSuccess; g++) {
      auto& w = mg->workers[(size_t)g];
      w.device = devices[g];
      DeviceScope scope(w.device);
      st = scope.status();
      if (st == cudaSuccess) st = cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking);
      if (st == cudaSuccess) st = cudaEventCreate(&w.e0);
      if (st == cudaSuccess) st = cudaEventCreate(&w.e1);
      if (st != cudaSuccess) break;
      if (g == 0 || devices[g] == devices[0]) {
        w.peerOk = true;
      } else {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, w.device, devices[0]) == cudaSuccess && can) {
          const cudaError_t pe = cudaDeviceEnablePeerAccess(devices[0], 0);
          w.peerOk = (pe == cudaSuccess || pe == cudaErrorPeerAccessAlreadyEnabled);
          (void)cudaGetLastError();
        }
      }
    }
    if (st == cudaSuccess) {
      for (int g = 0; g < numDevices; g++) mg->workers[(size_t)g].thread = std::thread(multiGpuWorker, mg, g);
    }
  } catch (...) {
    st = cudaErrorUnknown;
  }
  if (st != cudaSuccess) {
    gsdrMultiGpuDestroy(mg);
    return st;
  }
  *executor = mg;
  return cudaSuccess;
}

GSDR_C_LINKAGE int gsdrMultiGpuPeerOk(const gsdrMultiGpu* mg, int g) GSDR_NO_EXCEPT {
  if (!mg || g < 0 || g >= (int)mg->workers.size()) return 0;
  return mg->workers[(size_t)g].peerOk ? 1 : 0;
}

GSDR_C_LINKAGE cudaError_t gsdrFirFCMultiGpu(gsdrMultiGpu* mg, float sampleRate, float frequencyShift,
                                             size_t firstSampleIndex, size_t decimation, const float* const* taps,
                                             size_t tapCount, const cuComplex* const* inputs, cuComplex* const* outputs,
                                             cuComplex* gatherOutput, size_t numOutputs, int repeats,
                                             float* elapsedMs) GSDR_NO_EXCEPT {
  if (!mg || !taps || !inputs || decimation == 0 || (!outputs && !gatherOutput)) return cudaErrorInvalidValue;
  if (repeats < 1) repeats = 1;
  const int n = (int)mg->workers.size();
  if (gatherOutput) {
    for (int g = 0; g < n; g++)
      if (!mg->workers[(size_t)g].peerOk) return cudaErrorPeerAccessUnsupported;
  }
  const bool nco = !(sampleRate == 0.0f && frequencyShift == 0.0f);
  try {
    const cudaError_t st = multiGpuRun(mg, [&](int g) {
      auto& w = mg->workers[(size_t)g];
      w.result = cudaSuccess;
      w.ms = 0.f;
      gsdrShard sh;
      if (gsdrShardPlanTime(numOutputs, decimation, tapCount, firstSampleIndex, (uint32_t)n, (uint32_t)g, &sh) != 0) {
        w.result = cudaErrorInvalidValue;
        return;
      }
      FirCall c;
      c.type = kFirFC;
      c.nco = nco ? kNcoExact : kNcoNone;
      c.decimation = decimation;
      c.taps = taps[g];
      c.tapCount = tapCount;
      c.input = inputs[g];
      c.output = gatherOutput ? (void*)(gatherOutput + sh.firstOutput) : (void*)outputs[g];
      c.numOutputs = (size_t)sh.numOutputs;
      c.sampleRate = sampleRate;
      c.frequencyShift = frequencyShift;
      c.firstSampleIndex = (size_t)sh.firstSampleIndex;
      cudaError_t st = cudaEventRecord(w.e0, w.stream);
      for (int r = 0; r < repeats && st == cudaSuccess; r++) st = enqueueFir(c, w.stream);
      if (st == cudaSuccess) st = cudaEventRecord(w.e1, w.stream);
      const cudaError_t sy = cudaStreamSynchronize(w.stream);
      if (st == cudaSuccess) st = sy;
      if (st == cudaSuccess) st = cudaEventElapsedTime(&w.ms, w.e0, w.e1);
      w.result = st;
    });
    if (elapsedMs) {
      float worst = 0.f;
      for (const auto& w : mg->workers) worst = std::max(worst, w.ms);
      *elapsedMs = worst / (float)repeats;
    }
    return st;
  } catch (...) {
    return cudaErrorUnknown;
  }
}

GSDR_C_LINKAGE cudaError_t gsdrMultiGpuGather(gsdrMultiGpu* mg, size_t decimation, size_t tapCount,
                                              const cuComplex* const* outputs, cuComplex* dst, size_t numOutputs,
                                              float* elapsedMs) GSDR_NO_EXCEPT {
  if (!mg || !outputs || !dst || decimation == 0) return cudaErrorInvalidValue;
  const int n = (int)mg->workers.size();
  // all copies are issued on device 0's stream (the destination), each shard straight to its final offset
  auto& w0 = mg->workers[0];
  DeviceScope scope(w0.device);
  if (scope.status() != cudaSuccess) return scope.status();
  cudaError_t st = cudaEventRecord(w0.e0, w0.stream);
  for (int g = 0; g < n && st == cudaSuccess; g++) {
    gsdrShard sh;
    if (gsdrShardPlanTime(numOutputs, decimation, tapCount, 0, (uint32_t)n, (uint32_t)g, &sh) != 0)
      return cudaErrorInvalidValue;
    if (sh.numOutputs == 0) continue;
    st = cudaMemcpyPeerAsync(dst + sh.firstOutput, w0.device, outputs[g], mg->workers[(size_t)g].device,
                             (size_t)sh.numOutputs * sizeof(cuComplex), w0.stream);
  }
  if (st == cudaSuccess) st = cudaEventRecord(w0.e1, w0.stream);
  const cudaError_t sy = cudaStreamSynchronize(w0.stream);
  if (st == cudaSuccess) st = sy;
  if (st == cudaSuccess && elapsedMs) st = cudaEventElapsedTime(elapsedMs, w0.e0, w0.e1);
  return st;
}

// ---- shared output buffers (one process per GPU) -----------------------------------------------------------------

GSDR_C_LINKAGE cudaError_t gsdrSharedBufferCreate(size_t bytes, int32_t cudaDevice, void** devicePointer,
                                                  unsigned char handle[64]) GSDR_NO_EXCEPT {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the handle travels as 64 opaque bytes");
  if (!devicePointer || !handle || bytes == 0) return cudaErrorInvalidValue;
  DeviceScope scope(cudaDevice);
  if (scope.status() != cudaSuccess) return scope.status();
  void* p = nullptr;
  cudaError_t st = cudaMalloc(&p, bytes);
  if (st != cudaSuccess) return st;
  cudaIpcMemHandle_t h;
  st = cudaIpcGetMemHandle(&h, p);
  if (st != cudaSuccess) {
    cudaFree(p);
    return st;
  }
  std::copy((const unsigned char*)&h, (const unsigned char*)&h + 64, handle);
  *devicePointer = p;
  return cudaSuccess;
}

GSDR_C_LINKAGE cudaError_t gsdrSharedBufferOpen(const unsigned char handle[64], int32_t cudaDevice,
                                                void** devicePointer) GSDR_NO_EXCEPT {
  if (!devicePointer || !handle) return cudaErrorInvalidValue;
  DeviceScope scope(cudaDevice);
  if (scope.status() != cudaSuccess) return scope.status();
  cudaIpcMemHandle_t h;
  std::copy(handle, handle + 64, (unsigned char*)&h);
  return cudaIpcOpenMemHandle(devicePointer, h, cudaIpcMemLazyEnablePeerAccess);
}

GSDR_C_LINKAGE cudaError_t gsdrSharedBufferClose(void* devicePointer, int32_t cudaDevice) GSDR_NO_EXCEPT {
  DeviceScope scope(cudaDevice);
  if (scope.status() != cudaSuccess) return scope.status();
  return cudaIpcCloseMemHandle(devicePointer);
}

GSDR_C_LINKAGE cudaError_t gsdrSharedBufferDestroy(void* devicePointer, int32_t cudaDevice) GSDR_NO_EXCEPT {
  DeviceScope scope(cudaDevice);
  if (scope.status() != cudaSuccess) return scope.status();
  return cudaFree(devicePointer);
}
