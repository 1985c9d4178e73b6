// gsdr_fir.cu — C-ABI entry points of <gsdr/fir.h> and <gsdr/adjust_frequency.h> and the launch logic behind
// them.  Replaces the host wrappers at ref: src/fir.cu:73-171 (and the per-thread k_AdjustFrequency call at
// ref: src/fm.cu:46-56).  No CPU fallback exists: every path ends in a kernel launch on the caller's stream.
#include <gsdr/adjust_frequency.h>
#include <gsdr/b200.h>
#include <gsdr/fir.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "fir_kernels.cuh"
#include "launch.h"

namespace gsdr_b200 {

// ---------------------------------------------------------------------------------------------------------
// device scope + error reporting (one stderr line on failure, ref: include/gsdr/cuda_util.h:59-82)
// ---------------------------------------------------------------------------------------------------------
static cudaError_t report(cudaError_t st, const char* what) noexcept {
  if (st != cudaSuccess) {
    std::fprintf(stderr, "gsdr-b200: error %d - %s - %s\n", (int)st, cudaGetErrorName(st), what);
    (void)cudaGetLastError();  // the error is returned to the caller; do not leave it in the runtime's last-error slot
  }
  return st;
}

DeviceScope::DeviceScope(int32_t device) noexcept {
  status_ = cudaGetDevice(&previous_);
  if (status_ != cudaSuccess) return;
  if (previous_ != device) {
    status_ = report(cudaSetDevice(device), "cudaSetDevice(cudaDevice)");
    switched_ = (status_ == cudaSuccess);
  }
}

DeviceScope::~DeviceScope() noexcept {
  if (switched_) report(cudaSetDevice(previous_), "cudaSetDevice(previousCudaDevice)");
}

uint64_t ncoPhaseStep(float frequencyShift, float sampleRate) noexcept {
  const double r = (double)frequencyShift / (double)sampleRate;
  if (!std::isfinite(r)) return 0;
  const double frac = r - std::floor(r + 0.5);  // [-0.5, 0.5)
  return (uint64_t)(int64_t)std::llrint(frac * 18446744073709551616.0);
}

// ---------------------------------------------------------------------------------------------------------
// polyphase kernel variants
// ---------------------------------------------------------------------------------------------------------
struct PolyVariant {
  int R, tg, psplit, nbuf, minBlocks;
  int threads() const { return tg * psplit; }
};
// X(id, R, TG, PSPLIT, NBUF, MINB)
#define GSDR_POLY_VARIANTS(X) \
  X(0, 8, 128, 1, 1, 2)       \
  X(1, 8, 64, 1, 1, 3)        \
  X(2, 8, 32, 1, 1, 4)        \
  X(3, 8, 64, 2, 2, 2)        \
  X(4, 8, 64, 2, 1, 2)        \
  X(5, 8, 128, 1, 2, 1)       \
  X(6, 8, 64, 1, 2, 2)        \
  X(7, 8, 32, 4, 2, 2)        \
  X(8, 16, 64, 1, 1, 2)       \
  X(9, 8, 32, 2, 2, 4)        \
  X(10, 8, 128, 2, 1, 2)      \
  X(11, 8, 64, 4, 2, 2)

static constexpr PolyVariant kVariants[] = {
#define X(id, r, tg, ps, nb, mb) {r, tg, ps, nb, mb},
    GSDR_POLY_VARIANTS(X)
#undef X
};
static constexpr int kNumVariants = (int)(sizeof(kVariants) / sizeof(kVariants[0]));

static std::atomic<int> gDebugFlags{0};
static std::atomic<int> gForcedVariant{-1};  // -1 auto, -2 direct kernel, >=0 variant id (test/tuning hook)

struct DeviceInfo {
  std::once_flag once;
  int smCount = 0;
  int maxSmemOptin = 0;
  cudaError_t status = cudaSuccess;
};
static DeviceInfo gDevices[64];

static const DeviceInfo* deviceInfo(int dev) noexcept {
  if (dev < 0 || dev >= 64) return nullptr;
  DeviceInfo& d = gDevices[dev];
  std::call_once(d.once, [&]() {
    d.status = cudaDeviceGetAttribute(&d.smCount, cudaDevAttrMultiProcessorCount, dev);
    if (d.status == cudaSuccess)
      d.status = cudaDeviceGetAttribute(&d.maxSmemOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  });
  return &d;
}

struct PolyGeom {
  unsigned Jpad, pitch, stageElems;
  size_t smemBytes;
};

static bool polyGeometry(const PolyVariant& v, size_t D, size_t T, PolyGeom* g) noexcept {
  const size_t R = (size_t)v.R;
  if (D > 4096 || D < (size_t)v.psplit) return false;
  const size_t J = (T + D - 1) / D;
  const size_t Jpad = (J + 2 * R - 1) / (2 * R) * (2 * R);  // the tap loop is unrolled over two register blocks
  const size_t bout = R * (size_t)v.tg;
  const size_t rowStaged = bout + Jpad;
  const size_t rowAlloc = rowStaged + R;  // the refills run one block of samples past the last one used
  size_t pitch = polyPos((unsigned)(rowAlloc - 1), (unsigned)R) + 1;
  while (pitch % 4 != 2) pitch++;  // even (16-byte rows) and == 2 mod 4 (staging stores spread over banks)
  if (Jpad > (1u << 20)) return false;
  const size_t smem = (D * Jpad + 2 * R) * 4 + (size_t)v.nbuf * D * pitch * 8;
  if (smem > (size_t)1 << 20) return false;
  g->Jpad = (unsigned)Jpad;
  g->pitch = (unsigned)pitch;
  g->stageElems = (unsigned)(rowStaged * D);
  g->smemBytes = smem;
  return true;
}

template <int MODE, int R, int TG, int PSPLIT, int NBUF, int MINB>
static cudaError_t launchPolyT(PolyParams& P, size_t smem, int dev, int smCount, cudaStream_t stream) noexcept {
  static std::atomic<unsigned long long> configured{0};  // bit per device
  static std::atomic<int> occCache[64];                  // CTAs per SM at the max dynamic smem actually used
  auto kernel = firPolyKernel<MODE, R, TG, PSPLIT, NBUF, MINB>;
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(configured.load(std::memory_order_acquire) & bit)) {
    cudaError_t st = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (st != cudaSuccess) return report(st, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    configured.fetch_or(bit, std::memory_order_release);
  }
  (void)occCache;
  int perSm = 0;
  cudaError_t st = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, TG * PSPLIT, smem);
  if (st != cudaSuccess) return report(st, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
  if (perSm < 1) return cudaErrorInvalidConfiguration;
  const unsigned long long resident = (unsigned long long)perSm * (unsigned)smCount;
  const unsigned grid = (unsigned)(P.totalTiles < resident ? P.totalTiles : resident);
  void* args[] = {(void*)&P};
  return cudaLaunchKernel((const void*)kernel, dim3(grid), dim3(TG * PSPLIT), args, smem, stream);
}

template <int MODE>
static cudaError_t launchPolyMode(int variant, PolyParams& P, size_t smem, int dev, int smCount,
                                  cudaStream_t stream) noexcept {
  switch (variant) {
#define X(id, r, tg, ps, nb, mb) \
  case id: return launchPolyT<MODE, r, tg, ps, nb, mb>(P, smem, dev, smCount, stream);
    GSDR_POLY_VARIANTS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}

// Automatic choice: first variant of the preference list whose shared memory fits.
static int choosePolyVariant(size_t D, size_t T, size_t nOut, int maxSmem, PolyGeom* geom) noexcept {
  const int forced = gForcedVariant.load(std::memory_order_relaxed);
  if (forced == -2) return -1;
  if (forced >= 0 && forced < kNumVariants) {
    return (polyGeometry(kVariants[forced], D, T, geom) && geom->smemBytes <= (size_t)maxSmem) ? forced : -1;
  }
  (void)nOut;
  static const int order[] = {1, 2};
  const size_t budgets[] = {(size_t)44 * 1024, (size_t)(226 * 1024) / 2, (size_t)maxSmem};
  for (size_t limit : budgets) {
    for (int id : order) {
      PolyGeom g;
      if (!polyGeometry(kVariants[id], D, T, &g)) continue;
      if (g.smemBytes <= limit && g.smemBytes <= (size_t)maxSmem) {
        *geom = g;
        return id;
      }
    }
  }
  return -1;
}

template <class IN_T, class OUT_T, class TAP_T>
static cudaError_t launchDirect(const FirCall& c, cudaStream_t stream) noexcept {
  DirectParams P;
  P.x = c.input;
  P.h = c.taps;
  P.y = c.output;
  P.nOut = c.numOutputs;
  P.D = c.decimation;
  P.T = c.tapCount;
  P.xStride = c.inputStride;
  P.yStride = c.outputStride;
  P.hStride = c.tapStride;
  const unsigned long long bpc = (c.numOutputs + kDirectThreads - 1) / kDirectThreads;
  const unsigned long long grid = bpc * c.numChannels;
  if (bpc > 0x7fffffffull || grid > 0x7fffffffull) return cudaErrorInvalidValue;
  P.blocksPerChannel = (unsigned)bpc;
  void* args[] = {(void*)&P};
  return cudaLaunchKernel((const void*)firDirectKernel<IN_T, OUT_T, TAP_T>, dim3((unsigned)grid),
                          dim3(kDirectThreads), args, 0, stream);
}

static size_t outElemBytes(FirType t) noexcept { return t == kFirFF ? 4 : 8; }

cudaError_t enqueueFir(const FirCall& c, cudaStream_t stream) noexcept {
  if (c.numOutputs == 0 || c.numChannels == 0) return cudaSuccess;
  if (c.decimation == 0) return cudaErrorInvalidValue;
  if (c.nco != kNcoNone && c.type != kFirFC) return cudaErrorInvalidValue;
  if (c.tapCount == 0) {
    // the reference writes zero<OUT_T>() when the tap loop does not run (ref: src/fir.cu:67-70)
    for (size_t ch = 0; ch < c.numChannels; ch++) {
      cudaError_t st = cudaMemsetAsync((unsigned char*)c.output + ch * c.outputStride * outElemBytes(c.type), 0,
                                       c.numOutputs * outElemBytes(c.type), stream);
      if (st != cudaSuccess) return st;
    }
    return cudaSuccess;
  }
  int dev = 0;
  cudaError_t st = cudaGetDevice(&dev);
  if (st != cudaSuccess) return st;
  const DeviceInfo* info = deviceInfo(dev);
  if (!info || info->status != cudaSuccess) return info ? info->status : cudaErrorInvalidDevice;

  const bool polyType = (c.type == kFirFC || c.type == kFirFF);
  PolyGeom geom{};
  const int variant = polyType ? choosePolyVariant(c.decimation, c.tapCount, c.numOutputs, info->maxSmemOptin, &geom) : -1;
  if (variant < 0) {
    if (c.nco != kNcoNone) return cudaErrorInvalidValue;  // TODO(round 2): phase-chunked kernel for huge D*T
    switch (c.type) {
      case kFirFC: return launchDirect<float2, float2, float>(c, stream);
      case kFirFF: return launchDirect<float, float, float>(c, stream);
      case kFirCC: return launchDirect<float2, float2, float2>(c, stream);
      case kFirCF: return launchDirect<float, float2, float2>(c, stream);
    }
    return cudaErrorInvalidValue;
  }

  const PolyVariant& v = kVariants[variant];
  const size_t bout = (size_t)v.R * v.tg * (c.type == kFirFF ? 2 : 1);
  const unsigned long long tiles = (c.numOutputs + bout - 1) / bout;
  const unsigned long long total = tiles * c.numChannels;
  if (tiles > 0x7fffffffull || total > 0x7fffffffull) return cudaErrorInvalidValue;

  PolyParams P{};
  P.x = c.input;
  P.h = (const float*)c.taps;
  P.y = c.output;
  P.nOut = c.numOutputs;
  P.nIn = (unsigned long long)(c.numOutputs - 1) * c.decimation + c.tapCount;
  P.xStride = c.inputStride;
  P.yStride = c.outputStride;
  P.hStride = c.tapStride;
  P.tilesPerChannel = (unsigned)tiles;
  P.totalTiles = (unsigned)total;
  P.D = (unsigned)c.decimation;
  P.T = (unsigned)c.tapCount;
  P.Jpad = geom.Jpad;
  P.pitch = geom.pitch;
  P.stageElems = geom.stageElems;
  const unsigned nt = (unsigned)v.threads();
  P.dp = (unsigned)(nt % c.decimation);
  P.dm = (unsigned)(nt / c.decimation);
  P.fastStage = (P.dp == 0 && P.dm % v.R == 0 && P.dm > 0) ? 1u : 0u;
  P.posStep = P.dm + kPolyPad * (P.dm / v.R);
  const size_t oe = outElemBytes(c.type);
  P.y16 = ((uintptr_t)c.output % 16 == 0 && (c.numChannels == 1 || (c.outputStride * oe) % 16 == 0)) ? 1u : 0u;
  P.dbg = (unsigned)gDebugFlags.load(std::memory_order_relaxed);
  P.ncoStep = ncoPhaseStep(c.frequencyShift, c.sampleRate);
  P.ncoFirst = c.firstSampleIndex;
  P.ncoFirst32 = (uint32_t)fmodf((float)c.firstSampleIndex, c.sampleRate);  // ref: src/fm.cu:202
  P.ncoFs = c.sampleRate;
  P.ncoF = c.frequencyShift;

  const int sms = info->smCount;
  if (c.type == kFirFF) return launchPolyMode<kPolyFF>(variant, P, geom.smemBytes, dev, sms, stream);
  switch (c.nco) {
    case kNcoNone: return launchPolyMode<kPolyFC>(variant, P, geom.smemBytes, dev, sms, stream);
    case kNcoExact: return launchPolyMode<kPolyNcoExact>(variant, P, geom.smemBytes, dev, sms, stream);
    case kNcoLiteral: return launchPolyMode<kPolyNcoLiteral>(variant, P, geom.smemBytes, dev, sms, stream);
  }
  return cudaErrorInvalidValue;
}

static cudaError_t firEntry(FirType type, size_t decimation, const void* taps, size_t tapCount, const void* input,
                            void* output, size_t numOutputs, int32_t cudaDevice, cudaStream_t stream) noexcept {
  DeviceScope scope(cudaDevice);
  if (scope.status() != cudaSuccess) return scope.status();
  FirCall c;
  c.type = type;
  c.decimation = decimation;
  c.taps = taps;
  c.tapCount = tapCount;
  c.input = input;
  c.output = output;
  c.numOutputs = numOutputs;
  return enqueueFir(c, stream);
}

static cudaError_t ncoEntry(NcoMode mode, float sampleRate, float frequencyShift, size_t firstSampleIndex,
                            size_t decimation, const float* taps, size_t tapCount, const cuComplex* input,
                            cuComplex* output, size_t numOutputs, int32_t cudaDevice, cudaStream_t stream) noexcept {
  DeviceScope scope(cudaDevice);
  if (scope.status() != cudaSuccess) return scope.status();
  FirCall c;
  c.type = kFirFC;
  c.nco = mode;
  c.decimation = decimation;
  c.taps = taps;
  c.tapCount = tapCount;
  c.input = input;
  c.output = output;
  c.numOutputs = numOutputs;
  c.sampleRate = sampleRate;
  c.frequencyShift = frequencyShift;
  c.firstSampleIndex = firstSampleIndex;
  return enqueueFir(c, stream);
}

}  // namespace gsdr_b200

using namespace gsdr_b200;

// ---- <gsdr/fir.h> ----------------------------------------------------------------------------------------

GSDR_C_LINKAGE cudaError_t gsdrFirFC(size_t decimation, const float* taps, size_t tapCount, const cuComplex* input,
                                     cuComplex* output, size_t numOutputs, int32_t cudaDevice,
                                     cudaStream_t cudaStream) GSDR_NO_EXCEPT {
  return firEntry(kFirFC, decimation, taps, tapCount, input, output, numOutputs, cudaDevice, cudaStream);
}

GSDR_C_LINKAGE cudaError_t gsdrFirFF(size_t decimation, const float* taps, size_t tapCount, const float* input,
                                     float* output, size_t numOutputs, int32_t cudaDevice,
                                     cudaStream_t cudaStream) GSDR_NO_EXCEPT {
  return firEntry(kFirFF, decimation, taps, tapCount, input, output, numOutputs, cudaDevice, cudaStream);
}

GSDR_C_LINKAGE cudaError_t gsdrFirCC(size_t decimation, const cuComplex* taps, size_t tapCount,
                                     const cuComplex* input, cuComplex* output, size_t numOutputs,
                                     int32_t cudaDevice, cudaStream_t cudaStream) GSDR_NO_EXCEPT {
  return firEntry(kFirCC, decimation, taps, tapCount, input, output, numOutputs, cudaDevice, cudaStream);
}

GSDR_C_LINKAGE cudaError_t gsdrFirCF(size_t decimation, const cuComplex* taps, size_t tapCount, const float* input,
                                     cuComplex* output, size_t numOutputs, int32_t cudaDevice,
                                     cudaStream_t cudaStream) GSDR_NO_EXCEPT {
  return firEntry(kFirCF, decimation, taps, tapCount, input, output, numOutputs, cudaDevice, cudaStream);
}

// ---- <gsdr/adjust_frequency.h> ---------------------------------------------------------------------------

GSDR_C_LINKAGE cudaError_t gsdrAdjustFrequencyFirFC(float sampleRate, float frequencyShift, size_t firstSampleIndex,
                                                    size_t decimation, const float* taps, size_t tapCount,
                                                    const cuComplex* input, cuComplex* output, size_t numOutputs,
                                                    int32_t cudaDevice, cudaStream_t cudaStream) GSDR_NO_EXCEPT {
  return ncoEntry(kNcoExact, sampleRate, frequencyShift, firstSampleIndex, decimation, taps, tapCount, input, output,
                  numOutputs, cudaDevice, cudaStream);
}

GSDR_C_LINKAGE cudaError_t gsdrAdjustFrequencyFirFCLiteral(float sampleRate, float frequencyShift,
                                                           size_t firstSampleIndex, size_t decimation,
                                                           const float* taps, size_t tapCount, const cuComplex* input,
                                                           cuComplex* output, size_t numOutputs, int32_t cudaDevice,
                                                           cudaStream_t cudaStream) GSDR_NO_EXCEPT {
  return ncoEntry(kNcoLiteral, sampleRate, frequencyShift, firstSampleIndex, decimation, taps, tapCount, input,
                  output, numOutputs, cudaDevice, cudaStream);
}

GSDR_C_LINKAGE uint64_t gsdrNcoPhaseStep(float frequencyShift, float sampleRate) GSDR_NO_EXCEPT {
  return ncoPhaseStep(frequencyShift, sampleRate);
}

// ---- tuning / introspection hooks of <gsdr/b200.h> -------------------------------------------------------

GSDR_C_LINKAGE int gsdrB200SetKernelVariant(int variant) GSDR_NO_EXCEPT {
  if (variant < -2 || variant >= kNumVariants) return -1;
  gForcedVariant.store(variant, std::memory_order_relaxed);
  return 0;
}

GSDR_C_LINKAGE int gsdrB200NumKernelVariants(void) GSDR_NO_EXCEPT { return kNumVariants; }

GSDR_C_LINKAGE int gsdrB200SetDebugFlags(int flags) GSDR_NO_EXCEPT {
  gDebugFlags.store(flags & 3, std::memory_order_relaxed);
  return 0;
}

GSDR_C_LINKAGE int gsdrB200DescribeKernel(int firType, size_t decimation, size_t tapCount, size_t numOutputs,
                                          int32_t cudaDevice, gsdrB200KernelInfo* info) GSDR_NO_EXCEPT {
  if (!info || decimation == 0) return -1;
  DeviceScope scope(cudaDevice);
  if (scope.status() != cudaSuccess) return -1;
  const DeviceInfo* di = deviceInfo(cudaDevice);
  if (!di || di->status != cudaSuccess) return -1;
  *info = gsdrB200KernelInfo{};
  info->variant = -1;
  info->smCount = di->smCount;
  if (tapCount == 0 || numOutputs == 0) return 0;
  PolyGeom g{};
  const bool polyType = (firType == kFirFC || firType == kFirFF);
  const int v = polyType ? choosePolyVariant(decimation, tapCount, numOutputs, di->maxSmemOptin, &g) : -1;
  info->variant = v;
  if (v >= 0) {
    const size_t bout = (size_t)kVariants[v].R * kVariants[v].tg * (firType == kFirFF ? 2 : 1);
    info->outputsPerThread = kVariants[v].R;
    info->threadsPerBlock = kVariants[v].threads();
    info->phaseGroups = kVariants[v].psplit;
    info->windowBuffers = kVariants[v].nbuf;
    info->outputsPerBlock = bout;
    info->sharedBytesPerBlock = g.smemBytes;
    info->numBlocks = (numOutputs + bout - 1) / bout;
  } else {
    info->outputsPerThread = 1;
    info->threadsPerBlock = kDirectThreads;
    info->outputsPerBlock = kDirectThreads;
    info->sharedBytesPerBlock = kDirectTapChunk * (firType == kFirFC || firType == kFirFF ? 4 : 8);
    info->numBlocks = (numOutputs + kDirectThreads - 1) / kDirectThreads;
  }
  return 0;
}
