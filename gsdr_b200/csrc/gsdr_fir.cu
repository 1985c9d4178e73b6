// gsdr_fir.cu — C-ABI entry points of <gsdr/fir.h> and <gsdr/adjust_frequency.h> and the launch logic behind
// them.  Replaces the host wrappers at ref: src/fir.cu:73-171 (and the per-thread k_AdjustFrequency call at
// ref: src/fm.cu:46-56).  No CPU fallback exists: every path ends in a kernel launch on the caller's stream.
#include <gsdr/adjust_frequency.h>
#include <gsdr/b200.h>
#include <gsdr/conversion.h>
#include <gsdr/fir.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "fir_launch.cuh"
#include "fir_tc_kernel.cuh"

namespace gsdr_b200 {

// ---------------------------------------------------------------------------------------------------------
// device scope + error reporting (one stderr line on failure, ref: include/gsdr/cuda_util.h:59-82)
// ---------------------------------------------------------------------------------------------------------
cudaError_t report(cudaError_t st, const char* what) noexcept {
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

constexpr int kForceFusedChannelizer = -5;  // tuning hook value: gsdrChannelizeFC through firTmaChannelizerKernel
constexpr int kForceTensorCore = -4;  // tuning hook value: tensor-core kernel wherever its shape rules allow
constexpr int kForceNoTensorCore = -3;  // tuning hook value: automatic choice among the FFMA2 kernels only
// launchTma's answer when the driver rejects the tensor map: the caller goes on to the kernels that need none
constexpr cudaError_t kTmaEncodeFailed = (cudaError_t)0x7f0000e1;
// Tuning build only (-DGSDR_B200_TUNING): process-wide variant override and work-skipping measurement flags.  The
// release library has neither the state nor the setters; forcedVariant() / debugFlags() fold to constants there.
#ifdef GSDR_B200_TUNING
static std::atomic<int> gDebugFlags{0};
static std::atomic<int> gForcedVariant{-1};  // -1 auto, -2 direct kernel, >=0 variant id
static inline int forcedVariant() noexcept { return gForcedVariant.load(std::memory_order_relaxed); }
static inline unsigned debugFlags() noexcept { return (unsigned)gDebugFlags.load(std::memory_order_relaxed); }
#else
static inline int forcedVariant() noexcept { return -1; }
static inline unsigned debugFlags() noexcept { return 0u; }
#endif

// gsdrB200SetFirTensorCores: 1 = the tensor-core kernel where it was measured faster (default), 0 = FFMA2 kernels only
static std::atomic<int> gTensorCores{1};

// the override as the FFMA2-kernel choosers see it: the two tensor-core values mean "automatic" to them
static inline int forcedFfmaVariant() noexcept {
  const int f = forcedVariant();
  return (f == kForceTensorCore || f == kForceNoTensorCore || f == kForceFusedChannelizer) ? -1 : f;
}

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
  static std::atomic<size_t> configured[64];  // per device: dynamic shared memory the kernel has been opted in to
  auto kernel = firPolyKernel<MODE, R, TG, PSPLIT, NBUF, MINB>;
  if (smem > configured[dev & 63].load(std::memory_order_acquire)) {
    cudaError_t st = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (st != cudaSuccess) return report(st, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    configured[dev & 63].store(smem, std::memory_order_release);
  }
  int perSm = 0;
  static OccCache occ;
  cudaError_t st = occupancyCached(occ, kernel, TG * PSPLIT, smem, dev, &perSm);
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
static int choosePolyVariant(size_t D, size_t T, size_t nOut, int maxSmem, PolyGeom* geom,
                             bool doubleBuffered = false) noexcept {
  const int forced = forcedFfmaVariant();
  if (forced == -2) return -1;
  if (forced >= 0 && forced < kNumVariants) {
    return (polyGeometry(kVariants[forced], D, T, geom) && geom->smemBytes <= (size_t)maxSmem) ? forced : -1;
  }
  if (forced >= kNumVariants) return -1;  // a TMA variant was requested and did not qualify: direct kernel
  (void)nOut;
  static const int orderSingle[] = {1, 2};
  static const int orderDouble[] = {6, 1};  // real input, decimation 1: 64 threads, two window buffers
  const int* order = doubleBuffered ? orderDouble : orderSingle;
  const size_t budgets[] = {(size_t)44 * 1024, (size_t)(226 * 1024) / 2, (size_t)maxSmem};
  for (size_t limit : budgets) {
    for (int k = 0; k < 2; k++) {
      const int id = order[k];
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

// ---------------------------------------------------------------------------------------------------------
// TMA-fed fast path (fir_tma_kernel.cuh)
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encodeTiled() noexcept {
  static std::once_flag once;
  static EncodeTiledFn fn = nullptr;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = (EncodeTiledFn)p;
    } else {
      (void)cudaGetLastError();
    }
  });
  return fn;
}

struct TmaGeom {
  bool staticD;
  unsigned Jpad, mhp, planeBytes, swzShift, swzMask;
  CUtensorMapSwizzle swizzle;
  size_t smemBytes;
};

static bool tmaSupportedDecimation(size_t D) noexcept {
  // rows of 16..128 bytes whose bank pattern is conflict free (odd chunk count, or a TMA swizzle mode exists),
  // and wider rows that split into 128-byte segments
  return D == 2 || D == 4 || D == 6 || D == 8 || D == 10 || D == 14 || (D >= 16 && D % 16 == 0 && D <= 256);
}

// Decimations with a compile-time-geometry instantiation (the BASELINE shapes).
static bool tmaStaticDecimation(size_t D) noexcept { return D == 4 || D == 8 || D == 10 || D == 32; }

static bool tmaGeometry(const TmaVariant& v, size_t D, size_t T, TmaGeom* g) noexcept {
  if (!tmaSupportedDecimation(D) || (D / 2) < (size_t)v.psplit) return false;
  const size_t J = (T + D - 1) / D;
  const size_t Jpad = (J + 15) / 16 * 16;
  const size_t G = tmaSegBytes((unsigned)D);
  g->staticD = tmaStaticDecimation(D) && Jpad <= kTmaJpadCap;
  g->swizzle = CU_TENSOR_MAP_SWIZZLE_NONE;
  g->swzShift = 0;
  g->swzMask = 0;
  if (G == 32) g->swizzle = CU_TENSOR_MAP_SWIZZLE_32B, g->swzShift = 2, g->swzMask = 1;
  if (G == 64) g->swizzle = CU_TENSOR_MAP_SWIZZLE_64B, g->swzShift = 1, g->swzMask = 3;
  if (G == 128) g->swizzle = CU_TENSOR_MAP_SWIZZLE_128B, g->swzShift = 0, g->swzMask = 7;
  // must agree with the kernel's compile-time tmaPlaneRows(TG, kTmaJpadCap, D) when staticD
  const size_t mhp = tmaPlaneRows((unsigned)v.tg, (unsigned)(g->staticD ? kTmaJpadCap : Jpad), (unsigned)D);
  if (mhp > 256) return false;  // TMA box dimension limit
  g->Jpad = (unsigned)Jpad;
  g->mhp = (unsigned)mhp;
  g->planeBytes = (unsigned)(mhp * G);
  g->smemBytes = 1024 + (size_t)v.nbuf * (8 * D / G) * 8 * (size_t)g->planeBytes +
                 2 * (size_t)(v.psplit - 1) * v.tg * 64 + (D * Jpad + 32) * 4 +
                 // NCO: row anchors, then rotation table + fine / coarse anchor tables
                 ((size_t)kTmaR * v.tg + Jpad + ncoTableFloat2s((unsigned)D, (unsigned)(kTmaR * v.tg + Jpad))) * 8;
  return true;
}

static cudaError_t launchSpec(int variant, bool staticD, const CUtensorMap& map, TmaParams& P, size_t smem, int dev,
                              int smCount, cudaStream_t stream) noexcept {
  if (staticD) {
    switch (P.D) {
      case 8: return launchSpecDt8(variant, map, P, smem, dev, smCount, stream);
      case 10: return launchSpecDt10(variant, map, P, smem, dev, smCount, stream);
      case 32: return launchSpecDt32(variant, map, P, smem, dev, smCount, stream);
      default: break;
    }
  }
  return launchSpecDt0(variant, map, P, smem, dev, smCount, stream);
}

template <int MODE>
static cudaError_t launchTmaMode(int variant, bool staticD, const CUtensorMap& map, TmaParams& P, size_t smem,
                                 int dev, int smCount, cudaStream_t stream) noexcept {
  if (staticD) {
    switch (P.D) {
      case 4: return launchTmaDt4(MODE, variant, map, P, smem, dev, smCount, stream);
      case 8: return launchTmaDt8(MODE, variant, map, P, smem, dev, smCount, stream);
      case 10: return launchTmaDt10(MODE, variant, map, P, smem, dev, smCount, stream);
      case 32: return launchTmaDt32(MODE, variant, map, P, smem, dev, smCount, stream);
      default: break;
    }
  }
  return launchTmaDt0(MODE, variant, map, P, smem, dev, smCount, stream);
}

// TMA-kernel variant ids: [0, kNumTmaVariants) = firTmaKernel, then the warp-specialised fused-NCO kernel.
static bool tmaVariantShape(int id, TmaVariant* v, int* mixw) noexcept {
  if (id >= 0 && id < kNumTmaVariants) {
    *v = kTmaVariants[id];
    *mixw = 0;
    return true;
  }
  const int sid = id - kNumTmaVariants;
  if (sid >= 0 && sid < kNumSpecVariants) {
    *v = TmaVariant{kSpecVariants[sid].tg, kSpecVariants[sid].psplit, 2, kSpecVariants[sid].minBlocks};
    *mixw = kSpecVariants[sid].mixw;
    return true;
  }
  return false;
}

static bool tmaVariantFits(int id, const FirCall& c, int maxSmem, TmaGeom* geom) noexcept {
  TmaVariant v;
  int mixw = 0;
  if (!tmaVariantShape(id, &v, &mixw)) return false;
  if (mixw > 0) {
    // the specialised kernel: exact NCO, one tap set for every channel
    if (c.nco != kNcoExact) return false;
    if (c.numChannels > 1 && c.tapStride != 0) return false;
  }
  return tmaGeometry(v, c.decimation, c.tapCount, geom) && geom->smemBytes <= (size_t)maxSmem;
}

// Returns the TMA variant to use for this call, or -1 when the call does not qualify.
static int chooseTmaVariant(const FirCall& c, int maxSmem, TmaGeom* geom) noexcept {
  if (c.type != kFirFC) return -1;
  if (c.numOutputs >= 0xfff00000ull) return -1;  // the kernel compares output-row indices in 32 bits
  if (!tmaSupportedDecimation(c.decimation) || !encodeTiled()) return -1;
  if ((uintptr_t)c.input % 16 != 0) return -1;                       // TMA needs a 16-byte aligned base
  if (c.numChannels > 1 && (c.inputStride % 2) != 0) return -1;       // ... and 16-byte strides
  if (c.numChannels > 0x7fffffffull) return -1;
  const int forced = forcedFfmaVariant();
  if (forced >= kNumVariants) {
    const int id = forced - kNumVariants;
    return tmaVariantFits(id, c, maxSmem, geom) ? id : -1;
  }
  if (forced != -1) return -1;
  // an odd number of branch pairs (D = 6, 10, 14) cannot be split evenly over two filter groups: the single-group,
  // single-buffered kernel then matches (D = 10: 0.262 vs 0.263 ms) or beats (D = 6: 0.411 vs 0.478 ms, 255 taps,
  // 2^26 samples) the warp-specialised one
  const bool oddPairCount = c.decimation <= 16 && ((c.decimation / 2) % 2 == 1);
  // (D = 4: two branch pairs cannot keep two filter groups busy: 64 x 1 single-buffered 0.268 ms against 0.410 ms for
  //  the warp-specialised kernel, 127 taps, 2^26 samples: profiles/r02/sweep_nco_d4.jsonl)
  if (c.nco == kNcoExact && !oddPairCount && c.decimation >= 8 && c.epilogue != kFirEpiFmDemod) {
    // fused NCO: copy + mix on dedicated warps, overlapped with the FIR of the previous tile
    // (tools/sweep.py --nco: 64 x 2 filter threads + 4 mixer warps for narrow rows, 32 x 4 + 4 for wide rows)
    const int sid = kNumTmaVariants + (c.decimation > 16 ? 1 : 2);
    if (tmaVariantFits(sid, c, maxSmem, geom)) return sid;
  }
  // hand-tuned preference (tools/sweep.py): narrow rows -> 32 outputs-threads x 2 branch groups, double buffered;
  // wide rows (many branch pairs, big windows) -> 4 or 8 branch groups on a single buffer, 2 CTAs per SM
  // With the fused NCO every CTA alternates a mix pass (latency bound) and the FIR: more, single-buffered CTAs per
  // SM overlap one CTA's mix with another's FIR better than double buffering does.
  // ids: 1 = (32 threads x 2 groups, double buffered), 8 = (32 x 1, double buffered), 5 / 10 = single-buffered
  // 2 / 4 groups, 4 = (32 x 4, double buffered), 11 = (32 x 8, single buffered)
  static const int orderEvenPairs[] = {1, 8, 5, 0};   // D = 8, 16: the branch pairs split evenly over 2 groups
  static const int orderOddPairs[] = {8, 1, 5, 0};    // D = 2, 4, 6, 10, 14: one group keeps all pairs
  static const int orderWide[] = {10, 4, 11, 1};       // D >= 16 (1023 taps, D = 32, 2^27: 0.391 / 0.401 ms; 511 taps,
                                                       // D = 16, 2^26: 0.205 / 0.206 ms, id 1: 0.219 ms)
  static const int orderNarrowNcoEven[] = {6, 9, 5, 1};  // 6 = (64 x 2, single buffer), 9 = (64 x 1, single buffer)
  static const int orderNarrowNcoOdd[] = {9, 6, 5, 1};
  static const int orderWideNco[] = {10, 11, 4, 1};
  const bool wide = c.decimation >= 16;
  const bool nco = c.nco != kNcoNone;
  const bool evenPairs = ((c.decimation / 2) % 2 == 0) && c.decimation >= 8;
  const int* order = wide ? (nco ? orderWideNco : orderWide)
                          : (nco ? (evenPairs ? orderNarrowNcoEven : orderNarrowNcoOdd)
                                 : (evenPairs ? orderEvenPairs : orderOddPairs));
  const int orderLen = 4;
  for (int k = 0; k < orderLen; k++) {
    const int id = order[k];
    TmaGeom g;
    if (tmaVariantFits(id, c, maxSmem, &g)) {
      *geom = g;
      return id;
    }
  }
  return -1;
}

// ---- wide rows (D = 32), segment-pipelined: firTmaWideKernel ----
static int firstWideVariantId() noexcept;

static bool wideGeometry(const SpecVariant& v, size_t D, size_t T, bool nco, TmaGeom* g) noexcept {
  if (D != 32) return false;  // the compile-time instantiations
  TmaVariant shape{v.tg, v.psplit, 2, v.minBlocks};
  if (!tmaGeometry(shape, D, T, g) || !g->staticD) return false;
  // stage buffers (three with the NCO, two without; the partial-sum exchange reuses one), taps, row anchors + table
  g->smemBytes = 1024 + (nco ? 3 : 2) * 8 * (size_t)g->planeBytes + (D * g->Jpad + 32) * 4 +
                 ((size_t)kTmaR * v.tg + g->Jpad + ncoTableFloat2s((unsigned)D, (unsigned)(kTmaR * v.tg + g->Jpad))) * 8;
  return true;
}

// Returns the wide-row variant for this call, or -1 when the call does not qualify.
static int chooseWideVariant(const FirCall& c, int maxSmem, TmaGeom* geom) noexcept {
  if (c.type != kFirFC || c.nco == kNcoLiteral || c.epilogue == kFirEpiFmDemod) return -1;
  if (c.numOutputs >= 0xfff00000ull || c.decimation != 32 || !encodeTiled()) return -1;
  if ((uintptr_t)c.input % 16 != 0) return -1;
  if (c.numChannels > 1 && ((c.inputStride % 2) != 0 || c.tapStride != 0)) return -1;
  if (c.numChannels > 0x7fffffffull) return -1;
  const bool nco = c.nco == kNcoExact;
  auto fits = [&](int id, TmaGeom* g) {
    return id >= 0 && id < kNumWideVariants && wideGeometry(kWideVariants[id], c.decimation, c.tapCount, nco, g) &&
           g->smemBytes <= (size_t)maxSmem;
  };
  const int forced = forcedFfmaVariant();
  if (forced >= firstWideVariantId()) return fits(forced - firstWideVariantId(), geom) ? forced - firstWideVariantId() : -1;
  if (forced != -1) return -1;
  // 512-output tiles on one CTA per SM: worth it once every SM gets a few tiles
  if ((c.numOutputs + 511) / 512 * c.numChannels < 296) return -1;
  const int id = 1;  // 64 x 4 filter threads + 4 producer warps for both (tools/sweep.py: plain FIR 0.649 vs 0.665 ms with 1)
  return fits(id, geom) ? id : -1;
}

// ---- complex taps (gsdrFirCC): two tap planes on the same windows ----
static bool ccGeometry(const TmaVariant& v, size_t D, size_t T, TmaGeom* g) noexcept {
  TmaVariant half = v;
  half.psplit = v.psplit / 2;  // groups per tap plane: they split the branch pairs
  if (!tmaGeometry(half, D, T, g)) return false;
  if (g->staticD && D != 8) {
    // only D = 8 has a compile-time-geometry instantiation of this kernel: recompute the plane pitch for run time
    g->staticD = false;
    const size_t G = tmaSegBytes((unsigned)D);
    const size_t mhp = tmaPlaneRows((unsigned)v.tg, g->Jpad, (unsigned)D);
    if (mhp > 256) return false;
    g->mhp = (unsigned)mhp;
    g->planeBytes = (unsigned)(mhp * G);
  }
  g->smemBytes = 1024 + (size_t)v.nbuf * (8 * D / tmaSegBytes((unsigned)D)) * 8 * (size_t)g->planeBytes +
                 2 * (size_t)(v.psplit - 1) * v.tg * 64 + 2 * (D * g->Jpad + 32) * 4;
  return true;
}

static int firstCcVariantId() noexcept { return kNumVariants + kNumTmaVariants + kNumSpecVariants + kNumRealVariants; }

// Returns the complex-tap variant for this call, or -1 when the call does not qualify.
static int chooseCcVariant(const FirCall& c, int maxSmem, TmaGeom* geom) noexcept {
  if (c.type != kFirCC || c.nco != kNcoNone) return -1;
  if (c.numOutputs >= 0xfff00000ull) return -1;  // the kernel compares output-row indices in 32 bits
  if (!tmaSupportedDecimation(c.decimation) || !encodeTiled()) return -1;
  if ((uintptr_t)c.input % 16 != 0) return -1;
  if (c.numChannels > 1 && (c.inputStride % 2) != 0) return -1;
  if (c.numChannels > 0x7fffffffull) return -1;
  auto fits = [&](int id, TmaGeom* g) {
    return id >= 0 && id < kNumCcVariants && ccGeometry(kCcVariants[id], c.decimation, c.tapCount, g) &&
           g->smemBytes <= (size_t)maxSmem;
  };
  const int forced = forcedFfmaVariant();
  if (forced >= firstCcVariantId()) return fits(forced - firstCcVariantId(), geom) ? forced - firstCcVariantId() : -1;
  if (forced != -1) return -1;
  static const int orderNarrow[] = {0, 1, 3, 2};
  static const int orderWide[] = {1, 3, 0, 2};
  const int* order = c.decimation > 16 ? orderWide : orderNarrow;
  for (int k = 0; k < 4; k++) {
    TmaGeom g;
    if (fits(order[k], &g)) {
      *geom = g;
      return order[k];
    }
  }
  return -1;
}

// variant: a TMA / fused-NCO variant id, or (ccVariant >= 0) a complex-tap variant
// one input, several frequency shifts in one launch (firTmaChannelizerKernel)
struct ChanExtra {
  int tg, psplit;
  unsigned numShifts;
  unsigned long long yShiftStride;
  const unsigned long long* steps;
};

static cudaError_t launchTma(const FirCall& c, int variant, const TmaGeom& geom, int dev, int smCount,
                             cudaStream_t stream, int ccVariant = -1, int wideVariant = -1,
                             const ChanExtra* chan = nullptr) noexcept {
  TmaVariant v;
  int mixw = 0;
  if (chan) {
    v = TmaVariant{chan->tg, chan->psplit, 2, 1};
  } else if (wideVariant >= 0) {
    if (wideVariant >= kNumWideVariants) return cudaErrorInvalidValue;
    v = TmaVariant{kWideVariants[wideVariant].tg, kWideVariants[wideVariant].psplit, 2, 1};
  } else if (ccVariant >= 0) {
    if (ccVariant >= kNumCcVariants) return cudaErrorInvalidValue;
    v = kCcVariants[ccVariant];
  } else if (!tmaVariantShape(variant, &v, &mixw)) {
    return cudaErrorInvalidValue;
  }
  const size_t bout = (size_t)kTmaR * v.tg;
  const bool fm = c.epilogue == kFirEpiFmDemod;
  if (fm && (wideVariant >= 0 || ccVariant >= 0 || mixw > 0)) return cudaErrorNotSupported;
  // FM: numOutputs phase steps need numOutputs + 1 low-pass values; tiles overlap by their last row group
  const size_t firOutputs = c.numOutputs + (fm ? 1 : 0);
  const size_t tileOut = fm ? bout - kTmaR : bout;
  const unsigned long long tiles = (c.numOutputs + tileOut - 1) / tileOut;
  const unsigned long long total = tiles * c.numChannels;
  if (tiles > 0x7fffffffull || total > 0x7fffffffull) return cudaErrorInvalidValue;
  const size_t D = c.decimation;
  const unsigned long long nIn = (unsigned long long)(firOutputs - 1) * D + c.tapCount;
  const unsigned long long tmaRows = (nIn / (8 * D)) * 8;  // complete groups of 8 rows only: TMA never reads past nIn
  if (tmaRows / 8 > 0xffffffffull) return cudaErrorInvalidValue;

  TmaParams P{};
  P.x = (const float2*)c.input;
  P.h = (const float*)c.taps;
  P.y = (float2*)c.output;
  P.nOut = firOutputs;
  P.nIn = nIn;
  P.epi = (unsigned)c.epilogue;
  P.epiGain = c.epilogueGain;
  P.tileOut = (unsigned)tileOut;
  P.xStride = c.inputStride;
  P.yStride = c.outputStride;
  P.hStride = c.tapStride;
  P.tilesPerChannel = (unsigned)tiles;
  P.totalTiles = (unsigned)total;
  P.numChannels = (unsigned)c.numChannels;
  P.D = (unsigned)D;
  P.T = (unsigned)c.tapCount;
  P.Jpad = geom.Jpad;
  P.rowBytes = (unsigned)(8 * D);
  P.segBytes = tmaSegBytes((unsigned)D);
  P.mhp = geom.mhp;
  P.planeBytes = geom.planeBytes;
  P.swzShift = geom.swzShift;
  P.swzMask = geom.swzMask;
  P.tmaRows = (unsigned)(tmaRows > 0xffffffffull ? 0xffffffffull : tmaRows);
  {
    const size_t oe = c.epilogue == kFirEpiNone ? 8 : 4;  // real outputs with a fused stage
    P.y16 = ((uintptr_t)c.output % 16 == 0 && (c.numChannels == 1 || (c.outputStride * oe) % 16 == 0)) ? 1u : 0u;
  }
  P.dbg = debugFlags();
  P.ncoStep = ncoPhaseStep(c.frequencyShift, c.sampleRate);
  P.ncoFirst = c.firstSampleIndex;
  P.ncoFirst32 = (uint32_t)fmodf((float)c.firstSampleIndex, c.sampleRate);  // ref: src/fm.cu:202
  P.ncoFs = c.sampleRate;
  P.ncoF = c.frequencyShift;
  P.ncoRow0 = (unsigned long long)c.firstSampleIndex / D;
  P.ncoRho = (unsigned)((unsigned long long)c.firstSampleIndex % D);

  alignas(64) CUtensorMap map;
  {
    // dims: (floats of one row | groups of 8 rows | row within the group | channel); when fewer than 8 rows are
    // visible every tile takes the cp.async path and the map is never dereferenced (but must still encode).
    const cuuint64_t gdim[4] = {(cuuint64_t)(2 * D), (cuuint64_t)(tmaRows >= 8 ? tmaRows / 8 : 1), 8,
                                (cuuint64_t)c.numChannels};
    const cuuint64_t gstride[3] = {(cuuint64_t)(8 * D * 8), (cuuint64_t)(D * 8),
                                   (cuuint64_t)(c.numChannels > 1 ? c.inputStride * 8 : 8 * D * 8)};
    const cuuint32_t box[4] = {(cuuint32_t)(tmaSegBytes((unsigned)D) / 4), (cuuint32_t)geom.mhp, 8, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = encodeTiled()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)c.input, gdim, gstride, box,
                                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, geom.swizzle,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return kTmaEncodeFailed;  // e.g. a channel stride beyond the tensor map's 2^40 bytes
    if (tmaRows < 8) P.tmaRows = 0;
  }
#ifdef GSDR_B200_TUNING
  if (chan) {
    ChanParams CP{};
    static_cast<TmaParams&>(CP) = P;
    CP.numShifts = chan->numShifts;
    CP.yShiftStride = chan->yShiftStride;
    if (chan->numShifts > 1 && (chan->yShiftStride * 8) % 16 != 0) CP.y16 = 0;  // odd stride: 8-byte stores
    for (unsigned k = 0; k < chan->numShifts; k++) CP.steps[k] = chan->steps[k];
    return launchChan((int)D, map, CP, geom.smemBytes, dev, smCount, stream);
  }
#endif
  if (wideVariant >= 0) {
    return launchWideDt32(c.nco == kNcoExact ? kPolyNcoExact : kPolyFC, wideVariant, map, P, geom.smemBytes, dev, smCount,
                          stream);
  }
  if (ccVariant >= 0) {
    if (geom.staticD && D == 8) return launchCcDt8(ccVariant, map, P, geom.smemBytes, dev, smCount, stream);
    return launchCcDt0(ccVariant, map, P, geom.smemBytes, dev, smCount, stream);
  }
  if (mixw > 0) {
    return launchSpec(variant - kNumTmaVariants, geom.staticD, map, P, geom.smemBytes, dev, smCount, stream);
  }
  switch (c.nco) {
    case kNcoNone:
      return launchTmaMode<kPolyFC>(variant, geom.staticD, map, P, geom.smemBytes, dev, smCount, stream);
    case kNcoExact:
      return launchTmaMode<kPolyNcoExact>(variant, geom.staticD, map, P, geom.smemBytes, dev, smCount, stream);
    case kNcoLiteral:
      return launchTmaMode<kPolyNcoLiteral>(variant, geom.staticD, map, P, geom.smemBytes, dev, smCount, stream);
  }
  return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------------------------
// Real input (gsdrFirFF) on the TMA kernel's inner loop: firTmaRealKernel
// ---------------------------------------------------------------------------------------------------------
struct RealGeom {
  bool staticD;
  unsigned Jpad, mhp, planeBytes, swzShift, swzMask, rawFloats;
  size_t smemBytes;
};

static bool realStaticDecimation(size_t D2) noexcept { return D2 == 2 || D2 == 10; }

// nplane: 1 = real taps (FF), 2 = complex taps (CF: the groups split into two tap planes)
static bool realGeometry(const RealVariant& v, size_t Dreal, size_t T, RealGeom* g, int nplane = 1) noexcept {
  const size_t D = 2 * Dreal;  // decimation of the output-pair stream
  if (!tmaSupportedDecimation(D) || Dreal < (size_t)(v.psplit / nplane) || v.psplit % nplane != 0) return false;
  const size_t J = (T + D - 1) / D;
  const size_t Jpad = J <= 8 ? 8 : (J + 15) / 16 * 16;
  const size_t G = tmaSegBytes((unsigned)D);
  g->staticD = realStaticDecimation(D) && Jpad <= kTmaJpadCap;
  g->swzShift = 0;
  g->swzMask = 0;
  if (G == 32) g->swzShift = 2, g->swzMask = 1;
  if (G == 64) g->swzShift = 1, g->swzMask = 3;
  if (G == 128) g->swzShift = 0, g->swzMask = 7;
  const size_t mhp = tmaPlaneRows((unsigned)v.tg, (unsigned)(g->staticD ? kTmaJpadCap : Jpad), (unsigned)D);
  const size_t rows = (size_t)kTmaR * v.tg + Jpad;
  const size_t rawFloats = (rows * D + Dreal + 3) / 4 * 4;
  if (rawFloats * 4 > 0xfffff0u) return false;  // mbarrier transaction-count range
  g->Jpad = (unsigned)Jpad;
  g->mhp = (unsigned)mhp;
  g->planeBytes = (unsigned)(mhp * G + kRealPlanePad);
  g->rawFloats = (unsigned)rawFloats;
  g->smemBytes = (size_t)v.nwin * (8 * D / G) * 8 * (size_t)g->planeBytes + 2 * (size_t)(v.psplit - 1) * v.tg * 64 +
                 (size_t)nplane * (D * Jpad + 32) * 4 + (size_t)v.nraw * rawFloats * 4;
  return true;
}

static bool realVariantFits(int id, const FirCall& c, int maxSmem, RealGeom* geom) noexcept {
  if (id < 0 || id >= kNumRealVariants) return false;
  return realGeometry(kRealVariants[id], c.decimation, c.tapCount, geom) && geom->smemBytes <= (size_t)maxSmem;
}

// Returns the real-input variant for this call, or -1 when the call does not qualify.
static int chooseRealVariant(const FirCall& c, int maxSmem, RealGeom* geom) noexcept {
  if (c.type != kFirFF || c.nco != kNcoNone) return -1;
  if (!tmaSupportedDecimation(2 * c.decimation)) return -1;
  if ((uintptr_t)c.input % 16 != 0) return -1;                    // bulk copies need a 16-byte aligned source
  if (c.numChannels > 1 && (c.inputStride % 4) != 0) return -1;   // ... in every channel
  if (c.numChannels > 0x7fffffffull) return -1;
  const int firstId = kNumVariants + kNumTmaVariants + kNumSpecVariants;
  const int forced = forcedFfmaVariant();
  if (forced >= firstId) return realVariantFits(forced - firstId, c, maxSmem, geom) ? forced - firstId : -1;
  if (forced != -1) return -1;
  // decimation 1 (one branch pair, half the FFMA2s per tile of the complex kernel) is faster on the cp.async kernel:
  // 0.209 ms against 0.219 ms for 2^26 samples x 63 taps (tools/sweep.py --kind ff)
  if (c.decimation == 1) return -1;
  // one branch pair (decimation 1): four filter warps + four producer warps; more pairs: smaller tiles, more CTAs
  // (tools/sweep.py --kind ff: D = 1 -> id 0 at 0.209 ms for 2^26 samples x 63 taps, D = 5 -> id 7 at 0.217 ms for 2^27)
  static const int orderOnePair[] = {0, 1, 2, 3};
  static const int orderMore[] = {7, 5, 4, 1};
  const int* order = c.decimation == 1 ? orderOnePair : orderMore;
  for (int k = 0; k < 4; k++) {
    RealGeom g;
    if (realVariantFits(order[k], c, maxSmem, &g)) {
      *geom = g;
      return order[k];
    }
  }
  return -1;
}

static int firstCfVariantId() noexcept { return firstCcVariantId() + kNumCcVariants; }
static int firstWideVariantId() noexcept { return firstCfVariantId() + kNumCfVariants; }

// Returns the real-input x complex-taps variant for this call, or -1 when the call does not qualify.
static int chooseCfVariant(const FirCall& c, int maxSmem, RealGeom* geom) noexcept {
  if (c.type != kFirCF || c.nco != kNcoNone) return -1;
  if (!tmaSupportedDecimation(2 * c.decimation)) return -1;
  if ((uintptr_t)c.input % 16 != 0) return -1;
  if (c.numChannels > 1 && (c.inputStride % 4) != 0) return -1;
  if (c.numChannels > 0x7fffffffull) return -1;
  auto fits = [&](int id, RealGeom* g) {
    return id >= 0 && id < kNumCfVariants && realGeometry(kCfVariants[id], c.decimation, c.tapCount, g, 2) &&
           g->smemBytes <= (size_t)maxSmem;
  };
  const int forced = forcedFfmaVariant();
  if (forced >= firstCfVariantId()) return fits(forced - firstCfVariantId(), geom) ? forced - firstCfVariantId() : -1;
  if (forced != -1) return -1;
  // tools/sweep.py --kind cf: 63 complex taps, D = 5, 2^27 samples: id 2 0.321 ms, id 3 0.349, id 0 0.353, id 1 0.381
  // (direct kernel 0.484, reference 0.691); D = 1, 2^26 samples: id 1 0.486 ms, id 2 0.504, id 0 0.524 (direct 1.08)
  static const int orderOnePair[] = {1, 2, 0, 3};
  static const int orderMore[] = {2, 3, 0, 1};
  const int* order = c.decimation == 1 ? orderOnePair : orderMore;
  for (int k = 0; k < 4; k++) {
    RealGeom g;
    if (fits(order[k], &g)) {
      *geom = g;
      return order[k];
    }
  }
  return -1;
}

// cfVariant >= 0: complex taps (gsdrFirCF) — `variant` is then ignored
static cudaError_t launchReal(const FirCall& c, int variant, const RealGeom& geom, int dev, int smCount,
                              cudaStream_t stream, int cfVariant = -1) noexcept {
  const RealVariant& v = cfVariant >= 0 ? kCfVariants[cfVariant] : kRealVariants[variant];
  const size_t bout = (size_t)kTmaR * v.tg;                      // output pairs per tile
  const unsigned long long pairs = (c.numOutputs + 1) / 2;
  const unsigned long long tiles = (pairs + bout - 1) / bout;
  const unsigned long long total = tiles * c.numChannels;
  if (tiles > 0x7fffffffull || total > 0x7fffffffull) return cudaErrorInvalidValue;
  const size_t D = 2 * c.decimation;
  RealParams P{};
  P.x = (const float2*)c.input;
  P.h = (const float*)c.taps;
  P.y = (float2*)c.output;
  P.nOut = pairs;
  P.nOutReal = c.numOutputs;
  P.nIn = (unsigned long long)(c.numOutputs - 1) * c.decimation + c.tapCount;  // floats
  P.xStride = c.inputStride;
  P.yStride = c.outputStride;
  P.hStride = c.tapStride;
  P.tilesPerChannel = (unsigned)tiles;
  P.totalTiles = (unsigned)total;
  P.numChannels = (unsigned)c.numChannels;
  P.D = (unsigned)D;
  P.T = (unsigned)c.tapCount;
  P.Jpad = geom.Jpad;
  P.rowBytes = (unsigned)(8 * D);
  P.segBytes = tmaSegBytes((unsigned)D);
  P.mhp = geom.mhp;
  P.planeBytes = geom.planeBytes;
  P.swzShift = geom.swzShift;
  P.swzMask = geom.swzMask;
  P.rawFloats = geom.rawFloats;
  P.y16 = ((uintptr_t)c.output % 16 == 0 && (c.numChannels == 1 || (c.outputStride * 4) % 16 == 0)) ? 1u : 0u;
  P.dbg = debugFlags();
  if (cfVariant >= 0) {
    P.y16 = ((uintptr_t)c.output % 16 == 0 && (c.numChannels == 1 || (c.outputStride * 8) % 16 == 0)) ? 1u : 0u;
    if (geom.staticD && D == 2) return launchCfDt2(cfVariant, P, geom.smemBytes, dev, smCount, stream);
    if (geom.staticD && D == 10) return launchCfDt10(cfVariant, P, geom.smemBytes, dev, smCount, stream);
    return launchCfDt0(cfVariant, P, geom.smemBytes, dev, smCount, stream);
  }
  if (geom.staticD) {
    switch (D) {
      case 2: return launchRealDt2(variant, P, geom.smemBytes, dev, smCount, stream);
      case 10: return launchRealDt10(variant, P, geom.smemBytes, dev, smCount, stream);
      default: break;
    }
  }
  return launchRealDt0(variant, P, geom.smemBytes, dev, smCount, stream);
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

static cudaError_t launchDirectNco(const FirCall& c, cudaStream_t stream) noexcept {
  DirectNcoParams P{};
  P.x = (const float2*)c.input;
  P.h = (const float*)c.taps;
  P.y = (float2*)c.output;
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
  P.ncoStep = ncoPhaseStep(c.frequencyShift, c.sampleRate);
  P.ncoFirst = c.firstSampleIndex;
  P.ncoFirst32 = (uint32_t)fmodf((float)c.firstSampleIndex, c.sampleRate);  // ref: src/fm.cu:202
  P.ncoFs = c.sampleRate;
  P.ncoF = c.frequencyShift;
  void* args[] = {(void*)&P};
  const void* kernel = c.nco == kNcoExact ? (const void*)firDirectNcoKernel<kPolyNcoExact>
                                          : (const void*)firDirectNcoKernel<kPolyNcoLiteral>;
  return cudaLaunchKernel(kernel, dim3((unsigned)grid), dim3(kDirectThreads), args, 0, stream);
}

static size_t outElemBytes(FirType t) noexcept { return t == kFirFF ? 4 : 8; }

// ---------------------------------------------------------------------------------------------------------
// Tensor-core path (fir_tc_kernel.cuh): FC, decimation 4 / 8 / 16, up to 33 taps per output
// ---------------------------------------------------------------------------------------------------------
size_t tcSharedBytes(unsigned D, unsigned tablePitch) noexcept;
cudaError_t launchTc(unsigned D, TcParams& P, int dev, int smCount, cudaStream_t stream) noexcept;

// Tiles (32 windows of 32 outputs; of 64 at decimation 4) per channel when the call goes to the tensor-core kernel,
// else 0.
//
// Where it is used was decided by measurement (profiles/r02/tc_f16_sweep.txt, 2^26 samples):
//   * decimation 8, 129..264 taps: 0.132-0.146 ms against the FFMA2 kernel's 0.170 ms (0.2415 ms at 264 taps); up
//     to 128 taps the FFMA2 kernel is itself close to the HBM time (0.107 vs 0.133 ms at 127 taps);
//   * decimation 4, 65..260 taps: 0.152-0.176 ms against 0.170 ms (65..128 taps), 0.239 ms (..192), 0.307 ms (..256),
//     0.4445 ms (260); up to 64 taps FFMA2 wins (0.142 vs 0.153 ms);
//   * decimation 16, 257..528 taps: 0.167-0.194 ms against 0.202 ms (0.2806 ms at 528); up to 256 taps FFMA2 wins
//     (0.129 vs 0.155-0.167 ms).  (A tile's samples leave room for one CTA per SM there; what made it pay is a ring of
//     sixteen TMEM stages — with four it took 0.215 ms at 511 taps.)
//   * at least 65536 outputs per channel — the size from which gsdrShardPlanTime aligns shards to the kernel's tiles,
//     so that a call and its shards take the same kernel and agree bit for bit.  (Up to ~300 tiles the launch is
//     latency-bound and the two kernels tie; around 512 tiles the 444 resident CTAs leave a tail, 0.0170 vs
//     0.0154 ms; from 1024 tiles on the tensor-core kernel wins.)  The rule looks at ONE channel's outputs: a batched
//     call and the same channels filtered one by one take the same kernel.
static unsigned long long tcTilesPerChannel(const FirCall& c, int maxSmem, TcParams* P) noexcept {
  if (c.type != kFirFC || c.nco != kNcoNone || c.epilogue != kFirEpiNone) return 0;
  const size_t D = c.decimation, T = c.tapCount;
  if (D != 4 && D != 8 && D != 16) return 0;
  const int forced = forcedVariant();
  if (forced != kForceTensorCore) {
    if (forced != -1 || gTensorCores.load(std::memory_order_relaxed) == 0) return 0;
    if (!((D == 8 && T > 128) || (D == 4 && T > 64) || (D == 16 && T > 256)) || c.numOutputs < 65536) return 0;
  }
  const unsigned S = (unsigned)tcWindowOutputs((int)D), tileOut = S * kTcWindows;
  const size_t SD = (size_t)S * D;
  if (T > SD + D) return 0;                                // a window must fit two segments
  if ((uintptr_t)c.input % 16 != 0) return 0;             // bulk copies need 16-byte aligned sources
  if (c.numChannels > 1 && (c.tapStride != 0 || (c.inputStride % 2) != 0)) return 0;
  const unsigned long long tiles = (c.numOutputs + tileOut - 1) / tileOut;
  if (tiles * c.numChannels > 0x7fffffffull) return 0;
  const unsigned K = (unsigned)((S - 1) * D + T);
  P->numStages = (K + 31u) / 32u;
  P->aMax = (16u * (2u * P->numStages - 1u)) / (unsigned)D;
  P->tablePitch = (P->aMax + S) * 16u;
  if (tcSharedBytes((unsigned)D, P->tablePitch) > (size_t)maxSmem) return 0;
  return tiles;
}

static cudaError_t launchTcTiles(const FirCall& c, unsigned long long tiles, TcParams& P, int dev, int smCount,
                                 cudaStream_t stream) noexcept {
  P.x = (const float2*)c.input;
  P.h = (const float*)c.taps;
  P.y = (float2*)c.output;
  P.xStride = c.inputStride;
  P.yStride = c.outputStride;
  P.nOut = c.numOutputs;
  P.nIn = (unsigned long long)(c.numOutputs - 1) * c.decimation + c.tapCount;
  P.tilesPerChannel = (unsigned)tiles;
  P.totalTiles = (unsigned)(tiles * c.numChannels);
  P.T = (unsigned)c.tapCount;
  return launchTc((unsigned)c.decimation, P, dev, smCount, stream);
}

cudaError_t enqueueFir(const FirCall& c, cudaStream_t stream) noexcept {
  if (c.numOutputs == 0 || c.numChannels == 0) return cudaSuccess;
  if (c.decimation == 0) return cudaErrorInvalidValue;
  if (c.nco != kNcoNone && c.type != kFirFC) return cudaErrorInvalidValue;
  if (c.epilogue != kFirEpiNone && (c.type != kFirFC || c.tapCount == 0 || c.nco != kNcoExact)) return cudaErrorNotSupported;
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

  {
    TcParams tp{};
    const unsigned long long tcTiles = tcTilesPerChannel(c, info->maxSmemOptin, &tp);
    if (tcTiles > 0) return launchTcTiles(c, tcTiles, tp, dev, info->smCount, stream);
  }
  {
    TmaGeom tg{};
    // a tensor map the driver will not encode (kTmaEncodeFailed) sends the call on to the kernels that need none
    const int wv = chooseWideVariant(c, info->maxSmemOptin, &tg);
    if (wv >= 0) {
      st = launchTma(c, -1, tg, dev, info->smCount, stream, -1, wv);
      if (st != kTmaEncodeFailed) return st;
    }
    const int tv = wv >= 0 ? -1 : chooseTmaVariant(c, info->maxSmemOptin, &tg);
    if (tv >= 0) {
      st = launchTma(c, tv, tg, dev, info->smCount, stream);
      if (st != kTmaEncodeFailed) return st;
    }
    if (c.epilogue != kFirEpiNone) return cudaErrorNotSupported;  // only the TMA-fed kernels have fused output stages
    RealGeom rg{};
    const int rv = chooseRealVariant(c, info->maxSmemOptin, &rg);
    if (rv >= 0) return launchReal(c, rv, rg, dev, info->smCount, stream);
    const int cv = chooseCcVariant(c, info->maxSmemOptin, &tg);
    if (cv >= 0) {
      st = launchTma(c, -1, tg, dev, info->smCount, stream, cv);
      if (st != kTmaEncodeFailed) return st;
    }
    const int fv = chooseCfVariant(c, info->maxSmemOptin, &rg);
    if (fv >= 0) return launchReal(c, -1, rg, dev, info->smCount, stream, fv);
  }
  const bool polyType = (c.type == kFirFC || c.type == kFirFF);
  PolyGeom geom{};
  const int variant = polyType ? choosePolyVariant(c.decimation, c.tapCount, c.numOutputs, info->maxSmemOptin, &geom,
                                                   c.type == kFirFF && c.decimation == 1)
                               : -1;
  if (variant < 0) {
    if (c.nco != kNcoNone) return launchDirectNco(c, stream);
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
  P.dbg = debugFlags();
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

// ---------------------------------------------------------------------------------------------------------
// int8 IQ input (<gsdr/conversion.h>): firTmaInt8Kernel, or the direct kernel when the call does not qualify
// ---------------------------------------------------------------------------------------------------------
cudaError_t enqueueFirInt8(bool nco, float sampleRate, float frequencyShift, size_t firstSampleIndex, size_t decimation,
                           const float* taps, size_t tapCount, const signed char* input, float2* output,
                           size_t numOutputs, cudaStream_t stream) noexcept {
  if (numOutputs == 0) return cudaSuccess;
  if (decimation == 0) return cudaErrorInvalidValue;
  if (tapCount == 0) return cudaMemsetAsync(output, 0, numOutputs * sizeof(cuComplex), stream);
  int dev = 0;
  cudaError_t st = cudaGetDevice(&dev);
  if (st != cudaSuccess) return st;
  const DeviceInfo* info = deviceInfo(dev);
  if (!info || info->status != cudaSuccess) return info ? info->status : cudaErrorInvalidDevice;
  const size_t D = decimation, T = tapCount;
  const unsigned long long nIn = (unsigned long long)(numOutputs - 1) * D + T;
  const int forced = forcedFfmaVariant();

  if (forced != -2 && tmaSupportedDecimation(D) && (uintptr_t)input % 16 == 0) {
    const size_t J = (T + D - 1) / D;
    const size_t Jpad = (J + 15) / 16 * 16;
    const size_t G = tmaSegBytes((unsigned)D);
    const bool staticD = (D == 8 || D == 10) && Jpad <= kTmaJpadCap;
    static const int orderNarrow[] = {0, 2, 1};
    static const int orderWide[] = {1, 0, 2};
    const int* order = D > 16 ? orderWide : orderNarrow;
    for (int k = 0; k < kNumInt8Variants; k++) {
      const int id = order[k];
      const SpecVariant& v = kInt8Variants[id];
      if ((D / 2) < (size_t)v.psplit) continue;
      const size_t mhp = tmaPlaneRows((unsigned)v.tg, (unsigned)(staticD ? kTmaJpadCap : Jpad), (unsigned)D);
      const size_t planeBytes = mhp * G + kRealPlanePad;
      const size_t rows = (size_t)kTmaR * v.tg + Jpad;
      const size_t rawBytes = (rows * D * 2 + 15) / 16 * 16;
      const size_t smem = 2 * (8 * D / G) * 8 * planeBytes + 2 * (size_t)(v.psplit - 1) * v.tg * 64 +
                          (D * Jpad + 32) * 4 + (rows + D + 1) * 8 + 2 * rawBytes;
      if (smem > (size_t)info->maxSmemOptin || rawBytes > 0xfffff0u) continue;
      const size_t bout = (size_t)kTmaR * v.tg;
      const unsigned long long tiles = (numOutputs + bout - 1) / bout;
      if (tiles > 0x7fffffffull) break;
      Int8Params P{};
      P.x = (const float2*)input;
      P.h = taps;
      P.y = (float2*)output;
      P.nOut = numOutputs;
      P.nIn = nIn;
      P.tilesPerChannel = (unsigned)tiles;
      P.totalTiles = (unsigned)tiles;
      P.numChannels = 1;
      P.D = (unsigned)D;
      P.T = (unsigned)T;
      P.Jpad = (unsigned)Jpad;
      P.rowBytes = (unsigned)(8 * D);
      P.segBytes = (unsigned)G;
      P.mhp = (unsigned)mhp;
      P.planeBytes = (unsigned)planeBytes;
      if (G == 32) P.swzShift = 2, P.swzMask = 1;
      if (G == 64) P.swzShift = 1, P.swzMask = 3;
      if (G == 128) P.swzShift = 0, P.swzMask = 7;
      P.y16 = ((uintptr_t)output % 16 == 0) ? 1u : 0u;
      P.ncoStep = ncoPhaseStep(frequencyShift, sampleRate);
      P.ncoFirst = firstSampleIndex;
      P.rawBytes = (unsigned)rawBytes;
      P.tapScale = 1.0f / 127.0f;
      if (staticD && D == 8) return launchInt8Dt8(nco, id, P, smem, dev, info->smCount, stream);
      if (staticD && D == 10) return launchInt8Dt10(nco, id, P, smem, dev, info->smCount, stream);
      return launchInt8Dt0(nco, id, P, smem, dev, info->smCount, stream);
    }
  }

  DirectInt8Params P{};
  P.x = (const signed char*)input;
  P.h = taps;
  P.y = (float2*)output;
  P.nOut = numOutputs;
  P.D = D;
  P.T = T;
  const unsigned long long bpc = (numOutputs + kDirectThreads - 1) / kDirectThreads;
  if (bpc > 0x7fffffffull) return cudaErrorInvalidValue;
  P.blocksPerChannel = (unsigned)bpc;
  P.ncoStep = ncoPhaseStep(frequencyShift, sampleRate);
  P.ncoFirst = firstSampleIndex;
  void* args[] = {(void*)&P};
  const void* kernel = nco ? (const void*)firDirectInt8Kernel<true> : (const void*)firDirectInt8Kernel<false>;
  return cudaLaunchKernel(kernel, dim3((unsigned)bpc), dim3(kDirectThreads), args, 0, stream);
}

// One input, numShifts frequency shifts: the fused kernel where the shape allows (window fetched once per tile, mixed
// and filtered per shift), otherwise one fused NCO + FIR call per shift.  Results are those of the separate calls.
cudaError_t enqueueChannelizer(float sampleRate, const float* frequencyShifts, size_t numShifts, size_t firstSampleIndex,
                               size_t decimation, const float* taps, size_t tapCount, const cuComplex* input,
                               cuComplex* output, size_t outputStride, size_t numOutputs, cudaStream_t stream) noexcept {
  if (numOutputs == 0 || numShifts == 0) return cudaSuccess;
  if (decimation == 0 || !frequencyShifts) return cudaErrorInvalidValue;
  FirCall c;
  c.type = kFirFC;
  c.nco = kNcoExact;
  c.decimation = decimation;
  c.taps = taps;
  c.tapCount = tapCount;
  c.input = input;
  c.numOutputs = numOutputs;
  c.sampleRate = sampleRate;
  c.firstSampleIndex = firstSampleIndex;
  int dev = 0;
  cudaError_t st = cudaGetDevice(&dev);
  if (st != cudaSuccess) return st;
  const DeviceInfo* info = deviceInfo(dev);
  if (!info || info->status != cudaSuccess) return info ? info->status : cudaErrorInvalidDevice;
  const ChanShape* shape = nullptr;
  for (const ChanShape& sh : kChanShapes)
    if ((size_t)sh.D == decimation) shape = &sh;
  TmaGeom geom{};
  // Measured and NOT adopted (profiles/r02/channelizer.jsonl: 0.84-1.02 x the speed of the per-shift launches —
  // the FIR is issue-bound for every tap count, so fetching the window once buys nothing): the fused kernel exists in
  // the tuning build only, behind the override.
#ifdef GSDR_B200_TUNING
  bool fused = forcedVariant() == kForceFusedChannelizer && shape && tapCount > 0 && encodeTiled() &&
               (uintptr_t)input % 16 == 0 && numOutputs < 0xfff00000ull &&
               tmaGeometry(TmaVariant{shape->tg, shape->psplit, 2, 1}, decimation, tapCount, &geom) && geom.staticD;
#else
  bool fused = false;
#endif
  if (fused) {
    // single partial-sum scratch instead of the double one; one rotation table per shift
    const size_t maxShifts = numShifts < kChanMaxShifts ? numShifts : kChanMaxShifts;
    geom.smemBytes += (maxShifts - 1) * decimation * 8;
    fused = geom.smemBytes <= (size_t)info->maxSmemOptin;
  }
  for (size_t k0 = 0; k0 < numShifts; k0 += kChanMaxShifts) {
    const size_t n = numShifts - k0 < kChanMaxShifts ? numShifts - k0 : kChanMaxShifts;
    bool done = false;
    if (fused) {
      unsigned long long steps[kChanMaxShifts];
      for (size_t k = 0; k < n; k++) steps[k] = ncoPhaseStep(frequencyShifts[k0 + k], sampleRate);
      ChanExtra ex{shape->tg, shape->psplit, (unsigned)n, (unsigned long long)outputStride, steps};
      c.output = output + k0 * outputStride;
      st = launchTma(c, -1, geom, dev, info->smCount, stream, -1, -1, &ex);
      if (st == cudaSuccess) done = true;
      else if (st != kTmaEncodeFailed) return st;
    }
    for (size_t k = 0; !done && k < n; k++) {
      c.frequencyShift = frequencyShifts[k0 + k];
      c.output = output + (k0 + k) * outputStride;
      st = enqueueFir(c, stream);
      if (st != cudaSuccess) return st;
    }
  }
  return cudaSuccess;
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

GSDR_C_LINKAGE cudaError_t gsdrChannelizeFC(float sampleRate, const float* frequencyShifts, size_t numShifts,
                                            size_t firstSampleIndex, size_t decimation, const float* taps,
                                            size_t tapCount, const cuComplex* input, cuComplex* output,
                                            size_t outputStride, size_t numOutputs, int32_t cudaDevice,
                                            cudaStream_t cudaStream) GSDR_NO_EXCEPT {
  DeviceScope scope(cudaDevice);
  if (scope.status() != cudaSuccess) return scope.status();
  return enqueueChannelizer(sampleRate, frequencyShifts, numShifts, firstSampleIndex, decimation, taps, tapCount, input,
                            output, outputStride, numOutputs, cudaStream);
}

// ---- <gsdr/conversion.h> ---------------------------------------------------------------------------------

GSDR_C_LINKAGE cudaError_t gsdrInt8ToNormFloat(const int8_t* input, float* output, size_t numElements,
                                               int32_t cudaDevice, cudaStream_t cudaStream) GSDR_NO_EXCEPT {
  DeviceScope scope(cudaDevice);
  if (scope.status() != cudaSuccess) return scope.status();
  if (numElements == 0) return cudaSuccess;
  const unsigned long long blocks = (numElements + 255) / 256;
  if (blocks > 0x7fffffffull) return cudaErrorInvalidValue;
  unsigned long long n = numElements;
  const signed char* in = (const signed char*)input;
  void* args[] = {(void*)&in, (void*)&output, (void*)&n};
  return cudaLaunchKernel((const void*)int8ToNormFloatKernel, dim3((unsigned)blocks), dim3(256), args, 0, cudaStream);
}

GSDR_C_LINKAGE cudaError_t gsdrFirFCInt8(size_t decimation, const float* taps, size_t tapCount, const int8_t* input,
                                         cuComplex* output, size_t numOutputs, int32_t cudaDevice,
                                         cudaStream_t cudaStream) GSDR_NO_EXCEPT {
  DeviceScope scope(cudaDevice);
  if (scope.status() != cudaSuccess) return scope.status();
  return enqueueFirInt8(false, 1.0f, 0.0f, 0, decimation, taps, tapCount, (const signed char*)input, (float2*)output,
                        numOutputs, cudaStream);
}

GSDR_C_LINKAGE cudaError_t gsdrAdjustFrequencyFirFCInt8(float sampleRate, float frequencyShift,
                                                        size_t firstSampleIndex, size_t decimation, const float* taps,
                                                        size_t tapCount, const int8_t* input, cuComplex* output,
                                                        size_t numOutputs, int32_t cudaDevice,
                                                        cudaStream_t cudaStream) GSDR_NO_EXCEPT {
  DeviceScope scope(cudaDevice);
  if (scope.status() != cudaSuccess) return scope.status();
  return enqueueFirInt8(true, sampleRate, frequencyShift, firstSampleIndex, decimation, taps, tapCount,
                        (const signed char*)input, (float2*)output, numOutputs, cudaStream);
}

// ---- tuning / introspection hooks of <gsdr/b200.h> -------------------------------------------------------

#ifdef GSDR_B200_TUNING
GSDR_C_LINKAGE int gsdrB200SetKernelVariant(int variant) GSDR_NO_EXCEPT {
  if (variant < -5 || variant >= firstWideVariantId() + kNumWideVariants) return -1;
  gForcedVariant.store(variant, std::memory_order_relaxed);
  return 0;
}

GSDR_C_LINKAGE int gsdrB200SetDebugFlags(int flags) GSDR_NO_EXCEPT {
  gDebugFlags.store(flags & 7, std::memory_order_relaxed);
  return 0;
}
GSDR_C_LINKAGE int gsdrB200HasTuningHooks(void) GSDR_NO_EXCEPT { return 1; }
#else
GSDR_C_LINKAGE int gsdrB200HasTuningHooks(void) GSDR_NO_EXCEPT { return 0; }
#endif

GSDR_C_LINKAGE int gsdrB200SetFirTensorCores(int enable) GSDR_NO_EXCEPT {
  return gTensorCores.exchange(enable ? 1 : 0, std::memory_order_relaxed);
}

GSDR_C_LINKAGE int gsdrB200NumKernelVariants(void) GSDR_NO_EXCEPT {
  return firstWideVariantId() + kNumWideVariants;
}
GSDR_C_LINKAGE int gsdrB200NumPolyphaseVariants(void) GSDR_NO_EXCEPT { return kNumVariants; }

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
  if (firType == kFirFC || firType == 4) {
    // assumes what the common call has: one channel, 16-byte aligned input
    FirCall probe;
    probe.type = kFirFC;
    probe.nco = (firType == 4) ? kNcoExact : kNcoNone;
    probe.decimation = decimation;
    probe.tapCount = tapCount;
    probe.numOutputs = numOutputs;
    if (firType == kFirFC) {
      TcParams tp{};
      const unsigned long long tiles = tcTilesPerChannel(probe, di->maxSmemOptin, &tp);
      if (tiles > 0) {
        info->variant = firstWideVariantId() + kNumWideVariants;  // == gsdrB200NumKernelVariants(): the tensor-core kernel
        info->outputsPerThread = 0;
        info->threadsPerBlock = kTcThreads;
        info->phaseGroups = 1;
        info->windowBuffers = 1;
        info->outputsPerBlock = (size_t)tcWindowOutputs((int)decimation) * kTcWindows;
        info->sharedBytesPerBlock = tcSharedBytes((unsigned)decimation, tp.tablePitch);
        info->numBlocks = (size_t)tiles;
        return 0;
      }
    }
    TmaGeom tg{};
    const int wv = chooseWideVariant(probe, di->maxSmemOptin, &tg);
    if (wv >= 0) {
      const SpecVariant& ws = kWideVariants[wv];
      const size_t bout = (size_t)kTmaR * ws.tg;
      info->variant = firstWideVariantId() + wv;
      info->outputsPerThread = kTmaR;
      info->threadsPerBlock = ws.threads();
      info->phaseGroups = ws.psplit;
      info->windowBuffers = 2;
      info->outputsPerBlock = bout;
      info->sharedBytesPerBlock = tg.smemBytes;
      info->numBlocks = (numOutputs + bout - 1) / bout;
      return 0;
    }
    const int tv = chooseTmaVariant(probe, di->maxSmemOptin, &tg);
    if (tv >= 0) {
      TmaVariant vs;
      int mixw = 0;
      tmaVariantShape(tv, &vs, &mixw);
      const size_t bout = (size_t)kTmaR * vs.tg;
      info->variant = kNumVariants + tv;
      info->outputsPerThread = kTmaR;
      info->threadsPerBlock = vs.threads() + 32 * mixw;
      info->phaseGroups = vs.psplit;
      info->windowBuffers = vs.nbuf;
      info->outputsPerBlock = bout;
      info->sharedBytesPerBlock = tg.smemBytes;
      info->numBlocks = (numOutputs + bout - 1) / bout;
      return 0;
    }
  }
  if (firType == kFirCC) {
    FirCall probe;
    probe.type = kFirCC;
    probe.decimation = decimation;
    probe.tapCount = tapCount;
    probe.numOutputs = numOutputs;
    TmaGeom cg{};
    const int cv = chooseCcVariant(probe, di->maxSmemOptin, &cg);
    if (cv >= 0) {
      const TmaVariant& vs = kCcVariants[cv];
      const size_t bout = (size_t)kTmaR * vs.tg;
      info->variant = firstCcVariantId() + cv;
      info->outputsPerThread = kTmaR;
      info->threadsPerBlock = vs.threads();
      info->phaseGroups = vs.psplit;
      info->windowBuffers = vs.nbuf;
      info->outputsPerBlock = bout;
      info->sharedBytesPerBlock = cg.smemBytes;
      info->numBlocks = (numOutputs + bout - 1) / bout;
      return 0;
    }
  }
  if (firType == kFirCF) {
    FirCall probe;
    probe.type = kFirCF;
    probe.decimation = decimation;
    probe.tapCount = tapCount;
    probe.numOutputs = numOutputs;
    RealGeom fg{};
    const int fv = chooseCfVariant(probe, di->maxSmemOptin, &fg);
    if (fv >= 0) {
      const RealVariant& vs = kCfVariants[fv];
      const size_t bout = 2 * (size_t)kTmaR * vs.tg;
      info->variant = firstCfVariantId() + fv;
      info->outputsPerThread = 2 * kTmaR;
      info->threadsPerBlock = vs.threads();
      info->phaseGroups = vs.psplit;
      info->windowBuffers = vs.nwin;
      info->outputsPerBlock = bout;
      info->sharedBytesPerBlock = fg.smemBytes;
      info->numBlocks = (numOutputs + bout - 1) / bout;
      return 0;
    }
  }
  if (firType == kFirFF) {
    FirCall probe;
    probe.type = kFirFF;
    probe.decimation = decimation;
    probe.tapCount = tapCount;
    probe.numOutputs = numOutputs;
    RealGeom rg{};
    const int rv = chooseRealVariant(probe, di->maxSmemOptin, &rg);
    if (rv >= 0) {
      const RealVariant& vs = kRealVariants[rv];
      const size_t bout = 2 * (size_t)kTmaR * vs.tg;
      info->variant = kNumVariants + kNumTmaVariants + kNumSpecVariants + rv;
      info->outputsPerThread = 2 * kTmaR;
      info->threadsPerBlock = vs.threads();
      info->phaseGroups = vs.psplit;
      info->windowBuffers = vs.nwin;
      info->outputsPerBlock = bout;
      info->sharedBytesPerBlock = rg.smemBytes;
      info->numBlocks = (numOutputs + bout - 1) / bout;
      return 0;
    }
  }
  PolyGeom g{};
  const bool polyType = (firType == kFirFC || firType == kFirFF || firType == 4);
  const int v = polyType ? choosePolyVariant(decimation, tapCount, numOutputs, di->maxSmemOptin, &g,
                                             firType == kFirFF && decimation == 1)
                         : -1;
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
