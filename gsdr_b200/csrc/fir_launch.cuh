// fir_launch.cuh — kernel variant tables and the launch templates of the TMA-fed kernels, shared by gsdr_fir.cu (which
// chooses a variant) and the fir_inst_*.cu translation units (which instantiate the kernels: one unit per compile-time
// decimation, so that the ~300 instantiations compile in parallel).
#pragma once

#include <atomic>
#include <cstdio>

#include "fir_kernels.cuh"
#include "fir_tma_kernel.cuh"
#include "launch.h"

namespace gsdr_b200 {

// one stderr line on failure, clears the runtime's last-error slot (defined in gsdr_fir.cu)
cudaError_t report(cudaError_t st, const char* what) noexcept;

// The occupancy answer for a (kernel instantiation, device, dynamic shared memory) triple never changes: cache the last
// one per launch site and device (one query per call was ~10 % of the host-side cost of a call).
struct OccCache {
  std::atomic<unsigned long long> entry[64];  // (smem << 8) | CTAs per SM, 0 = empty
};
template <class K>
static cudaError_t occupancyCached(OccCache& cache, K kernel, int threads, size_t smem, int dev, int* perSm) noexcept {
  const unsigned long long e = cache.entry[dev & 63].load(std::memory_order_acquire);
  if (e != 0 && (e >> 8) == (unsigned long long)smem) {
    *perSm = (int)(e & 0xffu);
    return cudaSuccess;
  }
  const cudaError_t st = cudaOccupancyMaxActiveBlocksPerMultiprocessor(perSm, kernel, threads, smem);
  if (st == cudaSuccess && *perSm > 0 && *perSm < 256) {
    cache.entry[dev & 63].store(((unsigned long long)smem << 8) | (unsigned)*perSm, std::memory_order_release);
  }
  return st;
}

struct TmaVariant {
  int tg, psplit, nbuf, minBlocks;
  int threads() const { return tg * psplit; }
};
// X(id, TG, PSPLIT, NBUF, MINB) — ids continue after the polyphase variants
#ifndef GSDR_EXP_MINB1
#define GSDR_EXP_MINB1 4
#endif
#define GSDR_TMA_VARIANTS(X) \
  X(0, 64, 2, 2, 2)          \
  X(1, 32, 2, 2, GSDR_EXP_MINB1) \
  X(2, 64, 1, 2, 2)          \
  X(3, 128, 2, 2, 1)         \
  X(4, 32, 4, 2, 3)          \
  X(5, 32, 2, 1, 6)          \
  X(6, 64, 2, 1, 4)          \
  X(7, 32, 1, 1, 8)          \
  X(8, 32, 1, 2, 5)          \
  X(9, 64, 1, 1, 4)          \
  X(10, 32, 4, 1, 2)         \
  X(11, 32, 8, 1, 1)

static constexpr TmaVariant kTmaVariants[] = {
#define X(id, tg, ps, nb, mb) {tg, ps, nb, mb},
    GSDR_TMA_VARIANTS(X)
#undef X
};
static constexpr int kNumTmaVariants = (int)(sizeof(kTmaVariants) / sizeof(kTmaVariants[0]));

// Warp-specialised fused-NCO kernel: X(id, TG, PSPLIT, MIXW, MINB); ids continue after the TMA variants
#define GSDR_SPEC_VARIANTS(X) \
  X(0, 32, 2, 2, 4)           \
  X(1, 32, 4, 4, 1)           \
  X(2, 64, 2, 4, 2)           \
  X(3, 32, 1, 1, 4)           \
  X(4, 32, 2, 1, 4)           \
  X(5, 32, 4, 2, 2)

struct SpecVariant {
  int tg, psplit, mixw, minBlocks;
  int threads() const { return tg * psplit + 32 * mixw; }
};
static constexpr SpecVariant kSpecVariants[] = {
#define X(id, tg, ps, mw, mb) {tg, ps, mw, mb},
    GSDR_SPEC_VARIANTS(X)
#undef X
};
static constexpr int kNumSpecVariants = (int)(sizeof(kSpecVariants) / sizeof(kSpecVariants[0]));

// Real-input kernel (firTmaRealKernel): X(id, TG, PSPLIT, MIXW, NWIN, NRAW, MINB); ids continue after the fused-NCO
// variants.  NRAW - 1 bulk copies are in flight per CTA: low-rate (HBM-bound) shapes want 3 or 4.
#define GSDR_REAL_VARIANTS(X) \
  X(0, 128, 1, 4, 2, 2, 2)    \
  X(1, 64, 1, 2, 2, 2, 4)     \
  X(2, 32, 1, 1, 2, 3, 8)     \
  X(3, 32, 1, 1, 1, 3, 8)     \
  X(4, 64, 2, 4, 2, 3, 2)     \
  X(5, 64, 1, 2, 1, 3, 4)     \
  X(6, 32, 1, 1, 1, 4, 8)     \
  X(7, 128, 1, 4, 1, 3, 2)    \
  X(8, 64, 1, 2, 1, 4, 4)     \
  X(9, 128, 1, 4, 1, 4, 2)

struct RealVariant {
  int tg, psplit, mixw, nwin, nraw, minBlocks;
  int threads() const { return tg * psplit + 32 * mixw; }
};
static constexpr RealVariant kRealVariants[] = {
#define X(id, tg, ps, mw, nw, nr, mb) {tg, ps, mw, nw, nr, mb},
    GSDR_REAL_VARIANTS(X)
#undef X
};
static constexpr int kNumRealVariants = (int)(sizeof(kRealVariants) / sizeof(kRealVariants[0]));

// Complex-tap kernel (firTmaCcKernel): X(id, TG, PSPLIT, NBUF, MINB), PSPLIT even (two tap planes); ids continue
// after the real-input variants
#define GSDR_CC_VARIANTS(X) \
  X(0, 32, 2, 2, 4)         \
  X(1, 32, 4, 2, 3)         \
  X(2, 64, 2, 2, 2)         \
  X(3, 32, 4, 1, 2)

static constexpr TmaVariant kCcVariants[] = {
#define X(id, tg, ps, nb, mb) {tg, ps, nb, mb},
    GSDR_CC_VARIANTS(X)
#undef X
};
static constexpr int kNumCcVariants = (int)(sizeof(kCcVariants) / sizeof(kCcVariants[0]));

// Segment-pipelined kernel for wide rows (firTmaWideKernel): X(id, TG, PSPLIT, MIXW, MINB); global ids continue after
// the real-input x complex-tap variants.  MIXW = 1: plain FIR (the extra warp only issues the tensor copies).
#define GSDR_WIDE_VARIANTS(X) \
  X(0, 64, 4, 1, 1)           \
  X(1, 64, 4, 4, 1)           \
  X(2, 64, 4, 2, 1)

static constexpr SpecVariant kWideVariants[] = {
#define X(id, tg, ps, mw, mb) {tg, ps, mw, mb},
    GSDR_WIDE_VARIANTS(X)
#undef X
};
static constexpr int kNumWideVariants = (int)(sizeof(kWideVariants) / sizeof(kWideVariants[0]));

template <int MODE, int TG, int PSPLIT, int DT, int NBUF, int MINB>
static cudaError_t launchTmaT(const CUtensorMap& map, TmaParams& P, size_t smem, int dev, int smCount,
                              cudaStream_t stream) noexcept {
  static std::atomic<size_t> configured[64];
  auto kernel = firTmaKernel<MODE, TG, PSPLIT, DT, NBUF, MINB>;
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
  // Warps are pinned to one of the SM's four sub-partitions; with a static tile assignment the kernel runs at the
  // pace of the fullest one, so keep the resident warp count per SM a multiple of 4 (profiles/r01_notes.md).
  const int fullPerSm = perSm;
  {
    const int warpsPerCta = (TG * PSPLIT) / 32;
    int balanced = perSm;
    while (balanced > 1 && (balanced * warpsPerCta) % 4 != 0) balanced--;
    if ((balanced * warpsPerCta) % 4 == 0) perSm = balanced;
  }
  const unsigned long long resident = (unsigned long long)perSm * (unsigned)smCount;
  const unsigned grid = (unsigned)(P.totalTiles < resident ? P.totalTiles : resident);
  P.strideChan = grid / P.tilesPerChannel;
  P.strideTile = grid % P.tilesPerChannel;
  void* args[] = {(void*)&map, (void*)&P};
#ifndef GSDR_NO_PDL
  // Programmatic dependent launch lets the next launch's CTAs become resident as this grid drains.  Only when this
  // grid fills every CTA slot of an SM: where a slot is left free on purpose (sub-partition balance above), early
  // dependents would sit in it for the whole run (config 4: 15.2 ms instead of 12.9 ms).
  if (perSm != fullPerSm) {
    return cudaLaunchKernel((const void*)kernel, dim3(grid), dim3(TG * PSPLIT), args, smem, stream);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(TG * PSPLIT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelExC(&cfg, (const void*)kernel, args);
#else
  return cudaLaunchKernel((const void*)kernel, dim3(grid), dim3(TG * PSPLIT), args, smem, stream);
#endif
}

template <int TG, int PSPLIT, int DT, int MIXW, int MINB>
static cudaError_t launchSpecT(const CUtensorMap& map, TmaParams& P, size_t smem, int dev, int smCount,
                               cudaStream_t stream) noexcept {
  static std::atomic<size_t> configured[64];
  auto kernel = firTmaNcoSpecKernel<TG, PSPLIT, DT, MIXW, MINB>;
  constexpr int kThreads = TG * PSPLIT + 32 * MIXW;
  if (smem > configured[dev & 63].load(std::memory_order_acquire)) {
    cudaError_t st = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (st != cudaSuccess) return report(st, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    configured[dev & 63].store(smem, std::memory_order_release);
  }
  int perSm = 0;
  static OccCache occ;
  cudaError_t st = occupancyCached(occ, kernel, kThreads, smem, dev, &perSm);
  if (st != cudaSuccess) return report(st, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
  if (perSm < 1) return cudaErrorInvalidConfiguration;
  {
    const int warpsPerCta = kThreads / 32;
    int balanced = perSm;
    while (balanced > 1 && (balanced * warpsPerCta) % 4 != 0) balanced--;
    if ((balanced * warpsPerCta) % 4 == 0) perSm = balanced;
  }
  const unsigned long long resident = (unsigned long long)perSm * (unsigned)smCount;
  const unsigned grid = (unsigned)(P.totalTiles < resident ? P.totalTiles : resident);
  P.strideChan = grid / P.tilesPerChannel;
  P.strideTile = grid % P.tilesPerChannel;
  void* args[] = {(void*)&map, (void*)&P};
  return cudaLaunchKernel((const void*)kernel, dim3(grid), dim3(kThreads), args, smem, stream);
}

template <int MODE, int TG, int PSPLIT, int DT, int MIXW, int MINB>
static cudaError_t launchWideT(const CUtensorMap& map, TmaParams& P, size_t smem, int dev, int smCount,
                               cudaStream_t stream) noexcept {
  static std::atomic<size_t> configured[64];
  auto kernel = firTmaWideKernel<MODE, TG, PSPLIT, DT, MIXW, MINB>;
  constexpr int kThreads = TG * PSPLIT + 32 * MIXW;
  if (smem > configured[dev & 63].load(std::memory_order_acquire)) {
    cudaError_t st = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (st != cudaSuccess) return report(st, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    configured[dev & 63].store(smem, std::memory_order_release);
  }
  // two 72 KB stage buffers: one CTA per SM
  const unsigned grid = (unsigned)(P.totalTiles < (unsigned)smCount ? P.totalTiles : (unsigned)smCount);
  P.strideChan = grid / P.tilesPerChannel;
  P.strideTile = grid % P.tilesPerChannel;
  void* args[] = {(void*)&map, (void*)&P};
  return cudaLaunchKernel((const void*)kernel, dim3(grid), dim3(kThreads), args, smem, stream);
}

template <int DT>
static cudaError_t launchWideD(int mode, int variant, const CUtensorMap& map, TmaParams& P, size_t smem, int dev,
                               int smCount, cudaStream_t stream) noexcept {
  switch (variant) {
#define X(id, tg, ps, mw, mb)                                                                                      \
  case id:                                                                                                         \
    return mode == kPolyFC ? launchWideT<kPolyFC, tg, ps, DT, mw, mb>(map, P, smem, dev, smCount, stream)          \
                           : launchWideT<kPolyNcoExact, tg, ps, DT, mw, mb>(map, P, smem, dev, smCount, stream);
    GSDR_WIDE_VARIANTS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}

template <int TG, int PSPLIT, int DT, int MINB>
static cudaError_t launchChanT(const CUtensorMap& map, ChanParams& P, size_t smem, int dev, int smCount,
                               cudaStream_t stream) noexcept {
  static std::atomic<size_t> configured[64];
  auto kernel = firTmaChannelizerKernel<TG, PSPLIT, DT, MINB>;
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
  {
    const int warpsPerCta = (TG * PSPLIT) / 32;
    int balanced = perSm;
    while (balanced > 1 && (balanced * warpsPerCta) % 4 != 0) balanced--;
    if ((balanced * warpsPerCta) % 4 == 0) perSm = balanced;
  }
  const unsigned long long resident = (unsigned long long)perSm * (unsigned)smCount;
  const unsigned grid = (unsigned)(P.tilesPerChannel < resident ? P.tilesPerChannel : resident);
  void* args[] = {(void*)&map, (void*)&P};
  return cudaLaunchKernel((const void*)kernel, dim3(grid), dim3(TG * PSPLIT), args, smem, stream);
}

template <int DT>
static cudaError_t launchSpecD(int variant, const CUtensorMap& map, TmaParams& P, size_t smem, int dev, int smCount,
                               cudaStream_t stream) noexcept {
  switch (variant) {
#define X(id, tg, ps, mw, mb) \
  case id: return launchSpecT<tg, ps, DT, mw, mb>(map, P, smem, dev, smCount, stream);
    GSDR_SPEC_VARIANTS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}

template <int MODE, int DT>
static cudaError_t launchTmaModeD(int variant, const CUtensorMap& map, TmaParams& P, size_t smem, int dev,
                                  int smCount, cudaStream_t stream) noexcept {
  switch (variant) {
#define X(id, tg, ps, nb, mb) \
  case id: return launchTmaT<MODE, tg, ps, DT, nb, mb>(map, P, smem, dev, smCount, stream);
    GSDR_TMA_VARIANTS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}

template <int TG, int PSPLIT, int DT, int NBUF, int MINB>
static cudaError_t launchCcT(const CUtensorMap& map, TmaParams& P, size_t smem, int dev, int smCount,
                             cudaStream_t stream) noexcept {
  static std::atomic<size_t> configured[64];
  auto kernel = firTmaCcKernel<TG, PSPLIT, DT, NBUF, MINB>;
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
  {
    const int warpsPerCta = (TG * PSPLIT) / 32;
    int balanced = perSm;
    while (balanced > 1 && (balanced * warpsPerCta) % 4 != 0) balanced--;
    if ((balanced * warpsPerCta) % 4 == 0) perSm = balanced;
  }
  const unsigned long long resident = (unsigned long long)perSm * (unsigned)smCount;
  const unsigned grid = (unsigned)(P.totalTiles < resident ? P.totalTiles : resident);
  P.strideChan = grid / P.tilesPerChannel;
  P.strideTile = grid % P.tilesPerChannel;
  void* args[] = {(void*)&map, (void*)&P};
  return cudaLaunchKernel((const void*)kernel, dim3(grid), dim3(TG * PSPLIT), args, smem, stream);
}

template <int DT>
static cudaError_t launchCcD(int variant, const CUtensorMap& map, TmaParams& P, size_t smem, int dev, int smCount,
                             cudaStream_t stream) noexcept {
  switch (variant) {
#define X(id, tg, ps, nb, mb) \
  case id: return launchCcT<tg, ps, DT, nb, mb>(map, P, smem, dev, smCount, stream);
    GSDR_CC_VARIANTS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}

template <int TG, int PSPLIT, int DT, int MIXW, int NWIN, int NRAW, int MINB, int NPLANE = 1>
static cudaError_t launchRealT(RealParams& P, size_t smem, int dev, int smCount, cudaStream_t stream) noexcept {
  static std::atomic<size_t> configured[64];
  auto kernel = firTmaRealKernel<TG, PSPLIT, DT, MIXW, NWIN, NRAW, MINB, NPLANE>;
  constexpr int kThreads = TG * PSPLIT + 32 * MIXW;
  if (smem > configured[dev & 63].load(std::memory_order_acquire)) {
    cudaError_t st = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (st != cudaSuccess) return report(st, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    configured[dev & 63].store(smem, std::memory_order_release);
  }
  int perSm = 0;
  static OccCache occ;
  cudaError_t st = occupancyCached(occ, kernel, kThreads, smem, dev, &perSm);
  if (st != cudaSuccess) return report(st, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
  if (perSm < 1) return cudaErrorInvalidConfiguration;
  {
    const int warpsPerCta = kThreads / 32;
    int balanced = perSm;
    while (balanced > 1 && (balanced * warpsPerCta) % 4 != 0) balanced--;
    if ((balanced * warpsPerCta) % 4 == 0) perSm = balanced;
  }
  const unsigned long long resident = (unsigned long long)perSm * (unsigned)smCount;
  const unsigned grid = (unsigned)(P.totalTiles < resident ? P.totalTiles : resident);
  P.strideChan = grid / P.tilesPerChannel;
  P.strideTile = grid % P.tilesPerChannel;
  void* args[] = {(void*)&P};
  return cudaLaunchKernel((const void*)kernel, dim3(grid), dim3(kThreads), args, smem, stream);
}

template <int DT>
static cudaError_t launchRealD(int variant, RealParams& P, size_t smem, int dev, int smCount,
                               cudaStream_t stream) noexcept {
  switch (variant) {
#define X(id, tg, ps, mw, nw, nr, mb) \
  case id: return launchRealT<tg, ps, DT, mw, nw, nr, mb>(P, smem, dev, smCount, stream);
    GSDR_REAL_VARIANTS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}

// Real input x complex taps (gsdrFirCF: firTmaRealKernel with two tap planes):
// X(id, TG, PSPLIT, MIXW, NWIN, NRAW, MINB), PSPLIT even
#define GSDR_CF_VARIANTS(X) \
  X(0, 64, 2, 4, 1, 3, 2)   \
  X(1, 32, 2, 2, 2, 2, 4)   \
  X(2, 64, 2, 4, 2, 2, 2)   \
  X(3, 32, 4, 4, 1, 3, 2)

static constexpr RealVariant kCfVariants[] = {
#define X(id, tg, ps, mw, nw, nr, mb) {tg, ps, mw, nw, nr, mb},
    GSDR_CF_VARIANTS(X)
#undef X
};
static constexpr int kNumCfVariants = (int)(sizeof(kCfVariants) / sizeof(kCfVariants[0]));

template <int DT>
static cudaError_t launchCfD(int variant, RealParams& P, size_t smem, int dev, int smCount,
                             cudaStream_t stream) noexcept {
  switch (variant) {
#define X(id, tg, ps, mw, nw, nr, mb) \
  case id: return launchRealT<tg, ps, DT, mw, nw, nr, mb, 2>(P, smem, dev, smCount, stream);
    GSDR_CF_VARIANTS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}

// int8-input kernel (firTmaInt8Kernel): X(id, TG, PSPLIT, MIXW, MINB)
#define GSDR_INT8_VARIANTS(X) \
  X(0, 64, 2, 4, 2)           \
  X(1, 32, 4, 4, 1)           \
  X(2, 32, 2, 2, 4)

static constexpr SpecVariant kInt8Variants[] = {
#define X(id, tg, ps, mw, mb) {tg, ps, mw, mb},
    GSDR_INT8_VARIANTS(X)
#undef X
};
static constexpr int kNumInt8Variants = (int)(sizeof(kInt8Variants) / sizeof(kInt8Variants[0]));

template <int TG, int PSPLIT, int DT, int MIXW, bool NCO, int MINB>
static cudaError_t launchInt8T(Int8Params& P, size_t smem, int dev, int smCount, cudaStream_t stream) noexcept {
  static std::atomic<size_t> configured[64];
  auto kernel = firTmaInt8Kernel<TG, PSPLIT, DT, MIXW, NCO, MINB>;
  constexpr int kThreads = TG * PSPLIT + 32 * MIXW;
  if (smem > configured[dev & 63].load(std::memory_order_acquire)) {
    cudaError_t st = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (st != cudaSuccess) return report(st, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    configured[dev & 63].store(smem, std::memory_order_release);
  }
  int perSm = 0;
  static OccCache occ;
  cudaError_t st = occupancyCached(occ, kernel, kThreads, smem, dev, &perSm);
  if (st != cudaSuccess) return report(st, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
  if (perSm < 1) return cudaErrorInvalidConfiguration;
  {
    const int warpsPerCta = kThreads / 32;
    int balanced = perSm;
    while (balanced > 1 && (balanced * warpsPerCta) % 4 != 0) balanced--;
    if ((balanced * warpsPerCta) % 4 == 0) perSm = balanced;
  }
  const unsigned long long resident = (unsigned long long)perSm * (unsigned)smCount;
  const unsigned grid = (unsigned)(P.totalTiles < resident ? P.totalTiles : resident);
  P.strideChan = grid / P.tilesPerChannel;
  P.strideTile = grid % P.tilesPerChannel;
  void* args[] = {(void*)&P};
  return cudaLaunchKernel((const void*)kernel, dim3(grid), dim3(kThreads), args, smem, stream);
}

template <int DT, bool NCO>
static cudaError_t launchInt8D(int variant, Int8Params& P, size_t smem, int dev, int smCount,
                               cudaStream_t stream) noexcept {
  switch (variant) {
#define X(id, tg, ps, mw, mb) \
  case id: return launchInt8T<tg, ps, DT, mw, NCO, mb>(P, smem, dev, smCount, stream);
    GSDR_INT8_VARIANTS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}

// ---- per-decimation entry points (fir_inst_*.cu); mode is a PolyMode ----
#define GSDR_TMA_ARGS const CUtensorMap &map, TmaParams &P, size_t smem, int dev, int smCount, cudaStream_t stream
#define GSDR_REAL_ARGS RealParams &P, size_t smem, int dev, int smCount, cudaStream_t stream

#define GSDR_DECLARE_TMA_DT(DT) cudaError_t launchTmaDt##DT(int mode, int variant, GSDR_TMA_ARGS) noexcept;
#define GSDR_DECLARE_SPEC_DT(DT) cudaError_t launchSpecDt##DT(int variant, GSDR_TMA_ARGS) noexcept;
#define GSDR_DECLARE_CC_DT(DT) cudaError_t launchCcDt##DT(int variant, GSDR_TMA_ARGS) noexcept;
// K-shift channeliser (firTmaChannelizerKernel): one tile shape per compile-time decimation
struct ChanShape {
  int D, tg, psplit;
};
static constexpr ChanShape kChanShapes[] = {{4, 64, 1}, {8, 32, 2}, {10, 64, 1}};
cudaError_t launchChan(int D, const CUtensorMap& map, ChanParams& P, size_t smem, int dev, int smCount,
                       cudaStream_t stream) noexcept;
#define GSDR_DECLARE_WIDE_DT(DT) cudaError_t launchWideDt##DT(int mode, int variant, GSDR_TMA_ARGS) noexcept;
#define GSDR_DEFINE_WIDE_DT(DT)                                                  \
  cudaError_t launchWideDt##DT(int mode, int variant, GSDR_TMA_ARGS) noexcept {  \
    return launchWideD<DT>(mode, variant, map, P, smem, dev, smCount, stream);   \
  }
GSDR_DECLARE_WIDE_DT(32)
#define GSDR_DECLARE_REAL_DT(DT) cudaError_t launchRealDt##DT(int variant, GSDR_REAL_ARGS) noexcept;
GSDR_DECLARE_TMA_DT(0) GSDR_DECLARE_TMA_DT(4) GSDR_DECLARE_TMA_DT(8) GSDR_DECLARE_TMA_DT(10) GSDR_DECLARE_TMA_DT(32)
GSDR_DECLARE_SPEC_DT(0) GSDR_DECLARE_SPEC_DT(8) GSDR_DECLARE_SPEC_DT(10) GSDR_DECLARE_SPEC_DT(32)
GSDR_DECLARE_CC_DT(0) GSDR_DECLARE_CC_DT(8)
GSDR_DECLARE_REAL_DT(0) GSDR_DECLARE_REAL_DT(2) GSDR_DECLARE_REAL_DT(10)
#define GSDR_DECLARE_CF_DT(DT) cudaError_t launchCfDt##DT(int variant, GSDR_REAL_ARGS) noexcept;
GSDR_DECLARE_CF_DT(0) GSDR_DECLARE_CF_DT(2) GSDR_DECLARE_CF_DT(10)
#define GSDR_DEFINE_CF_DT(DT)                                                  \
  cudaError_t launchCfDt##DT(int variant, GSDR_REAL_ARGS) noexcept {           \
    return launchCfD<DT>(variant, P, smem, dev, smCount, stream);              \
  }
#define GSDR_INT8_ARGS Int8Params &P, size_t smem, int dev, int smCount, cudaStream_t stream
#define GSDR_DECLARE_INT8_DT(DT) cudaError_t launchInt8Dt##DT(bool nco, int variant, GSDR_INT8_ARGS) noexcept;
GSDR_DECLARE_INT8_DT(0) GSDR_DECLARE_INT8_DT(8) GSDR_DECLARE_INT8_DT(10)
#define GSDR_DEFINE_INT8_DT(DT)                                                             \
  cudaError_t launchInt8Dt##DT(bool nco, int variant, GSDR_INT8_ARGS) noexcept {            \
    return nco ? launchInt8D<DT, true>(variant, P, smem, dev, smCount, stream)              \
               : launchInt8D<DT, false>(variant, P, smem, dev, smCount, stream);            \
  }

#define GSDR_DEFINE_TMA_DT(DT)                                                                                  \
  cudaError_t launchTmaDt##DT(int mode, int variant, GSDR_TMA_ARGS) noexcept {                                  \
    switch (mode) {                                                                                             \
      case kPolyFC: return launchTmaModeD<kPolyFC, DT>(variant, map, P, smem, dev, smCount, stream);            \
      case kPolyNcoExact: return launchTmaModeD<kPolyNcoExact, DT>(variant, map, P, smem, dev, smCount, stream); \
      case kPolyNcoLiteral:                                                                                     \
        return launchTmaModeD<kPolyNcoLiteral, DT>(variant, map, P, smem, dev, smCount, stream);                \
      default: return cudaErrorInvalidValue;                                                                    \
    }                                                                                                           \
  }
#define GSDR_DEFINE_SPEC_DT(DT)                                                  \
  cudaError_t launchSpecDt##DT(int variant, GSDR_TMA_ARGS) noexcept {            \
    return launchSpecD<DT>(variant, map, P, smem, dev, smCount, stream);         \
  }
#define GSDR_DEFINE_CC_DT(DT)                                                    \
  cudaError_t launchCcDt##DT(int variant, GSDR_TMA_ARGS) noexcept {              \
    return launchCcD<DT>(variant, map, P, smem, dev, smCount, stream);           \
  }
#define GSDR_DEFINE_REAL_DT(DT)                                                  \
  cudaError_t launchRealDt##DT(int variant, GSDR_REAL_ARGS) noexcept {           \
    return launchRealD<DT>(variant, P, smem, dev, smCount, stream);              \
  }

}  // namespace gsdr_b200
