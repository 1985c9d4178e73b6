// fir_inst_tc.cu — kernel instantiations and launcher of the tensor-core FIR (fir_tc_kernel.cuh).
// (DESIGN.md §4.3b holds the measurements and the decision where it is used.)
#include <atomic>
#include <cstdio>
#include <cstdlib>

#include "fir_launch.cuh"
#include "fir_tc_kernel.cuh"

namespace gsdr_b200 {

template <int D, int MINB>
static cudaError_t launchTcT(TcParams& P, size_t smem, int dev, int smCount, cudaStream_t stream) noexcept {
  static std::atomic<size_t> configured[64];
  static std::atomic<int> perSmCache[64];
  auto kernel = firTcKernel<D, MINB>;
  if (smem > configured[dev & 63].load(std::memory_order_acquire)) {
    cudaError_t st = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (st != cudaSuccess) return report(st, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    configured[dev & 63].store(smem, std::memory_order_release);
    perSmCache[dev & 63].store(0, std::memory_order_release);
  }
  int perSm = perSmCache[dev & 63].load(std::memory_order_acquire);
  if (perSm <= 0) {
    cudaError_t st = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, kTcThreads, configured[dev & 63].load());
    if (st != cudaSuccess) return report(st, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if (perSm < 1) return cudaErrorInvalidConfiguration;
    // the occupancy query has been seen to answer 1 for this kernel; shared memory is what really limits it
    int smemPerSm = 0;
    if (cudaDeviceGetAttribute(&smemPerSm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev) == cudaSuccess) {
      const int bySmem = (int)((size_t)smemPerSm / (smem + 1024 + 256));
      if (bySmem > perSm) perSm = bySmem;
    }
    // every CTA owns kTcTmemCols (+ the ring's second allocation at 64 outputs per window) of the SM's 512 columns
    const int tmemLimit = 512 / (int)(kTcTmemCols + tcTmemCols2(D));
    if (perSm > tmemLimit) perSm = tmemLimit;
    if (perSm > MINB) perSm = MINB;
    perSmCache[dev & 63].store(perSm, std::memory_order_release);
  }
  const unsigned long long resident = (unsigned long long)perSm * (unsigned)smCount;
  const unsigned grid = (unsigned)(P.totalTiles < resident ? P.totalTiles : resident);
  void* args[] = {(void*)&P};
  return cudaLaunchKernel((const void*)kernel, dim3(grid), dim3(kTcThreads), args, smem, stream);
}

size_t tcSharedBytes(unsigned D, unsigned tablePitch) noexcept {
  const size_t raw = D == 4 ? TcGeom<4>::rawBytes : D == 8 ? TcGeom<8>::rawBytes : TcGeom<16>::rawBytes;
  return raw + tcTableBytes(tablePitch);
}

cudaError_t launchTc(unsigned D, TcParams& P, int dev, int smCount, cudaStream_t stream) noexcept {
  const size_t smem = tcSharedBytes(D, P.tablePitch);
  switch (D) {
    case 4: return launchTcT<4, 2>(P, smem, dev, smCount, stream);
    case 8: return launchTcT<8, 3>(P, smem, dev, smCount, stream);
    case 16: return launchTcT<16, 1>(P, smem, dev, smCount, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace gsdr_b200
