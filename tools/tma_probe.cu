// tma_probe.cu — measurement tool: how fast can TMA (cp.async.bulk.tensor) stage the FIR sample window in the
// de-interleaved layouts the polyphase kernel wants, compared with a plain 1-D bulk copy of the same bytes?
//
// Input: N cuComplex samples viewed as rows of D samples (row m = samples m*D .. m*D+D-1).
// Layout A ("pair planes"): per phase pair pp, smem[ml][mh][16 B] with m = 8*mh + ml — box (4 floats, MH, 8) of a
//   4-D tensor (2D floats | mh stride 8*D*8 B | ml stride D*8 B), one TMA op per phase pair per tile.
// Layout B ("rows"): smem[ml][mh][D*8 B] — box (2D floats, MH, 8): whole rows, one TMA op per tile.
// Layout C: 1-D cp.async.bulk of the contiguous window (no de-interleave), the upper bound.
// Each persistent CTA double-buffers tiles; compute is a token read so the copy engine is the only limit.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ unsigned smemAddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count));
}
__device__ __forceinline__ void mbarExpectTx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smemAddr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmaLoad3(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smemAddr(dst)), "l"(map), "r"(smemAddr(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulkLoad1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smemAddr(dst)), "l"(src), "r"(bytes), "r"(smemAddr(bar)) : "memory");
}

// mode 0: pair planes (D/2 ops per tile), 1: whole rows (1 op), 2: 1-D bulk.
template <int MODE>
__global__ void __launch_bounds__(128) k_probe(const __grid_constant__ CUtensorMap map, const float2* x, float* out,
                                               int D, int MH, int rowsPerTile, int numTiles) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar[2];
  const unsigned tileBytes = (unsigned)MH * 8u * (unsigned)D * 8u;
  if (threadIdx.x == 0) {
    mbarInit(&bar[0], 1);
    mbarInit(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int tile, int buf) {
    unsigned char* dst = smem + (size_t)buf * tileBytes;
    mbarExpectTx(&bar[buf], tileBytes);
    const int mh0 = tile * (rowsPerTile / 8);
    if (MODE == 0) {
      for (int pp = 0; pp < D / 2; pp++) tmaLoad3(dst + (size_t)pp * MH * 8 * 16, &map, &bar[buf], pp * 4, mh0, 0);
    } else if (MODE == 1) {
      tmaLoad3(dst, &map, &bar[buf], 0, mh0, 0);
    } else {
      bulkLoad1d(dst, x + (size_t)tile * rowsPerTile * D, tileBytes, &bar[buf]);
    }
  };
  int tile = blockIdx.x;
  if (threadIdx.x == 0 && tile < numTiles) issue(tile, 0);
  float acc = 0.f;
  for (int it = 0; tile < numTiles; tile += gridDim.x, it++) {
    const int buf = it & 1;
    if (threadIdx.x == 0 && tile + (int)gridDim.x < numTiles) issue(tile + gridDim.x, buf ^ 1);
    mbarWait(&bar[buf], (it >> 1) & 1);
    acc += reinterpret_cast<const float*>(smem + (size_t)buf * tileBytes)[threadIdx.x * 4];
    __syncthreads();
  }
  if (acc == 123.456f) out[0] = acc;
}

int main(int argc, char** argv) {
  const int D = argc > 1 ? atoi(argv[1]) : 8;
  const size_t N = (size_t)1 << 26;
  const int rowsPerTile = 1024;  // outputs per tile (one row per output), window overlap ignored here
  const int MH = rowsPerTile / 8;
  float2* x;
  float* out;
  CK(cudaMalloc(&x, N * 8 + 65536));
  CK(cudaMemset(x, 0, N * 8 + 65536));
  CK(cudaMalloc(&out, 4));
  const size_t rows = N / D;
  const int numTiles = (int)(rows / rowsPerTile);

  EncodeTiledFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }

  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t smemBytes = 2 * (size_t)MH * 8 * D * 8;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));

  for (int mode = 0; mode < 3; mode++) {
    for (int promo = 0; promo < (mode == 2 ? 1 : 3); promo++) {
      CUtensorMap map;
      const cuuint64_t gdim[3] = {(cuuint64_t)2 * D, (cuuint64_t)rows / 8, 8};
      const cuuint64_t gstride[2] = {(cuuint64_t)8 * D * 8, (cuuint64_t)D * 8};  // bytes, dims 1..2
      const cuuint32_t box[3] = {(cuuint32_t)(mode == 0 ? 4 : 2 * D), (cuuint32_t)MH, 8};
      const cuuint32_t estr[3] = {1, 1, 1};
      const CUtensorMapL2promotion pr = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                      : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                   : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
      CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, x, gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d (mode %d)\n", (int)r, mode); continue; }
      for (int perSm = 1; perSm <= 2; perSm++) {
        if (perSm * smemBytes > 220 * 1024) continue;
        const int grid = sms * perSm;
        float best = 1e9f;
        for (int rep = 0; rep < 6; rep++) {
          CK(cudaEventRecord(e0));
          if (mode == 0) {
            CK(cudaFuncSetAttribute(k_probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
            k_probe<0><<<grid, 128, smemBytes>>>(map, x, out, D, MH, rowsPerTile, numTiles);
          } else if (mode == 1) {
            CK(cudaFuncSetAttribute(k_probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
            k_probe<1><<<grid, 128, smemBytes>>>(map, x, out, D, MH, rowsPerTile, numTiles);
          } else {
            CK(cudaFuncSetAttribute(k_probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
            k_probe<2><<<grid, 128, smemBytes>>>(map, x, out, D, MH, rowsPerTile, numTiles);
          }
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          CK(cudaGetLastError());
          float ms;
          CK(cudaEventElapsedTime(&ms, e0, e1));
          if (rep > 0 && ms < best) best = ms;
        }
        const double bytes = (double)numTiles * rowsPerTile * D * 8;
        printf("mode %d (%s) l2promo %d ctas/SM %d: %.3f ms  %.1f GB/s\n", mode,
               mode == 0 ? "pair planes, 16B inner" : mode == 1 ? "whole rows" : "1-D bulk", promo, perSm, best,
               bytes / best / 1e6);
      }
    }
  }
  return 0;
}
