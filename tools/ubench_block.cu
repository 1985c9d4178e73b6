// ubench_block.cu — measurement tool: throughput of the production steady block (firPairBlock<kBlkSteady>) in
// isolation: no TMA, no tiles, no barriers — just the 128 FFMA2 + 8 sample LDS.128 + 4 tap LDS.128 of the inner
// loop running back to back on valid shared memory.  Tells how much of the kernel's gap to the FP32 peak is the
// loop body itself.
#include <cuda_runtime.h>
#include <cstdio>
#include "fir_tma_kernel.cuh"

using namespace gsdr_b200;

template <int WITH_LDS>
__global__ void __launch_bounds__(128) k_block(float* out, int iters) {
  extern __shared__ __align__(1024) unsigned char sm[];
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 1e-3f * (i & 255);
  __syncthreads();
  float2 acc[8];
  float4 q[8];
  float hAP[8], hAQ[8], hBP[8], hBQ[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    acc[i] = make_float2(0.f, 0.f);
    q[i] = reinterpret_cast<const float4*>(sm)[threadIdx.x + 128 * i];
    hAP[i] = sm[i]; hAQ[i] = sm[8 + i]; hBP[i] = sm[16 + i]; hBQ[i] = sm[24 + i];
  }
  const unsigned planeBytes = 4096;  // 8 planes x 4 KB = 32 KB of "window"
  const unsigned char* a0 = sm + threadIdx.x * 16;
  const unsigned char* a1 = a0 + 2048;
  const float* taps = reinterpret_cast<const float*>(sm + 40 * 1024);
  for (int it = 0; it < iters; it++) {
    if (WITH_LDS) {
      firPairBlock<kBlkSteady, true>(acc, q, hAP, hAQ, hBP, hBQ, a0, a1, planeBytes, taps);
      firPairBlock<kBlkSteady, true>(acc, q, hBP, hBQ, hAP, hAQ, a1, a0, planeBytes, taps + 16);
    } else {
      // same FFMA2s, operands stay in registers
#pragma unroll
      for (int rep = 0; rep < 2; rep++)
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const float2 xP = make_float2(q[i].x, q[i].y), xQ = make_float2(q[i].z, q[i].w);
#pragma unroll
          for (int r = 0; r < 8; r++) acc[r] = macTap(xP, r <= i ? hBP[i - (r <= i ? r : 0)] : hAP[8 + i - r], acc[r]);
#pragma unroll
          for (int r = 0; r < 8; r++) acc[r] = macTap(xQ, r <= i ? hBQ[i - (r <= i ? r : 0)] : hAQ[8 + i - r], acc[r]);
        }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; i++) s += acc[i].x + acc[i].y;
  if (s == 123.456f) out[0] = s;
}

// The whole per-tile FIR (prologue, steady, tail blocks, 4 branch pairs, D = 8, 32 taps per branch) on a static
// window: what the kernel would do with free copies, no barriers and no stores.
__global__ void __launch_bounds__(128) k_tile(float* out, int iters, TmaParams P) {
  extern __shared__ __align__(1024) unsigned char sm[];
  const unsigned planeBytes = tmaPlaneRows(128, kTmaJpadCap, 8) * 64u;  // 144 rows x 64 B
  for (unsigned i = threadIdx.x; i < (8 * planeBytes + 4096) / 4; i += blockDim.x)
    reinterpret_cast<float*>(sm)[i] = 1e-3f * (i & 255);
  __syncthreads();
  const float* hs = reinterpret_cast<const float*>(sm + 8 * planeBytes);
  float s = 0.f;
  for (int it = 0; it < iters; it++) {
    float2 acc[8];
#pragma unroll
    for (int r = 0; r < 8; r++) acc[r] = make_float2(0.f, 0.f);
    firComputePairs<8>(acc, sm, hs, threadIdx.x, 0, 4, 32, planeBytes, P);
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i].x + acc[i].y;
  }
  if (s == 123.456f) out[0] = s;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  cudaMalloc(&out, 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaFuncSetAttribute(k_block<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  cudaFuncSetAttribute(k_block<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  const int iters = 4000;
  for (int mode = 0; mode < 2; mode++) {
    for (int bps = 1; bps <= 4; bps++) {
      float best = 1e9f;
      for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        if (mode) k_block<1><<<sms * bps, 128, 48 * 1024>>>(out, iters);
        else k_block<0><<<sms * bps, 128, 48 * 1024>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
      }
      const double flops = 4.0 * 256.0 * iters * 128.0 * sms * bps;
      printf("%s, %d x 128-thread CTAs/SM: %.2f TFLOP/s\n", mode ? "steady block with LDS (production code)" : "same FFMA2, registers only",
             bps, flops / best / 1e9);
    }
  }
  {
    const size_t smem = 8 * tmaPlaneRows(128, kTmaJpadCap, 8) * 64 + 4096;
    cudaFuncSetAttribute(k_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    TmaParams P{};
    for (int bps = 1; bps <= 2; bps++) {
      float best = 1e9f;
      const int it2 = 200;
      for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k_tile<<<sms * bps, 128, smem>>>(out, it2, P);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
      }
      const double flops = 4.0 * 8.0 * 256.0 * it2 * 128.0 * sms * bps;  // 8 outputs x 256 (padded) taps per thread
      printf("whole per-tile FIR (4 pairs, P+3S+T each), %d x 128-thread CTAs/SM: %.2f TFLOP/s\n", bps, flops / best / 1e9);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
