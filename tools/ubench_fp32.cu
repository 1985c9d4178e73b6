// ubench_fp32.cu — measurement tool (not part of the product library): peak FP32 FMA throughput of the device
// with scalar FFMA and with packed FFMA2 (fma.rn.f32x2), the denominators of the FP32 side of the FIR roofline
// (MEASURED_PEAKS.json holds only HBM and bf16 tensor peaks).  Dependent-free: 16 independent accumulators per
// thread, operands in registers, no memory traffic inside the loop.
#include <cuda_runtime.h>
#include <cstdio>

template <int PACKED>
__global__ void __launch_bounds__(256) k_fma(float* out, int iters, float a, float b) {
  float2 acc[16];
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = make_float2(threadIdx.x * 1e-3f + i, blockIdx.x * 1e-3f - i);
  const float2 m = make_float2(a, a);
  const float2 c = make_float2(b, b);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int i = 0; i < 16; i++) {
        if (PACKED) {
          acc[i] = __ffma2_rn(acc[i], m, c);
        } else {
          acc[i].x = __fmaf_rn(acc[i].x, a, b);
          acc[i].y = __fmaf_rn(acc[i].y, a, b);
        }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; i++) s += acc[i].x + acc[i].y;
  if (s == 123.456f) out[0] = s;  // never true; keeps the loop alive
}

// Returns TFLOP/s (2 flops per FMA lane) or a negative cudaError_t.
extern "C" double ubenchFp32Tflops(int packed, int device, int iters, int reps, int blocksPerSm) {
  int prev = 0, sms = 0;
  if (cudaGetDevice(&prev) != cudaSuccess) return -1.0;
  if (cudaSetDevice(device) != cudaSuccess) return -2.0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  float* out = nullptr;
  cudaMalloc(&out, 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = sms * blocksPerSm;
  double best = 0.0;
  for (int r = 0; r < reps + 1; r++) {
    cudaEventRecord(e0);
    if (packed) k_fma<1><<<grid, 256>>>(out, iters, 0.999f, 1e-3f);
    else k_fma<0><<<grid, 256>>>(out, iters, 0.999f, 1e-3f);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { best = -3.0; break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 2.0 * 16.0 * 4.0 * (double)iters * 256.0 * (double)grid;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (r > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  cudaSetDevice(prev);
  return best;
}

#ifdef UBENCH_MAIN
int main() {
  for (int bps = 1; bps <= 8; bps *= 2) {
    printf("blocks/SM %d: FFMA %.2f TFLOP/s, FFMA2 %.2f TFLOP/s\n", bps, ubenchFp32Tflops(0, 0, 20000, 3, bps),
           ubenchFp32Tflops(1, 0, 20000, 3, bps));
  }
  return 0;
}
#endif
