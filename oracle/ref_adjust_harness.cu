/*
 * TEST INFRASTRUCTURE ONLY — launcher around the reference's internal __device__ function
 * k_AdjustFrequency (ref: src/adjustFrequency.cuh:27-33), which has no public entry point of its own.
 * It reproduces the call shape of the reference's only in-tree callers (ref: src/fm.cu:43-56,
 * src/am.cu:34-47): one thread per decimated output, firstSampleIndex + decimation*outputIndex.
 *
 * Linked (-rdc) against a scratch copy of the reference's adjustFrequency.cu that has the missing
 * `return sample;` added (ref: src/adjustFrequency.cu:55-56 — without it the function is undefined behaviour and
 * nvcc reduces it to a bare RET).  The scratch copy lives in a temp dir and is never stored in this repository;
 * see oracle/build_ref.sh.
 */
#include <cuComplex.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "adjustFrequency.cuh" /* from /root/reference/src via -I */

__global__ void k_refAdjustFrequencyFir(
    float frequencyShift, uint32_t firstSampleIndex, float sampleRate, uint32_t decimation, const cuComplex* input,
    const float* taps, uint32_t numTaps, cuComplex* output, uint32_t numOutputs) {
  const uint32_t o = blockDim.x * blockIdx.x + threadIdx.x;
  if (o >= numOutputs) {
    return;
  }
  const uint32_t first = decimation * o;
  output[o] = k_AdjustFrequency(frequencyShift, firstSampleIndex + first, sampleRate, input + first, taps, numTaps);
}

extern "C" cudaError_t refAdjustFrequencyFirFC(
    float sampleRate, float frequencyShift, size_t firstSampleIndex, size_t decimation, const float* taps,
    size_t tapCount, const cuComplex* input, cuComplex* output, size_t numOutputs, int32_t cudaDevice,
    cudaStream_t cudaStream) {
  int prev = 0;
  cudaError_t st = cudaGetDevice(&prev);
  if (st != cudaSuccess) return st;
  st = cudaSetDevice(cudaDevice);
  if (st != cudaSuccess) return st;
  /* same host-side pre-reduction as the reference's callers, ref: src/fm.cu:202 */
  const uint32_t first32 = (uint32_t)fmodf((float)firstSampleIndex, sampleRate);
  const unsigned blocks = (unsigned)((numOutputs + 31) / 32);
  if (blocks) {
    k_refAdjustFrequencyFir<<<blocks, 32, 0, cudaStream>>>(
        frequencyShift, first32, sampleRate, (uint32_t)decimation, input, taps, (uint32_t)tapCount, output,
        (uint32_t)numOutputs);
  }
  st = cudaGetLastError();
  cudaSetDevice(prev);
  return st;
}
