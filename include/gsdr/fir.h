/*
 * gsdr/fir.h — decimating FIR family, C ABI.
 *
 * Source-compatible replacement for the reference's include/gsdr/fir.h: the four entry points below have the
 * same names, argument order and types as ref: include/gsdr/fir.h:30-68, and replace the host wrappers at
 * ref: src/fir.cu:73-96 (FC), :98-121 (FF), :123-146 (CC), :148-171 (CF).
 *
 * Naming is <TapType><InputType>: F = float, C = cuComplex.  Every variant computes, for n in [0, numOutputs):
 *
 *     output[n] = sum_{i=0}^{tapCount-1} input[n * decimation + i] * taps[i]
 *
 * i.e. a correlation with the taps in the order given (the caller pre-reverses them for a convolution), keeping
 * every decimation-th result (ref: src/fir.cu:57-70).
 *
 * Contract (same as the reference unless marked NEW):
 *  - taps, input and output are DEVICE pointers valid on cudaDevice.  The caller owns them and must keep them
 *    alive and unchanged until cudaStream has passed the enqueued work.
 *  - input must hold at least (numOutputs - 1) * decimation + tapCount elements.
 *  - Work is enqueued on cudaStream only; the call does not synchronize, allocate or use other streams, and can
 *    be captured into a CUDA graph.
 *  - The calling thread's current device is saved and restored.
 *  - Returns cudaSuccess or the CUDA error of the first failing runtime call.
 *  - NEW: output must not alias input or taps.
 *  - NEW: indices are 64-bit (the reference narrows to uint32_t, ref: src/fir.cu:30-32,53-58); results are
 *    identical wherever the reference is defined.
 *  - NEW: numOutputs == 0 returns cudaSuccess without launching; decimation == 0 returns cudaErrorInvalidValue;
 *    launch errors are reported (the reference never calls cudaGetLastError).
 *  - Floating point: FP32 FMA accumulation in a different order than the reference's single ascending chain;
 *    max |err| <= 1e-5 * sum|taps| * max|input| against it.
 *  - NEW: the staged kernels pad the tap set with zeros up to a multiple of 8 or 16 taps per polyphase branch, so an
 *    Inf or NaN input sample can reach outputs up to 16 * decimation samples earlier than the reference's window
 *    would (0 * Inf = NaN); finite inputs are unaffected.
 *  - NEW: gsdrFirFC with decimation 8 and 129..264 taps (or decimation 4 and 65..260, or 16 and 257..528 taps), >= 65536 outputs and a
 *    16-byte aligned input runs on the tensor
 *    cores (FP16 operands with error compensation, FP32 accumulation): the same error bound (measured ~1e-6 of
 *    sum|taps| * max|input|), but an Inf or NaN sample there reaches every output of the up to three windows (32 outputs;
 *    64 at decimation 4) that read its segment of 256 (decimation 16: 512) samples.  gsdrB200SetFirTensorCores(0) (gsdr/b200.h) keeps every call on the FMA kernels.
 */
#ifndef GSDR_B200_INCLUDE_GSDR_FIR_H_
#define GSDR_B200_INCLUDE_GSDR_FIR_H_

#include <cuComplex.h>
#include <cuda_runtime.h>
#include <gsdr/gsdr_export.h>
#include <gsdr/util.h>
#include <stddef.h>
#include <stdint.h>

/* float taps, cuComplex input -> cuComplex output.  Replaces ref: src/fir.cu:73-96. */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrFirFC(
    size_t decimation,
    const float* taps,
    size_t tapCount,
    const cuComplex* input,
    cuComplex* output,
    size_t numOutputs,
    int32_t cudaDevice,
    cudaStream_t cudaStream) GSDR_NO_EXCEPT;

/* float taps, float input -> float output.  Replaces ref: src/fir.cu:98-121. */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrFirFF(
    size_t decimation,
    const float* taps,
    size_t tapCount,
    const float* input,
    float* output,
    size_t numOutputs,
    int32_t cudaDevice,
    cudaStream_t cudaStream) GSDR_NO_EXCEPT;

/* cuComplex taps, cuComplex input -> cuComplex output.  Replaces ref: src/fir.cu:123-146. */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrFirCC(
    size_t decimation,
    const cuComplex* taps,
    size_t tapCount,
    const cuComplex* input,
    cuComplex* output,
    size_t numOutputs,
    int32_t cudaDevice,
    cudaStream_t cudaStream) GSDR_NO_EXCEPT;

/* cuComplex taps, float input -> cuComplex output.  Replaces ref: src/fir.cu:148-171. */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrFirCF(
    size_t decimation,
    const cuComplex* taps,
    size_t tapCount,
    const float* input,
    cuComplex* output,
    size_t numOutputs,
    int32_t cudaDevice,
    cudaStream_t cudaStream) GSDR_NO_EXCEPT;

#endif /* GSDR_B200_INCLUDE_GSDR_FIR_H_ */
