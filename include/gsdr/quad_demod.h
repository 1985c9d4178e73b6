/*
 * gsdr/quad_demod.h — quadrature demodulators, C ABI.
 * Source-compatible with the reference's include/gsdr/quad_demod.h (ref: include/gsdr/quad_demod.h:25-45); replaces
 * the kernels and wrappers at ref: src/quad_demod.cu:23-74.  Same in-stream / device / error conventions as
 * <gsdr/fir.h>; 64-bit indexing; numOutputElements == 0 returns cudaSuccess without a launch.
 *
 *   gsdrQuadFmDemod: output[i] = gain * atan2f(Im(m), Re(m)),  m = input[i+1] * conj(input[i])
 *                    (input must hold numOutputElements + 1 elements).  Bit-identical to the reference kernel.
 *   gsdrQuadAmDemod: output[i] = 2 * saturate(hypotf(re, im)) - 1.
 */
#ifndef GSDR_B200_INCLUDE_GSDR_QUAD_DEMOD_H_
#define GSDR_B200_INCLUDE_GSDR_QUAD_DEMOD_H_

#include <cuComplex.h>
#include <cuda_runtime.h>
#include <gsdr/gsdr_export.h>
#include <gsdr/util.h>
#include <stddef.h>
#include <stdint.h>

/* Replaces ref: src/quad_demod.cu:56-66.  gain = sampleRate / (2*pi*deviation) for FM (ref: src/fm.cu:203). */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrQuadFmDemod(
    const cuComplex* input,
    float* output,
    float gain,
    size_t numOutputElements,
    int32_t cudaDevice,
    cudaStream_t cudaStream) GSDR_NO_EXCEPT;

/* Replaces ref: src/quad_demod.cu:68-74. */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrQuadAmDemod(
    const cuComplex* input,
    float* output,
    size_t numOutputElements,
    int32_t cudaDevice,
    cudaStream_t cudaStream) GSDR_NO_EXCEPT;

#endif /* GSDR_B200_INCLUDE_GSDR_QUAD_DEMOD_H_ */
