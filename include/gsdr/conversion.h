/*
 * gsdr/conversion.h — int8 sample input, C ABI.
 *
 * gsdrInt8ToNormFloat is source-compatible with ref: include/gsdr/conversion.h:24-35 and replaces ref:
 * src/conversion.cu:20-35: output = max(-1, input / 127) (-128 and -127 give -1, 127 gives 1), bit-identical to the
 * reference kernel; unlike it, element numElements is not written (ref: src/conversion.cu:22 tests x > numElements).
 *
 * NEW: gsdrFirFCInt8 / gsdrAdjustFrequencyFirFCInt8 take the receiver's int8 IQ stream directly —
 * interleaved I, Q, 2 bytes per complex sample — and give the result of gsdrInt8ToNormFloat followed by gsdrFirFC /
 * gsdrAdjustFrequencyFirFC, without the float copy of the input ever existing: a quarter of the HBM (and PCIe) bytes
 * per sample.  The division by 127 is folded into the taps, so the result matches convert-then-filter within the FP32
 * tolerance of <gsdr/fir.h> (max |err| <= 1e-5 * sum|taps| * max|input|), not bit for bit.  Same in-stream / device /
 * error conventions as <gsdr/fir.h>.
 */
#ifndef GSDR_B200_INCLUDE_GSDR_CONVERSION_H_
#define GSDR_B200_INCLUDE_GSDR_CONVERSION_H_

#include <cuComplex.h>
#include <cuda_runtime.h>
#include <gsdr/gsdr_export.h>
#include <gsdr/util.h>
#include <stddef.h>
#include <stdint.h>

/* Replaces ref: src/conversion.cu:29-35. */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrInt8ToNormFloat(
    const int8_t* input,
    float* output,
    size_t numElements,
    int32_t cudaDevice,
    cudaStream_t cudaStream) GSDR_NO_EXCEPT;

/* input: (numOutputs - 1) * decimation + tapCount complex samples = twice as many int8 (I0, Q0, I1, Q1, ...). */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrFirFCInt8(
    size_t decimation,
    const float* taps,
    size_t tapCount,
    const int8_t* input,
    cuComplex* output,
    size_t numOutputs,
    int32_t cudaDevice,
    cudaStream_t cudaStream) GSDR_NO_EXCEPT;

/* As gsdrAdjustFrequencyFirFC (<gsdr/adjust_frequency.h>: exact 64-bit-phase NCO) on int8 IQ input. */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrAdjustFrequencyFirFCInt8(
    float sampleRate,
    float frequencyShift,
    size_t firstSampleIndex,
    size_t decimation,
    const float* taps,
    size_t tapCount,
    const int8_t* input,
    cuComplex* output,
    size_t numOutputs,
    int32_t cudaDevice,
    cudaStream_t cudaStream) GSDR_NO_EXCEPT;

#endif /* GSDR_B200_INCLUDE_GSDR_CONVERSION_H_ */
