/*
 * gsdr/adjust_frequency.h — NCO mix-down fused with the decimating real-tap FIR (channel select), C ABI.
 *
 * NEW, additive header.  The reference has no public adjustFrequency symbol: its NCO is the internal
 * __device__ function k_AdjustFrequency (ref: src/adjustFrequency.cuh:27-33, src/adjustFrequency.cu:25-56),
 * reachable only through gsdrFmDemod (ref: src/fm.cu:46-56) and gsdrAmDemod (ref: src/am.cu:34-47).  The entry
 * points below expose exactly that stage — "mix each input sample by the NCO, low-pass with real taps, keep
 * every decimation-th result" — with the reference's argument tail
 * (decimation, taps, tapCount, input, output, numOutputs, cudaDevice, cudaStream) and the same in-stream
 * semantics as <gsdr/fir.h>:
 *
 *     output[n] = sum_{i<tapCount} input[n*D + i] * nco(firstSampleIndex + n*D + i) * taps[i]
 *
 * The mixed signal is produced on chip while the sample window is staged into shared memory; it never exists
 * in HBM.  frequencyShift is what the reference's callers pass: tuningFrequency - channelFrequency
 * (ref: src/fm.cu:204).
 *
 * Two phase laws are provided:
 *
 *  gsdrAdjustFrequencyFirFC         nco(n) = exp(j*2*pi*frequencyShift*n/sampleRate), generated from a 64-bit
 *                                   fixed-point phase accumulator: phase(n) = (n * step) mod 2^64 with
 *                                   step = round(frac(frequencyShift/sampleRate) * 2^64).  Exact in n (no drift,
 *                                   any 64-bit firstSampleIndex), so time shards reproduce the unsharded result.
 *                                   This is what the reference documents the stage to do (ref: include/gsdr/fm.h:25-41).
 *
 *  gsdrAdjustFrequencyFirFCLiteral  the reference's per-tap arithmetic taken literally, bugs included
 *                                   (ref: src/adjustFrequency.cu:23,35-50): firstSampleIndex is first reduced to
 *                                   (uint32_t)fmodf((float)firstSampleIndex, sampleRate) (ref: src/fm.cu:202), the
 *                                   sample index is a wrapping uint32_t, and the angle is
 *                                   2*pi*fmodf(fmodf((float)idx, fs)/fs, 1/f) — a time in seconds, not a fraction
 *                                   of a period.  Provided so results can be compared against the reference's own
 *                                   kernel; not physically meaningful.
 */
#ifndef GSDR_B200_INCLUDE_GSDR_ADJUST_FREQUENCY_H_
#define GSDR_B200_INCLUDE_GSDR_ADJUST_FREQUENCY_H_

#include <cuComplex.h>
#include <cuda_runtime.h>
#include <gsdr/gsdr_export.h>
#include <gsdr/util.h>
#include <stddef.h>
#include <stdint.h>

/* Replaces the k_AdjustFrequency call made per output thread at ref: src/fm.cu:46-56 / src/am.cu:41-47. */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrAdjustFrequencyFirFC(
    float sampleRate,
    float frequencyShift,
    size_t firstSampleIndex,
    size_t decimation,
    const float* taps,
    size_t tapCount,
    const cuComplex* input,
    cuComplex* output,
    size_t numOutputs,
    int32_t cudaDevice,
    cudaStream_t cudaStream) GSDR_NO_EXCEPT;

/* Bug-compatible phase law of ref: src/adjustFrequency.cu:35-50 (see above). */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrAdjustFrequencyFirFCLiteral(
    float sampleRate,
    float frequencyShift,
    size_t firstSampleIndex,
    size_t decimation,
    const float* taps,
    size_t tapCount,
    const cuComplex* input,
    cuComplex* output,
    size_t numOutputs,
    int32_t cudaDevice,
    cudaStream_t cudaStream) GSDR_NO_EXCEPT;

/*
 * The 64-bit phase increment per sample used by gsdrAdjustFrequencyFirFC (host arithmetic, no CUDA call):
 * round(frac(frequencyShift / sampleRate) * 2^64) with frac() in [-0.5, 0.5).  Exposed so callers and the
 * sharding layer can reason about phase continuity; bit-exact by definition.
 */
GSDR_C_LINKAGE GSDR_PUBLIC uint64_t gsdrNcoPhaseStep(float frequencyShift, float sampleRate) GSDR_NO_EXCEPT;

/*
 * One input, numShifts channels (the idea of the reference's dead k_Fm4x, ref: src/fm.cu:71-179):
 *     output[k * outputStride + n] = gsdrAdjustFrequencyFirFC(sampleRate, frequencyShifts[k], ...)[n]
 * frequencyShifts is a HOST array; everything else as in gsdrAdjustFrequencyFirFC.  One fused NCO + FIR launch per
 * shift is enqueued on cudaStream (capturable), so the outputs ARE those of the separate calls.  A kernel that fetches
 * each tile's window once and mixes + filters it per shift was built and measured (decimations 4, 8, 10, up to 16
 * shifts per launch; bit-identical outputs): 0.84-1.02 x the speed of the per-shift launches, because the FIR is bound
 * by FP32 issue slots, not by HBM, for every tap count — it is kept in the tuning build only.
 */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrChannelizeFC(
    float sampleRate,
    const float* frequencyShifts,
    size_t numShifts,
    size_t firstSampleIndex,
    size_t decimation,
    const float* taps,
    size_t tapCount,
    const cuComplex* input,
    cuComplex* output,
    size_t outputStride,
    size_t numOutputs,
    int32_t cudaDevice,
    cudaStream_t cudaStream) GSDR_NO_EXCEPT;

#endif /* GSDR_B200_INCLUDE_GSDR_ADJUST_FREQUENCY_H_ */
