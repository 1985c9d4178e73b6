/*
 * gsdr/fm.h — FM receive stage: NCO mix-down, real-tap low-pass, decimate, quadrature demodulate.  C ABI.
 * Same symbol, argument order and types as the reference's include/gsdr/fm.h (ref: include/gsdr/fm.h:42-55);
 * replaces ref: src/fm.cu:21-69 (k_Fm) and :181-218 (gsdrFmDemod).
 *
 *   lp[n]     = sum_{i<numLowPassTaps} input[n*D + i] * nco(firstSampleIndex + n*D + i) * lowPassTaps[i]
 *   output[n] = gain * atan2f(Im(m), Re(m)),  m = lp[n+1] * conj(lp[n]),  n in [0, numOutputs)
 *   gain = rfSampleRate / (2*pi*frequencyDeviation)            (ref: src/fm.cu:203)
 *   nco  = exp(j*2*pi*(tuningFrequency - channelFrequency)*k/rfSampleRate)   (frequencyShift, ref: src/fm.cu:204)
 *
 * This is the stage as the reference documents it.  The reference's own kernel cannot serve as an oracle here:
 * k_AdjustFrequency returns no value (ref: src/adjustFrequency.cu:55-56), the grid is short by 1/32 of the
 * outputs and lanes exit before a full-mask shuffle (ref: src/fm.cu:34-36,58-64).  The NCO is the exact
 * 64-bit-phase one of <gsdr/adjust_frequency.h>; numOutputs + 1 low-pass values are produced internally, so
 *
 *   input must hold numOutputs * decimation + numLowPassTaps elements
 *
 * (the reference's header says (numOutputs + 1) * decimation, which is what its kernel would read only if
 * numLowPassTaps <= decimation).  Repeated calls need the usual overlap (ref: include/gsdr/fm.h:26) and the
 * running firstSampleIndex.  Scratch for the low-pass values is taken from a library-private stream-ordered memory
 * pool (cudaMallocFromPoolAsync / cudaFreeAsync on cudaStream): no synchronisation, capturable in a CUDA graph,
 * and the device's default pool is not touched.
 */
#ifndef GSDR_B200_INCLUDE_GSDR_FM_H_
#define GSDR_B200_INCLUDE_GSDR_FM_H_

#include <cuComplex.h>
#include <cuda_runtime.h>
#include <gsdr/gsdr_export.h>
#include <gsdr/util.h>
#include <stddef.h>
#include <stdint.h>

GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrFmDemod(
    float rfSampleRate,
    float tuningFrequency,
    float channelFrequency,
    float frequencyDeviation,
    uint32_t decimation,
    size_t firstSampleIndex,
    const float* lowPassTaps,
    size_t numLowPassTaps,
    const cuComplex* input,
    float* output,
    size_t numOutputs,
    int32_t cudaDevice,
    cudaStream_t cudaStream) GSDR_NO_EXCEPT;

/*
 * Additive: the same stage with CALLER-OWNED scratch, for callers that hold the reference's "the library allocates
 * nothing" contract (ref: include/gsdr/fir.h:25-29).  workspace: device memory on cudaDevice, 16-byte aligned, at least
 * gsdrFmDemodWorkspaceBytes(numOutputs) bytes, not read or written by anyone else until the stream has passed the call.
 */
GSDR_C_LINKAGE GSDR_PUBLIC size_t gsdrFmDemodWorkspaceBytes(size_t numOutputs) GSDR_NO_EXCEPT;
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrFmDemodWorkspace(
    float rfSampleRate,
    float tuningFrequency,
    float channelFrequency,
    float frequencyDeviation,
    uint32_t decimation,
    size_t firstSampleIndex,
    const float* lowPassTaps,
    size_t numLowPassTaps,
    const cuComplex* input,
    float* output,
    size_t numOutputs,
    void* workspace,
    size_t workspaceBytes,
    int32_t cudaDevice,
    cudaStream_t cudaStream) GSDR_NO_EXCEPT;
/*
 * Additive: the stage as ONE kernel — the quadrature demodulator runs in the FIR's store path, the low-pass values
 * never reach HBM and nothing is allocated.  Measured 4.5 % SLOWER than gsdrFmDemod on BASELINE config 5 (the
 * arctangent costs the issue-bound FIR kernel more than the separate launch costs), hence not the default.  Returns
 * cudaErrorNotSupported, with nothing enqueued, for shapes outside the TMA-fed kernels (odd decimations, unaligned
 * input, tap sets beyond shared memory): use gsdrFmDemod / gsdrFmDemodWorkspace there.
 */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrFmDemodFused(
    float rfSampleRate,
    float tuningFrequency,
    float channelFrequency,
    float frequencyDeviation,
    uint32_t decimation,
    size_t firstSampleIndex,
    const float* lowPassTaps,
    size_t numLowPassTaps,
    const cuComplex* input,
    float* output,
    size_t numOutputs,
    int32_t cudaDevice,
    cudaStream_t cudaStream) GSDR_NO_EXCEPT;
/*
 * Returns the memory gsdrFmDemod's private scratch pool on cudaDevice is holding to the driver (blocks still in use by
 * enqueued work are kept).  A long-running receiver calls this after its largest block size shrinks; never required.
 */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrB200ReleaseScratch(int32_t cudaDevice) GSDR_NO_EXCEPT;

#endif /* GSDR_B200_INCLUDE_GSDR_FM_H_ */
