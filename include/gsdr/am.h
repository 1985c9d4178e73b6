/*
 * gsdr/am.h — AM receive stage, source-compatible with kernrj/gsdr's include/gsdr/am.h:25-37.
 *
 * output[n] = 2 * saturate(|y[n]|) - 1 with y = (input mixed down by the NCO) filtered by lowPassTaps and decimated
 * (ref: src/am.cu:21-50: k_AdjustFrequency per output, then __saturatef(hypotf(re, im)) * 2 - 1).  Here the whole
 * stage is ONE kernel for the shapes the TMA-fed FIR kernels cover: the envelope is taken in the FIR's store path, the
 * mixed and the filtered signals never exist in HBM.  Other shapes run the fused NCO + FIR into library scratch
 * (stream-ordered pool, see gsdr/fm.h) followed by gsdrQuadAmDemod's kernel.  The NCO is the exact one
 * (gsdr/adjust_frequency.h): the reference's k_AdjustFrequency returns nothing.
 *
 * input: numElements * decimation + numLowPassTaps - decimation cuComplex samples ((numElements - 1) * decimation +
 * numLowPassTaps); firstSampleIndex: index of input[0] in the capture (running across calls, ref: include/gsdr/am.h).
 */
#ifndef GSDR_B200_INCLUDE_GSDR_AM_H_
#define GSDR_B200_INCLUDE_GSDR_AM_H_

#include <cuComplex.h>
#include <cuda_runtime.h>
#include <gsdr/gsdr_export.h>
#include <gsdr/util.h>
#include <stddef.h>
#include <stdint.h>

GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrAmDemod(
    float rfSampleRate,
    float tuningFrequency,
    float channelFrequency,
    uint32_t decimation,
    size_t firstSampleIndex,
    const float* lowPassTaps,
    size_t numLowPassTaps,
    const cuComplex* input,
    float* output,
    size_t numElements,
    int32_t cudaDevice,
    cudaStream_t cudaStream) GSDR_NO_EXCEPT;

#endif /* GSDR_B200_INCLUDE_GSDR_AM_H_ */
