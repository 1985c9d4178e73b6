/*
 * gsdr/stream.h — block-streaming state for the decimating FIR family, C ABI (additive; not in the reference).
 *
 * The reference's FIR entry points are stateless: a caller that feeds fixed-size blocks has to keep the last
 * tapCount - 1 samples of every block in front of the next one and, for the NCO, carry the running
 * firstSampleIndex itself (the contract stated at ref: include/gsdr/fm.h:26,34 and applied at ref: src/fm.cu:202).
 * gsdrFirStream does that bookkeeping over the same kernels:
 *
 *   - blocks of any length, in order, each call enqueues on the caller's stream and never synchronises or allocates;
 *   - the outputs of all pushes, concatenated, are the outputs of ONE gsdrFirFC / gsdrFirFF /
 *     gsdrAdjustFrequencyFirFC call over the concatenated input (same decimation phase, same NCO phase);
 *   - outputs whose window starts in samples carried over from earlier blocks are computed from a small staging
 *     buffer (carry + the first samples of the new block); all other outputs are computed straight from the
 *     caller's block, so the block itself is never copied.
 *
 * All pushes of one stream object must be enqueued on the same CUDA stream (the carry buffers are reused in stream
 * order).  Integer bookkeeping is exposed as a pure function, gsdrFirStreamPlan, and is bit-exact by test.
 */
#ifndef GSDR_B200_INCLUDE_GSDR_STREAM_H_
#define GSDR_B200_INCLUDE_GSDR_STREAM_H_

#include <cuComplex.h>
#include <cuda_runtime.h>
#include <gsdr/gsdr_export.h>
#include <gsdr/util.h>
#include <stddef.h>
#include <stdint.h>

typedef struct gsdrFirStream gsdrFirStream;

enum {
  GSDR_STREAM_FIR_FC = 0,     /* cuComplex input, float taps (gsdrFirFC) */
  GSDR_STREAM_FIR_FF = 1,     /* float input, float taps (gsdrFirFF) */
  GSDR_STREAM_FIR_FC_NCO = 4, /* gsdrAdjustFrequencyFirFC: NCO mix-down fused in front of the FIR */
  GSDR_STREAM_FIR_FC_INT8 = 5,     /* gsdrFirFCInt8: interleaved int8 I/Q input (2 bytes per sample), cuComplex output */
  GSDR_STREAM_FIR_FC_NCO_INT8 = 6  /* gsdrAdjustFrequencyFirFCInt8 */
};

/* What one push does, as a function of the counters alone. */
typedef struct gsdrStreamPlan {
  uint64_t numOutputs;      /* outputs this push produces */
  uint64_t skippedInputs;   /* leading samples of the block that no window needs (tapCount < decimation only) */
  uint64_t headOutputs;     /* computed from the staging buffer [carry | first headNewInputs samples of the block] */
  uint64_t headNewInputs;   /* block samples (after the skipped ones) appended to the carry for them */
  uint64_t bodyOutputs;     /* computed from the caller's block */
  uint64_t bodyOffset;      /* index in the block of the first body window's first sample */
  uint64_t carryLength;     /* samples carried INTO this push */
  uint64_t newCarryLength;  /* samples carried out of it */
  uint64_t newNextStart;    /* absolute index of the first sample of the next output's window */
} gsdrStreamPlan;

/*
 * totalInputs: samples pushed so far; nextStart: absolute index of the first sample of the next output's window
 * (0 for a new stream); align: body windows start on a multiple of `align` samples of the block when a few extra
 * head outputs can achieve it (2 for cuComplex, 4 for float, 8 for int8 I/Q: 16-byte alignment keeps the bulk-copy
 * kernels eligible), 1 to disable.  Returns 0, or -1 on invalid arguments.
 */
GSDR_C_LINKAGE GSDR_PUBLIC int gsdrFirStreamPlan(
    uint64_t decimation,
    uint64_t tapCount,
    uint64_t totalInputs,
    uint64_t nextStart,
    uint64_t numInputs,
    uint32_t align,
    gsdrStreamPlan* plan) GSDR_NO_EXCEPT;

/*
 * taps is a DEVICE pointer (as in gsdrFirFC); the tapCount floats are copied into the object.  sampleRate,
 * frequencyShift and firstSampleIndex are used by GSDR_STREAM_FIR_FC_NCO only.  Allocates (and may synchronise);
 * the push calls do neither.
 */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrFirStreamCreate(
    gsdrFirStream** stream,
    int firType,
    size_t decimation,
    const float* taps,
    size_t tapCount,
    float sampleRate,
    float frequencyShift,
    size_t firstSampleIndex,
    int32_t cudaDevice) GSDR_NO_EXCEPT;

GSDR_C_LINKAGE GSDR_PUBLIC void gsdrFirStreamDestroy(gsdrFirStream* stream) GSDR_NO_EXCEPT;

/* Forgets the carried samples and restarts the sample count (the NCO restarts at firstSampleIndex). */
GSDR_C_LINKAGE GSDR_PUBLIC void gsdrFirStreamReset(gsdrFirStream* stream) GSDR_NO_EXCEPT;

/* Outputs the next push of numInputs samples will produce (so that the caller can size `output`). */
GSDR_C_LINKAGE GSDR_PUBLIC size_t gsdrFirStreamNumOutputs(const gsdrFirStream* stream, size_t numInputs) GSDR_NO_EXCEPT;

/*
 * input: numInputs samples (device; cuComplex, float or int8 I/Q pairs according to firType).  output: at least
 * gsdrFirStreamNumOutputs(stream, numInputs) elements.  *numOutputs (may be NULL) receives the count — known on the
 * host when the call returns; the data follows in stream order.  input must stay valid until the work enqueued
 * here has run, like every device pointer passed to the stateless entry points.
 */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrFirStreamPush(
    gsdrFirStream* stream,
    const void* input,
    size_t numInputs,
    void* output,
    size_t* numOutputs,
    cudaStream_t cudaStream) GSDR_NO_EXCEPT;

#endif /* GSDR_B200_INCLUDE_GSDR_STREAM_H_ */
