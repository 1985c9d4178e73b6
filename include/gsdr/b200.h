/*
 * gsdr/b200.h — NEW, additive host-side layer around the <gsdr/fir.h> / <gsdr/adjust_frequency.h> kernels:
 *   - channel batching (many independent filters in one launch),
 *   - shard planning for multi-GPU runs (independent channels, or contiguous time blocks of one long capture
 *     with a (tapCount - decimation)-sample overlap), pure integer arithmetic, bit-exact,
 *   - a host-buffer pipeline (pinned or pageable host memory in, host memory out) that overlaps H2D copies,
 *     kernels and D2H copies chunk by chunk,
 *   - a single-process multi-GPU executor that runs one pipeline per device.
 * Nothing here exists in the reference (it is single-GPU, one kernel per call, device pointers only:
 * ref: src/fir.cu:73-96); the reference's contract for repeated calls — the caller supplies the overlap and the
 * running firstSampleIndex (ref: include/gsdr/fm.h:26,34) — is what the planner automates.
 *
 * All functions are extern "C", noexcept, and keep the reference's conventions: cudaError_t results, caller-owned
 * buffers, no hidden global state besides per-device attribute caches.
 */
#ifndef GSDR_B200_INCLUDE_GSDR_B200_H_
#define GSDR_B200_INCLUDE_GSDR_B200_H_

#include <cuComplex.h>
#include <cuda_runtime.h>
#include <gsdr/gsdr_export.h>
#include <gsdr/util.h>
#include <stddef.h>
#include <stdint.h>

/* ---- size arithmetic (bit-exact integers) -------------------------------------------------------------- */

/* Largest numOutputs whose last window fits: floor((numInputs - tapCount) / decimation) + 1, or 0. */
GSDR_C_LINKAGE GSDR_PUBLIC size_t gsdrFirNumOutputs(size_t numInputs, size_t tapCount, size_t decimation) GSDR_NO_EXCEPT;
/* Input elements a call with numOutputs outputs reads: (numOutputs - 1) * decimation + tapCount, or 0. */
GSDR_C_LINKAGE GSDR_PUBLIC size_t gsdrFirNumInputs(size_t numOutputs, size_t tapCount, size_t decimation) GSDR_NO_EXCEPT;

/* ---- channel batching ---------------------------------------------------------------------------------- */
/*
 * numChannels independent filters in ONE launch.  Channel c reads input + c*inputStride, writes
 * output + c*outputStride (strides in elements) and uses taps + c*tapStride (tapStride == 0: one tap set shared
 * by all channels).  Per channel the result is exactly that of the corresponding <gsdr/fir.h> call.
 */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrFirFCBatched(
    size_t decimation,
    const float* taps,
    size_t tapCount,
    size_t tapStride,
    const cuComplex* input,
    size_t inputStride,
    cuComplex* output,
    size_t outputStride,
    size_t numOutputs,
    size_t numChannels,
    int32_t cudaDevice,
    cudaStream_t cudaStream) GSDR_NO_EXCEPT;

GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrFirFFBatched(
    size_t decimation,
    const float* taps,
    size_t tapCount,
    size_t tapStride,
    const float* input,
    size_t inputStride,
    float* output,
    size_t outputStride,
    size_t numOutputs,
    size_t numChannels,
    int32_t cudaDevice,
    cudaStream_t cudaStream) GSDR_NO_EXCEPT;

/* ---- shard planning ------------------------------------------------------------------------------------ */

typedef struct gsdrShard {
  uint64_t firstOutput;      /* first output index owned by the shard */
  uint64_t numOutputs;       /* outputs owned (may be 0 when there are more shards than outputs) */
  uint64_t firstInput;       /* firstOutput * decimation */
  uint64_t numInputs;        /* (numOutputs - 1) * decimation + tapCount, 0 for an empty shard */
  uint64_t firstSampleIndex; /* NCO index of the shard's first input: firstSampleIndex + firstInput */
} gsdrShard;

/*
 * Time-block sharding of one capture: shard s of numShards owns outputs
 * [floor(s*N/S), floor((s+1)*N/S)).  Splitting on OUTPUT indices keeps the decimation phase exact; neighbouring
 * shards overlap by tapCount - decimation input samples (read from each shard's own resident copy — nothing is
 * exchanged at run time).  Returns 0, or -1 on invalid arguments.
 */
GSDR_C_LINKAGE GSDR_PUBLIC int gsdrShardPlanTime(
    uint64_t numOutputs,
    uint64_t decimation,
    uint64_t tapCount,
    uint64_t firstSampleIndex,
    uint32_t numShards,
    uint32_t shardIndex,
    gsdrShard* shard) GSDR_NO_EXCEPT;

/* Channel sharding: shard s owns channels [floor(s*C/S), floor((s+1)*C/S)). */
GSDR_C_LINKAGE GSDR_PUBLIC int gsdrShardPlanChannels(
    uint64_t numChannels,
    uint32_t numShards,
    uint32_t shardIndex,
    uint64_t* firstChannel,
    uint64_t* channelCount) GSDR_NO_EXCEPT;

/* ---- host-buffer pipeline ------------------------------------------------------------------------------ */

typedef struct gsdrHostPipeline gsdrHostPipeline;

/*
 * Creates a pipeline on cudaDevice with numBuffers (>= 2) device staging slots, each able to hold a chunk of
 * chunkInputBytes of input (plus the matching output).  Allocation happens here, never in the execute calls.
 */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrHostPipelineCreate(
    int32_t cudaDevice, size_t chunkInputBytes, int numBuffers, gsdrHostPipeline** pipeline) GSDR_NO_EXCEPT;
GSDR_C_LINKAGE GSDR_PUBLIC void gsdrHostPipelineDestroy(gsdrHostPipeline* pipeline) GSDR_NO_EXCEPT;

/*
 * Same result as gsdrFirFC / gsdrFirFF / gsdrAdjustFrequencyFirFC, but taps, input and output are HOST pointers.
 * The capture is cut into time blocks (gsdrShardPlanTime arithmetic); block k+1 is copied in while block k is
 * filtered and block k-1 is copied out.  Returns after the last output byte has landed (the call synchronizes
 * its own streams only).  Pinned host memory gives full overlap; pageable memory works but serialises.
 */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrFirFCHost(
    gsdrHostPipeline* pipeline,
    size_t decimation,
    const float* taps,
    size_t tapCount,
    const cuComplex* input,
    cuComplex* output,
    size_t numOutputs) GSDR_NO_EXCEPT;

GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrFirFFHost(
    gsdrHostPipeline* pipeline,
    size_t decimation,
    const float* taps,
    size_t tapCount,
    const float* input,
    float* output,
    size_t numOutputs) GSDR_NO_EXCEPT;

GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrAdjustFrequencyFirFCHost(
    gsdrHostPipeline* pipeline,
    float sampleRate,
    float frequencyShift,
    size_t firstSampleIndex,
    size_t decimation,
    const float* taps,
    size_t tapCount,
    const cuComplex* input,
    cuComplex* output,
    size_t numOutputs) GSDR_NO_EXCEPT;

/*
 * Single-process multi-GPU: time-shards one host-resident capture over the given pipelines (one per device,
 * driven by one host thread each).  Output lands in one contiguous host array.  No inter-GPU traffic.
 */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrFirFCMultiGpuHost(
    gsdrHostPipeline* const* pipelines,
    int numPipelines,
    size_t decimation,
    const float* taps,
    size_t tapCount,
    const cuComplex* input,
    cuComplex* output,
    size_t numOutputs) GSDR_NO_EXCEPT;

/* ---- kernel selection: introspection and test/tuning hook ---------------------------------------------- */

typedef struct gsdrB200KernelInfo {
  int variant;          /* kernel variant id (see gsdrB200NumPolyphaseVariants), or -1 for the direct kernel */
  int outputsPerThread; /* R */
  int threadsPerBlock;
  int phaseGroups;      /* thread groups that split the polyphase branches of a tile (partial sums added on chip) */
  int windowBuffers;    /* 2: the next tile's window is copied in while the current one is filtered */
  int smCount;
  size_t outputsPerBlock;
  size_t sharedBytesPerBlock;
  size_t numBlocks;
} gsdrB200KernelInfo;

/* firType: 0 = FC, 1 = FF, 2 = CC, 3 = CF, 4 = FC with the fused NCO.  Reports the kernel a call of that shape
 * (one channel, 16-byte aligned pointers) would launch. */
GSDR_C_LINKAGE GSDR_PUBLIC int gsdrB200DescribeKernel(
    int firType,
    size_t decimation,
    size_t tapCount,
    size_t numOutputs,
    int32_t cudaDevice,
    gsdrB200KernelInfo* info) GSDR_NO_EXCEPT;

/*
 * Process-wide override for tests and tuning sweeps: -1 = automatic (default), -2 = always the direct kernel,
 * k >= 0 = polyphase variant k whenever it fits (else the direct kernel).  Returns 0, or -1 if out of range.
 */
GSDR_C_LINKAGE GSDR_PUBLIC int gsdrB200SetKernelVariant(int variant) GSDR_NO_EXCEPT;
GSDR_C_LINKAGE GSDR_PUBLIC int gsdrB200NumKernelVariants(void) GSDR_NO_EXCEPT;
/* Variant ids [0, gsdrB200NumPolyphaseVariants()) are the cp.async-staged polyphase kernel (any decimation, FC and
 * FF); ids from there up to gsdrB200NumKernelVariants() are the TMA-fed kernel (FC, even decimation <= 16). */
GSDR_C_LINKAGE GSDR_PUBLIC int gsdrB200NumPolyphaseVariants(void) GSDR_NO_EXCEPT;
/*
 * Measurement hook (results are WRONG while set): bit 0 skips the global->shared window copies, bit 1 skips the
 * FIR loop.  Lets a profiler time the two halves of the kernel separately.  0 restores normal operation.
 */
GSDR_C_LINKAGE GSDR_PUBLIC int gsdrB200SetDebugFlags(int flags) GSDR_NO_EXCEPT;

#endif /* GSDR_B200_INCLUDE_GSDR_B200_H_ */
