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
 * Time-block sharding of one capture: shard s of numShards owns outputs [p(s), p(s+1)), p(s) = floor(s*N/S) — rounded
 * down to a multiple of 2048 when the shards hold at least 65536 outputs each (the tensor-core FIR kernel computes in
 * tiles of 1024 or 2048 outputs counted from a call's first output; with aligned shards every output is computed exactly as in
 * the unsharded call, so shards stay bit-identical to it whichever kernel runs).  Splitting on OUTPUT indices keeps
 * the decimation phase exact; neighbouring shards overlap by tapCount - decimation input samples (read from each
 * shard's own resident copy — nothing is exchanged at run time).  Returns 0, or -1 on invalid arguments.
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

/* The same for int8 I/Q host input (2 bytes per complex sample over PCIe; <gsdr/conversion.h> semantics). */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrFirFCInt8Host(
    gsdrHostPipeline* pipeline,
    size_t decimation,
    const float* taps,
    size_t tapCount,
    const int8_t* input,
    cuComplex* output,
    size_t numOutputs) GSDR_NO_EXCEPT;

GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrAdjustFrequencyFirFCInt8Host(
    gsdrHostPipeline* pipeline,
    float sampleRate,
    float frequencyShift,
    size_t firstSampleIndex,
    size_t decimation,
    const float* taps,
    size_t tapCount,
    const int8_t* input,
    cuComplex* output,
    size_t numOutputs) GSDR_NO_EXCEPT;

/* Time-sharded fused NCO + FIR over several pipelines (the NCO index of every shard is advanced exactly). */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrAdjustFrequencyFirFCMultiGpuHost(
    gsdrHostPipeline* const* pipelines,
    int numPipelines,
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
 * Channel-sharded: numChannels independent host-resident channels (strides in elements, one shared tap set), channel
 * c handled by pipeline gsdrShardPlanChannels assigns it to.
 */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrFirFCChannelsMultiGpuHost(
    gsdrHostPipeline* const* pipelines,
    int numPipelines,
    size_t decimation,
    const float* taps,
    size_t tapCount,
    const cuComplex* input,
    size_t inputStride,
    cuComplex* output,
    size_t outputStride,
    size_t numOutputs,
    size_t numChannels) GSDR_NO_EXCEPT;

/* ---- device-resident multi-GPU executor with fused gather ----------------------------------------------- */
/*
 * One process, several GPUs, every shard already resident in its GPU's HBM (the layout gsdrShardPlanTime
 * describes).  One persistent host thread and one stream per device.  No collective on the compute path.
 *
 * Gather ("collect the decimated outputs on one GPU when the caller asks for it"): when gatherOutput is non-NULL
 * it must be a buffer of numOutputs elements on devices[0]; peer access is enabled at creation and EVERY device's
 * FIR kernel stores its block straight into gatherOutput + firstOutput over NVLink — compute and gather are the
 * same kernel, nothing is staged or copied afterwards.  With gatherOutput == NULL shard g's outputs go to
 * outputs[g] on device g, and gsdrMultiGpuGather() can collect them later (cudaMemcpyPeerAsync, one copy per
 * shard, each to its final offset).
 */
typedef struct gsdrMultiGpu gsdrMultiGpu;

GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrMultiGpuCreate(
    const int32_t* devices, int numDevices, gsdrMultiGpu** executor) GSDR_NO_EXCEPT;
GSDR_C_LINKAGE GSDR_PUBLIC void gsdrMultiGpuDestroy(gsdrMultiGpu* executor) GSDR_NO_EXCEPT;
/* 1 when devices[g] can store into devices[0]'s memory (always 1 for g == 0) */
GSDR_C_LINKAGE GSDR_PUBLIC int gsdrMultiGpuPeerOk(const gsdrMultiGpu* executor, int g) GSDR_NO_EXCEPT;

/*
 * taps[g], inputs[g], outputs[g]: device pointers on devices[g]; inputs[g] holds shard g's block (its first element
 * is input sample gsdrShard.firstInput of the capture).  frequencyShift == 0 && sampleRate == 0: plain gsdrFirFC;
 * otherwise the fused exact NCO with firstSampleIndex the index of the capture's first sample.  Blocks until every
 * device has finished (elapsedMs, when non-NULL, receives the longest per-device kernel time measured with CUDA
 * events around `repeats` back-to-back launches, divided by repeats).
 */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrFirFCMultiGpu(
    gsdrMultiGpu* executor,
    float sampleRate,
    float frequencyShift,
    size_t firstSampleIndex,
    size_t decimation,
    const float* const* taps,
    size_t tapCount,
    const cuComplex* const* inputs,
    cuComplex* const* outputs,
    cuComplex* gatherOutput,
    size_t numOutputs,
    int repeats,
    float* elapsedMs) GSDR_NO_EXCEPT;

/* Collects shard outputs (outputs[g] on devices[g], the shard plan's counts) into dst on devices[0]. */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrMultiGpuGather(
    gsdrMultiGpu* executor,
    size_t decimation,
    size_t tapCount,
    const cuComplex* const* outputs,
    cuComplex* dst,
    size_t numOutputs,
    float* elapsedMs) GSDR_NO_EXCEPT;

/* ---- shared output buffers for one-process-per-GPU runs ------------------------------------------------- */
/*
 * The same fused gather when every GPU is driven by its own process (torchrun): rank 0 creates the buffer and
 * publishes the 64-byte handle; the other ranks open it and pass the mapped pointer (plus their firstOutput) as
 * `output` of gsdrFirFC.  Thin wrappers over cudaMalloc / cudaIpcGetMemHandle / cudaIpcOpenMemHandle.
 */
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrSharedBufferCreate(
    size_t bytes, int32_t cudaDevice, void** devicePointer, unsigned char handle[64]) GSDR_NO_EXCEPT;
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrSharedBufferOpen(
    const unsigned char handle[64], int32_t cudaDevice, void** devicePointer) GSDR_NO_EXCEPT;
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrSharedBufferClose(void* devicePointer, int32_t cudaDevice) GSDR_NO_EXCEPT;
GSDR_C_LINKAGE GSDR_PUBLIC cudaError_t gsdrSharedBufferDestroy(void* devicePointer, int32_t cudaDevice) GSDR_NO_EXCEPT;

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

GSDR_C_LINKAGE GSDR_PUBLIC int gsdrB200NumKernelVariants(void) GSDR_NO_EXCEPT;
/* Variant ids [0, gsdrB200NumPolyphaseVariants()) are the cp.async-staged polyphase kernel (any decimation, FC and
 * FF); ids from there up to gsdrB200NumKernelVariants() are the TMA-fed kernels. */
GSDR_C_LINKAGE GSDR_PUBLIC int gsdrB200NumPolyphaseVariants(void) GSDR_NO_EXCEPT;
/*
 * gsdrFirFC calls with decimation 8 and 129..264 taps, decimation 4 and 65..260 taps or decimation 16 and 257..528
 * taps, and at least 65536 outputs per channel (16-byte aligned input) run on the tensor cores: FP16 operands with error compensation (gsdr_b200/csrc/fir_tc_kernel.cuh), FP32-grade results
 * (the same 1e-5 * sum|h| * max|x| bound, measured error ~1e-6 of it relative to the FFMA2 kernels) — but a different
 * rounding than the FFMA2 kernels', depending on an output's position in its tile of 1024 (2048 at decimation 4).  enable = 0 keeps every
 * call on the FFMA2 kernels (bit-identical results whatever the call's size, ~17 % slower on those shapes);
 * process-wide, default 1.  Returns the previous setting.
 */
GSDR_C_LINKAGE GSDR_PUBLIC int gsdrB200SetFirTensorCores(int enable) GSDR_NO_EXCEPT;
/* 1 when this library is the tuning build (the two hooks below exist), 0 for the release build. */
GSDR_C_LINKAGE GSDR_PUBLIC int gsdrB200HasTuningHooks(void) GSDR_NO_EXCEPT;

#ifdef GSDR_B200_TUNING
/*
 * TUNING BUILD ONLY (libgsdr_b200_tuning.so, compiled with -DGSDR_B200_TUNING; used by the variant-coverage tests
 * and tools/sweep.py).  The release library neither exports these symbols nor contains the code behind them.
 *
 * Process-wide override: -1 = automatic (default), -2 = always the direct kernel, k >= 0 = kernel variant k
 * whenever it fits (else the direct kernel).  Returns 0, or -1 if out of range.
 */
GSDR_C_LINKAGE GSDR_PUBLIC int gsdrB200SetKernelVariant(int variant) GSDR_NO_EXCEPT;
/*
 * Measurement hook (results are WRONG while set): bit 0 skips the global->shared window copies, bit 1 skips the
 * FIR loop, bit 2 the output stores.  Lets a profiler time the parts of the kernel separately.  0 restores normal
 * operation.
 */
GSDR_C_LINKAGE GSDR_PUBLIC int gsdrB200SetDebugFlags(int flags) GSDR_NO_EXCEPT;
#endif /* GSDR_B200_TUNING */

#endif /* GSDR_B200_INCLUDE_GSDR_B200_H_ */
