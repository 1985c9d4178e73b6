"""gsdr_b200 — B200-native (sm_100a) decimating FIR + NCO mix-down behind kernrj/gsdr's C ABI.

The product is the shared library `gsdr_b200/csrc/libgsdr_b200.so` (headers in `include/gsdr/`).  This package
is the thin Python mirror of that ABI used by the tests and the benchmark: same function names, same argument
order and meaning as the reference's `include/gsdr/fir.h`, with torch tensors standing in for device pointers.
"""
from .api import (  # noqa: F401
    CudaError,
    MultiGpu,
    FirStream,
    HostPipeline,
    describe_kernel,
    fir_num_inputs,
    fir_num_outputs,
    gsdrAdjustFrequencyFirFC,
    gsdrAdjustFrequencyFirFCInt8,
    gsdrAdjustFrequencyFirFCLiteral,
    gsdrFirCC,
    gsdrFirCF,
    gsdrFirFC,
    gsdrFirFCBatched,
    gsdrFirFCInt8,
    gsdrFirFF,
    gsdrFirFFBatched,
    gsdrAmDemod,
    gsdrFmDemod,
    gsdrFmDemodFused,
    gsdrFmDemodWorkspace,
    fm_demod_workspace_bytes,
    release_scratch,
    fm_chain_launches,
    has_tuning_hooks,
    library_path,
    set_debug_flags,
    gsdrFirFCMultiGpuHost,
    gsdrAdjustFrequencyFirFCMultiGpuHost,
    gsdrFirFCChannelsMultiGpuHost,
    shared_buffer_create,
    shared_buffer_open,
    shared_buffer_close,
    shared_buffer_destroy,
    gsdrInt8ToNormFloat,
    gsdrQuadAmDemod,
    gsdrQuadFmDemod,
    nco_phase_step,
    set_kernel_variant,
    num_kernel_variants,
    num_polyphase_variants,
    shard_plan_channels,
    shard_plan_time,
    stream_plan,
)

__all__ = [n for n in dir() if not n.startswith("_")]
