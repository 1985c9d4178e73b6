"""Size and shard arithmetic of the FIR path in pure Python (no native code is loaded).

Mirrors gsdrFirNumOutputs / gsdrFirNumInputs / gsdrShardPlanTime / gsdrShardPlanChannels of include/gsdr/b200.h
(gsdr_b200/csrc/gsdr_host.cu); tests/test_shard_plan.py checks the two implementations against each other.
The reference's implicit contract (ref: src/fir.cu:57-70): the caller owns (numOutputs-1)*decimation + tapCount
input samples.
"""
from __future__ import annotations

from dataclasses import dataclass


def fir_num_outputs(num_inputs: int, tap_count: int, decimation: int) -> int:
    if decimation == 0 or tap_count == 0 or num_inputs < tap_count:
        return 0
    return (num_inputs - tap_count) // decimation + 1


def fir_num_inputs(num_outputs: int, tap_count: int, decimation: int) -> int:
    return 0 if num_outputs == 0 else (num_outputs - 1) * decimation + tap_count


@dataclass(frozen=True)
class Shard:
    firstOutput: int
    numOutputs: int
    firstInput: int
    numInputs: int
    firstSampleIndex: int


def _time_split(n: int, shards: int, s: int) -> int:
    """Interior split points of shards of >= 64Ki outputs sit on multiples of 2048 outputs (the tensor-core kernel's
    largest tile: aligned shards reproduce the unsharded call's bits; gsdr_host.cu timeSplitPoint)."""
    p = n * s // shards
    if s == 0 or s >= shards or n // shards < 65536:
        return p
    return p - p % 2048


def shard_plan_time(num_outputs: int, decimation: int, tap_count: int, first_sample_index: int, num_shards: int,
                    shard_index: int) -> Shard:
    """Split on OUTPUT indices; the (taps - decimation)-sample overlap is read from the shard's own copy."""
    if num_shards <= 0 or not 0 <= shard_index < num_shards or decimation <= 0:
        raise ValueError("bad shard request")
    a = _time_split(num_outputs, num_shards, shard_index)
    b = _time_split(num_outputs, num_shards, shard_index + 1)
    return Shard(a, b - a, a * decimation, (b - a - 1) * decimation + tap_count if b > a else 0,
                 first_sample_index + a * decimation)


def shard_plan_channels(num_channels: int, num_shards: int, shard_index: int) -> tuple[int, int]:
    if num_shards <= 0 or not 0 <= shard_index < num_shards:
        raise ValueError("bad shard request")
    a = num_channels * shard_index // num_shards
    b = num_channels * (shard_index + 1) // num_shards
    return a, b - a
