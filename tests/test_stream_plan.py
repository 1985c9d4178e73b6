"""gsdrFirStreamPlan (include/gsdr/stream.h): the integer bookkeeping of block streaming, checked without a GPU
against first principles — the pushes of any block sequence must tile the outputs of one call over the concatenated
input, every window must lie inside the buffer it is read from, and the carry must stay below tapCount."""
import random

import pytest

import gsdr_b200 as g


def _num_outputs(n_in, T, D):
    return 0 if n_in < T else (n_in - T) // D + 1


def _simulate(D, T, blocks, align):
    total = next_start = 0
    outputs = 0
    for n in blocks:
        p = g.stream_plan(D, T, total, next_start, n, align)
        carry = max(0, total - next_start)
        assert p.carryLength == carry < max(T, 1)
        assert p.numOutputs == _num_outputs(total + n - next_start, T, D) if total + n >= next_start else p.numOutputs == 0
        assert p.headOutputs + p.bodyOutputs == p.numOutputs
        fresh = n - p.skippedInputs
        assert p.skippedInputs == min(n, max(0, next_start - total))
        if p.headOutputs:
            span = (p.headOutputs - 1) * D + T           # staging samples read by the head outputs
            assert span == carry + p.headNewInputs or (p.headNewInputs == 0 and span <= carry)
            assert p.headNewInputs <= fresh
            assert carry + p.headNewInputs <= 2 * T + (align + 2) * D + 8  # the object's staging capacity
        else:
            assert p.headNewInputs == 0
        if p.bodyOutputs:
            # first body window starts exactly where the head outputs stop, inside the block
            assert p.bodyOffset == p.skippedInputs + p.headOutputs * D - carry
            assert p.bodyOffset + (p.bodyOutputs - 1) * D + T <= n
            if align > 1 and p.bodyOffset % align:
                # unaligned only when no head count within reach fixes it
                assert all((p.skippedInputs + (p.headOutputs + e) * D - carry) % align for e in range(align))
        assert p.newNextStart == next_start + p.numOutputs * D
        total += n
        next_start = p.newNextStart
        assert p.newCarryLength == max(0, total - next_start)
        outputs += p.numOutputs
        assert outputs == _num_outputs(total, T, D)
    return outputs


@pytest.mark.parametrize("D,T", [(1, 1), (1, 63), (8, 255), (32, 1023), (5, 63), (7, 5), (10, 3), (3, 1000), (64, 2)])
@pytest.mark.parametrize("align", [1, 2, 4])
def test_block_streams_tile_the_one_shot_outputs(D, T, align):
    rng = random.Random(1000 * D + T + align)
    for _ in range(20):
        blocks = [rng.choice([0, 1, 2, 3, D, T - 1, T, T + 1, 2 * T + D, 4096, 4097, rng.randrange(1, 20000)])
                  for _ in range(rng.randrange(1, 40))]
        _simulate(D, T, [max(0, b) for b in blocks], align)


def test_aligned_blocks_keep_the_body_aligned():
    """Power-of-two blocks of cuComplex: every body window starts on a 16-byte boundary (the bulk-copy kernels stay
    eligible); only the few head outputs come from the staging buffer."""
    D, T = 8, 255
    total = next_start = 0
    for _ in range(50):
        p = g.stream_plan(D, T, total, next_start, 65536, 2)
        assert p.bodyOffset % 2 == 0 and p.headOutputs <= (T + D - 1) // D + 1
        total += 65536
        next_start = p.newNextStart


def test_invalid_arguments():
    with pytest.raises(ValueError):
        g.stream_plan(0, 5, 0, 0, 10)
    with pytest.raises(ValueError):
        g.stream_plan(4, 0, 0, 0, 10)
