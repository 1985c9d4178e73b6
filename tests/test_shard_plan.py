"""Host-side sharding arithmetic (bit-exact integers) and the N>1 path over gloo on CPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gsdr_b200 as g
from gsdr_b200 import dist as gd
from gsdr_b200 import synth
from oracle import oracle


def test_time_shards_partition_outputs_exactly():
    for n_out in list(range(0, 40)) + [8_388_577, 214_748_340]:
        for D in (1, 3, 8):
            for T in (1, 7, 255):
                for S in (1, 2, 3, 8):
                    nxt = 0
                    for s in range(S):
                        sh = g.shard_plan_time(n_out, D, T, 1000, S, s)
                        assert sh.firstOutput == nxt
                        nxt += sh.numOutputs
                        assert sh.firstInput == sh.firstOutput * D
                        assert sh.firstSampleIndex == 1000 + sh.firstInput
                        if sh.numOutputs:
                            assert sh.numInputs == (sh.numOutputs - 1) * D + T
                            assert sh.firstInput + sh.numInputs <= max((n_out - 1) * D + T, 0)
                        else:
                            assert sh.numInputs == 0
                        if n_out // S < 65536:
                            assert abs(sh.numOutputs - n_out / S) < 1
                        else:   # big shards start on the tensor-core kernel's tile grid (2048 outputs)
                            assert abs(sh.numOutputs - n_out / S) <= 2048
                            assert sh.firstOutput % 2048 == 0
                    assert nxt == n_out


def test_pure_python_plan_equals_the_c_abi_plan():
    """harness/plan.py (used by bench.py's reference arm, which must not map the product library) restates
    gsdrShardPlanTime / gsdrShardPlanChannels / gsdrFirNum*: identical integers."""
    from harness import plan

    for n_out in (0, 1, 5, 1000, 8_388_577, 214_748_340, 2 ** 40 + 17):
        for D, T in ((1, 63), (8, 255), (32, 1023), (50, 885)):
            for S in (1, 2, 3, 8):
                for s in range(S):
                    a, b = g.shard_plan_time(n_out, D, T, 12345, S, s), plan.shard_plan_time(n_out, D, T, 12345, S, s)
                    assert (a.firstOutput, a.numOutputs, a.firstInput, a.numInputs, a.firstSampleIndex) == (
                        b.firstOutput, b.numOutputs, b.firstInput, b.numInputs, b.firstSampleIndex)
    for C in (0, 1, 7, 1024):
        for S in (1, 3, 8):
            for s in range(S):
                assert g.shard_plan_channels(C, S, s) == plan.shard_plan_channels(C, S, s)
    for n_in in (0, 62, 63, 1 << 20, 1 << 26, (1 << 31) + 5):
        for D, T in ((1, 63), (8, 255), (10, 255)):
            assert g.fir_num_outputs(n_in, T, D) == plan.fir_num_outputs(n_in, T, D)
            assert g.fir_num_inputs(n_in, T, D) == plan.fir_num_inputs(n_in, T, D)


def test_neighbouring_time_shards_overlap_by_taps_minus_decimation():
    n_out, D, T, S = 1000, 8, 255, 4
    shards = [g.shard_plan_time(n_out, D, T, 0, S, s) for s in range(S)]
    for a, b in zip(shards, shards[1:]):
        assert a.firstInput + a.numInputs - b.firstInput == T - D


def test_fm_chain_halo_is_835_samples():
    """BASELINE config 5: window 10*63 + 255 = 885 inputs per final output, stride 50 => overlap 835."""
    D1, T1, D3, T3 = 10, 255, 5, 63
    window = g.fir_num_inputs(g.fir_num_inputs(1, T3, D3) + 1, T1, D1)  # +1: quad demod needs y[n+1]
    assert window == 885
    assert window - D1 * D3 == 835


def test_channel_shards_partition():
    for C in (0, 1, 7, 1024):
        for S in (1, 2, 4, 8):
            nxt = 0
            for s in range(S):
                a, n = g.shard_plan_channels(C, S, s)
                assert a == nxt
                nxt += n
            assert nxt == C


def test_invalid_plans_are_rejected():
    with pytest.raises(ValueError):
        g.shard_plan_time(10, 0, 3, 0, 2, 0)
    with pytest.raises(ValueError):
        g.shard_plan_time(10, 1, 3, 0, 2, 2)
    with pytest.raises(ValueError):
        g.shard_plan_channels(10, 0, 0)


def test_sharded_oracle_equals_unsharded_bit_exact():
    """What the shard plan promises: computing each block separately reproduces the single-call bits."""
    D, T = 8, 255
    h = synth.lowpass_taps(T, D)
    x = synth.tone_plus_noise(0, 40_000, seed=11)
    n_out = g.fir_num_outputs(x.shape[0], T, D)
    whole = oracle.fir("fc", D, h, x, n_out)
    for S in (2, 3, 8):
        parts = []
        for s in range(S):
            sh = g.shard_plan_time(n_out, D, T, 0, S, s)
            parts.append(oracle.fir("fc", D, h, x[sh.firstInput:sh.firstInput + sh.numInputs], sh.numOutputs))
        assert np.concatenate(parts).tobytes() == whole.tobytes()


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank: int, world: int, port: int, n_in: int, ret):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, w, _ = gd.init_from_env(backend="gloo")
    D, T, first = 8, 255, 123_456_789
    fs, f = 2.4e6, 29520.0
    h = synth.lowpass_taps(T, D)
    n_out = g.fir_num_outputs(n_in, T, D)
    sh = gd.time_shard(n_out, D, T, first, w, r)
    counts = [gd.time_shard(n_out, D, T, first, w, q).numOutputs for q in range(w)]
    # every rank regenerates only ITS block (+halo) of the capture from the global sample index
    x_local = synth.tone_plus_noise(sh.firstInput, sh.numInputs, seed=21)
    # the kernel's stand-in on CPU is the oracle (tests may call it); the plumbing under test is plan + gather
    y_local = oracle.adjust_frequency_fir_fc(oracle.NCO_EXACT, fs, f, sh.firstSampleIndex, D, h, x_local, sh.numOutputs)
    t = gd.max_over_ranks(float(r + 1))
    assert t == float(w)
    got = gd.gather_outputs(torch.from_numpy(y_local), counts, dst=0)
    # the unpadded point-to-point gather: every block received straight at its final offset
    firsts = [gd.time_shard(n_out, D, T, first, w, q).firstOutput for q in range(w)]
    full = torch.zeros(n_out, dtype=torch.complex64) if r == 0 else None
    gd.gather_outputs_p2p(torch.from_numpy(y_local), counts, firsts, full, dst=0)
    if r == 0:
        ret["p2p_ok"] = bool(full.numpy().tobytes() == got.numpy().tobytes())
    if r == 0:
        x = synth.tone_plus_noise(0, n_in, seed=21)
        want = oracle.adjust_frequency_fir_fc(oracle.NCO_EXACT, fs, f, first, D, h, x, n_out)
        ret["ok"] = bool(got.numpy().tobytes() == want.tobytes())
        ret["n"] = int(got.shape[0])
    else:
        assert got is None
    dist.destroy_process_group()


def test_two_rank_gloo_time_shard_and_gather_is_bit_exact():
    world, n_in = 2, 30_011
    port = _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_gloo_worker, args=(world, port, n_in, ret), nprocs=world, join=True)
        assert ret["ok"] and ret["n"] == g.fir_num_outputs(n_in, 255, 8)
        assert ret["p2p_ok"]
