"""The host-side batching / sharding layer of <gsdr/b200.h> through the C ABI: int8 host pipeline, the multi-GPU
executors (host-buffer and device-resident, with the fused gather), shared output buffers, and gsdrFmDemod's
caller-workspace variant.  Multi-device cases skip on a one-GPU box (run with `gpurun --gpus 2`)."""
import numpy as np
import pytest
import torch

import gsdr_b200 as g
from gsdr_b200 import synth
from oracle import oracle

pytestmark = pytest.mark.gpu


def _devices(n):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} CUDA devices")
    return list(range(n))


def test_int8_host_pipeline_equals_device_call_bit_exact(cuda_device):
    D, T, n_in = 8, 255, 2_000_003
    taps = synth.lowpass_taps(T, D)
    x = synth.tone_plus_noise(0, n_in, seed=90)
    xi8 = np.clip(np.round(np.stack([x.real, x.imag], axis=-1) * 127.0), -128, 127).astype(np.int8).reshape(-1)
    n_out = g.fir_num_outputs(n_in, T, D)
    hin = torch.from_numpy(xi8).pin_memory()
    yout = torch.zeros(n_out, dtype=torch.complex64).pin_memory()
    pipe = g.HostPipeline(0, chunkInputBytes=1 << 18, numBuffers=3)  # many chunks, 16-byte aligned chunk starts
    pipe.gsdrFirFCInt8Host(D, taps, T, hin, yout, n_out)
    dt = torch.from_numpy(taps).to(cuda_device)
    dy = torch.zeros(n_out, dtype=torch.complex64, device=cuda_device)
    g.gsdrFirFCInt8(D, dt, T, hin.to(cuda_device), dy, n_out, 0, None)
    torch.cuda.synchronize()
    assert torch.equal(yout, dy.cpu())
    # and against the oracle on the converted samples: max(-1, v / 127) (ref: src/conversion.cu:26)
    xf = np.maximum(xi8.astype(np.float32) / np.float32(127.0), np.float32(-1.0)).view(np.complex64)
    want = oracle.fir("fc", D, taps, xf, n_out, f64=True)
    assert np.abs(yout.numpy() - want).max() <= 1e-5 * float(np.abs(taps).sum()) * float(np.abs(xf).max())
    z = torch.zeros(n_out, dtype=torch.complex64).pin_memory()
    pipe.gsdrAdjustFrequencyFirFCInt8Host(2.4e6, 29520.0, 2 ** 33 + 5, D, taps, T, hin, z, n_out)
    dz = torch.zeros_like(dy)
    g.gsdrAdjustFrequencyFirFCInt8(2.4e6, 29520.0, 2 ** 33 + 5, D, dt, T, hin.to(cuda_device), dz, n_out, 0, None)
    torch.cuda.synchronize()
    assert torch.equal(z, dz.cpu())
    pipe.close()


def test_ff_host_pipeline_equals_device_call_bit_exact(cuda_device):
    D, T, n_in = 5, 63, 1_500_007
    taps = synth.lowpass_taps(T, D)
    x = synth.tone_plus_noise(0, n_in, seed=91, real=True)
    n_out = g.fir_num_outputs(n_in, T, D)
    xin = torch.from_numpy(x).pin_memory()
    yout = torch.zeros(n_out, dtype=torch.float32).pin_memory()
    pipe = g.HostPipeline(0, chunkInputBytes=1 << 20, numBuffers=2)
    pipe.gsdrFirFFHost(D, taps, T, xin, yout, n_out)
    want = oracle.fir("ff", D, taps, x, n_out, f64=True)
    assert np.abs(yout.numpy() - want).max() <= 1e-5 * float(np.abs(taps).sum()) * float(np.abs(x).max())
    pipe.close()


@pytest.mark.parametrize("ndev", [1, 2])
def test_multi_gpu_host_executor_equals_single_device_bit_exact(ndev, cuda_device):
    devs = _devices(ndev)
    D, T, n_in = 8, 255, 3_000_017
    taps = synth.lowpass_taps(T, D)
    x = synth.tone_plus_noise(0, n_in, seed=92)
    n_out = g.fir_num_outputs(n_in, T, D)
    xin = torch.from_numpy(x).pin_memory()
    pipes = [g.HostPipeline(d, chunkInputBytes=1 << 20, numBuffers=3) for d in devs]
    single = torch.zeros(n_out, dtype=torch.complex64).pin_memory()
    pipes[0].gsdrFirFCHost(D, taps, T, xin, single, n_out)
    multi = torch.zeros(n_out, dtype=torch.complex64).pin_memory()
    g.gsdrFirFCMultiGpuHost(pipes, D, taps, T, xin, multi, n_out)
    assert torch.equal(single, multi)
    z1 = torch.zeros(n_out, dtype=torch.complex64).pin_memory()
    zn = torch.zeros(n_out, dtype=torch.complex64).pin_memory()
    pipes[0].gsdrAdjustFrequencyFirFCHost(2.4e6, 29520.0, 77, D, taps, T, xin, z1, n_out)
    g.gsdrAdjustFrequencyFirFCMultiGpuHost(pipes, 2.4e6, 29520.0, 77, D, taps, T, xin, zn, n_out)
    assert torch.equal(z1, zn)
    # channel-sharded: 5 channels of a strided host array
    C, n_c = 5, 400_000
    n_out_c = g.fir_num_outputs(n_c, T, D)
    xc = torch.from_numpy(synth.tone_plus_noise(0, C * n_c, seed=93)).view(C, n_c).pin_memory()
    yc = torch.zeros((C, n_out_c), dtype=torch.complex64).pin_memory()
    g.gsdrFirFCChannelsMultiGpuHost(pipes, D, taps, T, xc, n_c, yc, n_out_c, n_out_c, C)
    for c in range(C):
        one = torch.zeros(n_out_c, dtype=torch.complex64).pin_memory()
        pipes[0].gsdrFirFCHost(D, taps, T, xc[c], one, n_out_c)
        assert torch.equal(one, yc[c])
    for p in pipes:
        p.close()


@pytest.mark.parametrize("ndev", [1, 2])
@pytest.mark.parametrize("nco", [False, True])
def test_device_resident_executor_fused_gather_equals_single_call(ndev, nco, cuda_device):
    """gsdrFirFCMultiGpu: every device filters its resident shard and stores the block straight into the gather buffer
    on devices[0] (peer stores); the result must equal one unsharded call bit for bit, and gsdrMultiGpuGather of
    per-device outputs must give the same."""
    devs = _devices(ndev)
    D, T, n_in, first = 8, 255, 2_000_011, 2 ** 34 + 3
    fs, shift = (2.4e6, 29520.0) if nco else (0.0, 0.0)
    taps = synth.lowpass_taps(T, D)
    x = synth.tone_plus_noise(0, n_in, seed=94)
    n_out = g.fir_num_outputs(n_in, T, D)
    d0 = torch.device("cuda", devs[0])
    want = torch.zeros(n_out, dtype=torch.complex64, device=d0)
    dt0, dx0 = torch.from_numpy(taps).to(d0), torch.from_numpy(x).to(d0)
    if nco:
        g.gsdrAdjustFrequencyFirFC(fs, shift, first, D, dt0, T, dx0, want, n_out, devs[0], None)
    else:
        g.gsdrFirFC(D, dt0, T, dx0, want, n_out, devs[0], None)
    torch.cuda.synchronize(d0)
    mg = g.MultiGpu(devs)
    assert all(mg.peer_ok(i) for i in range(ndev)), "peer access to devices[0] is required for the fused gather"
    shards = [g.shard_plan_time(n_out, D, T, first, ndev, s) for s in range(ndev)]
    tapsd = [torch.from_numpy(taps).to(torch.device("cuda", d)) for d in devs]
    xs = [torch.from_numpy(x[sh.firstInput: sh.firstInput + sh.numInputs]).to(torch.device("cuda", d))
          for sh, d in zip(shards, devs)]
    gathered = torch.zeros(n_out, dtype=torch.complex64, device=d0)
    ms = mg.gsdrFirFCMultiGpu(fs, shift, first, D, tapsd, T, xs, None, gathered, n_out, repeats=2)
    assert ms > 0.0
    assert torch.equal(gathered, want)
    outs = [torch.zeros(sh.numOutputs, dtype=torch.complex64, device=torch.device("cuda", d))
            for sh, d in zip(shards, devs)]
    mg.gsdrFirFCMultiGpu(fs, shift, first, D, tapsd, T, xs, outs, None, n_out)
    later = torch.zeros(n_out, dtype=torch.complex64, device=d0)
    mg.gsdrMultiGpuGather(D, T, outs, later, n_out)
    assert torch.equal(later, want)
    mg.close()


def test_shared_buffer_roundtrip_in_one_process(cuda_device):
    """gsdrSharedBufferCreate gives memory gsdrFirFC can write (the cross-process open is exercised by bench.py N>1)."""
    D, T, n_in = 4, 127, 300_001
    taps = synth.lowpass_taps(T, D)
    n_out = g.fir_num_outputs(n_in, T, D)
    dx = synth.tone_plus_noise(0, n_in, seed=95, device=cuda_device)
    dt = torch.from_numpy(taps).to(cuda_device)
    ptr, handle = g.shared_buffer_create(n_out * 8, 0)
    assert len(handle) == 64 and ptr % 256 == 0
    g.gsdrFirFC(D, dt, T, dx, ptr, n_out, 0, None)
    want = torch.zeros(n_out, dtype=torch.complex64, device=cuda_device)
    g.gsdrFirFC(D, dt, T, dx, want, n_out, 0, None)
    torch.cuda.synchronize()
    from gsdr_b200 import dist as gd
    po = gd.PeerOutput.__new__(gd.PeerOutput)
    po.rank, po.local, po.ptr, po.nbytes = 0, 0, ptr, n_out * 8
    assert torch.equal(po.as_tensor(n_out), want)
    po.close()


def test_fm_demod_with_caller_workspace_equals_pool_variant(cuda_device):
    fs, tuning, channel, dev_hz = 2.4e6, 100.0e6, 100.3e6, 75e3
    D, T, n_out = 10, 255, 30_001
    n_in = n_out * D + T
    x = synth.tone_plus_noise(0, n_in, seed=96, tone_cycles_per_sample=0.125, device=cuda_device)
    dt = torch.from_numpy(synth.lowpass_taps(T, D)).to(cuda_device)
    a = torch.zeros(n_out, dtype=torch.float32, device=cuda_device)
    b = torch.zeros_like(a)
    g.gsdrFmDemod(fs, tuning, channel, dev_hz, D, 5, dt, T, x, a, n_out, 0, None)
    nbytes = g.fm_demod_workspace_bytes(n_out)
    assert nbytes == (n_out + 1) * 8
    ws = torch.empty(nbytes, dtype=torch.uint8, device=cuda_device)
    g.gsdrFmDemodWorkspace(fs, tuning, channel, dev_hz, D, 5, dt, T, x, b, n_out, ws, nbytes, 0, None)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    with pytest.raises(g.CudaError):
        g.gsdrFmDemodWorkspace(fs, tuning, channel, dev_hz, D, 5, dt, T, x, b, n_out, ws, nbytes - 8, 0, None)
    g.release_scratch(0)  # trims the private pool; the next pool call still works
    g.gsdrFmDemod(fs, tuning, channel, dev_hz, D, 5, dt, T, x, b, n_out, 0, None)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
