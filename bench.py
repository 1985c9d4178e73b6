#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: input Msamples/s of the 255-tap decimate-by-8 complex FIR (gsdrFirFC).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|cpu] [--workload cfg2|...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one capture: 2^26 cuComplex samples per GPU (BASELINE config 2).  At
N > 1 the capture is N*2^26 samples long and time-sharded with gsdrShardPlanTime: every rank holds its own block
plus the (taps - decimation)-sample overlap resident in HBM, so there is no collective on the compute path
(weak scaling).  Timing is on the device (CUDA events on the launching stream), K steps bracketed by a barrier +
synchronize, max over ranks.  The input (537 MB per GPU) is larger than the 126 MB L2, so every step streams
from HBM.  After the timed loop the output that the timed launches wrote is checked against the oracle
(prefix / middle / suffix windows, `parity` in the line).

The single JSON line also carries (impl ours):
  other_configs  BASELINE configs 1, 3, 4, 5, each timed the same way with its own roofline, clocks and parity check
                 (time-boxed; cfg4 is channel-sharded and cfg5 time-sharded with the 835-sample halo at N > 1);
  strong         (N > 1) the fixed 2^26-sample capture of config 2 split N ways, CUDA-graph launch, per-rank times;
  gather         (N > 1) decimated outputs collected on rank 0: NCCL send/recv straight into the final buffer, and
                 the fused form (each rank's kernel stores its outputs into rank 0's buffer over NVLink).

impl:
  ours       libgsdr_b200.so through the C ABI (`value`: device-resident buffers; `e2e`: pinned host buffers in,
             pinned host buffers out through gsdrFirFCHost, H2D/D2H inside the timed region).
  reference  the reference's OWN CUDA kernels (oracle/_ref/libgsdr_ref.so, compiled for sm_100 from
             /root/reference/src/fir.cu by oracle/build_ref.sh) on the same buffers, rank 0 only.  gsdr has no
             CPU implementation; its e2e is what a user of the reference must do: cudaMemcpy in, gsdrFirFC,
             cudaMemcpy out.  This arm never imports gsdr_b200: the product library is not mapped.
  cpu        the scalar C oracle (restating ref: src/fir.cu:57-70) on all host cores, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from harness import benchlib as bl  # noqa: E402  (pure Python: loads no native code)
from harness import plan, synth  # noqa: E402

WORKLOADS = bl.WORKLOADS
UNIT = bl.UNIT


def _parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference", "cpu"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="cfg2")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the other_configs / strong / gather blocks")
    ap.add_argument("--others-budget-s", type=float, default=150.0)
    ap.add_argument("--variant", type=int, default=-1, help="force a kernel variant (needs the tuning build)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    return args


# ---------------------------------------------------------------------------------------------------------------------
# reference and cpu arms: harness + oracle only
# ---------------------------------------------------------------------------------------------------------------------

def _run_cpu(args):
    wl = WORKLOADS[args.workload]
    taps = synth.lowpass_taps(wl["T"], wl["D"])
    cb = bl.cpu_baseline(wl["D"], wl["T"], taps, wl["n_in"], target_seconds=15.0)
    bl.emit({"metric": bl.METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": 0, "steps": 1, "warmup": 0, "impl": "cpu",
             "higher_is_better": True, "dtype": "f32", "data": "synthetic", "config": bl.config_block(args.workload, 1),
             "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                                          "d2h_bytes_per_step": 0}})


def _run_reference(args):
    """The reference's own kernel (unmodified src/fir.cu built for sm_100) on one B200, rank 0 only."""
    import torch

    from oracle import ref_cuda

    wl = WORKLOADS[args.workload]
    if wl["kind"] != "fc" or "channels" in wl:
        bl.emit({"impl": "reference", "unavailable": f"the reference arm times single-call FC workloads, not {args.workload}"})
        return
    if not ref_cuda.available():
        bl.emit({"impl": "reference", "unavailable": "oracle/_ref/libgsdr_ref.so was not built (needs /root/reference "
                                                      "at build time)"})
        return
    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    D, T, n_in = wl["D"], wl["T"], wl["n_in"]
    taps = synth.lowpass_taps(T, D)
    n_out = plan.fir_num_outputs(n_in, T, D)
    x = synth.tone_plus_noise(0, n_in, seed=0x5EED0002, device=dev)
    dtaps = torch.from_numpy(taps).to(dev)
    y = torch.zeros(n_out, dtype=torch.complex64, device=dev)
    stream = torch.cuda.Stream(device=dev)
    if wl["nco"]:
        def step():
            ref_cuda.adjust_frequency_fir_fc(bl.NCO_FS, bl.NCO_SHIFT, 0, D, dtaps, T, x, y, n_out, local,
                                             stream.cuda_stream)
    else:
        def step():
            ref_cuda.fir("fc", D, dtaps, T, x, y, n_out, local, stream.cuda_stream)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize(dev)
    sampler = bl.ClockSampler(local).start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    sampler.stop()
    ms_step = e0.elapsed_time(e1) / args.steps
    value = n_in / (ms_step * 1e-3) / 1e6
    parity = bl.check_fir_windows("fc", D, taps, x, y, n_out) if not wl["nco"] else None
    peak, peak_src, ffma, ffma2 = bl.fp32_peak(local)
    roof = bl.roofline(8 * n_in + 8 * n_out + 4 * T, 4.0 * T * n_out, ms_step * 1e-3, peak, peak_src)
    # end to end as a user of the reference has to write it: copy in, call, copy out
    e2e = None
    if not args.no_e2e:
        xin = torch.empty(n_in, dtype=torch.complex64).pin_memory()
        xin.copy_(x)
        yout = torch.zeros(n_out, dtype=torch.complex64).pin_memory()
        htaps = torch.from_numpy(taps).pin_memory()

        def e2e_step():
            with torch.cuda.stream(stream):
                dtaps.copy_(htaps, non_blocking=True)
                x.copy_(xin, non_blocking=True)
                step()
                yout.copy_(y, non_blocking=True)
            stream.synchronize()
        e2e_step()
        n = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(n):
            e2e_step()
        dt = (time.perf_counter() - t0) / n
        e2e = {"value": n_in / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": n_in * 8 + T * 4,
               "d2h_bytes_per_step": n_out * 8, "ms_per_step": dt * 1e3, "steps": n,
               "timing": "host wall clock around copy in + gsdrFirFC + copy out + stream sync"}
    cpu = None if args.no_cpu else bl.cpu_baseline(D, T, taps, n_in)
    bl.emit({"metric": bl.METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
             "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
             "data": "synthetic", "impl": "reference", "config": bl.config_block(args.workload, 1),
             "kernel": "reference k_FirDecimate<float2,float2,float> (32-thread blocks, one thread per output), "
                       "oracle/_ref/libgsdr_ref.so",
             "roofline": roof, "parity": parity, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": args.steps,
             "clocks": sampler.summary()})


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------

class Ctx:
    """Per-process state of the `ours` arm."""

    def __init__(self, args):
        import torch

        import gsdr_b200 as g
        from gsdr_b200 import dist as gd

        self.torch, self.g, self.gd, self.args = torch, g, gd, args
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        self.rank, self.world, self.local = gd.init_from_env()
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist_on = self.world > 1
        self.numa_node = gd.bind_to_gpu_numa_node(self.local) if self.dist_on else None
        self.stream = torch.cuda.Stream(device=self.dev)
        self.fp32_peak, self.fp32_src, self.ffma_tf, self.ffma2_tf = bl.fp32_peak(self.local)
        self._flush = None

    def barrier(self):
        t = self.torch
        t.cuda.synchronize(self.dev)
        if self.dist_on:
            t.distributed.barrier()
        t.cuda.synchronize(self.dev)

    def max_ranks(self, v: float) -> float:
        return self.gd.max_over_ranks(v, device=self.dev) if self.dist_on else v

    def all_ranks(self, v: float):
        if not self.dist_on:
            return [v]
        t = self.torch
        out = [t.zeros(1, dtype=t.float64, device=self.dev) for _ in range(self.world)]
        t.distributed.all_gather(out, t.tensor([v], dtype=t.float64, device=self.dev))
        return [float(o.item()) for o in out]

    def flush_l2(self):
        """Writes 256 MiB (2x the L2) on the timing stream."""
        t = self.torch
        if self._flush is None:
            self._flush = t.empty(256 << 20, dtype=t.uint8, device=self.dev)
        with t.cuda.stream(self.stream):
            self._flush.zero_()

    def time_steps(self, step, steps, warmup, flush_between=False):
        """W warm-up steps, then K timed steps bracketed by barrier + synchronize; returns (this rank's ms per step,
        max-over-ranks ms per step, clock summary)."""
        t = self.torch
        sampler = bl.ClockSampler(self.local).start()  # (NVML set-up and thread start cost milliseconds: not here ↓)
        for _ in range(warmup):
            step()
        self.barrier()
        sampler.begin()
        if flush_between:
            total = 0.0
            for _ in range(steps):
                self.flush_l2()
                e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
                e0.record(self.stream)
                step()
                e1.record(self.stream)
                self.stream.synchronize()
                total += e0.elapsed_time(e1)
            self.barrier()
        else:
            e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            for _ in range(steps):
                step()
            e1.record(self.stream)
            self.barrier()
            total = e0.elapsed_time(e1)
        sampler.stop()
        mine = total / steps
        return mine, self.max_ranks(mine), sampler.summary()


def _kernel_text(g, info) -> str:
    if info.variant == g.num_kernel_variants():
        return (f"tensor-core kernel (tcgen05.mma kind::f16, FP16 head + remainder operands, FP32 accumulators in "
                f"TMEM; gsdr_b200/csrc/fir_tc_kernel.cuh): {info.threadsPerBlock} threads/CTA, "
                f"{info.outputsPerBlock} outputs/tile, {info.sharedBytesPerBlock} B smem, {info.numBlocks} tiles")
    fam = "TMA-fed" if info.variant >= g.num_polyphase_variants() else "cp.async-staged"
    return (f"{fam} persistent polyphase kernel, variant {info.variant}: {info.threadsPerBlock} threads/CTA, "
            f"{info.outputsPerThread} outputs/thread, {info.phaseGroups} branch groups, {info.sharedBytesPerBlock} B "
            f"smem, {info.numBlocks} tiles")


def case_fc(c: Ctx, name: str, steps: int, warmup: int, keep=False) -> dict:
    """Single-call FC workloads (cfg2, cfg3, ...), time-sharded at N > 1 (weak scaling)."""
    t, g = c.torch, c.g
    wl = WORKLOADS[name]
    D, T, n_in_gpu = wl["D"], wl["T"], wl["n_in"]
    taps = synth.lowpass_taps(T, D)
    n_in_total = n_in_gpu * c.world
    n_out_total = g.fir_num_outputs(n_in_total, T, D)
    sh = g.shard_plan_time(n_out_total, D, T, 0, c.world, c.rank)
    x = synth.tone_plus_noise(sh.firstInput, sh.numInputs, seed=0x5EED0002, device=c.dev)
    dtaps = t.from_numpy(taps).to(c.dev)
    y = t.zeros(sh.numOutputs, dtype=t.complex64, device=c.dev)
    if wl["nco"]:
        def step():
            g.gsdrAdjustFrequencyFirFC(bl.NCO_FS, bl.NCO_SHIFT, sh.firstSampleIndex, D, dtaps, T, x, y, sh.numOutputs,
                                       c.local, c.stream)
    else:
        def step():
            g.gsdrFirFC(D, dtaps, T, x, y, sh.numOutputs, c.local, c.stream)
    mine, ms, clocks = c.time_steps(step, steps, warmup)
    y_timed = y.clone()  # what the timed launches wrote
    sustained = None
    if keep:
        # the timed region of K steps lasts a few milliseconds: too short for more than a couple of clock samples.  A
        # second, longer loop of the same launches (about 0.3 s) shows what the clocks do under sustained load.
        k_long = max(steps, int(300.0 / max(ms, 1e-3)))
        s_mine, s_ms, s_clocks = c.time_steps(step, k_long, 0)
        sustained = {"steps": k_long, "ms_per_step": s_ms, "value": n_in_total / (s_ms * 1e-3) / 1e6, "clocks": s_clocks}
    parity = bl.check_fir_windows("fc", D, taps, x, y_timed, sh.numOutputs,
                                  nco=(bl.NCO_FS, bl.NCO_SHIFT, sh.firstSampleIndex) if wl["nco"] else None)
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get(name)
        except Exception:
            traffic = None
    info = g.describe_kernel(4 if wl["nco"] else 0, D, T, sh.numOutputs, c.local)
    on_tensor_cores = info.variant == g.num_kernel_variants()
    roof = bl.roofline(8 * sh.numInputs + 8 * sh.numOutputs + 4 * T, 4.0 * T * sh.numOutputs, mine * 1e-3,
                       c.fp32_peak, c.fp32_src, traffic, tensor_core=on_tensor_cores)
    roof["fp32"]["ffma_tflops"], roof["fp32"]["ffma2_tflops"] = c.ffma_tf, c.ffma2_tf
    out = {"ms_per_step": ms, "value": n_in_total / (ms * 1e-3) / 1e6, "unit": UNIT, "scaling": "weak",
           "roofline": roof, "parity": parity, "clocks": clocks, "gpu_launches": steps, "kernel": _kernel_text(g, info),
           "config": bl.config_block(name, c.world)}
    if sustained is not None:
        out["sustained"] = sustained
    if keep:
        out["_state"] = dict(x=x, y=y, y_timed=y_timed, dtaps=dtaps, taps=taps, sh=sh, step=step,
                             n_in_total=n_in_total, n_out_total=n_out_total)
    return out


def case_cfg1(c: Ctx, steps: int, warmup: int) -> dict:
    """BASELINE config 1: one 63-tap real FIR over 1Mi samples — a few microseconds of kernel, so the launch itself
    is a large share.  Timed per launch with an L2 flush before each (cold), and back to back (warm L2)."""
    t, g = c.torch, c.g
    wl = WORKLOADS["cfg1"]
    D, T, n_in = wl["D"], wl["T"], wl["n_in"]
    n_out = g.fir_num_outputs(n_in, T, D)
    taps = synth.lowpass_taps(T, D)
    x = synth.tone_plus_noise(0, n_in, seed=0x5EED0001, device=c.dev, real=True)
    dtaps = t.from_numpy(taps).to(c.dev)
    y = t.zeros(n_out, dtype=t.float32, device=c.dev)
    tiny = t.zeros(8, dtype=t.float32, device=c.dev)

    def step():
        g.gsdrFirFF(D, dtaps, T, x, y, n_out, c.local, c.stream)

    def step_floor():  # the same entry point with one output: what a launch costs with no work in it
        g.gsdrFirFF(D, dtaps, T, x, tiny, 1, c.local, c.stream)

    k = max(steps, 50)
    cold, _, clocks = c.time_steps(step, k, warmup, flush_between=True)
    warm, _, _ = c.time_steps(step, 4 * k, warmup)
    floor, _, _ = c.time_steps(step_floor, 4 * k, warmup)
    # the same launches replayed from a CUDA graph: what the device needs per launch once the host is out of the way
    graph = t.cuda.CUDAGraph()
    with t.cuda.graph(graph, stream=c.stream):
        for _ in range(100):
            step()
    with t.cuda.stream(c.stream):
        graph.replay()
    c.stream.synchronize()
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    with t.cuda.stream(c.stream):
        e0.record(c.stream)
        for _ in range(5):
            graph.replay()
        e1.record(c.stream)
    c.stream.synchronize()
    graphed = e0.elapsed_time(e1) / 500.0
    want = None
    from oracle import oracle
    want = oracle.fir("ff", D, taps, x.cpu().numpy(), n_out, f64=True)
    err = float(abs(y.cpu().numpy().astype("float64") - want).max())
    xmax = float(x.abs().max().item())
    tol = 1e-5 * float(abs(taps).sum()) * xmax
    roof = bl.roofline(4 * n_in + 4 * n_out + 4 * T, 2.0 * T * n_out, cold * 1e-3, c.fp32_peak, c.fp32_src)
    info = g.describe_kernel(1, D, T, n_out, c.local)
    return {"ms_per_step": cold, "value": n_in / (cold * 1e-3) / 1e6, "unit": UNIT, "scaling": "single", "n_gpus": 1,
            "us_per_launch_cold_l2": cold * 1e3, "us_per_launch_back_to_back": warm * 1e3,
            "us_per_empty_launch_back_to_back": floor * 1e3, "launch_latency_share": floor / warm if warm > 0 else None,
            "us_per_launch_cuda_graph": graphed * 1e3, "value_cuda_graph": n_in / (graphed * 1e-3) / 1e6,
            "roofline": roof, "parity": {"max_err": err, "tol": tol, "ok": bool(err <= tol), "windows": 1,
                                         "outputs_per_window": n_out, "against": "oracle f64, every output"},
            "clocks": clocks, "gpu_launches": k, "kernel": _kernel_text(g, info), "config": bl.config_block("cfg1", 1)}


def case_cfg4(c: Ctx, steps: int, warmup: int) -> dict:
    """BASELINE config 4: 1024 channels sharded across the ranks, one batched launch per rank (strong scaling)."""
    t, g = c.torch, c.g
    wl = WORKLOADS["cfg4"]
    D, T, n_in, chans_total = wl["D"], wl["T"], wl["n_in"], wl["channels"]
    taps = synth.lowpass_taps(T, D)
    c0, cn = g.shard_plan_channels(chans_total, c.world, c.rank)
    n_out = g.fir_num_outputs(n_in, T, D)
    base = synth.tone_plus_noise(0, n_in * 8, seed=0x5EED0004, device=c.dev).view(8, n_in)
    xb = t.empty((cn, n_in), dtype=t.complex64, device=c.dev)
    for ch in range(cn):
        xb[ch].copy_(base[(c0 + ch) % 8])
        xb[ch, :1024] *= 1.0 + 0.001 * (c0 + ch)  # channels are not bit-identical copies
    del base
    yb = t.zeros((cn, n_out), dtype=t.complex64, device=c.dev)
    dtaps = t.from_numpy(taps).to(c.dev)

    def step():
        g.gsdrFirFCBatched(D, dtaps, T, 0, xb, n_in, yb, n_out, n_out, cn, c.local, c.stream)

    mine, ms, clocks = c.time_steps(step, steps, warmup)
    worst = {"max_err": 0.0, "tol": float("inf"), "ok": True}
    for ch in sorted({0, cn // 2, cn - 1}):
        p = bl.check_fir_windows("fc", D, taps, xb[ch], yb[ch], n_out, width=256)
        if p["max_err"] / p["tol"] >= worst["max_err"] / worst["tol"]:
            worst = p
        worst["ok"] = worst["ok"] and p["ok"]
    worst["channels_checked"] = len({0, cn // 2, cn - 1})
    info = g.describe_kernel(0, D, T, n_out, c.local)
    roof = bl.roofline(cn * (8 * n_in + 8 * n_out) + 4 * T, 4.0 * T * n_out * cn, mine * 1e-3, c.fp32_peak, c.fp32_src,
                       tensor_core=info.variant == g.num_kernel_variants())
    return {"ms_per_step": ms, "value": n_in * chans_total / (ms * 1e-3) / 1e6, "unit": UNIT, "scaling": "strong",
            "channels_this_rank": cn, "roofline": roof, "parity": worst, "clocks": clocks, "gpu_launches": steps,
            "kernel": _kernel_text(g, info), "config": bl.config_block("cfg4", c.world)}


def case_cfg5(c: Ctx, steps: int, warmup: int) -> dict:
    """BASELINE config 5: the FM receive chain; the composite stage is a window-885 / stride-50 FIR for planning."""
    t, g = c.torch, c.g
    wl = WORKLOADS["cfg5"]
    D, T, n_in_gpu = wl["D"], wl["T"], wl["n_in"]
    D3, T3 = wl["chain"]["D3"], wl["chain"]["T3"]
    fs, tuning, channel, dev_hz = 2.4e6, 100.0e6, 100.3e6, 75e3
    window, stride = D * T3 + T, D * D3  # 885, 50
    n_in_total = n_in_gpu * c.world
    n3_total = (n_in_total - window) // stride + 1
    sh3 = g.shard_plan_time(n3_total, stride, window, 0, c.world, c.rank)
    n3 = sh3.numOutputs
    n2 = g.fir_num_inputs(n3, T3, D3)
    h1, h3 = synth.lowpass_taps(T, D), synth.lowpass_taps(T3, D3)
    x5 = synth.tone_plus_noise(sh3.firstInput, sh3.numInputs, seed=0x5EED0005, device=c.dev,
                               tone_cycles_per_sample=300e3 / 2.4e6)
    d1, d3 = t.from_numpy(h1).to(c.dev), t.from_numpy(h3).to(c.dev)
    dm = t.zeros(n2, dtype=t.float32, device=c.dev)
    au = t.zeros(n3, dtype=t.float32, device=c.dev)

    def step():
        g.gsdrFmDemod(fs, tuning, channel, dev_hz, D, sh3.firstSampleIndex, d1, T, x5, dm, n2, c.local, c.stream)
        g.gsdrFirFF(D3, d3, T3, dm, au, n3, c.local, c.stream)

    mine, ms, clocks = c.time_steps(step, steps, warmup)
    gain = float(__import__("numpy").float32(fs) / (__import__("numpy").float32(2.0 * math.pi) *
                                                      __import__("numpy").float32(dev_hz)))
    parity = bl.check_chain_windows(D, h1, D3, h3, fs, tuning - channel, sh3.firstSampleIndex, gain, x5, au, n3)
    # roofline of the dominant kernel (stage 1: fused mix + 255-tap FIR): bytes in + low-pass samples out
    n1 = n2 + 1
    roof = bl.roofline(8 * sh3.numInputs + 8 * n1 + 4 * T, 4.0 * T * n1, mine * 1e-3, c.fp32_peak, c.fp32_src)
    roof["note"] = ("time is the whole chain (3 launches); algorithmic work is stage 1's only, so frac is a lower "
                    "bound for the dominant kernel")
    return {"ms_per_step": ms, "value": n_in_total / (ms * 1e-3) / 1e6, "unit": UNIT, "scaling": "weak",
            "halo_input_samples": window - stride, "audio_outputs_total": n3_total, "roofline": roof, "parity": parity,
            "clocks": clocks, "gpu_launches": steps * g.fm_chain_launches(), "config": bl.config_block("cfg5", c.world)}


def strong_scaling(c: Ctx, steps: int, warmup: int) -> dict:
    """The fixed 2^26-sample capture of config 2 split over the ranks (strong scaling): K launches captured in one
    CUDA graph per rank, rotating over input copies that together exceed twice the L2."""
    t, g = c.torch, c.g
    wl = WORKLOADS["cfg2"]
    D, T, n_in_total = wl["D"], wl["T"], wl["n_in"]
    taps = synth.lowpass_taps(T, D)
    n_out_total = g.fir_num_outputs(n_in_total, T, D)
    sh = g.shard_plan_time(n_out_total, D, T, 0, c.world, c.rank)
    copies = max(2, math.ceil((256 << 20) / (sh.numInputs * 8)) + 1)
    x0 = synth.tone_plus_noise(sh.firstInput, sh.numInputs, seed=0x5EED0002, device=c.dev)
    xs = [x0] + [x0.clone() for _ in range(copies - 1)]
    dtaps = t.from_numpy(taps).to(c.dev)
    y = t.zeros(sh.numOutputs, dtype=t.complex64, device=c.dev)
    for i in range(warmup):
        g.gsdrFirFC(D, dtaps, T, xs[i % copies], y, sh.numOutputs, c.local, c.stream)
    c.stream.synchronize()
    graph = t.cuda.CUDAGraph()
    with t.cuda.graph(graph, stream=c.stream):
        for i in range(steps):
            g.gsdrFirFC(D, dtaps, T, xs[i % copies], y, sh.numOutputs, c.local, c.stream)
    with t.cuda.stream(c.stream):
        graph.replay()  # warm the graph
    c.barrier()
    sampler = bl.ClockSampler(c.local).start()
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    with t.cuda.stream(c.stream):
        e0.record(c.stream)
        graph.replay()
        e1.record(c.stream)
    c.barrier()
    sampler.stop()
    mine = e0.elapsed_time(e1) / steps
    ms = c.max_ranks(mine)
    parity = bl.check_fir_windows("fc", D, taps, x0, y, sh.numOutputs)
    info = g.describe_kernel(0, D, T, sh.numOutputs, c.local)
    roof = bl.roofline(8 * sh.numInputs + 8 * sh.numOutputs + 4 * T, 4.0 * T * sh.numOutputs, mine * 1e-3, c.fp32_peak,
                       c.fp32_src, tensor_core=info.variant == g.num_kernel_variants())
    # the 67 MB of outputs of this capture collected on rank 0 (NCCL send/recv into the final buffer)
    counts = [g.shard_plan_time(n_out_total, D, T, 0, c.world, r).numOutputs for r in range(c.world)]
    firsts = [g.shard_plan_time(n_out_total, D, T, 0, c.world, r).firstOutput for r in range(c.world)]
    full = t.zeros(n_out_total, dtype=t.complex64, device=c.dev) if c.rank == 0 else None
    c.gd.gather_outputs_p2p(y, counts, firsts, full)
    c.barrier()
    g0, g1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(5):
        c.gd.gather_outputs_p2p(y, counts, firsts, full)
    g1.record()
    c.barrier()
    gather_ms = c.max_ranks(g0.elapsed_time(g1) / 5)
    return {"ms_per_step": ms, "gather_ms": gather_ms, "gather_bytes_to_rank0": 8 * (n_out_total - counts[0]), "value": n_in_total / (ms * 1e-3) / 1e6, "unit": UNIT, "scaling": "strong",
            "input_samples_total": n_in_total, "per_rank_us": [v * 1e3 for v in c.all_ranks(mine)],
            "launch": f"one CUDA graph of {steps} gsdrFirFC launches per rank",
            "l2": f"{copies} rotating input copies of {sh.numInputs * 8 >> 20} MiB per rank (> 2x the 126 MB L2 together)",
            "roofline": roof, "parity": parity, "clocks": sampler.summary(), "gpu_launches": steps,
            "kernel": _kernel_text(g, info)}


def gather_block(c: Ctx, st: dict, steps: int) -> dict:
    """Decimated outputs of the headline run collected on rank 0 (opt-in in the API; reported apart from `value`).
    (a) NCCL send/recv, grouped, every shard straight into its final offset of rank 0's buffer (no padding, no
    concatenation); (b) fused: each rank's FIR kernel stores its outputs directly into rank 0's buffer through a
    peer mapping (CUDA IPC) over NVLink — the gather costs no extra pass."""
    t, g, gd = c.torch, c.g, c.gd
    sh, D, T = st["sh"], WORKLOADS["cfg2"]["D"], WORKLOADS["cfg2"]["T"]
    n_out_total = st["n_out_total"]
    counts = [g.shard_plan_time(n_out_total, D, T, 0, c.world, r).numOutputs for r in range(c.world)]
    firsts = [g.shard_plan_time(n_out_total, D, T, 0, c.world, r).firstOutput for r in range(c.world)]
    out = {}
    # (a) NCCL P2P
    full = t.zeros(n_out_total, dtype=t.complex64, device=c.dev) if c.rank == 0 else None
    gd.gather_outputs_p2p(st["y_timed"], counts, firsts, full)
    c.barrier()
    reps = 5
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gd.gather_outputs_p2p(st["y_timed"], counts, firsts, full)
    e1.record()
    c.barrier()
    ms = c.max_ranks(e0.elapsed_time(e1) / reps)
    ok = True
    if c.rank == 0:
        ok = bool(t.equal(full[firsts[0]: firsts[0] + counts[0]], st["y_timed"]))
        probe = full[firsts[-1] + counts[-1] - 1].item()
        ok = ok and probe != 0
    out["nccl_p2p"] = {"ms": ms, "bytes_to_rank0": 8 * (n_out_total - counts[0]),
                       "gb_per_s": 8 * (n_out_total - counts[0]) / (ms * 1e-3) / 1e9, "rank0_block_matches": ok,
                       "how": "torch.distributed.batch_isend_irecv (ncclGroupStart / ncclSend / ncclRecv / "
                              "ncclGroupEnd), receives land at their final offsets"}
    # (b) fused: kernels store through a peer mapping of rank 0's output buffer
    try:
        fused = gd.PeerOutput(n_out_total * 8, c.rank, c.local, c.dev)
        dst = fused.ptr + firsts[c.rank] * 8
        dtaps, x = st["dtaps"], st["x"]

        def step():
            g.gsdrFirFC(D, dtaps, T, x, dst, sh.numOutputs, c.local, c.stream)
        mine, ms_f, clocks = c.time_steps(step, steps, 3)
        same = True
        if c.rank == 0:
            whole = fused.as_tensor(n_out_total)
            same = bool(t.equal(whole, full))
        out["fused_peer_store"] = {"ms_per_step": ms_f, "value": st["n_in_total"] / (ms_f * 1e-3) / 1e6, "unit": UNIT,
                                   "equals_nccl_gather": same, "clocks": clocks,
                                   "how": "cudaIpc mapping of rank 0's output buffer in every rank; each rank's "
                                          "gsdrFirFC writes its block at its final offset over NVLink, so compute and "
                                          "gather are one kernel per rank"}
        c.barrier()
        fused.close()
    except Exception as e:  # noqa: BLE001 - report, do not lose the headline
        out["fused_peer_store"] = {"error": repr(e)}
        c.barrier()
    return out


def e2e_block(c: Ctx, st: dict, steps: int, nco: bool) -> dict:
    """The same metric through the host-buffer API: pinned host memory in and out, copies inside the timed region."""
    t, g = c.torch, c.g
    sh, taps = st["sh"], st["taps"]
    D, T = WORKLOADS[c.args.workload]["D"], WORKLOADS[c.args.workload]["T"]
    n = max(2, min(steps, 5))
    xin = t.empty(sh.numInputs, dtype=t.complex64).pin_memory()
    xin.copy_(st["x"])
    yout = t.zeros(sh.numOutputs, dtype=t.complex64).pin_memory()
    pipe = g.HostPipeline(c.local, chunkInputBytes=32 << 20, numBuffers=3)
    if nco:
        def e2e_step():
            pipe.gsdrAdjustFrequencyFirFCHost(bl.NCO_FS, bl.NCO_SHIFT, sh.firstSampleIndex, D, taps, T, xin, yout,
                                              sh.numOutputs)
    else:
        def e2e_step():
            pipe.gsdrFirFCHost(D, taps, T, xin, yout, sh.numOutputs)
    e2e_step()
    c.barrier()
    t0 = time.perf_counter()
    for _ in range(n):
        e2e_step()
    t.cuda.synchronize(c.dev)
    dt = c.max_ranks(time.perf_counter() - t0)
    ok = bool(t.equal(yout.to(c.dev), st["y_timed"]))
    # What the platform gives: the same input bytes as plain pinned-host -> device copies, all ranks at once, no kernel.
    # e2e within a few percent of this = the host's PCIe / memory system is the limit, not the pipeline or the GPUs.
    dst = t.empty_like(st["x"])
    with t.cuda.stream(c.stream):
        dst.copy_(xin, non_blocking=True)
    c.barrier()
    t0 = time.perf_counter()
    with t.cuda.stream(c.stream):
        for _ in range(n):
            dst.copy_(xin, non_blocking=True)
    c.stream.synchronize()
    dt_copy = c.max_ranks(time.perf_counter() - t0)
    del dst
    copy_gbs = c.world * xin.numel() * 8 * n / dt_copy / 1e9
    res = {"value": st["n_in_total"] / (dt / n) / 1e6, "unit": UNIT, "h2d_bytes_per_step": xin.numel() * 8 + T * 4,
           "d2h_bytes_per_step": yout.numel() * 8, "ms_per_step": dt / n * 1e3, "steps": n,
           "timing": "host wall clock around the blocking API call (copies + kernels + sync), max over ranks",
           "matches_device_path": ok, "numa_node_rank0": c.numa_node,
           "h2d_gb_s_all_ranks": c.world * (xin.numel() * 8) / (dt / n) / 1e9,
           "plain_h2d_copy_gb_s_all_ranks": copy_gbs,
           "fraction_of_plain_copy": (c.world * (xin.numel() * 8) / (dt / n) / 1e9) / copy_gbs}
    # the same capture as int8 I/Q (2 bytes per sample over PCIe instead of 8) through gsdrFirFCInt8Host
    if not nco and hasattr(pipe, "gsdrFirFCInt8Host"):
        xi8 = t.view_as_real(st["x"]).mul(127.0).round().clamp(-127, 127).to(t.int8).reshape(-1)
        hin = t.empty(xi8.numel(), dtype=t.int8).pin_memory()
        hin.copy_(xi8)
        y8 = t.zeros(sh.numOutputs, dtype=t.complex64).pin_memory()
        pipe.gsdrFirFCInt8Host(D, taps, T, hin, y8, sh.numOutputs)
        c.barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            pipe.gsdrFirFCInt8Host(D, taps, T, hin, y8, sh.numOutputs)
        t.cuda.synchronize(c.dev)
        dt8 = c.max_ranks(time.perf_counter() - t0)
        dy8 = t.zeros(sh.numOutputs, dtype=t.complex64, device=c.dev)
        g.gsdrFirFCInt8(D, st["dtaps"], T, xi8, dy8, sh.numOutputs, c.local, c.stream)
        c.stream.synchronize()
        res["int8_input"] = {"value": st["n_in_total"] / (dt8 / n) / 1e6, "unit": UNIT,
                             "h2d_bytes_per_step": hin.numel() + T * 4, "d2h_bytes_per_step": y8.numel() * 8,
                             "ms_per_step": dt8 / n * 1e3, "matches_device_path": bool(t.equal(y8.to(c.dev), dy8))}
    pipe.close()
    return res


def _run_ours(args):
    c = Ctx(args)
    t, g = c.torch, c.g
    if args.variant != -1:
        g.set_kernel_variant(args.variant)
    t_start = time.perf_counter()
    name = args.workload
    wl = WORKLOADS[name]

    # ---- headline ----
    if name == "cfg1":
        head = case_cfg1(c, args.steps, args.warmup)
    elif name == "cfg4":
        head = case_cfg4(c, args.steps, args.warmup)
    elif name == "cfg5":
        head = case_cfg5(c, args.steps, args.warmup)
    else:
        head = case_fc(c, name, args.steps, args.warmup, keep=True)
    st = head.pop("_state", None)

    e2e = None
    if st is not None and not args.no_e2e:
        e2e = e2e_block(c, st, args.steps, wl["nco"])

    extras = {}
    if not args.no_others and name == "cfg2":
        def agreed_time_left() -> bool:
            left = 1.0 if time.perf_counter() - t_start < args.others_budget_s else 0.0
            if c.dist_on:
                flag = t.tensor([left], device=c.dev)
                t.distributed.broadcast(flag, src=0)
                left = float(flag.item())
            return left > 0.5

        if c.dist_on:
            for key, fn in (("gather", lambda: gather_block(c, st, min(args.steps, 20))),
                            ("strong", lambda: strong_scaling(c, min(args.steps, 50), args.warmup))):
                if agreed_time_left():
                    extras[key] = fn()
        st = None
        t.cuda.empty_cache()
        others = {}
        k_other = max(3, min(args.steps, 10))
        for key, fn in (("cfg3", lambda: case_fc(c, "cfg3", k_other, 3)),
                        ("cfg5", lambda: case_cfg5(c, k_other, 3)),
                        ("cfg4", lambda: case_cfg4(c, max(3, min(args.steps, 5)), 3)),
                        ("cfg1", (lambda: case_cfg1(c, 50, 3)) if c.world == 1 else None)):
            if fn is None:
                continue
            if not agreed_time_left():
                others[key] = {"skipped": f"time budget of {args.others_budget_s:.0f} s used up"}
                continue
            try:
                others[key] = fn()
            except Exception as e:  # noqa: BLE001 - keep the headline
                others[key] = {"error": repr(e)}
                if c.dist_on:
                    raise
            t.cuda.empty_cache()
        extras["other_configs"] = others

    if c.rank != 0:
        if c.dist_on:
            t.distributed.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu and c.world == 1 and wl["kind"] == "fc":
        cpu = bl.cpu_baseline(wl["D"], wl["T"], synth.lowpass_taps(wl["T"], wl["D"]), wl["n_in"])
    launches = head.get("gpu_launches", args.steps)
    for v in extras.get("other_configs", {}).values():
        launches += v.get("gpu_launches", 0) if isinstance(v, dict) else 0
    line = {"metric": bl.METRIC, "value": head["value"], "unit": UNIT, "n_gpus": c.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": head["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "ours",
            "config": head["config"], "kernel": head.get("kernel"), "roofline": head["roofline"],
            "parity": head["parity"], "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": head.get("gpu_launches", args.steps),
            "gpu_launches_all_blocks": launches, "clocks": head["clocks"],
            "library": {"path": str(Path(g.library_path()).relative_to(ROOT)), "tuning_hooks": g.has_tuning_hooks()},
            "wall_s": None}
    for k, v in head.items():
        if k not in line and k not in ("unit",):
            line[k] = v
    line.update(extras)
    line["wall_s"] = time.perf_counter() - t_start
    bl.emit(line)
    if c.dist_on:
        t.distributed.destroy_process_group()


def main() -> None:
    bl.capture_stdout()
    args = _parse()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "cpu":
        if rank == 0:
            _run_cpu(args)
    elif args.impl == "reference":
        if rank == 0:  # the reference is single-GPU: rank 0 alone runs it, the others exit 0 without work
            _run_reference(args)
    else:
        _run_ours(args)


if __name__ == "__main__":
    main()
