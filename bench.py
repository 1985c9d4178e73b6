#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: input Msamples/s of the 255-tap decimate-by-8 complex FIR (gsdrFirFC).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|cpu] [--gather] [--workload cfg2|cfg3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one capture: 2^26 cuComplex samples per GPU (BASELINE config 2).  At
N > 1 the capture is N*2^26 samples long and time-sharded with gsdrShardPlanTime: every rank holds its own block
plus the (taps - decimation)-sample overlap resident in HBM, so there is no collective on the compute path
(weak scaling).  Timing is on the device (CUDA events on the launching stream), K steps bracketed by a barrier +
synchronize, max over ranks.  The input (537 MB per GPU) is larger than the 126 MB L2, so every step streams
from HBM.

impl:
  ours       libgsdr_b200.so through the C ABI (`value`: device-resident buffers; `e2e`: pinned host buffers in,
             pinned host buffers out through gsdrFirFCHost, H2D/D2H inside the timed region).
  reference  the reference's OWN CUDA kernels (oracle/_ref/libgsdr_ref.so, compiled for sm_100 from
             /root/reference/src/fir.cu by oracle/build_ref.sh) on the same buffers, rank 0 only.  gsdr has no
             CPU implementation; its e2e is what a user of the reference must do: cudaMemcpy in, gsdrFirFC,
             cudaMemcpy out.
  cpu        the scalar C oracle (restating ref: src/fir.cu:57-70) on all host cores, on a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "input Msamples/s, 255-tap decim-8 complex FIR at 1/2/4/8 B200; % roofline"
UNIT = "Msamples/s"
WORKLOADS = {
    # name: (decimation, taps, input samples per GPU, nco)
    "cfg2": dict(D=8, T=255, n_in=1 << 26, nco=False,
                 desc="complex FIR, 255 real taps, decimation 8, 64Mi cuComplex samples per GPU (BASELINE config 2)"),
    "cfg3": dict(D=32, T=1023, n_in=1 << 28, nco=True,
                 desc="fused NCO mix + 1023-tap decimate-by-32, 256Mi samples per GPU (BASELINE config 3)"),
    "cfg3-nomix": dict(D=32, T=1023, n_in=1 << 28, nco=False,
                       desc="1023-tap decimate-by-32 complex FIR without the NCO, 256Mi samples per GPU"),
    "cfg5s1": dict(D=10, T=255, n_in=1 << 28, nco=True,
                   desc="fused NCO mix + 255-tap decimate-by-10 (BASELINE config 5 stage 1 shape), 256Mi samples per GPU"),
    # the two below have their own step functions (see main)
    "cfg4": dict(D=4, T=127, n_in=1 << 22, nco=False, channels=1024,
                 desc="1024 independent channels x 4Mi samples, 127-tap decimate-by-4, sharded by channel (BASELINE "
                      "config 4; channels are split over the ranks: strong scaling)"),
    "cfg5": dict(D=10, T=255, n_in=1 << 28, nco=True, chain=dict(D3=5, T3=63),
                 desc="FM receive chain: NCO mix -> 255-tap FIR decim 10 -> quad demod -> 63-tap audio FIR decim 5, "
                      "256Mi samples per GPU of one capture, time-sharded with the 835-sample halo (BASELINE config 5)"),
}
PAPER_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 74.45


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTED = {"sw_power_cap": 0x4}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _once(self):
        nv = self._nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(
            nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for name, bit in {**self.BAD, **self.NOTED}.items():
            if mask & bit:
                self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._once()
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        if self._nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            try:
                self._once()
            except Exception:
                pass
            self._stop.set()
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def _fp32_peak(device_index: int):
    lib_path = ROOT / "tools" / "libubench_fp32.so"
    if not lib_path.exists():
        return None, None
    lib = ctypes.CDLL(str(lib_path))
    lib.ubenchFp32Tflops.restype = ctypes.c_double
    lib.ubenchFp32Tflops.argtypes = [ctypes.c_int] * 5
    ffma = lib.ubenchFp32Tflops(0, device_index, 4000, 3, 4)
    ffma2 = lib.ubenchFp32Tflops(1, device_index, 4000, 3, 4)
    return (ffma if ffma > 0 else None), (ffma2 if ffma2 > 0 else None)


def _cpu_baseline(D, T, taps, n_in_full, target_seconds=10.0):
    """Scalar C oracle (restating ref: src/fir.cu:57-70) on all host cores over a bounded sample of the workload:
    the first min(workload, 2^26) input samples, repeated until ~target_seconds of CPU work has been timed."""
    from gsdr_b200 import synth
    from oracle import oracle

    cores = os.cpu_count() or 1
    n_in = int(min(n_in_full, 1 << 26))
    n_out = (n_in - T) // D + 1
    x = synth.tone_plus_noise(0, n_in, seed=0x5EED0002)
    oracle.fir("fc", D, taps, x[: 1 << 20], threads=cores)  # page in, spin up
    reps, total = 0, 0.0
    while total < target_seconds and reps < 200:
        t0 = time.perf_counter()
        oracle.fir("fc", D, taps, x, n_out, threads=cores)
        total += time.perf_counter() - t0
        reps += 1
    dt = total / reps
    return {"value": n_in / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {n_in} input samples ({n_out} outputs) of the workload, {reps} passes of {dt:.3f} s, "
                      f"{cores} pthreads over contiguous output blocks, gcc -O2 -mfma scalar fmaf chain in the "
                      f"reference's accumulation order"}


# stdout carries exactly ONE JSON line.  Libraries loaded later write there too (NCCL prints its version line to fd 1
# whatever NCCL_DEBUG_FILE says), so fd 1 is pointed at stderr for the life of the process and the JSON line goes to the
# saved original.
_REAL_STDOUT = None


def _capture_stdout() -> None:
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(text: str) -> None:
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        print(text, flush=True)
    else:
        os.write(_REAL_STDOUT, (text + "\n").encode())


def main() -> None:
    _capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference", "cpu"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="cfg2")
    ap.add_argument("--gather", action="store_true", help="also time the optional NCCL gather of outputs to rank 0")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--variant", type=int, default=-1, help="force a polyphase kernel variant (tuning)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    wl = WORKLOADS[args.workload]
    D, T, n_in_gpu = wl["D"], wl["T"], wl["n_in"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    import numpy as np

    from gsdr_b200 import synth

    taps = synth.lowpass_taps(T, D)

    if args.impl == "cpu":
        if rank == 0:
            cb = _cpu_baseline(D, T, taps, n_in_gpu, target_seconds=15.0)
            _emit(json.dumps({"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": 0, "steps": 1, "warmup": 0,
                              "impl": "cpu", "higher_is_better": True, "dtype": "f32", "data": "synthetic",
                              "config": {"workload": wl["desc"]}, "cpu_baseline": cb,
                              "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                                      "d2h_bytes_per_step": 0}}))
        return
    if args.impl == "reference" and rank != 0:
        return  # the reference is single-GPU: rank 0 alone runs it

    import torch

    import gsdr_b200 as g
    from gsdr_b200 import dist as gd

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    # stdout carries exactly one JSON line: NCCL's version / debug lines go to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if args.impl == "reference":
        world = 1
    else:
        rank, world, local = gd.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist_on = world > 1
    # several ranks: keep each rank's pinned host buffers on the NUMA node of its GPU (e2e is host-memory bound)
    numa_node = gd.bind_to_gpu_numa_node(local) if dist_on else None

    # ---- this rank's block of the capture, resident in HBM before timing starts ----
    n_in_total = n_in_gpu * world
    n_out_total = g.fir_num_outputs(n_in_total, T, D)
    sh = g.shard_plan_time(n_out_total, D, T, 0, world, rank)
    x = synth.tone_plus_noise(sh.firstInput, sh.numInputs, seed=0x5EED0002, device=dev)
    dtaps = torch.from_numpy(taps).to(dev)
    y = torch.zeros(sh.numOutputs, dtype=torch.complex64, device=dev)
    stream = torch.cuda.Stream(device=dev)
    fs, fshift = 2.4e6, 29520.0

    special = None
    if args.impl == "ours" and args.workload == "cfg4":
        # ---- BASELINE config 4: channels sharded across ranks, one batched launch per rank ----
        chans_total = wl["channels"]
        c0, cn = g.shard_plan_channels(chans_total, world, rank)
        n_out_c = g.fir_num_outputs(n_in_gpu, T, D)
        base = synth.tone_plus_noise(0, n_in_gpu * 8, seed=0x5EED0004, device=dev).view(8, n_in_gpu)
        xb = torch.empty((cn, n_in_gpu), dtype=torch.complex64, device=dev)
        for c in range(cn):
            xb[c].copy_(base[(c0 + c) % 8])
            xb[c, : 1024] *= 1.0 + 0.001 * (c0 + c)  # channels are not bit-identical copies
        yb = torch.zeros((cn, n_out_c), dtype=torch.complex64, device=dev)
        del x, y, base

        def step():
            g.gsdrFirFCBatched(D, dtaps, T, 0, xb, n_in_gpu, yb, n_out_c, n_out_c, cn, local, stream)

        n_in_total = n_in_gpu * chans_total
        n_out_total = n_out_c * chans_total
        special = dict(units_in=cn * n_in_gpu, units_out=cn * n_out_c, scaling="strong",
                       sharding=f"channels: rank {rank} owns {cn} of {chans_total}; one batched launch per rank")
    elif args.impl == "ours" and args.workload == "cfg5":
        # ---- BASELINE config 5: the composite stage is a FIR with window 885 and stride 50 for planning ----
        D3, T3 = wl["chain"]["D3"], wl["chain"]["T3"]
        window, stride = D * T3 + T, D * D3  # 885, 50
        n3_total = (n_in_total - window) // stride + 1
        sh3 = g.shard_plan_time(n3_total, stride, window, 0, world, rank)
        n3 = sh3.numOutputs
        n2 = g.fir_num_inputs(n3, T3, D3)
        del x, y
        x5 = synth.tone_plus_noise(sh3.firstInput, sh3.numInputs, seed=0x5EED0005, device=dev,
                                   tone_cycles_per_sample=300e3 / 2.4e6)
        h3 = torch.from_numpy(synth.lowpass_taps(T3, D3)).to(dev)
        dm = torch.zeros(n2, dtype=torch.float32, device=dev)
        au = torch.zeros(n3, dtype=torch.float32, device=dev)

        def step():
            g.gsdrFmDemod(2.4e6, 100.0e6, 100.3e6, 75e3, D, sh3.firstSampleIndex, dtaps, T, x5, dm, n2, local, stream)
            g.gsdrFirFF(D3, h3, T3, dm, au, n3, local, stream)

        n_out_total = n3_total
        special = dict(units_in=sh3.numInputs, units_out=n2 + 1, scaling="weak",
                       sharding=f"time blocks of the final audio outputs; halo {window - stride} input samples")
        args.no_e2e = True
    if special is not None:
        args.no_e2e = True
        args.gather = False

    if special is not None:
        pass
    elif args.impl == "reference":
        from oracle import ref_cuda

        if not ref_cuda.available():
            _emit(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libgsdr_ref.so was not built "
                                                                  "(needs /root/reference at build time)"}))
            return
        if wl["nco"]:
            def step():
                ref_cuda.adjust_frequency_fir_fc(fs, fshift, sh.firstSampleIndex, D, dtaps, T, x, y, sh.numOutputs,
                                                 local, stream.cuda_stream)
        else:
            def step():
                ref_cuda.fir("fc", D, dtaps, T, x, y, sh.numOutputs, local, stream.cuda_stream)
    else:
        g.set_kernel_variant(args.variant)
        if wl["nco"]:
            def step():
                g.gsdrAdjustFrequencyFirFC(fs, fshift, sh.firstSampleIndex, D, dtaps, T, x, y, sh.numOutputs, local,
                                           stream)
        else:
            def step():
                g.gsdrFirFC(D, dtaps, T, x, y, sh.numOutputs, local, stream)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist_on:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    sampler.stop()
    ms_total = gd.max_over_ranks(e0.elapsed_time(e1), device=dev) if dist_on else e0.elapsed_time(e1)
    ms_step = ms_total / args.steps
    value = n_in_total / (ms_step * 1e-3) / 1e6

    # ---- roofline of the dominant (only) kernel: algorithmic bytes and flops per launch ----
    hbm_peak, hbm_src = _peaks()
    units_in = special["units_in"] if special else sh.numInputs
    units_out = special["units_out"] if special else sh.numOutputs
    bytes_alg = 8 * units_in + 8 * units_out + 4 * T  # each input once, each output once, taps once
    flops_alg = 4.0 * T * units_out                   # FC: 4*T flops per complex output (dominant kernel)
    kernel_s = (e0.elapsed_time(e1) / args.steps) * 1e-3      # this rank's average launch duration
    ffma_tf, ffma2_tf = _fp32_peak(local)
    fp32_peak = max([v for v in (ffma_tf, ffma2_tf) if v] or [PAPER_FP32_TFLOPS])
    t_mem, t_fp = bytes_alg / (hbm_peak * 1e9), flops_alg / (fp32_peak * 1e12)
    ach_gbs, ach_tf = bytes_alg / kernel_s / 1e9, flops_alg / kernel_s / 1e12
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get(args.workload)
        except Exception:
            traffic = None
    if t_fp >= t_mem:
        roof = {"bound": "fp32", "achieved": ach_tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach_tf / fp32_peak}
    else:
        roof = {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak}
    roof.update({
        "traffic": traffic,
        "kernel_us": kernel_s * 1e6,
        "roofline_us": max(t_mem, t_fp) * 1e6,
        "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak, "peak_source": hbm_src},
        "fp32": {"achieved": ach_tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach_tf / fp32_peak,
                 "peak_source": "FFMA/FFMA2 microbenchmark run in this process (tools/ubench_fp32.cu)",
                 "ffma_tflops": ffma_tf, "ffma2_tflops": ffma2_tf, "paper_peak": PAPER_FP32_TFLOPS,
                 "frac_of_paper": ach_tf / PAPER_FP32_TFLOPS},
        "algorithmic_bytes": bytes_alg, "algorithmic_flops": flops_alg,
    })

    # ---- optional gather of the decimated outputs to rank 0 (NCCL over NVLink), reported separately ----
    gather_ms = None
    if args.gather and dist_on and args.impl == "ours":
        counts = [g.shard_plan_time(n_out_total, D, T, 0, world, r).numOutputs for r in range(world)]
        gd.gather_outputs(y, counts, dst=0)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        gd.gather_outputs(y, counts, dst=0)
        g1.record()
        barrier()
        gather_ms = gd.max_over_ranks(g0.elapsed_time(g1), device=dev)

    # ---- end to end: pinned host buffers in and out, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(2, min(args.steps, 5))
        xin = torch.empty(sh.numInputs, dtype=torch.complex64).pin_memory()
        xin.copy_(x)
        yout = torch.zeros(sh.numOutputs, dtype=torch.complex64).pin_memory()
        h2d, d2h = xin.numel() * 8 + T * 4, yout.numel() * 8
        if args.impl == "ours":
            pipe = g.HostPipeline(local, chunkInputBytes=32 << 20, numBuffers=3)
            if wl["nco"]:
                def e2e_step():
                    pipe.gsdrAdjustFrequencyFirFCHost(fs, fshift, sh.firstSampleIndex, D, taps, T, xin, yout,
                                                      sh.numOutputs)
            else:
                def e2e_step():
                    pipe.gsdrFirFCHost(D, taps, T, xin, yout, sh.numOutputs)
        else:
            htaps = torch.from_numpy(taps).pin_memory()

            def e2e_step():
                with torch.cuda.stream(stream):
                    dtaps.copy_(htaps, non_blocking=True)
                    x.copy_(xin, non_blocking=True)
                    step()
                    yout.copy_(y, non_blocking=True)
                stream.synchronize()
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        dt = gd.max_over_ranks(dt, device=dev) if dist_on else dt
        ok = bool(torch.equal(yout.to(dev), y)) if args.impl == "ours" else True
        e2e = {"value": n_in_total / (dt / e2e_steps) / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": dt / e2e_steps * 1e3, "steps": e2e_steps,
               "timing": "host wall clock around the blocking API call (copies + kernels + sync), max over ranks",
               "matches_device_path": ok, "numa_node_rank0": numa_node}

    if rank != 0:
        if dist_on:
            torch.distributed.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu and world == 1:
        cpu = _cpu_baseline(D, T, taps, n_in_gpu)

    info = g.describe_kernel(4 if wl["nco"] else 0, D, T, sh.numOutputs, local) if args.impl == "ours" else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": special["scaling"] if special else "weak",
        "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": args.impl,
        "config": {
            "workload": wl["desc"], "decimation": D, "taps": T, "input_samples_per_gpu": n_in_gpu,
            "input_samples_total": n_in_total, "outputs_total": n_out_total,
            "sharding": special["sharding"] if special else (
                "time blocks of one capture, (taps-decimation)-sample overlap resident per rank, no collective"
                if world > 1 else "single GPU"),
            "l2": "input (537 MB/GPU) larger than the 126 MB L2; no explicit flush",
            "timing": "CUDA events on the launching stream around K back-to-back launches, max over ranks",
            "kernel": (f"{'TMA-fed' if info.variant >= g.num_polyphase_variants() else 'cp.async-staged'} persistent "
                       f"polyphase kernel, variant {info.variant}: {info.threadsPerBlock} threads/CTA, "
                       f"{info.outputsPerThread} outputs/thread, {info.phaseGroups} branch groups, "
                       f"{info.sharedBytesPerBlock} B smem, {info.numBlocks} tiles") if info else
            "reference k_FirDecimate<float2,float2,float> (32-thread blocks, one thread per output)",
        },
        "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": args.steps * (3 if args.workload == "cfg5" and args.impl == "ours" else 1),
        "clocks": sampler.summary(),
    }
    if gather_ms is not None:
        line["gather_ms"] = gather_ms
    _emit(json.dumps(line))
    if dist_on:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
