"""Shared pieces of bench.py that load NO product code: workload table, peaks, clock sampling, stdout discipline and
the oracle spot check of what a timed run produced.  `bench.py --impl reference|cpu` imports only this module, the
rest of harness/, oracle/ and torch — never gsdr_b200 (whose import maps libgsdr_b200.so).
"""
from __future__ import annotations

import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent

METRIC = "input Msamples/s, 255-tap decim-8 complex FIR at 1/2/4/8 B200; % roofline"
UNIT = "Msamples/s"
PAPER_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 74.45
NCO_FS, NCO_SHIFT = 2.4e6, 29520.0

# BASELINE.json `configs`, binary sizes.  n_in is per GPU for the weak-scaled workloads.
WORKLOADS = {
    "cfg1": dict(kind="ff", D=1, T=63, n_in=1 << 20, nco=False, scaling="single",
                 desc="real float FIR, 63 taps, decimation 1, 1Mi samples (BASELINE config 1; one launch, single GPU)"),
    "cfg2": dict(kind="fc", D=8, T=255, n_in=1 << 26, nco=False, scaling="weak",
                 desc="complex FIR, 255 real taps, decimation 8, 64Mi cuComplex samples per GPU (BASELINE config 2)"),
    "cfg3": dict(kind="fc", D=32, T=1023, n_in=1 << 28, nco=True, scaling="weak",
                 desc="fused NCO mix + 1023-tap decimate-by-32, 256Mi samples per GPU (BASELINE config 3)"),
    "cfg3-nomix": dict(kind="fc", D=32, T=1023, n_in=1 << 28, nco=False, scaling="weak",
                       desc="1023-tap decimate-by-32 complex FIR without the NCO, 256Mi samples per GPU"),
    "cfg5s1": dict(kind="fc", D=10, T=255, n_in=1 << 28, nco=True, scaling="weak",
                   desc="fused NCO mix + 255-tap decimate-by-10 (BASELINE config 5 stage 1 shape), 256Mi samples per GPU"),
    "cfg4": dict(kind="fc", D=4, T=127, n_in=1 << 22, nco=False, channels=1024, scaling="strong",
                 desc="1024 independent channels x 4Mi samples, 127-tap decimate-by-4, sharded by channel (BASELINE "
                      "config 4; the 1024 channels are split over the ranks: strong scaling)"),
    "cfg5": dict(kind="chain", D=10, T=255, n_in=1 << 28, nco=True, chain=dict(D3=5, T3=63), scaling="weak",
                 desc="FM receive chain: NCO mix -> 255-tap FIR decim 10 -> quad demod -> 63-tap audio FIR decim 5, "
                      "256Mi samples per GPU of one capture (2Gi at 8 GPUs), time-sharded with the 835-sample halo "
                      "(BASELINE config 5)"),
}


def config_block(name: str, world: int) -> dict:
    """The `config` object of the JSON line: identical for every --impl of the same workload and GPU count."""
    wl = WORKLOADS[name]
    n_total = wl["n_in"] * (wl.get("channels", 1) if name == "cfg4" else world)
    return {
        "workload": wl["desc"], "name": name, "decimation": wl["D"], "taps": wl["T"],
        "input_samples_per_gpu": n_total // world if name == "cfg4" else wl["n_in"], "input_samples_total": n_total,
        "sharding": ("single GPU" if world == 1 else
                     "channels split over the ranks, one batched launch per rank, no collective" if name == "cfg4" else
                     "time blocks of one capture, (taps-decimation)-sample overlap resident per rank, no collective"),
        "l2": "per-GPU input is larger than the 126 MB L2; no explicit flush" if wl["n_in"] * wl.get("channels", 1) * 8 >
              (126 << 20) else "input fits the L2: an L2 flush (256 MiB write) runs between timed launches",
        "timing": "CUDA events on the launching stream around K back-to-back steps, max over ranks",
        "units": "binary sizes: 1Mi = 2^20 samples",
    }


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def roofline(bytes_alg: float, flops_alg: float, kernel_s: float, fp32_peak_tf: float, fp32_src: str,
             traffic=None, tensor_core: bool = False) -> dict:
    """The slower of (bytes / HBM peak) and (flops / FP32 peak) bounds the kernel (BASELINE.json north_star).
    tensor_core: the multiply-accumulates run on the tensor cores (fir_tc_kernel.cuh), whose FP16 rate is ~30 x the
    FP32 pipes': the FP32 peak does not bound that kernel, HBM does."""
    hbm, hbm_src = hbm_peak()
    t_mem, t_fp = bytes_alg / (hbm * 1e9), flops_alg / (fp32_peak_tf * 1e12)
    ach_gbs, ach_tf = bytes_alg / kernel_s / 1e9, flops_alg / kernel_s / 1e12
    if tensor_core:
        t_fp = 0.0
    if t_fp >= t_mem:
        r = {"bound": "fp32", "achieved": ach_tf, "peak": fp32_peak_tf, "unit": "TFLOP/s", "frac": ach_tf / fp32_peak_tf}
    else:
        r = {"bound": "hbm", "achieved": ach_gbs, "peak": hbm, "unit": "GB/s", "frac": ach_gbs / hbm}
    r.update({"traffic": traffic, "kernel_us": kernel_s * 1e6, "roofline_us": max(t_mem, t_fp) * 1e6,
              "hbm": {"achieved": ach_gbs, "peak": hbm, "unit": "GB/s", "frac": ach_gbs / hbm, "peak_source": hbm_src},
              "fp32": {"achieved": ach_tf, "peak": fp32_peak_tf, "unit": "TFLOP/s", "frac": ach_tf / fp32_peak_tf,
                       "peak_source": fp32_src, "paper_peak": PAPER_FP32_TFLOPS,
                       "frac_of_paper": ach_tf / PAPER_FP32_TFLOPS},
              "algorithmic_bytes": bytes_alg, "algorithmic_flops": flops_alg})
    if tensor_core:
        r["fp32"]["note"] = ("informational: this kernel's multiply-accumulates run on the tensor cores (tcgen05 "
                             "kind::f16), so the FP32 FFMA peak does not bound it; the FFMA2 kernel it replaced on "
                             "this shape was FP32-issue bound")
    return r


def fp32_peak(device_index: int):
    """FFMA / FFMA2 peak measured in this process (tools/ubench_fp32.cu; MEASURED_PEAKS.json has no FP32 entry)."""
    import ctypes

    lib_path = ROOT / "tools" / "libubench_fp32.so"
    if not lib_path.exists():
        return PAPER_FP32_TFLOPS, "paper peak (tools/libubench_fp32.so not built)", None, None
    lib = ctypes.CDLL(str(lib_path))
    lib.ubenchFp32Tflops.restype = ctypes.c_double
    lib.ubenchFp32Tflops.argtypes = [ctypes.c_int] * 5
    ffma = lib.ubenchFp32Tflops(0, device_index, 4000, 3, 4)
    ffma2 = lib.ubenchFp32Tflops(1, device_index, 4000, 3, 4)
    vals = [v for v in (ffma, ffma2) if v and v > 0]
    if not vals:
        return PAPER_FP32_TFLOPS, "paper peak (microbenchmark failed)", None, None
    return max(vals), "FFMA/FFMA2 microbenchmark run in this process (tools/ubench_fp32.cu)", ffma, ffma2


class ClockSampler:
    """Samples SM clock, power and throttle reasons with NVML while a timed region runs."""

    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTED = {"sw_power_cap": 0x4}

    def __init__(self, index: int):
        self.samples, self.power, self.reasons, self.max_mhz = [], [], set(), None
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
        try:
            self.power.append(nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
        except Exception:
            pass
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
            time.sleep(0.001)

    def start(self):
        if self._nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def begin(self):
        """Forget what was sampled so far (warm-up): the summary covers the timed region only.  The thread is started
        BEFORE the warm-up so that nothing but this call sits between the barrier and the first timed launch."""
        self.samples, self.power, self.reasons = [], [], set()
        return self

    def stop(self):
        if self._thr is not None:
            try:
                self._once()
            except Exception:
                pass
            self._stop.set()
            self._thr.join()
        return self

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_min_mhz": min(self.samples),
                "sm_max_mhz": self.max_mhz, "power_w_max": max(self.power) if self.power else None,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# stdout carries exactly ONE JSON line.  Libraries loaded later write there too (NCCL prints its version line to fd 1
# whatever NCCL_DEBUG_FILE says), so fd 1 is pointed at stderr for the life of the process and the JSON line goes to
# the saved original.
_REAL_STDOUT = None


def capture_stdout() -> None:
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj) -> None:
    text = obj if isinstance(obj, str) else json.dumps(obj)
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        print(text, flush=True)
    else:
        os.write(_REAL_STDOUT, (text + "\n").encode())


# ---- oracle spot checks of what was timed -------------------------------------------------------------------------

def _windows(n_out: int, width: int):
    width = min(width, n_out)
    starts = sorted({0, max(0, (n_out - width) // 2), n_out - width})
    return [(s, width) for s in starts]


def check_fir_windows(kind: str, D: int, taps: np.ndarray, x_dev, y_dev, n_out: int, nco=None, width: int = 512) -> dict:
    """Compares prefix / middle / suffix windows of the device output `y_dev` (what the timed launches wrote) with the
    double-precision oracle evaluated on the device input `x_dev` copied back (ref: src/fir.cu:57-70; the tolerance is
    BASELINE.json's: max|err| <= 1e-5 * sum|h| * max|x|).  nco = (sampleRate, frequencyShift, firstSampleIndex)."""
    from oracle import oracle

    T = int(taps.shape[0])
    max_err, max_x = 0.0, 0.0
    for o0, w in _windows(n_out, width):
        n_in = (w - 1) * D + T
        xw = x_dev[o0 * D: o0 * D + n_in].cpu().numpy()
        yw = y_dev[o0: o0 + w].cpu().numpy()
        if nco is not None:
            fs, shift, first = nco
            ref = oracle.adjust_frequency_fir_fc(oracle.NCO_EXACT, fs, shift, first + o0 * D, D, taps, xw, w, f64=True)
        else:
            ref = oracle.fir(kind, D, taps, xw, w, f64=True)
        max_err = max(max_err, float(np.abs(yw.astype(ref.dtype) - ref).max()))
        max_x = max(max_x, float(np.abs(xw).max()))
    tol = 1e-5 * float(np.abs(taps).sum()) * max_x
    return {"max_err": max_err, "tol": tol, "ok": bool(max_err <= tol), "windows": len(_windows(n_out, width)),
            "outputs_per_window": min(width, n_out), "against": "oracle f64 on the device input copied back"}


def check_chain_windows(D1, h1, D3, h3, fs, shift, first, gain, x_dev, au_dev, n3: int, width: int = 256) -> dict:
    """FM chain: windows of the final audio output against the oracle's stages run in sequence."""
    from oracle import oracle

    T1, T3 = int(h1.shape[0]), int(h3.shape[0])
    max_err = 0.0
    for o3, w in _windows(n3, width):
        n2 = (w - 1) * D3 + T3           # demodulated samples the window needs
        n1 = n2 + 1                      # low-pass samples (quad demod looks one ahead)
        o1 = o3 * D3                     # first low-pass sample
        n_in = (n1 - 1) * D1 + T1
        xw = x_dev[o1 * D1: o1 * D1 + n_in].cpu().numpy()
        lp = oracle.adjust_frequency_fir_fc(oracle.NCO_EXACT, fs, shift, first + o1 * D1, D1, h1, xw, n1)
        dm = oracle.quad_fm_demod(lp, gain, n2)
        ref = oracle.fir("ff", D3, h3, dm, w)
        got = au_dev[o3: o3 + w].cpu().numpy()
        max_err = max(max_err, float(np.abs(got - ref).max()))
    tol = float(gain) * 6e-5 * float(np.abs(h3).sum())  # FIR error -> phase error of the demodulator, times the gain
    return {"max_err": max_err, "tol": tol, "ok": bool(max_err <= tol), "windows": len(_windows(n3, width)),
            "outputs_per_window": min(width, n3), "against": "oracle chain (mix+FIR, quad demod, audio FIR) on the device input"}


def cpu_baseline(D: int, T: int, taps: np.ndarray, n_in_full: int, target_seconds: float = 10.0) -> dict:
    """Scalar C oracle (restating ref: src/fir.cu:57-70) on all host cores over a bounded sample of the workload: the
    first min(workload, 2^25) input samples, repeated until ~target_seconds of CPU work has been timed."""
    from harness import synth
    from oracle import oracle

    cores = os.cpu_count() or 1
    n_in = int(min(n_in_full, 1 << 25))
    n_out = (n_in - T) // D + 1
    x = synth.tone_plus_noise(0, n_in, seed=0x5EED0002)
    oracle.fir("fc", D, taps, x[: 1 << 20], threads=cores)  # page in, spin up
    reps, total = 0, 0.0
    while total < target_seconds and reps < 400:
        t0 = time.perf_counter()
        oracle.fir("fc", D, taps, x, n_out, threads=cores)
        total += time.perf_counter() - t0
        reps += 1
    dt = total / reps
    return {"value": n_in / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {n_in} input samples ({n_out} outputs) of the workload, {reps} passes of {dt:.3f} s, "
                      f"{cores} pthreads over contiguous output blocks, gcc -O2 -mfma scalar fmaf chain in the "
                      f"reference's accumulation order"}
