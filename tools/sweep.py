#!/usr/bin/env python
"""Tuning sweep (GPU box): times every polyphase kernel variant, the direct kernel and the reference's CUDA
kernel on a BASELINE workload, and measures the FP32 FFMA/FFMA2 peaks.  Prints one JSON line per measurement."""
from __future__ import annotations

import argparse
import ctypes
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import gsdr_b200 as g  # noqa: E402
from gsdr_b200 import synth  # noqa: E402


def timeit(fn, stream, reps=20, warm=3):
    for _ in range(warm):
        fn()
    stream.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        stream.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--D", type=int, default=8)
    ap.add_argument("--T", type=int, default=255)
    ap.add_argument("--log2n", type=int, default=26)
    ap.add_argument("--kind", default="fc")
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--peaks", action="store_true")
    ap.add_argument("--nco", action="store_true", help="time gsdrAdjustFrequencyFirFC instead of gsdrFirFC")
    ap.add_argument("--split", action="store_true", help="also time copy-only and FIR-only halves of each variant")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    D, T, n_in = args.D, args.T, 1 << args.log2n
    real = args.kind in ("ff", "cf")  # real input
    n_out = g.fir_num_outputs(n_in, T, D)
    x = synth.tone_plus_noise(0, n_in, seed=1, device=dev, real=real)
    i8 = args.kind == "i8"
    if i8:  # int8 IQ input: the same signal quantised to 8 bits, 2 bytes per sample
        x = torch.view_as_real(x).mul(127.0).round().clamp(-127, 127).to(torch.int8).reshape(-1).contiguous()
    cc = args.kind in ("cc", "cf")
    taps = torch.from_numpy(synth.random_taps(T, 3, complex_taps=True) if cc else synth.lowpass_taps(T, D)).to(dev)
    y = torch.zeros(n_out, dtype=torch.float32 if args.kind == "ff" else torch.complex64, device=dev)
    stream = torch.cuda.Stream()
    fn = g.gsdrFirCF if args.kind == "cf" else (g.gsdrFirFF if real else (g.gsdrFirCC if cc else g.gsdrFirFC))
    if i8:
        fn = g.gsdrFirFCInt8
    if args.nco:
        nco_fn = g.gsdrAdjustFrequencyFirFCInt8 if i8 else g.gsdrAdjustFrequencyFirFC

        def fn(D_, taps_, T_, x_, y_, n_, dev_, stream_):  # noqa: E306
            nco_fn(2.4e6, 29520.0, 12345, D_, taps_, T_, x_, y_, n_, dev_, stream_)
    esz = 4 if real else 8
    bytes_alg = (2 if i8 else esz) * n_in + (4 if args.kind == "ff" else 8) * n_out + 4 * T
    flops = (4.0 if args.kind == "cf" else (2.0 if real else (8.0 if cc else 4.0))) * T * n_out
    if args.peaks:
        lib = ctypes.CDLL(str(ROOT / "tools" / "libubench_fp32.so"))
        lib.ubenchFp32Tflops.restype = ctypes.c_double
        lib.ubenchFp32Tflops.argtypes = [ctypes.c_int] * 5
        lib.ubenchFirShapeTflops.restype = ctypes.c_double
        lib.ubenchFirShapeTflops.argtypes = [ctypes.c_int] * 5
        for bps in (1, 2, 4, 8):
            print(json.dumps({"peak": "fir-shaped operands, 128-thread blocks", "blocks_per_sm": bps,
                              "ffma2_tflops": lib.ubenchFirShapeTflops(0, 0, 8000, 3, bps),
                              "ffma_tflops": lib.ubenchFirShapeTflops(1, 0, 8000, 3, bps)}), flush=True)
        for bps in (1, 2, 4, 8):
            print(json.dumps({"peak": "fp32", "blocks_per_sm": bps, "ffma_tflops": lib.ubenchFp32Tflops(0, 0, 8000, 3, bps),
                              "ffma2_tflops": lib.ubenchFp32Tflops(1, 0, 8000, 3, bps)}), flush=True)
    ref = None
    for v in ([-1, -2] if i8 else list(range(g.num_kernel_variants())) + [-2]):
        g.set_kernel_variant(v)
        info = g.describe_kernel(3 if args.kind == "cf" else (1 if real else (2 if cc else (4 if args.nco else 0))), D, T, n_out)
        if not i8 and v >= 0 and info.variant != v:
            print(json.dumps({"variant": v, "skipped": "does not fit"}), flush=True)
            continue
        y.zero_()
        try:
            med, best = timeit(lambda: fn(D, taps, T, x, y, n_out, 0, stream), stream, reps=5 if v == -2 else 20)
        except g.CudaError as e:
            print(json.dumps({"variant": v, "skipped": str(e)}), flush=True)
            continue
        if ref is None:
            ref = y.clone()
        diff = float((y - ref).abs().max())
        extra = {}
        if args.split and v >= 0:
            for name, flag in (("ms_copy_only", 2), ("ms_fir_only", 1), ("ms_fir_nostore", 5)):
                g.set_debug_flags(flag)
                extra[name] = timeit(lambda: fn(D, taps, T, x, y, n_out, 0, stream), stream, reps=10)[0]
            g.set_debug_flags(0)
        print(json.dumps({"variant": v, **extra, "threads": info.threadsPerBlock, "R": info.outputsPerThread,
                          "smem": info.sharedBytesPerBlock, "ctas": info.numBlocks, "ms_median": med, "ms_best": best,
                          "msamples_s": n_in / med / 1e3, "gbs": bytes_alg / med / 1e6, "tflops": flops / med / 1e9,
                          "maxdiff_vs_first": diff}), flush=True)
    g.set_kernel_variant(-1)
    if args.ref:
        from oracle import ref_cuda

        if ref_cuda.available():
            yr = torch.zeros_like(y)
            med, best = timeit(lambda: ref_cuda.fir(args.kind, D, taps, T, x, yr, n_out, 0, stream.cuda_stream),
                               stream, reps=5, warm=1)
            print(json.dumps({"variant": "reference-cuda", "ms_median": med, "ms_best": best,
                              "msamples_s": n_in / med / 1e3, "gbs": bytes_alg / med / 1e6,
                              "tflops": flops / med / 1e9,
                              "maxdiff_vs_ours": float((yr - ref).abs().max())}), flush=True)


if __name__ == "__main__":
    main()
