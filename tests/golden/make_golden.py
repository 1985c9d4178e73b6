"""Generates tests/golden/ref_*.npz: seeded inputs plus the outputs of the REFERENCE's own CUDA kernels.

Run on a GPU box (the reference is CUDA-only and this container has no GPU):

    gpurun -- python tests/golden/make_golden.py          # writes gpurun_out/golden/ref_*.npz
    cp gpurun_out/golden/ref_*.npz tests/golden/          # then commit them

The kernels come from oracle/_ref/libgsdr_ref.so, which oracle/build_ref.sh compiles (for sm_100) from
/root/reference/src/fir.cu and — with the missing `return sample;` added in a temp copy — src/adjustFrequency.cu.
Inputs are numpy default_rng streams with the seeds below, stored in the files so the fixtures are self-contained.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

# (decimation, tapCount, numOutputs)
FIR_CASES = [(1, 63, 300), (8, 255, 200), (5, 33, 17), (3, 1, 2), (10, 255, 64), (2, 16, 8), (1, 1, 1), (4, 127, 96)]
NCO_CASES = [  # (sampleRate, frequencyShift, firstSampleIndex, decimation, tapCount, numOutputs)
    (2.4e6, 1.0e5, 0, 8, 63, 64),
    (2.4e6, -3.1e5, 5_000_003, 10, 255, 40),
    (1.0e6, 12345.0, 2 ** 33 + 5, 4, 31, 33),
]


def _rand(rng, n, cplx):
    if cplx:
        return (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
    return rng.uniform(-1, 1, n).astype(np.float32)


def main() -> None:
    import torch

    from oracle import ref_cuda

    assert torch.cuda.is_available() and ref_cuda.available()
    out_dir = ROOT / "gpurun_out" / "golden"
    out_dir.mkdir(parents=True, exist_ok=True)
    dev = torch.device("cuda:0")
    for kind in ("ff", "fc", "cc", "cf"):
        for ci, (D, T, n_out) in enumerate(FIR_CASES):
            seed = 1000 * (1 + "ff fc cc cf".split().index(kind)) + ci
            rng = np.random.default_rng(seed)
            taps = _rand(rng, T, kind[0] == "c")
            x = _rand(rng, (n_out - 1) * D + T, kind[1] == "c")
            dt, dx = torch.from_numpy(taps).to(dev), torch.from_numpy(x).to(dev)
            dy = torch.zeros(n_out, dtype=torch.float32 if kind == "ff" else torch.complex64, device=dev)
            ref_cuda.fir(kind, D, dt, T, dx, dy, n_out)
            torch.cuda.synchronize()
            np.savez(out_dir / f"ref_{kind}_d{D}_t{T}_n{n_out}.npz", kind=kind, decimation=D, num_outputs=n_out,
                     seed=seed, taps=taps, input=x, output=dy.cpu().numpy())
    for ci, (fs, f, first, D, T, n_out) in enumerate(NCO_CASES):
        seed = 9000 + ci
        rng = np.random.default_rng(seed)
        taps = _rand(rng, T, False)
        x = _rand(rng, (n_out - 1) * D + T, True)
        dt, dx = torch.from_numpy(taps).to(dev), torch.from_numpy(x).to(dev)
        dy = torch.zeros(n_out, dtype=torch.complex64, device=dev)
        ref_cuda.adjust_frequency_fir_fc(fs, f, first, D, dt, T, dx, dy, n_out)
        torch.cuda.synchronize()
        np.savez(out_dir / f"ref_adjust_literal_{ci}.npz", kind="adjust_literal", sample_rate=fs, frequency_shift=f,
                 first_sample_index=first, decimation=D, num_outputs=n_out, seed=seed, taps=taps, input=x,
                 output=dy.cpu().numpy())
    for ci, n_out in enumerate((1, 257, 4099)):
        seed = 9500 + ci
        rng = np.random.default_rng(seed)
        x = _rand(rng, n_out + 1, True)
        dx = torch.from_numpy(x).to(dev)
        dy = torch.zeros(n_out, dtype=torch.float32, device=dev)
        rc = ref_cuda.lib().gsdrQuadFmDemod(dx.data_ptr(), dy.data_ptr(), 1.5, n_out, 0, 0)
        assert rc == 0
        torch.cuda.synchronize()
        np.savez(out_dir / f"ref_quadfm_{ci}.npz", kind="quad_fm", gain=1.5, num_outputs=n_out, seed=seed, input=x,
                 output=dy.cpu().numpy())
    print(f"wrote {len(list(out_dir.glob('ref_*.npz')))} golden files to {out_dir}")


if __name__ == "__main__":
    main()
