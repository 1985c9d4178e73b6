"""In-tree build of libgsdr_b200.so (sm_100a only) and of the test oracle.

The product library is `gsdr_b200/csrc/libgsdr_b200.so`; it links nothing but the CUDA runtime.  The oracle
(`oracle/libgsdr_oracle.so`) and the compiled reference (`oracle/_ref/libgsdr_ref.so`) are test infrastructure:
they are built here so they travel to the GPU box with the snapshot, but the product never loads them.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "gsdr_b200" / "csrc"
LIB = CSRC / "libgsdr_b200.so"                # the product: no tuning / debug hooks
LIB_TUNING = CSRC / "libgsdr_b200_tuning.so"  # same sources with -DGSDR_B200_TUNING (variant override, work-skipping
#                                               measurement flags): loaded only by variant-coverage tests and tools/sweep.py
SOURCES = [CSRC / "gsdr_fir.cu", CSRC / "gsdr_host.cu", CSRC / "gsdr_demod.cu", CSRC / "gsdr_stream.cu",
           *sorted(CSRC.glob("fir_inst_*.cu"))]  # fir_inst_*: kernel instantiations, one unit per compile-time decimation
HEADERS = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + sorted((ROOT / "include" / "gsdr").glob("*.h"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-Xptxas", "-v",
    # Take the register-hungry schedule of the FIR loop (135 registers instead of 127: +2.5 % on the headline shape;
    # residency is bound by shared memory, not registers).
    "-Xptxas", "--register-usage-level=10",
    # ptxas (only) over all host cores: ~300 template instantiations of the big kernels.  NOT nvcc's own
    # -split-compile: that also splits the module inside the NVVM optimiser, and the FIR loop then comes out in one
    # of two schedules (98 or 126 registers, 7 % apart) depending on unrelated source changes (profiles/r01_notes.md).
    "-Xptxas", "--split-compile=0",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libgsdr_b200.so cannot be built")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build_library(force: bool = False, verbose: bool = False, tuning: bool | None = None) -> Path:
    """Every translation unit is compiled to an object by its own nvcc process (in parallel), then linked.
    tuning=None builds both the release and the tuning library (one pool of compile jobs)."""
    flavours = [False, True] if tuning is None else [tuning]
    todo = [t for t in flavours
            if force or _stale(LIB_TUNING if t else LIB, SOURCES + HEADERS + [Path(__file__)])]
    if not todo:
        return LIB
    from concurrent.futures import ThreadPoolExecutor

    nvcc = _nvcc()

    def compile_one(job):
        src, tun = job
        objdir = CSRC / ("build_tuning" if tun else "build")
        objdir.mkdir(exist_ok=True)
        obj = objdir / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *(["-DGSDR_B200_TUNING"] if tun else []), "-c", "-I", str(ROOT / "include"), "-I",
               str(CSRC), "-o", str(obj), str(src)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return tun, obj, cmd, res

    # biggest units first so that they do not end up alone at the tail
    order = sorted(SOURCES, key=lambda p: (0 if p.name.startswith("fir_inst_") or p.name == "gsdr_fir.cu" else 1, p.name))
    jobs = [(src, tun) for src in order for tun in todo]
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
        results = list(pool.map(compile_one, jobs))
    log = ""
    failed = False
    for tun, obj, cmd, res in results:
        log += " ".join(cmd) + "\n" + res.stdout + res.stderr
        failed = failed or res.returncode != 0
    for tun in todo:
        if failed:
            break
        target = LIB_TUNING if tun else LIB
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(target),
               *[str(o) for t, o, _, _ in results if t == tun]]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log += " ".join(cmd) + "\n" + res.stdout + res.stderr
        failed = res.returncode != 0
    (CSRC / "build.log").write_text(log)
    if failed:
        sys.stderr.write(log[-8000:])
        raise RuntimeError("nvcc failed building libgsdr_b200.so")
    if verbose:
        print(log)
    return LIB


def build_tools(force: bool = False) -> Path:
    """tools/libubench_fp32.so — FP32 FFMA/FFMA2 peak microbenchmark used by bench.py for the roofline."""
    src = ROOT / "tools" / "ubench_fp32.cu"
    lib = ROOT / "tools" / "libubench_fp32.so"
    if force or _stale(lib, [src]):
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-Xcompiler", "-fPIC",
               "-shared", "-o", str(lib), str(src)]
        subprocess.run(cmd, check=True, capture_output=True)
    return lib


def build_oracle(force: bool = False) -> Path:
    odir = ROOT / "oracle"
    lib = odir / "libgsdr_oracle.so"
    if force or _stale(lib, [odir / "gsdr_oracle.c", odir / "gsdr_oracle.h", odir / "Makefile"]):
        subprocess.run(["make", "-C", str(odir), "-B"], check=True, capture_output=True)
    return lib


def build_reference() -> Path | None:
    """Compiles the reference's own kernels when /root/reference is present (never on the GPU box)."""
    odir = ROOT / "oracle"
    lib = odir / "_ref" / "libgsdr_ref.so"
    ref = Path(os.environ.get("GSDR_REFERENCE_DIR", "/root/reference"))
    if (ref / "src" / "fir.cu").exists():
        if _stale(lib, [odir / "build_ref.sh", odir / "ref_adjust_harness.cu"]):
            subprocess.run(["bash", str(odir / "build_ref.sh")], check=True, capture_output=True)
    return lib if lib.exists() else None


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
    print(build_tools())
    print(build_oracle())
    print(build_reference())
