"""The C-ABI library loads and exports every symbol include/gsdr/*.h declares (no compute calls: no GPU here)."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
INCLUDE = ROOT / "include" / "gsdr"


def _declared_symbols(tuning: bool = False):
    """Symbols the headers declare; the block under #ifdef GSDR_B200_TUNING only for the tuning build."""
    names = set()
    for h in sorted(INCLUDE.glob("*.h")):
        text = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        if not tuning:
            text = re.sub(r"#ifdef GSDR_B200_TUNING.*?#endif", "", text, flags=re.S)
        for m in re.finditer(r"GSDR_PUBLIC\s+[\w\s\*]+?\b(gsdr\w+)\s*\(", text):
            names.add(m.group(1))
    return sorted(names)


def test_headers_declare_the_reference_fir_entry_points():
    d = _declared_symbols()
    for name in ("gsdrFirFC", "gsdrFirFF", "gsdrFirCC", "gsdrFirCF", "gsdrAdjustFrequencyFirFC"):
        assert name in d


def test_library_exports_every_declared_symbol():
    from gsdr_b200 import _lib

    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/gsdr but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in gsdr_b200/_lib.py"


def _exports(path):
    out = subprocess.run(["nm", "-D", "--defined-only", str(path)], capture_output=True, text=True).stdout
    return {line.split()[-1] for line in out.splitlines() if " T " in line and line.split()[-1].startswith("gsdr")}


def test_no_undeclared_gsdr_exports():
    from gsdr_b200 import _lib

    assert _exports(_lib.LIB_PATH) == set(_declared_symbols())


def test_release_library_has_no_tuning_hooks_and_tuning_library_has_them():
    """gsdrB200SetKernelVariant / gsdrB200SetDebugFlags (process-wide, work-skipping) exist only in the tuning build."""
    from gsdr_b200 import _lib

    hooks = {"gsdrB200SetKernelVariant", "gsdrB200SetDebugFlags"}
    assert not (hooks & _exports(_lib.LIB_PATH))
    assert _exports(_lib.TUNING_LIB_PATH) == set(_declared_symbols(tuning=True))
    assert hooks <= _exports(_lib.TUNING_LIB_PATH)
    assert ctypes.CDLL(str(_lib.LIB_PATH)).gsdrB200HasTuningHooks() == 0
    assert ctypes.CDLL(str(_lib.TUNING_LIB_PATH)).gsdrB200HasTuningHooks() == 1


def test_compiled_c_consumer_links_and_resolves_the_reference_call_shape(tmp_path):
    """tests/abi_consumer.c is plain C written against include/gsdr/fir.h exactly as a user of the reference writes it
    (ref: include/gsdr/fir.h:30-38); it must compile with gcc and link against libgsdr_b200.so.  Run without a GPU it
    only checks the error path (the call must return a cudaError_t, not crash)."""
    from gsdr_b200 import _lib

    cuda_inc, cuda_lib = "/usr/local/cuda/include", "/usr/local/cuda/lib64"
    if not Path(cuda_inc, "cuda_runtime.h").exists():
        pytest.skip("CUDA headers not installed")
    exe = tmp_path / "abi_consumer"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-I", cuda_inc,
                    str(ROOT / "tests" / "abi_consumer.c"), "-o", str(exe), "-L", str(_lib.LIB_PATH.parent),
                    "-lgsdr_b200", "-L", cuda_lib, "-lcudart", f"-Wl,-rpath,{_lib.LIB_PATH.parent}",
                    f"-Wl,-rpath,{cuda_lib}"], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode in (0, 3), res.stdout + res.stderr  # 0: ran on a GPU and matched; 3: no device (error path)
    assert "gsdrFirFC" in res.stdout


def test_fir_signature_matches_reference_argument_order():
    """(decimation, taps, tapCount, input, output, numOutputs, cudaDevice, cudaStream), ref: include/gsdr/fir.h:30-38."""
    text = (INCLUDE / "fir.h").read_text()
    m = re.search(r"cudaError_t gsdrFirFC\((.*?)\)", text, flags=re.S)
    params = [p.strip().split()[-1].lstrip("*") for p in m.group(1).split(",")]
    assert params == ["decimation", "taps", "tapCount", "input", "output", "numOutputs", "cudaDevice", "cudaStream"]


def test_headers_compile_as_c_and_cxx(tmp_path):
    """fir.h must stay usable from plain C (the reference's is)."""
    cuda_inc = "/usr/local/cuda/include"
    if not Path(cuda_inc, "cuda_runtime.h").exists():
        pytest.skip("CUDA headers not installed")
    src = ''.join(f'#include <gsdr/{h.name}>\n' for h in sorted(INCLUDE.glob('*.h'))) + 'int main(void){return 0;}\n'
    for compiler, name in (("gcc", "t.c"), ("g++", "t.cpp")):
        f = tmp_path / name
        f.write_text(src)
        subprocess.run([compiler, "-fsyntax-only", "-I", str(ROOT / "include"), "-I", cuda_inc, str(f)], check=True)


def test_size_helpers_without_gpu():
    import gsdr_b200 as g

    assert g.fir_num_outputs(1 << 26, 255, 8) == 8_388_577  # BASELINE config 2
    assert g.fir_num_outputs(1 << 20, 63, 1) == 1_048_514   # config 1
    assert g.fir_num_outputs(1 << 28, 1023, 32) == 8_388_577  # config 3
    assert g.fir_num_outputs(1 << 22, 127, 4) == 1_048_545  # config 4
    assert g.fir_num_outputs(1 << 31, 255, 10) == 214_748_340  # config 5 stage 1
    assert g.fir_num_outputs(10, 11, 1) == 0
    assert g.fir_num_inputs(8_388_577, 255, 8) <= 1 << 26
    assert g.fir_num_inputs(0, 255, 8) == 0


def test_nco_phase_step_matches_oracle():
    import gsdr_b200 as g
    from oracle import oracle

    for f, fs in [(100e3, 2.4e6), (-100e3, 2.4e6), (0.0, 1e6), (1.2e6, 2.4e6), (29520.0, 2.4e6), (7.0e6, 2.4e6)]:
        assert g.nco_phase_step(f, fs) == oracle.nco_exact_phase_step(f, fs)


@pytest.mark.gpu
def test_compiled_c_consumer_runs_on_the_gpu(tmp_path):
    """The same plain-C program on a device: the impulse response must come back exactly (exit code 0)."""
    from gsdr_b200 import _lib

    cuda_inc, cuda_lib = "/usr/local/cuda/include", "/usr/local/cuda/lib64"
    exe = tmp_path / "abi_consumer"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-I", cuda_inc,
                    str(ROOT / "tests" / "abi_consumer.c"), "-o", str(exe), "-L", str(_lib.LIB_PATH.parent),
                    "-lgsdr_b200", "-L", cuda_lib, "-lcudart", f"-Wl,-rpath,{_lib.LIB_PATH.parent}",
                    f"-Wl,-rpath,{cuda_lib}"], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "impulse response ok" in res.stdout
