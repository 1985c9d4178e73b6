#!/usr/bin/env python
"""Builds experimental copies of libgsdr_b200.so with -D overrides into tools/exp/ (git-ignored), for A/B timing on
the GPU box:  python tools/exp_build.py name1:-DFOO=1,-DBAR=2 name2:...   then
GSDR_B200_LIB=tools/exp/libgsdr_b200_name1.so python bench.py ..."""
import importlib.util
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
spec = importlib.util.spec_from_file_location("_b", ROOT / "gsdr_b200" / "build.py")
b = importlib.util.module_from_spec(spec)
spec.loader.exec_module(b)

out = ROOT / "tools" / "exp"
out.mkdir(exist_ok=True)
procs = []
for arg in sys.argv[1:]:
    name, _, defs = arg.partition(":")
    lib = out / f"libgsdr_b200_{name}.so"
    flags = list(b.NVCC_FLAGS)
    cmd = [b._nvcc(), *flags, *[d for d in defs.split(",") if d], "-shared", "-I", str(ROOT / "include"),
           "-I", str(b.CSRC), "-o", str(lib), *map(str, b.SOURCES)]
    procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, p in procs:
    log = p.communicate()[0]
    print(name, "rc", p.returncode, log[-300:] if p.returncode else "")
