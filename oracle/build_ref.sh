#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# Builds the reference's own CUDA kernels for this path, from the sources where they lie under
# /root/reference, into oracle/_ref/libgsdr_ref.so (git-ignored; travels to the GPU box with the snapshot).
# The reference's CMake is not used (it does not configure with CMake 4.x, and iir.cu/qpsk*.cu do not compile).
#
#   libgsdr_ref.so exports, unmodified:   gsdrFirFC/FF/CC/CF   (src/fir.cu)
#                                         gsdrQuadFmDemod/gsdrQuadAmDemod (src/quad_demod.cu)
#   plus refAdjustFrequencyFirFC          (oracle/ref_adjust_harness.cu around src/adjustFrequency.cu)
#
# The only edit made to any reference source is in a throw-away temp copy of adjustFrequency.cu: the missing
# `return sample;` is appended to k_AdjustFrequency (without it the function returns garbage).  The copy is
# deleted after the build; no reference source is stored in this repository.
set -euo pipefail
REF=${GSDR_REFERENCE_DIR:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
  echo "build_ref.sh: $REF not present; keeping any prebuilt $OUT/libgsdr_ref.so" >&2
  exit 0
fi
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$TMP/inc/gsdr"
# what the reference's CMake would generate for a static, non-exported build (ref: CMakeLists.txt:52-65)
printf '#pragma once\n#define GSDR_PUBLIC\n#define GSDR_PRIVATE\n' > "$TMP/inc/gsdr/gsdr_export.h"
python3 - "$REF/src/adjustFrequency.cu" "$TMP/adjustFrequency_patched.cu" <<'PY'
import sys
src = open(sys.argv[1]).read()
i = src.rstrip().rfind('}')
assert 'return sample' not in src, "reference already has the return; drop the patch"
open(sys.argv[2], 'w').write(src[:i] + '  return sample;\n}\n')
PY
# same language level / optimisation as the reference (ref: CMakeLists.txt:24,33,130-146), sm_100 instead of sm_75
FLAGS=(-std=c++11 -O3 -Xcompiler -fPIC -gencode arch=compute_100,code=sm_100 -I"$REF/include" -I"$REF/src" -I"$TMP/inc")
"$NVCC" "${FLAGS[@]}" -c "$REF/src/fir.cu" -o "$TMP/fir.o"
"$NVCC" "${FLAGS[@]}" -c "$REF/src/quad_demod.cu" -o "$TMP/quad_demod.o"
"$NVCC" "${FLAGS[@]}" -rdc=true -c "$TMP/adjustFrequency_patched.cu" -o "$TMP/adjustFrequency.o" 2>/dev/null
"$NVCC" "${FLAGS[@]}" -rdc=true -c "$HERE/ref_adjust_harness.cu" -o "$TMP/harness.o"
"$NVCC" -gencode arch=compute_100,code=sm_100 -Xcompiler -fPIC -dlink "$TMP/adjustFrequency.o" "$TMP/harness.o" -o "$TMP/dlink.o"
CUDA_HOME="$(dirname "$(dirname "$NVCC")")"
${CXX:-g++} -shared -o "$OUT/libgsdr_ref.so" "$TMP/fir.o" "$TMP/quad_demod.o" "$TMP/adjustFrequency.o" "$TMP/harness.o" "$TMP/dlink.o" \
  -L"$CUDA_HOME/lib64" -lcudart_static -lrt -ldl -lpthread
echo "built $OUT/libgsdr_ref.so"
