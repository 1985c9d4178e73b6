#!/bin/bash
# Builds gsdr_b200/csrc/libgsdr_b200_tuning_x<N>.so: the tuning library with fir_inst_tc.cu compiled with
# -DGSDR_TC_EXPERIMENT=<N> (and -DGSDR_TC_PHASE_TIMING when TC_TIMING is set) (timing experiments on the tensor-core kernel; results may be wrong
# by design).  Select it with GSDR_B200_TUNING_LIB=<path>.  Needs an up-to-date build_tuning/ (gsdr_b200/build.py).
set -e
cd "$(dirname "$0")/../gsdr_b200/csrc"
for x in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden \
    -Xptxas --register-usage-level=10 -DGSDR_B200_TUNING ${TC_TIMING:+-DGSDR_TC_PHASE_TIMING=1} -DGSDR_TC_EXPERIMENT=$x \
    -c -I ../../include -I . -o /tmp/fir_inst_tc_x$x.o fir_inst_tc.cu
  objs=$(ls build_tuning/*.o | grep -v fir_inst_tc.o)
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o libgsdr_b200_tuning_x$x.so $objs /tmp/fir_inst_tc_x$x.o
done
