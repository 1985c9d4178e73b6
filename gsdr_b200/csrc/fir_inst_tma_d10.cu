// fir_inst_tma_d10.cu — kernel instantiations: firTmaKernel, compile-time decimation 10 (see fir_launch.cuh).
#include "fir_launch.cuh"

namespace gsdr_b200 {
GSDR_DEFINE_TMA_DT(10)
}  // namespace gsdr_b200
