// fir_inst_tma_d4.cu — kernel instantiations: firTmaKernel, compile-time decimation 4 (see fir_launch.cuh).
#include "fir_launch.cuh"

namespace gsdr_b200 {
GSDR_DEFINE_TMA_DT(4)
}  // namespace gsdr_b200
