// fir_inst_tma_d8.cu — kernel instantiations: firTmaKernel, compile-time decimation 8 (see fir_launch.cuh).
#include "fir_launch.cuh"

namespace gsdr_b200 {
GSDR_DEFINE_TMA_DT(8)
}  // namespace gsdr_b200
