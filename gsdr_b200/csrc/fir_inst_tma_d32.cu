// fir_inst_tma_d32.cu — kernel instantiations: firTmaKernel, compile-time decimation 32 (see fir_launch.cuh).
#include "fir_launch.cuh"

namespace gsdr_b200 {
GSDR_DEFINE_TMA_DT(32)
}  // namespace gsdr_b200
