// fir_inst_tma_d0.cu — kernel instantiations: firTmaKernel, run-time geometry (see fir_launch.cuh).
#include "fir_launch.cuh"

namespace gsdr_b200 {
GSDR_DEFINE_TMA_DT(0)
}  // namespace gsdr_b200
