// fir_inst_real.cu — kernel instantiations: firTmaRealKernel (see fir_launch.cuh).
#include "fir_launch.cuh"

namespace gsdr_b200 {
GSDR_DEFINE_REAL_DT(0)
GSDR_DEFINE_REAL_DT(2)
GSDR_DEFINE_REAL_DT(10)
}  // namespace gsdr_b200
