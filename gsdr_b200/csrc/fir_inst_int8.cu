// fir_inst_int8.cu — kernel instantiations: firTmaInt8Kernel (see fir_launch.cuh).
#include "fir_launch.cuh"

namespace gsdr_b200 {
GSDR_DEFINE_INT8_DT(0)
GSDR_DEFINE_INT8_DT(8)
GSDR_DEFINE_INT8_DT(10)
}  // namespace gsdr_b200
