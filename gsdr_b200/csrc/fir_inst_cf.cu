// fir_inst_cf.cu — kernel instantiations: firTmaRealKernel with two tap planes (gsdrFirCF; see fir_launch.cuh).
#include "fir_launch.cuh"

namespace gsdr_b200 {
GSDR_DEFINE_CF_DT(0)
GSDR_DEFINE_CF_DT(2)
GSDR_DEFINE_CF_DT(10)
}  // namespace gsdr_b200
