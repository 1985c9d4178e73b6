// fir_inst_spec_cc.cu — kernel instantiations: firTmaNcoSpecKernel and firTmaCcKernel (see fir_launch.cuh).
#include "fir_launch.cuh"

namespace gsdr_b200 {
GSDR_DEFINE_SPEC_DT(0)
GSDR_DEFINE_SPEC_DT(8)
GSDR_DEFINE_SPEC_DT(10)
GSDR_DEFINE_SPEC_DT(32)
GSDR_DEFINE_CC_DT(0)
GSDR_DEFINE_CC_DT(8)
}  // namespace gsdr_b200
