// fir_inst_wide.cu — kernel instantiations: firTmaWideKernel (segment-pipelined, rows wider than 128 bytes).
#include "fir_launch.cuh"

namespace gsdr_b200 {
GSDR_DEFINE_WIDE_DT(32)
}  // namespace gsdr_b200
