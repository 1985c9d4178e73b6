// fir_inst_chan.cu — kernel instantiations: firTmaChannelizerKernel (one input, K frequency shifts).
// Tuning build only: measured no faster than per-shift launches (profiles/r02/channelizer.jsonl).
#ifdef GSDR_B200_TUNING
#include "fir_launch.cuh"

namespace gsdr_b200 {

cudaError_t launchChan(int D, const CUtensorMap& map, ChanParams& P, size_t smem, int dev, int smCount,
                       cudaStream_t stream) noexcept {
  switch (D) {  // tile shapes: kChanShapes
    case 4: return launchChanT<64, 1, 4, 4>(map, P, smem, dev, smCount, stream);
    case 8: return launchChanT<32, 2, 8, 4>(map, P, smem, dev, smCount, stream);
    case 10: return launchChanT<64, 1, 10, 4>(map, P, smem, dev, smCount, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace gsdr_b200
#endif  // GSDR_B200_TUNING
