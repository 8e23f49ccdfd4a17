// conv_tc_kernel instantiations with 256-column N tiles (see conv_tc_kernel.cuh)
#include "conv_tc_kernel.cuh"

namespace ipk {

void tc_launch_bn256(bool split, int fused, bool halo, int cg, const TcMaps& m, TcArgs& a, cudaStream_t st) {
  IPK_CHECK(!halo, IPK_ERR_UNSUPPORTED, "conv_tc: halo mode needs 32- or 64-column tiles");
  if (cg == 2) IPK_TC_FAMILY(256, false, 2);   // CTA pairs: NICE conv2 and the other long-K wide-N contractions
  else IPK_TC_FAMILY(256, false, 1);
}

}  // namespace ipk
