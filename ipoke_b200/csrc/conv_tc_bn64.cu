// conv_tc_kernel instantiations with 64-column N tiles (see conv_tc_kernel.cuh)
#include "conv_tc_kernel.cuh"

namespace ipk {

void tc_launch_bn64(bool split, int fused, bool halo, int cg, const TcMaps& m, TcArgs& a, cudaStream_t st) {
  IPK_CHECK(cg == 1, IPK_ERR_UNSUPPORTED, "conv_tc: CTA pairs need 256-column tiles");
  if (halo) IPK_TC_FAMILY(64, true, 1);       // 3x3 on 128-wide images: the decoder's last conv2
  else IPK_TC_FAMILY(64, false, 1);
}

}  // namespace ipk
