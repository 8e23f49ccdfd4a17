// conv_tc_kernel instantiations with 32-column N tiles (see conv_tc_kernel.cuh)
#include "conv_tc_kernel.cuh"

namespace ipk {

void tc_launch_bn32(bool split, int fused, bool halo, int cg, const TcMaps& m, TcArgs& a, cudaStream_t st) {
  IPK_CHECK(cg == 1, IPK_ERR_UNSUPPORTED, "conv_tc: CTA pairs need 256-column tiles");
  if (halo) IPK_TC_FAMILY(32, true, 1);       // the decoder's out_conv on the tensor-core engine (IPK_OUTCONV_TC=1)
  else IPK_TC_FAMILY(32, false, 1);
}

}  // namespace ipk
