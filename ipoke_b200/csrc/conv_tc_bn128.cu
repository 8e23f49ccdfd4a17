// conv_tc_kernel instantiations with 128-column N tiles (see conv_tc_kernel.cuh)
#include "conv_tc_kernel.cuh"

namespace ipk {

void tc_launch_bn128(bool split, int fused, bool halo, int cg, const TcMaps& m, TcArgs& a, cudaStream_t st) {
  IPK_CHECK(cg == 1 && !halo, IPK_ERR_UNSUPPORTED, "conv_tc: 128-column tiles run on single CTAs without halo mode");
  IPK_TC_FAMILY(128, false, 1);
}

}  // namespace ipk
