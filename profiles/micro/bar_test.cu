// Does the final __syncthreads() of a warp-specialised kernel wait for slow warps?  Variants: bit0 = __syncwarp before the barrier in the
// lane-divergent warps, bit1 = griddepcontrol.launch_dependents first, bit2 = launched with programmatic stream serialization +
// griddepcontrol.wait, bit3 = spin on an mbarrier try_wait loop instead of the clock
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(320, 1) k(long long* out, int variant) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (variant & 4) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (variant & 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  long long t0 = clock64();
  if (warp == 0) {
    if (lane == 0) { while (clock64() - t0 < 2000) {} }
    if (variant & 1) __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) { while (clock64() - t0 < 4000) {} }
    if (variant & 1) __syncwarp();
  } else {
    while (clock64() - t0 < 40000) {}
    if (lane == 0 && warp == 2) out[1] = clock64() - t0;
  }
  __syncthreads();
  if (threadIdx.x == 0) out[0] = clock64() - t0;
  if (threadIdx.x == 32) out[2] = clock64() - t0;
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  for (int v = 0; v < 8; ++v) {
    cudaMemset(d, 0, 64);
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(1); cfg.blockDim = dim3(320);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = (v & 4) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, k, d, v);
    cudaLaunchKernelEx(&cfg, k, d + 4, v);
    long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("variant %d: first kernel: tid0 passes barrier at %lld, tid32 at %lld, slow warp done at %lld | second: %lld %lld %lld [%s]\n", v, h[0], h[2], h[1], h[4], h[6], h[5],
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
