// Conv3d on the tcgen05 engine (conv3d_tc.cu): NDHWC bf16 (hi [, lo]) planes in, fp32 NDHWC out (+ GroupNorm statistics).
#pragma once
#include "conv.cuh"

namespace ipk {

struct Conv3dShape {
  int Cin, Cout;                 // channels as stored
  int Ti, Hi, Wi;                // input volume
  int kt, ky, kx, st, sy, sx, pt, py, px;
  // optional byte strides of the input view (0 = dense NDHWC): element (c, x, y) of a time plane sits at c*2 + x*stride_x + y*stride_y.
  // stride_x < Cin*2 gives an OVERLAPPED view: the x taps of a small-channel input folded into K (see encoder.cu: stem).
  long long stride_x = 0, stride_y = 0;
};

// Cin, Cout multiples of 64; output rows tile into 128-voxel boxes of one time step (W_out a power of two <= 128 or a multiple of 128)
// pack OIDHW fp32 weights [Cout][CinSrc][kt][ky][kx] into a ConvW from conv_alloc(pool, engine, kt*ky*kx, Cin, Cout, false)
void conv3d_tc_pack(ConvW& dst, const float* w_oidhw, int Cout, int CinSrc, cudaStream_t st);
bool conv3d_tc_supported(const Conv3dShape& s);
// w: ConvW packed with ntaps = kt*ky*kx (tap = (dt*ky + dy)*kx + dx), K = Cin, N = Cout, tensor-core engine.
// in_hi / in_lo: [B][Ti][Hi][Wi][Cin] bf16 planes; out: [B][To][Ho][Wo][Cout] fp32;
// stats (optional, zeroed by the caller): [B][Cout][2] += (sum, sum of squares) of the output over each sample's volume.
void conv3d_tc_run(const ConvW& w, const Conv3dShape& s, const void* in_hi, const void* in_lo, int B, float* out, double* stats, cudaStream_t st);

}  // namespace ipk
