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
  // optional explicit output extents (0 = (in + 2*pad - k) / stride + 1): with them pt / py / px are the LOW-side paddings of an
  // asymmetric (TF "SAME") padding -- the high side is whatever the output extent implies, zero-filled by the TMA unit like the rest
  int To = 0, Ho = 0, Wo = 0;
};

// Epilogue of conv3d_tc_run_ex: out = act(acc * scale[n] + shift[n]) written as fp32 rows and / or bf16 (hi [, lo]) operand planes, both
// [voxel][cstride] with the first output channel at column coff (a channel slice of a wider tensor: the Inception concat).
struct Conv3dEpi {
  float* out_f32 = nullptr;
  __nv_bfloat16* out_hi = nullptr;
  __nv_bfloat16* out_lo = nullptr;
  int cstride = 0, coff = 0;
  const float* scale = nullptr;     // [Cout] or null (folded BatchNorm: gamma / sqrt(var + eps))
  const float* shift = nullptr;     // [Cout] or null (beta - mean * scale, or a conv bias)
  int relu = 0;
  double* stats = nullptr;          // as in conv3d_tc_run (needs >= 32 voxels of one sample per warp: power-of-two boxes)
};

// Cin, Cout multiples of 64; output rows tile into 128-voxel boxes of one time step (W_out a power of two <= 128 or a multiple of 128)
// pack OIDHW fp32 weights [Cout][CinSrc][kt][ky][kx] into a ConvW from conv_alloc(pool, engine, kt*ky*kx, Cin, Cout, false)
void conv3d_tc_pack(ConvW& dst, const float* w_oidhw, int Cout, int CinSrc, cudaStream_t st);
bool conv3d_tc_supported(const Conv3dShape& s);
// w: ConvW packed with ntaps = kt*ky*kx (tap = (dt*ky + dy)*kx + dx), K = Cin, N = Cout, tensor-core engine.
// in_hi / in_lo: [B][Ti][Hi][Wi][Cin] bf16 planes; out: [B][To][Ho][Wo][Cout] fp32;
// stats (optional, zeroed by the caller): [B][Cout][2] += (sum, sum of squares) of the output over each sample's volume.
void conv3d_tc_run(const ConvW& w, const Conv3dShape& s, const void* in_hi, const void* in_lo, int B, float* out, double* stats, cudaStream_t st);
// general form: Cin / Cout multiples of 8 (the K tail and the N tail are zero-filled / masked), any output extent (the 128-voxel box
// (bw, bh, bb) with the least out-of-range voxels is chosen; out-of-range rows are masked), fused scale / shift / ReLU, operand-plane output
void conv3d_tc_run_ex(const ConvW& w, const Conv3dShape& s, const void* in_hi, const void* in_lo, int B, const Conv3dEpi& epi, cudaStream_t st);
bool conv3d_tc_supported_ex(const Conv3dShape& s);

}  // namespace ipk
