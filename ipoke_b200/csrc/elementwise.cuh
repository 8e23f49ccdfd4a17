// Elementwise / reduction kernels of the first-stage decoder and layout conversion at the ABI edge.
#pragma once
#include "common.cuh"

namespace ipk {

// [B][C][P] fp32 -> [B][P][cstride] fp32 (channels [0,C) written)
void nchw_to_nhwc(const float* in, float* out, int B, int C, int P, int cstride, cudaStream_t st);
void nhwc_to_nchw(const float* in, float* out, int B, int C, int P, int cstride, cudaStream_t st);
// out[b][p][coff + c] = src[c-major broadcast tensor [C][P]]  (motion_bias repeated over the batch)
void broadcast_chw_to_nhwc(const float* src, float* out, int B, int C, int P, int cstride, int coff, cudaStream_t st);
// copy channels: dst[m][dcoff + c] = src[m][scoff + c], m < M, c < C
void copy_channels(const float* src, int scs, int scoff, float* dst, int dcs, int dcoff, long long M, int C, cudaStream_t st);

// per-(frame, channel) sum and sum of squares over the P pixels of a frame: sums[F][C][2] (double, accumulated; zero first)
void channel_stats(const float* x, int F, long long P, int C, double* sums, cudaStream_t st);
// mr[F][C][2] = (mean, rstd) per channel; groups == 0: InstanceNorm (per channel), else GroupNorm(groups) stats replicated
// to the channels of each group.  Biased variance, eps inside the sqrt (torch semantics).
void finalize_stats(const double* sums, float* mr, int F, long long P, int C, int groups, float eps, cudaStream_t st);

struct NormApply {
  const float* x = nullptr;      // [F][P][C]
  int F = 0, C = 0;
  long long P = 0;
  const float* mr = nullptr;     // [F][C][2] or null
  const float* w = nullptr;      // per-channel affine weight or null
  const float* b = nullptr;
  int act = ACT_NONE;
  const float* add = nullptr;    // residual [F][P][C] added after the activation
  const float* spade = nullptr;  // [B][P][2C]: (1 + gamma | beta), video index = f / T
  int T = 1;
  float* out_f32 = nullptr;      // any of the three outputs may be null
  __nv_bfloat16* out_hi = nullptr;
  __nv_bfloat16* out_lo = nullptr;
};
void norm_apply(const NormApply& a, cudaStream_t st);

// ConvGRUCell gate math (models/modules/motion_models/rnn.py:50-54)
// raw[M][2z] = (update_pre | reset_pre); xh[M][2z] holds (x | h); writes U[M][z] = sigmoid(update_pre) and
// xrh[M][z + c] = sigmoid(reset_pre) * h
void gru_gate1(const float* raw, const float* xh, float* U, float* xrh, long long M, int z, cudaStream_t st);
// raw[M][z] = out_pre; h' = h*(1-u) + tanh(out_pre)*u; writes h' to up to four destinations (null = skip)
struct GruDst { float* p; int cstride, coff; };
void gru_gate2(const float* raw, const float* U, const float* xh, long long M, int z, const GruDst* dst, int ndst,
               float* seq_out, int T, int t, cudaStream_t st);

// F.interpolate(mode='bilinear', align_corners=True): in [B][3][S][S] NCHW -> out [B][s][s][3] NHWC
void bilinear_nchw_to_nhwc(const float* in, float* out, int B, int C, int S, int s, cudaStream_t st);

// weight-norm row scale: oscale[n] = g[n] / ||v[n,:]||_2   (torch.nn.utils.weight_norm, dim=0)
void weight_norm_scale(const float* v, const float* g, float* oscale, int N, int row, cudaStream_t st);
// spectral-norm sigma = u^T W_mat v with W_mat = weight viewed [rows][cols] (dim 0) or permuted (dim 1, ConvTranspose)
void spectral_sigma(const float* w, const float* u, const float* v, float* sigma, int d0, int d1, int rest, bool dim1, cudaStream_t st);
// MCF weight packing into the canonical line layout of flow_segment.cu
void pack_mcf_shift(const float* w, float* dst, int hid, int C, int Cp, int kh, int kw, int order, cudaStream_t st);
// dst[k/4][o][k%4] = w[o][k_off + k] * oscale[o]  for k < K (K multiple of 4)
void pack_rows4(const float* w, const float* oscale, float* dst, int O, int row, int k_off, int K, cudaStream_t st);
void i64_to_i32(const long long* src, int* dst, int n, cudaStream_t st);

}  // namespace ipk
