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

// sample post-processing: fp32 frames [F][3][P] in [-1,1] -> uint8 [F][P][3] = trunc((x + 1) * 127.5) (second_stage_video.py:673-675)
void frames_to_u8(const float* frames_nchw, uint8_t* out_nhwc, long long F, int P, cudaStream_t st);

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
  const float* add = nullptr;    // residual [F][P][C] added after the activation (before it when act_last is set)
  bool act_last = false;         // out = act(norm(x) + add)  (BasicBlock of the 3-D encoder)
  const float* spade = nullptr;  // [B][P][2C]: (1 + gamma | beta), video index = f / T
  int T = 1;
  float* out_f32 = nullptr;      // any of the three outputs may be null
  __nv_bfloat16* out_hi = nullptr;
  __nv_bfloat16* out_lo = nullptr;
  double* stats_out = nullptr;   // optional [F][C][2]: sum / sum of squares of the values produced, accumulated (zero first);
                                 // with no output pointer set this is a pure statistics pass
};
void norm_apply(const NormApply& a, cudaStream_t st);

// Destination of a conv-input operand: fp32 rows, or bf16 (hi [, lo]) planes, channels [coff, coff + C) of rows of cstride
struct OperandDst {
  void* p = nullptr;      // float* or bf16* (hi plane)
  void* p_lo = nullptr;   // bf16 lo plane (OUT_BF16_SPLIT)
  int mode = OUT_F32_NHWC;
  int cstride = 0, coff = 0;
};
// dst[m][coff + c] = src[m][scoff + c]   (m < M, c < C), converted to the operand storage mode
void operand_copy(const float* src, int scs, int scoff, const OperandDst& dst, long long M, int C, cudaStream_t st);

// ConvGRUCell gate math (models/modules/motion_models/rnn.py:50-54); the hidden state lives in fp32 (Hf [M][z]) and the
// conv inputs (x | h), (x | r*h) in the engine's operand storage.
// raw[M][2z] = (update_pre | reset_pre): writes U[M][z] = sigmoid(update_pre) and xrh[m][coff + c] = sigmoid(reset_pre) * h
void gru_gate1(const float* raw, const float* Hf, float* U, const OperandDst& xrh, long long M, int z, cudaStream_t st);
// raw[M][z] = out_pre; h' = h*(1-u) + tanh(out_pre)*u; updates Hf in place and writes h' to up to three operand destinations
void gru_gate2(const float* raw, const float* U, float* Hf, long long M, int z, const OperandDst* dst, int ndst,
               float* seq_out, int T, int t, cudaStream_t st);

// 3x3 im2col of a small-channel NHWC image: dst[(b, y, x)][tap*C + c] = src[b][y+ky-1][x+kx-1][c] (zero outside, zero for
// k >= 9*C up to Kfill), tap = ky*3 + kx.  Feeds the SPADE 3->128 conv as a one-tap GEMM.
void im2col3x3_small(const float* src, int B, int s, int C, const OperandDst& dst, int Kfill, cudaStream_t st);

// F.interpolate(mode='bilinear', align_corners=True): in [B][3][S][S] NCHW -> out [B][s][s][3] NHWC
void bilinear_nchw_to_nhwc(const float* in, float* out, int B, int C, int S, int s, cudaStream_t st);

// Final decoder conv (3x3, 64 -> 3, tanh, NCHW frames): out_conv.cu
struct OutConvPlan;
constexpr int OUT_CONV_CIN = 64;
constexpr int OUT_CONV_PACKED_FLOATS = 9 * 64 * 3 + 3;
void out_conv_pack(const float* w_oihw, const float* sigma, const float* bias, float* dst, cudaStream_t st);
OutConvPlan* out_conv_plan_create(const float* w_dev_packed, cudaStream_t st);
void out_conv_plan_destroy(OutConvPlan* p);
// mr / spade (optional, both or none): the input is normalised in the staged tile first, y = (x - mean) * rstd * spade_gamma1 + spade_beta
// with mr [F][64][2] and spade [videos][S*S][128] = (1 + gamma | beta), video = frame / T
void out_conv_run(const OutConvPlan* p, const float* in_nhwc, float* out_nchw, int F, int S, cudaStream_t st, const float* mr = nullptr,
                  const float* spade = nullptr, int T = 1);

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
