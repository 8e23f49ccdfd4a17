// Second-stage training step of the conditional MaCow flow: forward (density direction) + log-det + FlowLoss + backward.
//   PokeMotionModel.training_step -> forward_density      models/second_stage_video.py:345-350
//   FlowLoss.forward                                      models/modules/INN/loss.py:13-31   (loss = mean_B(0.5 sum z^2) - mean_B(logdet))
//   the modules differentiated: macow2.py:97-151 (MaskedConvFlow.forward), 395-431 (NICE2d.forward), 490-513 (ActNorm2dFlow),
//   flow_blocks.py:314-326 (Shuffle), macow_utils.py:49-59 (Affine.fwd), 211-251 (Conv2dWeightNorm), 313-337, 427-434 (the nets)
//
// First cut of this row (parity first): in the density direction every op of the flow is a convolution or elementwise, so the
// whole step is expressed as contractions on the shared engines (conv.cuh: tcgen05 bf16x3 in `fp32` mode, FFMA in `fp32_simt`)
// plus small elementwise / gather kernels on global memory:
//   * forward walks the logical op list, keeping the input state of every op on a tape ([M = B*64][C0] fp32 each) and each
//     network's activations (MCF: pre-ELU hidden + params; coupling: im2col rows, both hidden activations, params; ~10 GB at
//     B = 32, IPK_TRAIN_RECOMPUTE=1 recomputes them in the backward pass instead);
//   * backward walks it in reverse:
//       affine:   dx_t = dy s,  dmu = dy,  dls = (dy x_t - (1/B)/s) * s (2 - s) / 2        (s = 1 + tanh(ls/2), logdet = sum log s)
//       dgrad:    the same engine with transposed (and tap-flipped) weight packings,
//       wgrad:    dW[n][k] = sum_m dY[m][n] X[m][k] as a GEMM whose reduction runs over the M pixels (transposed operands),
//       weight norm: (dv, dg) from dW_eff;  ActNorm / Shuffle / bias by direct reductions.
//   Weights are re-packed from the master fp32 parameters at the start of every step (they change every optimizer step): one launch
//   for all weight-norm scalings, one for all ~4 500 packings (job table built at finalize).  The step is a fixed sequence of ~30 000
//   launches and is replayed as one CUDA graph from its third call on (IPK_TRAIN_GRAPH=0 disables).
// The fused inference kernels (flow_segment.cu) are not used here; fusing this path is next-round work.
#include <map>
#include <string>
#include "conv.cuh"
#include "elementwise.cuh"
#include "flow_plan.cuh"

namespace ipk {

struct TrainRef { const void* p; float* g; int64_t numel; int dtype; };

struct TapOff { int n; int dy[9], dx[9]; };

// ------------------------------------------------------------------------------------------------ operand stores
__device__ __forceinline__ void put_op(const OperandDst& d, size_t i, float v) {
  if (d.mode == OUT_F32_NHWC) {
    ((float*)d.p)[i] = v;
  } else {
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    ((__nv_bfloat16*)d.p)[i] = hi;
    if (d.mode == OUT_BF16_SPLIT) ((__nv_bfloat16*)d.p_lo)[i] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

// dst[m][c] = c < C ? src[m*ld + c0 + c] : 0   for c < Cpad   (row-major operand, rows of dst.cstride)
__global__ void to_operand_kernel(const float* __restrict__ src, int ld, int c0, long long M, int C, int Cpad, OperandDst dst) {
  const long long total = M * Cpad;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % Cpad);
    const long long m = e / Cpad;
    put_op(dst, (size_t)m * dst.cstride + dst.coff + c, c < C ? src[m * ld + c0 + c] : 0.f);
  }
}
// dst[c][m] = src[m*ld + c0 + c]   (transposed operand: C rows of dst.cstride >= M), 32x32 smem tiles
__global__ void to_operand_T_kernel(const float* __restrict__ src, int ld, int c0, int M, int C, OperandDst dst) {
  __shared__ float t[32][33];
  const int m0 = blockIdx.x * 32, cc0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int m = m0 + r, c = cc0 + threadIdx.x;
    t[r][threadIdx.x] = (m < M && c < C) ? src[(size_t)m * ld + c0 + c] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = cc0 + r, m = m0 + threadIdx.x;
    if (c < C && m < M) put_op(dst, (size_t)c * dst.cstride + dst.coff + m, t[threadIdx.x][r]);
  }
}
// out[m][c*nt + t] = src[(b, y+dy_t, x+dx_t)][chan(c)] (0 outside the 8x8 grid); columns [nt*C, Kfill) zero.  m = b*64 + y*8 + x.
// Channel-major columns: a conv weight [N][C][taps] (OIHW) is then the GEMM weight [N][K] as it lies in memory.
__global__ void im2col_taps_kernel(const float* __restrict__ src, int ld, const int* __restrict__ idx, int C, TapOff taps, int sign,
                                   float* __restrict__ out, int ldo, int Kfill, long long M) {
  const long long total = M * Kfill;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e % Kfill);
    const long long m = e / Kfill;
    float v = 0.f;
    if (k < taps.n * C) {
      const int c = k / taps.n, t = k - c * taps.n;
      const int p = (int)(m & 63), y = (p >> 3) + sign * taps.dy[t], x = (p & 7) + sign * taps.dx[t];
      if (y >= 0 && y < 8 && x >= 0 && x < 8) v = src[((m & ~63LL) + y * 8 + x) * ld + (idx ? idx[c] : c)];
    }
    out[m * ldo + k] = v;
  }
}
// G[m][chan(c)] += sum_t dcol[(b, y-dy_t, x-dx_t)][c*nt + t]   (transpose of im2col_taps with sign = +1)
__global__ void col2im_taps_kernel(const float* __restrict__ dcol, int ldc, const int* __restrict__ idx, int C, TapOff taps,
                                   float* __restrict__ G, int ldg, long long M) {
  const long long total = M * C;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const long long m = e / C;
    const int p = (int)(m & 63);
    float acc = 0.f;
    for (int t = 0; t < taps.n; ++t) {
      const int y = (p >> 3) - taps.dy[t], x = (p & 7) - taps.dx[t];
      if (y >= 0 && y < 8 && x >= 0 && x < 8) acc += dcol[((m & ~63LL) + y * 8 + x) * ldc + c * taps.n + t];
    }
    G[m * ldg + (idx ? idx[c] : c)] += acc;
  }
}
// out[m][0..hid) = ELU(c1[m][..]);  out[m][hid..hid+hch) = E[m][..]
__global__ void elu_concat_kernel(const float* __restrict__ c1, int ld1, int hid, const float* __restrict__ E, int hch, float* __restrict__ out,
                                  int ldo, long long M) {
  const int K = hid + hch;
  const long long total = M * K;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e % K);
    const long long m = e / K;
    float v;
    if (k < hid) { v = c1[m * ld1 + k]; v = v > 0.f ? v : expm1f(v); } else v = E[m * hch + (k - hid)];
    out[m * ldo + k] = v;
  }
}
__global__ void elu_kernel(const float* __restrict__ x, float* __restrict__ out, long long n) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const float v = x[e];
    out[e] = v > 0.f ? v : expm1f(v);
  }
}
// d[m][c] *= ELU'(pre) expressed through the ELU output h: h > 0 ? 1 : h + 1
__global__ void elu_bwd_kernel(float* __restrict__ d, int ldd, const float* __restrict__ h, int ldh, int C, long long M) {
  const long long total = M * C;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const long long m = e / C;
    const float hv = h[m * ldh + c];
    d[m * ldd + c] *= hv > 0.f ? 1.0f : hv + 1.0f;
  }
}
// d[m][c] *= ELU'(pre[m][c]) = pre > 0 ? 1 : exp(pre)
__global__ void elu_bwd_pre_kernel(float* __restrict__ d, int ldd, const float* __restrict__ pre, int ldp, int C, long long M) {
  const long long total = M * C;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const long long m = e / C;
    const float x = pre[m * ldp + c];
    d[m * ldd + c] *= x > 0.f ? 1.0f : expf(x);
  }
}
// Affine.fwd on the transformed channels (macow_utils.py:49-59): blockIdx.x = sample, blockIdx.y = slice of its 64 * nt elements (the
// log-det partial sums of a sample's slices meet in one atomicAdd each); P[m][j] = mu, P[m][nt + j] = log-scale input
__global__ void affine_fwd_kernel(float* __restrict__ S, int ld, const int* __restrict__ idx, int nt, const float* __restrict__ P, int ldp,
                                  float* __restrict__ logdet) {
  const int b = blockIdx.x;
  float ldsum = 0.f;
  for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < 64 * nt; e += gridDim.y * blockDim.x) {
    const int j = e % nt;
    const size_t m = (size_t)b * 64 + e / nt;
    const float mu = P[m * ldp + j], sc = 1.0f + tanhf(0.5f * P[m * ldp + nt + j]);
    const int ch = idx ? idx[j] : j;
    S[m * ld + ch] = sc * S[m * ld + ch] + mu;
    ldsum += logf(sc);
  }
  __shared__ float red[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ldsum += __shfl_xor_sync(0xffffffffu, ldsum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ldsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) v += red[i];
    atomicAdd(&logdet[b], v);
  }
}
// backward of y_t = s x_t + mu with logdet = sum log s and dL/dlogdet_b = dld[b] (-1/B for FlowLoss):
//   G_t <- dy s;  dP = (dy | (dy x_t + dld_b / s) s (2 - s) / 2)
__global__ void affine_bwd_kernel(const float* __restrict__ X, float* __restrict__ G, int ld, const int* __restrict__ idx, int nt,
                                  const float* __restrict__ P, float* __restrict__ dP, int ldp, const float* __restrict__ dld, long long M) {
  const long long total = M * nt;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e % nt);
    const long long m = e / nt;
    const int ch = idx ? idx[j] : j;
    const float sc = 1.0f + tanhf(0.5f * P[m * ldp + nt + j]);
    const float dy = G[m * ld + ch], x = X[m * ld + ch];
    G[m * ld + ch] = dy * sc;
    dP[m * ldp + j] = dy;
    dP[m * ldp + nt + j] = (dy * x + dld[m >> 6] / sc) * 0.5f * sc * (2.0f - sc);
  }
}
// ActNorm2dFlow.forward (macow2.py:507-513): y = x exp(ls) + b on channels [coff, coff + cnt); logdet += 64 sum ls
__global__ void actnorm_fwd_kernel(float* __restrict__ S, int ld, int coff, int cnt, const float* __restrict__ ls, const float* __restrict__ bias,
                                   float* __restrict__ logdet, int B) {
  const long long M = (long long)B * 64, total = M * cnt;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % cnt);
    const long long m = e / cnt;
    S[m * ld + coff + c] = S[m * ld + coff + c] * expf(ls[c]) + bias[c];
  }
  if (blockIdx.x == 0) {
    float tot = 0.f;
    for (int c = 0; c < cnt; ++c) tot += ls[c];
    for (int b = threadIdx.x; b < B; b += blockDim.x) logdet[b] += 64.0f * tot;
  }
}
// one block per channel: dls = sum_m dy x exp(ls) + 64 sum_b dld_b (= -64 for FlowLoss), db = sum_m dy, G <- dy exp(ls)
__global__ void actnorm_bwd_kernel(const float* __restrict__ X, float* __restrict__ G, int ld, int coff, const float* __restrict__ ls,
                                   float* __restrict__ dls, float* __restrict__ dbias, const float* __restrict__ dld, long long M) {
  const int c = blockIdx.x;
  const float e = expf(ls[c]);
  float s1 = 0.f, s2 = 0.f;
  for (long long b = threadIdx.x; b < (M >> 6); b += blockDim.x) s1 += 64.0f * dld[b];
  for (long long m = threadIdx.x; m < M; m += blockDim.x) {
    const float dy = G[m * ld + coff + c];
    s1 += dy * X[m * ld + coff + c] * e;
    s2 += dy;
    G[m * ld + coff + c] = dy * e;
  }
  __shared__ float r1[32], r2[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = s1; r2[threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += r1[i]; b += r2[i]; }
    dls[c] = a;
    dbias[c] = b;
  }
}
// forward: out[m][c] = in[m][idx[c]] (c < C); backward (scatter = 1): out[m][idx[c]] = in[m][c]; channels >= C copied
__global__ void shuffle_kernel(const float* __restrict__ in, float* __restrict__ out, int ld, const int* __restrict__ idx, int C, int scatter, long long M) {
  const long long total = M * ld;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % ld);
    const long long m = e / ld;
    if (c >= C) out[e] = in[e];
    else if (scatter) out[m * ld + idx[c]] = in[e];
    else out[e] = in[m * ld + idx[c]];
  }
}
// out[c] = sum_m x[m][c]   (one block per column)
__global__ void colsum_kernel(const float* __restrict__ x, int ld, long long M, float* __restrict__ out) {
  const int c = blockIdx.x;
  float s = 0.f;
  for (long long m = threadIdx.x; m < M; m += blockDim.x) s += x[m * ld + c];
  __shared__ float r[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) a += r[i];
    out[c] = a;
  }
}
// legacy weight_norm (dim 0): Weff[o][:] = v[o][:] g[o] / ||v[o]||   (one block per output row)
__global__ void wn_apply_kernel(const float* __restrict__ v, const float* __restrict__ g, float* __restrict__ weff, int row) {
  const int o = blockIdx.x;
  float s = 0.f;
  for (int k = threadIdx.x; k < row; k += blockDim.x) { const float t = v[(size_t)o * row + k]; s = fmaf(t, t, s); }
  __shared__ float r[32];
  __shared__ float scale;
#pragma unroll
  for (int of = 16; of > 0; of >>= 1) s += __shfl_xor_sync(0xffffffffu, s, of);
  if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) a += r[i];
    scale = g[o] / sqrtf(a);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < row; k += blockDim.x) weff[(size_t)o * row + k] = v[(size_t)o * row + k] * scale;
}
struct WnJob { const float* v; const float* g; float* weff; int N, row; };
// all weight-normed layers in one launch: blockIdx.y = job, blockIdx.x = output row
__global__ void wn_apply_multi_kernel(const WnJob* __restrict__ jobs) {
  const WnJob j = jobs[blockIdx.y];
  const int o = blockIdx.x;
  if (o >= j.N) return;
  float s = 0.f;
  for (int k = threadIdx.x; k < j.row; k += blockDim.x) { const float t = j.v[(size_t)o * j.row + k]; s = fmaf(t, t, s); }
  __shared__ float r[32];
  __shared__ float scale;
#pragma unroll
  for (int of = 16; of > 0; of >>= 1) s += __shfl_xor_sync(0xffffffffu, s, of);
  if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) a += r[i];
    scale = j.g[o] / sqrtf(a);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < j.row; k += blockDim.x) j.weff[(size_t)o * j.row + k] = j.v[(size_t)o * j.row + k] * scale;
}
// P[m][j] = bias[j] + sum_s slices[s][m][j]   (split-K partial sums of a contraction)
__global__ void sum_slices_kernel(const float* __restrict__ slices, long long slice_stride, int ns, const float* __restrict__ bias, float* __restrict__ P,
                                  int ld, int N, long long M) {
  const long long total = M * N;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e % N);
    const long long m = e / N;
    float v = bias ? bias[j] : 0.f;
    for (int s2 = 0; s2 < ns; ++s2) v += slices[s2 * slice_stride + m * ld + j];
    P[m * ld + j] = v;
  }
}

// dWeff element (o, k) lives at dW[o*so + (k / inner)*sk + (k % inner)*si]  ->  dg[o] = <dW, v>/||v||,  dv = g/||v|| (dW - v <dW,v>/||v||^2)
__global__ void wn_bwd_kernel(const float* __restrict__ dW, long long so, long long sk, long long si, int inner, const float* __restrict__ v,
                              const float* __restrict__ g, float* __restrict__ dv, float* __restrict__ dg, int row) {
  const int o = blockIdx.x;
  float s = 0.f, d = 0.f;
  for (int k = threadIdx.x; k < row; k += blockDim.x) {
    const float t = v[(size_t)o * row + k];
    s = fmaf(t, t, s);
    d = fmaf(dW[o * so + (k / inner) * sk + (k % inner) * si], t, d);
  }
  __shared__ float r1[32], r2[32];
  __shared__ float nrm2, dot;
#pragma unroll
  for (int of = 16; of > 0; of >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, of); d += __shfl_xor_sync(0xffffffffu, d, of); }
  if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = s; r2[threadIdx.x >> 5] = d; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += r1[i]; b += r2[i]; }
    nrm2 = a; dot = b;
    dg[o] = b / sqrtf(a);
  }
  __syncthreads();
  const float inv = 1.0f / sqrtf(nrm2), gg = g[o];
  for (int k = threadIdx.x; k < row; k += blockDim.x) {
    const float w = dW[o * so + (k / inner) * sk + (k % inner) * si];
    dv[(size_t)o * row + k] = gg * inv * (w - v[(size_t)o * row + k] * dot / nrm2);
  }
}
// grad[o][k] = src[o*so + (k / inner)*sk + (k % inner)*si]   (plain re-layout of a weight gradient)
__global__ void relayout_kernel(const float* __restrict__ src, long long so, long long sk, long long si, int inner, float* __restrict__ grad, int row,
                                long long total) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long o = e / row;
    const int k = (int)(e % row);
    grad[e] = src[o * so + (k / inner) * sk + (k % inner) * si];
  }
}
// z = S: dz = z / B -> G;  dL/dlogdet_b = -1/B -> dld;  loss += (0.5 sum z^2 - sum_b logdet_b) / B     (*loss zeroed by the caller)
__global__ void loss_kernel(const float* __restrict__ S, float* __restrict__ G, long long n, const float* __restrict__ logdet, int B, float* __restrict__ loss,
                            float* __restrict__ dld) {
  const float invB = 1.0f / (float)B;
  if (blockIdx.x == 0)
    for (int b = threadIdx.x; b < B; b += blockDim.x) dld[b] = -invB;
  double acc = 0.0;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const float z = S[e];
    G[e] = z * invB;
    acc += 0.5 * (double)z * (double)z;
  }
  if (blockIdx.x == 0)
    for (int b = threadIdx.x; b < B; b += blockDim.x) acc -= (double)logdet[b];
  __shared__ double r[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) a += r[i];
    atomicAdd(loss, (float)(a * invB));
  }
}
// Adam / AMSGrad (torch.optim.Adam semantics, second_stage_video.py:633-636): in place on a contiguous shard
// step = lr / (1 - beta1^t), omb1 = 1 - beta1, omb2 = 1 - beta2, bc2s = sqrt(1 - beta2^t)
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float* vmax, float step, float b1, float b2, float omb1, float omb2, float eps,
                                         float wd, float bc2s, float gscale) {
  float gr = g * gscale;
  if (wd != 0.f) gr = fmaf(wd, p, gr);
  const float mm = m + (gr - m) * omb1;                  // exp_avg.lerp_(grad, 1 - beta1)
  const float vv = fmaf(omb2 * gr, gr, b2 * v);          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  m = mm;
  v = vv;
  float vh = vv;
  if (vmax) { vh = fmaxf(*vmax, vv); *vmax = vh; }
  p -= step * (mm / (sqrtf(vh) / bc2s + eps));
}
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, float* __restrict__ vmax,
                            long long n, float lr, float b1, float b2, float omb1, float omb2, float eps, float wd, float bc2, float gscale) {
  // HBM-bound (28 bytes per parameter with AMSGrad): 16-byte accesses when every buffer is 16-byte aligned, scalar tail
  const bool vec = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)vmax) & 15) == 0);
  const long long n4 = vec ? n / 4 : 0;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x) {
    float4 pp = ((float4*)p)[e], mm = ((float4*)m)[e], vv = ((float4*)v)[e];
    const float4 gg = ((const float4*)g)[e];
    float4 vm = vmax ? ((float4*)vmax)[e] : make_float4(0.f, 0.f, 0.f, 0.f);
    adam_one(pp.x, gg.x, mm.x, vv.x, vmax ? &vm.x : nullptr, lr, b1, b2, omb1, omb2, eps, wd, bc2, gscale);
    adam_one(pp.y, gg.y, mm.y, vv.y, vmax ? &vm.y : nullptr, lr, b1, b2, omb1, omb2, eps, wd, bc2, gscale);
    adam_one(pp.z, gg.z, mm.z, vv.z, vmax ? &vm.z : nullptr, lr, b1, b2, omb1, omb2, eps, wd, bc2, gscale);
    adam_one(pp.w, gg.w, mm.w, vv.w, vmax ? &vm.w : nullptr, lr, b1, b2, omb1, omb2, eps, wd, bc2, gscale);
    ((float4*)p)[e] = pp; ((float4*)m)[e] = mm; ((float4*)v)[e] = vv;
    if (vmax) ((float4*)vmax)[e] = vm;
  }
  for (long long e = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    adam_one(p[e], g[e], m[e], v[e], vmax ? vmax + e : nullptr, lr, b1, b2, omb1, omb2, eps, wd, bc2, gscale);
}

static inline int gridn(long long n) { return (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, 148LL * 16)); }

// ------------------------------------------------------------------------------------------------ small fp32 GEMM (MCF networks)
// The 800 masked-conv flows' contractions are tiny (M = B*64 rows, N, K <= 384): on the persistent tcgen05 engine each of them paid
// 16-50 us of fixed cost plus two operand-conversion passes for 1-2 us of math.  This kernel reads the fp32 tensors where they lie
// (the parameters in their checkpoint layout, activations row-major), with generic strides so that one kernel serves
//   forward  C[m][n] = sum_k A[m][k] W[n][k] (+ bias)        dgrad  C[m][k] = sum_n dY[m][n] W[n][k]
//   wgrad    C[n][k] (+)= sum_m dY[m][n] X[m][k]             (split over the M reduction, fp32 atomics)
// 64x64 output tile per CTA, 16-deep K slices staged in shared memory, 4x4 register tile per thread, plain FFMA: exact fp32.
struct SgemmArgs {
  const float* A; long long sAi, sAk;     // A(i, k) = A[i*sAi + k*sAk]
  const float* B; long long sBj, sBk;     // B(j, k) = B[j*sBj + k*sBk]
  float* C; long long ldc;                // C[i*ldc + j]
  const float* bias;                      // [J] added to every row (ksplit == 1 only), or null
  int I, J, K, kchunk, atomic;            // blockIdx.z covers k in [z*kchunk, (z+1)*kchunk); atomic: atomicAdd into C
};
// Loader: every thread fetches 4 elements of the A slice and 4 of the B slice per 16-deep step -- as ONE 16-byte load when the operand's
// unit-stride axis allows it (k-fast rows with ld % 4 == 0, or 4 consecutive rows of an i-fast operand), else as 4 scalar loads -- into
// registers BEFORE the FMAs of the current step, and stores them to the other shared-memory buffer afterwards (one barrier per step).
struct SgemmFrag { float v[4]; };
__device__ __forceinline__ SgemmFrag sgemm_fetch(const float* __restrict__ P, long long sI, long long sK, int i0, int I, int k0, int k_end, int tid, bool kfast,
                                                 bool vec) {
  SgemmFrag f;
  if (kfast) {           // thread -> (row = tid / 4, 4 consecutive k)
    const int ii = tid >> 2, kk = (tid & 3) * 4;
    const int gi = i0 + ii, gk = k0 + kk;
    if (vec && gi < I && gk + 3 < k_end) {
      const float4 t = *(const float4*)(P + gi * sI + gk);
      f.v[0] = t.x; f.v[1] = t.y; f.v[2] = t.z; f.v[3] = t.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) f.v[e] = (gi < I && gk + e < k_end) ? P[gi * sI + (gk + e) * sK] : 0.f;
    }
  } else {               // thread -> (k = tid / 16, 4 consecutive rows)
    const int kk = tid >> 4, ii = (tid & 15) * 4;
    const int gi = i0 + ii, gk = k0 + kk;
    if (vec && gi + 3 < I && gk < k_end) {
      const float4 t = *(const float4*)(P + gk * sK + gi);
      f.v[0] = t.x; f.v[1] = t.y; f.v[2] = t.z; f.v[3] = t.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) f.v[e] = (gi + e < I && gk < k_end) ? P[(gi + e) * sI + gk * sK] : 0.f;
    }
  }
  return f;
}
__device__ __forceinline__ void sgemm_stash(float (*S)[68], const SgemmFrag& f, int tid, bool kfast) {
  if (kfast) {
    const int ii = tid >> 2, kk = (tid & 3) * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) S[kk + e][ii] = f.v[e];
  } else {
    const int kk = tid >> 4, ii = (tid & 15) * 4;
    *(float4*)&S[kk][ii] = make_float4(f.v[0], f.v[1], f.v[2], f.v[3]);
  }
}
__global__ void __launch_bounds__(256) sgemm_kernel(const SgemmArgs a) {
  __shared__ __align__(16) float As[2][16][68], Bs[2][16][68];
  const int tid = threadIdx.x;
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  const int k_begin = blockIdx.z * a.kchunk, k_end = min(a.K, k_begin + a.kchunk);
  const int ti = tid >> 4, tj = tid & 15;          // thread computes rows ti*4..+3, cols tj*4..+3
  const bool a_kfast = a.sAk == 1, b_kfast = a.sBk == 1;
  // 16-byte loads need the unit-stride run of 4 elements to start on a 16-byte boundary for every thread
  const bool a_vec = a_kfast ? ((a.sAi & 3) == 0 && (((uintptr_t)a.A) & 15) == 0 && (k_begin & 3) == 0)
                             : (a.sAi == 1 && (a.sAk & 3) == 0 && (((uintptr_t)a.A) & 15) == 0);
  const bool b_vec = b_kfast ? ((a.sBj & 3) == 0 && (((uintptr_t)a.B) & 15) == 0 && (k_begin & 3) == 0)
                             : (a.sBj == 1 && (a.sBk & 3) == 0 && (((uintptr_t)a.B) & 15) == 0);
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  SgemmFrag fa = sgemm_fetch(a.A, a.sAi, a.sAk, i0, a.I, k_begin, k_end, tid, a_kfast, a_vec);
  SgemmFrag fb = sgemm_fetch(a.B, a.sBj, a.sBk, j0, a.J, k_begin, k_end, tid, b_kfast, b_vec);
  sgemm_stash(As[0], fa, tid, a_kfast);
  sgemm_stash(Bs[0], fb, tid, b_kfast);
  __syncthreads();
  int buf = 0;
  for (int k0 = k_begin; k0 < k_end; k0 += 16) {
    const bool more = k0 + 16 < k_end;
    if (more) {
      fa = sgemm_fetch(a.A, a.sAi, a.sAk, i0, a.I, k0 + 16, k_end, tid, a_kfast, a_vec);
      fb = sgemm_fetch(a.B, a.sBj, a.sBk, j0, a.J, k0 + 16, k_end, tid, b_kfast, b_vec);
    }
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 av = *(const float4*)&As[buf][kk][ti * 4];
      const float4 bv = *(const float4*)&Bs[buf][kk][tj * 4];
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(ar[r], br[c], acc[r][c]);
    }
    if (more) {
      sgemm_stash(As[buf ^ 1], fa, tid, a_kfast);
      sgemm_stash(Bs[buf ^ 1], fb, tid, b_kfast);
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int gi = i0 + ti * 4 + r;
    if (gi >= a.I) continue;
    float* crow = a.C + gi * a.ldc + j0 + tj * 4;
    if (!a.atomic && j0 + tj * 4 + 3 < a.J && (a.ldc & 3) == 0 && (((uintptr_t)a.C) & 15) == 0) {
      float4 o = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
      if (a.bias) { const float* bp = a.bias + j0 + tj * 4; o.x += bp[0]; o.y += bp[1]; o.z += bp[2]; o.w += bp[3]; }     // parameter views: 4-byte aligned only
      *(float4*)crow = o;
      continue;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int gj = j0 + tj * 4 + c;
      if (gj >= a.J) continue;
      const float v = acc[r][c];
      if (a.atomic) atomicAdd(&crow[c], v);
      else crow[c] = a.bias ? v + a.bias[gj] : v;
    }
  }
}
// ksplit > 1: the reduction is split over blockIdx.z and accumulated with atomics -- C must be zeroed by the caller (zero_first does it)
static void sgemm(const float* A, long long sAi, long long sAk, const float* B, long long sBj, long long sBk, float* C, long long ldc, const float* bias,
                  int I, int J, int K, int ksplit, bool zero_first, cudaStream_t st) {
  SgemmArgs a{A, sAi, sAk, B, sBj, sBk, C, ldc, bias, I, J, K, 0, 0};
  ksplit = std::max(1, std::min(ksplit, cdiv(K, 16)));
  a.kchunk = round_up(cdiv(K, ksplit), 16);
  ksplit = cdiv(K, a.kchunk);
  a.atomic = ksplit > 1 ? 1 : 0;
  if (a.atomic) {
    IPK_CHECK(bias == nullptr, IPK_ERR_STATE, "sgemm: bias with split-K");
    if (zero_first) {
      IPK_CHECK(ldc == J, IPK_ERR_STATE, "sgemm: zero_first needs a dense output");
      IPK_CUDA(cudaMemsetAsync(C, 0, (size_t)I * J * sizeof(float), st));
    }
  }
  sgemm_kernel<<<dim3(cdiv(J, 64), cdiv(I, 64), ksplit), 256, 0, st>>>(a);
  IPK_LAUNCH_CHECK();
}

struct McfTrain {
  std::string p; int C, Cp, order, hid, K1, C2;
  TapOff taps;
  ConvW ws, wsT, w1, w1T;             // shifted conv (6 taps) fwd / dgrad (flipped taps, transposed); 1x1 fwd (bias) / dgrad (first hid inputs)
  const float *v_ws, *v1, *g1, *b1;   // master parameters
  float *g_ws, *g_v1, *g_g1, *g_b1;   // gradient destinations
  float* weff;                        // [C2][K1]
};
struct NiceTrain {
  std::string p; int n_z, n_p, N3, N3p, K1, K1p;
  int *d_iz, *d_ip;
  ConvW c1, c1T, c2, c2T, c3, c3T;
  const float *w1, *w2, *v3, *g3, *b3;
  float *g_w1, *g_w2, *g_v3, *g_g3, *g_b3;
  float* weff3;                       // [N3][Hd*9]
};
struct TrainOp { int kind; int C; int a = -1; int coff = 0, cnt = 0; int* idx = nullptr; const float *ls = nullptr, *bias = nullptr; float *g_ls = nullptr, *g_bias = nullptr; };

}  // namespace ipk

using namespace ipk;

struct ipk_flowtrain {
  ipk_flow_config cfg;
  std::map<std::string, TrainRef> tensors;
  bool finalized = false;
  DevPool pool;
  Arena ws;
  int C0 = 0, Hd = 0, hch = 0, eng = 0, omode = OUT_F32_NHWC;
  std::vector<TrainOp> ops;
  std::vector<McfTrain> mcfs;
  std::vector<NiceTrain> nices;
  // workspace (M = max_batch * 64 rows)
  float *tape = nullptr, *G = nullptr, *Gtmp = nullptr, *logdet = nullptr, *Ecache = nullptr;
  float *c1 = nullptr, *E = nullptr, *P = nullptr, *dP = nullptr, *a1 = nullptr, *a2 = nullptr, *da = nullptr, *col = nullptr, *dcol = nullptr, *stack = nullptr,
        *wout = nullptr, *dstack = nullptr, *cond_nhwc = nullptr, *x_in = nullptr, *cond_in = nullptr, *loss_dev = nullptr, *z_dev = nullptr, *dld = nullptr, *dz_in = nullptr;
  bool use_graph = true;
  int fwd_batch = 0;                  // batch of the last ipk_flowtrain_forward whose tape is still valid
  float* slices = nullptr;            // [9][M][ldP] split-K partial sums of the NICE conv3
  // activations kept from the forward pass for the backward pass (IPK_TRAIN_RECOMPUTE=1 recomputes them instead, saving the memory):
  // per MCF the pre-ELU hidden c1 and the params P; per coupling the im2col rows, both hidden activations and P
  bool keep_acts = true;
  std::vector<float*> sv_mc1, sv_mP, sv_ncol, sv_na1, sv_na2, sv_nP;
  void* d_pack_jobs = nullptr; int n_pack_jobs = 0;
  WnJob* d_wn_jobs = nullptr; int n_wn_jobs = 0, wn_maxrows = 0;
  void *opA = nullptr, *opA_lo = nullptr, *opB = nullptr, *opB_lo = nullptr, *opT = nullptr, *opT_lo = nullptr, *opW = nullptr, *opW_lo = nullptr;
  size_t opA_elems = 0, opT_elems = 0, opW_elems = 0;
  int ldc1 = 0, ldE = 0, ldP = 0, ldcol = 0, ldstack = 0, ldwout = 0;
};

namespace ipk {

static const TrainRef& tneed(ipk_flowtrain* f, const std::string& name, int64_t numel, int dtype) {
  auto it = f->tensors.find(name);
  IPK_CHECK(it != f->tensors.end(), IPK_ERR_MISSING, "flow train: missing tensor '%s'", name.c_str());
  IPK_CHECK(it->second.numel == numel && it->second.dtype == dtype, IPK_ERR_SHAPE, "flow train: tensor '%s' has %lld elements / dtype %d, expected %lld / %d",
            name.c_str(), (long long)it->second.numel, it->second.dtype, (long long)numel, dtype);
  if (dtype == IPK_F32) IPK_CHECK(it->second.g != nullptr, IPK_ERR_MISSING, "flow train: no gradient buffer for '%s'", name.c_str());
  return it->second;
}

static TapOff mcf_taps(int order, int kh, int kw) {
  // ShiftedConv2d pad / cut table (macow_utils.py:465-484): offset of weight tap (ky, kx) relative to the output pixel
  TapOff t;
  t.n = kh * kw;
  for (int ky = 0; ky < kh; ++ky)
    for (int kx = 0; kx < kw; ++kx) {
      int dy, dx;
      switch (order) {
        case 0: dy = ky - kh; dx = kx - (kw - 1) / 2; break;         // A: rows above
        case 1: dy = ky + 1; dx = kx - (kw - 1) / 2; break;          // B: rows below
        case 2: dy = ky - (kh - 1) / 2; dx = kx - kw; break;         // C: columns to the left
        default: dy = ky - (kh - 1) / 2; dx = kx + 1; break;         // D: columns to the right
      }
      t.dy[ky * kw + kx] = dy; t.dx[ky * kw + kx] = dx;
    }
  return t;
}
static TapList taplist(const TapOff& t, int sign) {
  TapList l;
  l.n = t.n;
  for (int i = 0; i < t.n; ++i) { l.dy[i] = sign * t.dy[i]; l.dx[i] = sign * t.dx[i]; l.widx[i] = i; }
  return l;
}
static TapOff taps3x3() {
  TapOff t;
  t.n = 9;
  for (int i = 0; i < 9; ++i) { t.dy[i] = i / 3 - 1; t.dx[i] = i % 3 - 1; }
  return t;
}
static std::vector<int> iota(int n) { std::vector<int> v(n); for (int i = 0; i < n; ++i) v[i] = i; return v; }

static int* up_ints(ipk_flowtrain* f, const std::vector<int>& v, cudaStream_t st) {
  int* d = f->pool.alloc<int>(v.size());
  IPK_CUDA(cudaMemcpyAsync(d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  IPK_CUDA(cudaStreamSynchronize(st));
  return d;
}

// ---- operand plumbing -------------------------------------------------------------------------------------------------
static OperandDst opdst(ipk_flowtrain* f, void* p, void* lo, int cstride) {
  OperandDst d; d.p = p; d.p_lo = f->omode == OUT_BF16_SPLIT ? lo : nullptr; d.mode = f->omode; d.cstride = cstride; d.coff = 0;
  return d;
}
static void to_operand(ipk_flowtrain* f, const float* src, int ld, int c0, long long M, int C, int Cpad, void* p, void* lo, cudaStream_t st) {
  IPK_CHECK((size_t)M * Cpad <= f->opA_elems, IPK_ERR_STATE, "flow train: operand scratch too small (%lld x %d)", M, Cpad);
  to_operand_kernel<<<gridn(M * Cpad), 256, 0, st>>>(src, ld, c0, M, C, Cpad, opdst(f, p, lo, Cpad));
  IPK_LAUNCH_CHECK();
}
static void to_operand_T(ipk_flowtrain* f, const float* src, int ld, int c0, int M, int C, void* p, void* lo, size_t cap, cudaStream_t st) {
  IPK_CHECK((size_t)C * M <= cap, IPK_ERR_STATE, "flow train: transposed operand scratch too small (%d x %d)", C, M);
  dim3 g(cdiv(M, 32), cdiv(C, 32));
  to_operand_T_kernel<<<g, dim3(32, 8), 0, st>>>(src, ld, c0, M, C, opdst(f, p, lo, M));
  IPK_LAUNCH_CHECK();
}
// out[M][ldo] (fp32) = A[M][K] * W^T (+ bias), A a row-major operand with rows of `acs`
static void gemm(const ConvW& w, const void* a, const void* a_lo, int acs, long long M, float* out, int ldo, const float* bias, int act, cudaStream_t st) {
  ConvIn in; in.p = a; in.p_lo = a_lo; in.cstride = acs; in.F = (int)M; in.H = 1; in.W = 1;
  ConvOut o; o.p = out; o.mode = OUT_F32_NHWC; o.cstride = ldo; o.Ho = 1; o.Wo = 1; o.bias = bias; o.act = act;
  conv_run(w, in, o, taps_1x1(), 1, st);
}
// the same over the 8x8 latent grid with a tap list
static void conv8(const ConvW& w, const void* a, const void* a_lo, int acs, int B, const TapList& taps, float* out, int ldo, const float* bias, cudaStream_t st) {
  ConvIn in; in.p = a; in.p_lo = a_lo; in.cstride = acs; in.F = B; in.H = 8; in.W = 8;
  ConvOut o; o.p = out; o.mode = OUT_F32_NHWC; o.cstride = ldo; o.Ho = 8; o.Wo = 8; o.bias = bias;
  conv_run(w, in, o, taps, 1, st);
}
// out[N][ldo] = sum_m dY[m][n] X[m][k]: dYT = transposed operand of dY ([N][M] in opT), X fp32 [M][ldx] with Kx valid columns
static void wgrad(ipk_flowtrain* f, int N, const float* X, int ldx, int Kx, int M, float* out, int ldo, cudaStream_t st) {
  ConvW w;
  w.engine = f->eng; w.ntaps = 1; w.K = M; w.N = Kx;
  if (f->eng == IPK_PREC_FP32_SIMT) {
    w.Kpad = M; w.Npad = ldx; w.w_f32 = const_cast<float*>(X);        // SIMT layout [K = M][Npad = ldx] is X itself
    IPK_CHECK(ldx % 4 == 0 && ldo >= ldx, IPK_ERR_STATE, "flow train: wgrad strides (%d, %d)", ldx, ldo);
  } else {
    w.Kpad = round_up(M, 64); w.Npad = round_up(Kx, 16);
    IPK_CHECK(w.Kpad == M && ldo >= w.Npad, IPK_ERR_STATE, "flow train: wgrad needs M %% 64 == 0 and ldo >= %d (got %d, %d)", w.Npad, M, ldo);
    to_operand_T(f, X, ldx, 0, M, Kx, f->opW, f->opW_lo, f->opW_elems, st);      // X^T planes [Kx][M]
    w.w_hi = (__nv_bfloat16*)f->opW; w.w_lo = (__nv_bfloat16*)f->opW_lo;
  }
  gemm(w, f->opT, f->opT_lo, M, N, out, ldo, nullptr, ACT_NONE, st);
}

}  // namespace ipk

namespace ipk {

// ---- network forward passes (also the recomputation inside backward) ------------------------------------------------------
// MCFBlock on the taped input x [M][C0]: leaves c1 (pre-ELU hidden), E = [ELU(c1) | ELU(cond)] and P = params in the workspace.
// Both contractions read the parameters where they lie: the shifted conv is an im2col (channel-major columns c*taps + t, the OIHW
// order) followed by a GEMM against the raw weight tensor, the 1x1 a GEMM against the weight-normed matrix of this step.
static void mcf_net(ipk_flowtrain* f, const McfTrain& m, const float* x, int B, cudaStream_t st) {
  const long long M = (long long)B * 64;
  const int Kc = m.taps.n * m.C;
  im2col_taps_kernel<<<gridn(M * Kc), 256, 0, st>>>(x, f->C0, nullptr, m.C, m.taps, 1, f->stack, f->ldstack, Kc, M);
  IPK_LAUNCH_CHECK();
  sgemm(f->stack, f->ldstack, 1, m.v_ws, Kc, 1, f->c1, f->ldc1, nullptr, (int)M, m.hid, Kc, 1, false, st);
  elu_concat_kernel<<<gridn(M * m.K1), 256, 0, st>>>(f->c1, f->ldc1, m.hid, f->Ecache, f->hch, f->E, f->ldE, M);
  IPK_LAUNCH_CHECK();
  sgemm(f->E, f->ldE, 1, m.weff, m.K1, 1, f->P, f->ldP, m.b1, (int)M, m.C2, m.K1, 1, false, st);
}
// NICEConvBlock on the z channels of x: leaves col (im2col of z), a1, a2 (post-ELU) and P in the workspace
static void nice_net(ipk_flowtrain* f, const NiceTrain& n, const float* x, int B, cudaStream_t st) {
  const long long M = (long long)B * 64;
  const int Hd = f->Hd;
  im2col_taps_kernel<<<gridn(M * n.K1p), 256, 0, st>>>(x, f->C0, n.d_iz, n.n_z, taps3x3(), 1, f->col, f->ldcol, n.K1p, M);
  IPK_LAUNCH_CHECK();
  to_operand(f, f->col, f->ldcol, 0, M, n.K1, n.K1p, f->opA, f->opA_lo, st);
  gemm(n.c1, f->opA, f->opA_lo, n.K1p, M, f->a1, Hd, nullptr, ACT_ELU, st);
  to_operand(f, f->a1, Hd, 0, M, Hd, Hd, f->opA, f->opA_lo, st);
  gemm(n.c2, f->opA, f->opA_lo, Hd, M, f->a2, Hd, nullptr, ACT_ELU, st);
  to_operand(f, f->a2, Hd, 0, M, Hd, Hd, f->opA, f->opA_lo, st);
  // conv3 (K = 9 * Hd, N <= 64): split-K over the taps so that 9x as many CTAs work on it, partial sums reduced with the bias
  {
    ConvIn in; in.p = f->opA; in.p_lo = f->opA_lo; in.cstride = Hd; in.F = B; in.H = 8; in.W = 8;
    ConvOut o; o.p = f->slices; o.mode = OUT_F32_NHWC; o.cstride = f->ldP; o.Ho = 8; o.Wo = 8;
    o.split_stride = (long long)f->cfg.max_batch * 64 * f->ldP;
    const int ns = conv_run(n.c3, in, o, taplist(taps3x3(), 1), 9, st);
    sum_slices_kernel<<<gridn(M * n.N3), 256, 0, st>>>(f->slices, o.split_stride, ns, n.c3.bias, f->P, f->ldP, n.N3, M);
    IPK_LAUNCH_CHECK();
  }
}

static void train_forward(ipk_flowtrain* f, int B, cudaStream_t st) {
  const long long M = (long long)B * 64;
  const size_t slot = (size_t)f->cfg.max_batch * 64 * f->C0;
  for (size_t i = 0; i < f->ops.size(); ++i) {
    const TrainOp& op = f->ops[i];
    const float* x = f->tape + i * slot;
    float* y = f->tape + (i + 1) * slot;
    if (op.kind == L_SHUFFLE) {
      shuffle_kernel<<<gridn(M * f->C0), 256, 0, st>>>(x, y, f->C0, op.idx, op.C, 0, M);
      IPK_LAUNCH_CHECK();
      continue;
    }
    IPK_CUDA(cudaMemcpyAsync(y, x, M * f->C0 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (op.kind == L_ACTNORM) {
      actnorm_fwd_kernel<<<gridn(M * op.cnt), 256, 0, st>>>(y, f->C0, op.coff, op.cnt, op.ls, op.bias, f->logdet, B);
      IPK_LAUNCH_CHECK();
    } else if (op.kind == L_MCF) {
      ProfScope psm("train.fwd.mcf", st);
      const McfTrain& m = f->mcfs[op.a];
      float *c1s = f->c1, *Ps = f->P;
      if (f->keep_acts) { f->c1 = f->sv_mc1[op.a]; f->P = f->sv_mP[op.a]; }
      mcf_net(f, m, x, B, st);
      affine_fwd_kernel<<<dim3(B, cdiv(64 * m.C, 512)), 256, 0, st>>>(y, f->C0, nullptr, m.C, f->P, f->ldP, f->logdet);
      IPK_LAUNCH_CHECK();
      f->c1 = c1s; f->P = Ps;
    } else {
      ProfScope psn("train.fwd.nice", st);
      const NiceTrain& n = f->nices[op.a];
      float *cols = f->col, *a1s = f->a1, *a2s = f->a2, *Ps = f->P;
      if (f->keep_acts) { f->col = f->sv_ncol[op.a]; f->a1 = f->sv_na1[op.a]; f->a2 = f->sv_na2[op.a]; f->P = f->sv_nP[op.a]; }
      nice_net(f, n, x, B, st);
      affine_fwd_kernel<<<dim3(B, cdiv(64 * n.n_p, 512)), 256, 0, st>>>(y, f->C0, n.d_ip, n.n_p, f->P, f->ldP, f->logdet);
      IPK_LAUNCH_CHECK();
      f->col = cols; f->a1 = a1s; f->a2 = a2s; f->P = Ps;
    }
  }
}

static void mcf_backward(ipk_flowtrain* f, McfTrain& m, int ai, const float* x, int B, cudaStream_t st) {
  const int M = B * 64;
  float *c1s = f->c1, *Ps = f->P;
  struct Restore { ipk_flowtrain* f; float *c1, *P; ~Restore() { f->c1 = c1; f->P = P; } } restore{f, c1s, Ps};
  if (f->keep_acts) {        // c1 and P come from the forward pass; only E = [ELU(c1) | ELU(cond)] is rebuilt
    f->c1 = f->sv_mc1[ai]; f->P = f->sv_mP[ai];
    elu_concat_kernel<<<gridn((long long)M * m.K1), 256, 0, st>>>(f->c1, f->ldc1, m.hid, f->Ecache, f->hch, f->E, f->ldE, M);
    IPK_LAUNCH_CHECK();
  } else {
    mcf_net(f, m, x, B, st);                                                  // c1, E, P of this MCF
  }
  affine_bwd_kernel<<<gridn((long long)M * m.C), 256, 0, st>>>(x, f->G, f->C0, nullptr, m.C, f->P, f->dP, f->ldP, f->dld, M);
  IPK_LAUNCH_CHECK();
  colsum_kernel<<<m.C2, 256, 0, st>>>(f->dP, f->ldP, M, m.g_b1);
  IPK_LAUNCH_CHECK();
  const int Kc = m.taps.n * m.C;
  const int ks = std::max(1, M / 128);                // split of the M reduction of the weight gradients
  // 1x1: dW_eff[o][k] = sum_m dP[m][o] E[m][k]  -> (dv, dg);  dE = dP W_eff (only the first hid columns feed back)
  sgemm(f->dP, 1, f->ldP, f->E, 1, f->ldE, f->wout, m.K1, nullptr, m.C2, m.K1, M, ks, true, st);
  wn_bwd_kernel<<<m.C2, 128, 0, st>>>(f->wout, m.K1, 1, 0, 1, m.v1, m.g1, m.g_v1, m.g_g1, m.K1);
  IPK_LAUNCH_CHECK();
  sgemm(f->dP, f->ldP, 1, m.weff, 1, m.K1, f->E, f->ldE, nullptr, M, m.hid, m.C2, 1, false, st);       // E[:, :hid] <- dE
  // dc1 = dE * ELU'(c1), ELU'(x) = x > 0 ? 1 : exp(x), straight from the pre-activation
  elu_bwd_pre_kernel<<<gridn((long long)M * m.hid), 256, 0, st>>>(f->E, f->ldE, f->c1, f->ldc1, m.hid, M);
  IPK_LAUNCH_CHECK();
  // shifted conv: dWs[n][c*taps + t] = sum_m dc1[m][n] x[m + delta_t][c] (straight into the OIHW gradient);  dx += conv^T(dc1)
  im2col_taps_kernel<<<gridn((long long)M * Kc), 256, 0, st>>>(x, f->C0, nullptr, m.C, m.taps, 1, f->stack, f->ldstack, Kc, M);
  IPK_LAUNCH_CHECK();
  sgemm(f->E, 1, f->ldE, f->stack, 1, f->ldstack, m.g_ws, Kc, nullptr, m.hid, Kc, M, ks, true, st);
  sgemm(f->E, f->ldE, 1, m.v_ws, 1, Kc, f->dstack, f->ldstack, nullptr, M, Kc, m.hid, 1, false, st);      // d(im2col rows)
  col2im_taps_kernel<<<gridn((long long)M * m.C), 256, 0, st>>>(f->dstack, f->ldstack, nullptr, m.C, m.taps, f->G, f->C0, M);
  IPK_LAUNCH_CHECK();
}

static void nice_backward(ipk_flowtrain* f, NiceTrain& n, int ai, const float* x, int B, cudaStream_t st) {
  const int M = B * 64, Hd = f->Hd;
  struct Restore { ipk_flowtrain* f; float *col, *a1, *a2, *P; ~Restore() { f->col = col; f->a1 = a1; f->a2 = a2; f->P = P; } } restore{f, f->col, f->a1, f->a2, f->P};
  if (f->keep_acts) { f->col = f->sv_ncol[ai]; f->a1 = f->sv_na1[ai]; f->a2 = f->sv_na2[ai]; f->P = f->sv_nP[ai]; }
  else nice_net(f, n, x, B, st);                                              // col, a1, a2, P
  affine_bwd_kernel<<<gridn((long long)M * n.n_p), 256, 0, st>>>(x, f->G, f->C0, n.d_ip, n.n_p, f->P, f->dP, f->ldP, f->dld, M);
  IPK_LAUNCH_CHECK();
  colsum_kernel<<<n.N3, 256, 0, st>>>(f->dP, f->ldP, M, n.g_b3);
  IPK_LAUNCH_CHECK();
  // conv3 wgrad: rows (j, t) of the shifted stack of dP against a2
  ProfScope* ps3 = new ProfScope("train.bwd.nice.conv3", st);
  im2col_taps_kernel<<<gridn((long long)M * 9 * n.N3), 256, 0, st>>>(f->dP, f->ldP, nullptr, n.N3, taps3x3(), -1, f->stack, f->ldstack, 9 * n.N3, M);
  IPK_LAUNCH_CHECK();
  to_operand_T(f, f->stack, f->ldstack, 0, M, 9 * n.N3, f->opT, f->opT_lo, f->opT_elems, st);
  wgrad(f, 9 * n.N3, f->a2, Hd, Hd, M, f->wout, f->ldwout, st);                // wout[(j*9 + t)][c]
  wn_bwd_kernel<<<n.N3, 256, 0, st>>>(f->wout, 9LL * f->ldwout, 1, f->ldwout, 9, n.v3, n.g3, n.g_v3, n.g_g3, Hd * 9);
  IPK_LAUNCH_CHECK();
  // conv3 dgrad -> da2, times ELU'
  to_operand(f, f->dP, f->ldP, 0, M, n.N3, n.N3p, f->opA, f->opA_lo, st);
  conv8(n.c3T, f->opA, f->opA_lo, n.N3p, B, taplist(taps3x3(), -1), f->da, Hd, nullptr, st);
  elu_bwd_kernel<<<gridn((long long)M * Hd), 256, 0, st>>>(f->da, Hd, f->a2, Hd, Hd, M);
  IPK_LAUNCH_CHECK();
  delete ps3;
  // conv2: dW2 = da2pre^T a1 (straight into the gradient tensor), da1 = da2pre W2
  ProfScope* ps2 = new ProfScope("train.bwd.nice.conv2", st);
  to_operand_T(f, f->da, Hd, 0, M, Hd, f->opT, f->opT_lo, f->opT_elems, st);
  wgrad(f, Hd, f->a1, Hd, Hd, M, n.g_w2, Hd, st);
  to_operand(f, f->da, Hd, 0, M, Hd, Hd, f->opA, f->opA_lo, st);
  gemm(n.c2T, f->opA, f->opA_lo, Hd, M, f->a2, Hd, nullptr, ACT_NONE, st);       // a2 <- da1
  elu_bwd_kernel<<<gridn((long long)M * Hd), 256, 0, st>>>(f->a2, Hd, f->a1, Hd, Hd, M);
  IPK_LAUNCH_CHECK();
  delete ps2;
  // conv1: dW1 = da1pre^T col (OIHW as it lies), dcol = da1pre W1 -> col2im onto the z channels
  ProfScope ps1("train.bwd.nice.conv1", st);
  to_operand_T(f, f->a2, Hd, 0, M, Hd, f->opT, f->opT_lo, f->opT_elems, st);
  wgrad(f, Hd, f->col, f->ldcol, n.K1, M, f->wout, f->ldwout, st);
  relayout_kernel<<<gridn((long long)Hd * n.K1), 256, 0, st>>>(f->wout, f->ldwout, 1, 0, 1, n.g_w1, n.K1, (long long)Hd * n.K1);
  IPK_LAUNCH_CHECK();
  to_operand(f, f->a2, Hd, 0, M, Hd, Hd, f->opA, f->opA_lo, st);
  gemm(n.c1T, f->opA, f->opA_lo, Hd, M, f->dcol, f->ldcol, nullptr, ACT_NONE, st);
  col2im_taps_kernel<<<gridn((long long)M * n.n_z), 256, 0, st>>>(f->dcol, f->ldcol, n.d_iz, n.n_z, taps3x3(), f->G, f->C0, M);
  IPK_LAUNCH_CHECK();
}

static void train_backward(ipk_flowtrain* f, int B, cudaStream_t st) {
  const long long M = (long long)B * 64;
  const size_t slot = (size_t)f->cfg.max_batch * 64 * f->C0;
  for (size_t i = f->ops.size(); i-- > 0;) {
    const TrainOp& op = f->ops[i];
    const float* x = f->tape + i * slot;
    switch (op.kind) {
      case L_ACTNORM:
        actnorm_bwd_kernel<<<op.cnt, 256, 0, st>>>(x, f->G, f->C0, op.coff, op.ls, op.g_ls, op.g_bias, f->dld, M);
        IPK_LAUNCH_CHECK();
        break;
      case L_SHUFFLE:
        shuffle_kernel<<<gridn(M * f->C0), 256, 0, st>>>(f->G, f->Gtmp, f->C0, op.idx, op.C, 1, M);
        IPK_LAUNCH_CHECK();
        std::swap(f->G, f->Gtmp);
        break;
      case L_MCF: { ProfScope psm("train.bwd.mcf", st); mcf_backward(f, f->mcfs[op.a], op.a, x, B, st); break; }
      default: { ProfScope psn("train.bwd.nice", st); nice_backward(f, f->nices[op.a], op.a, x, B, st); break; }
    }
  }
}

}  // namespace ipk

// ----------------------------------------------------------------------------------------------- C ABI
extern "C" int ipk_flowtrain_create(const ipk_flow_config* cfg, ipk_flowtrain** out) {
  IPK_TRY
  IPK_CHECK(cfg && out, IPK_ERR_INVALID, "ipk_flowtrain_create: null argument");
  IPK_CHECK(cfg->n_levels > 0 && cfg->n_levels <= IPK_MAX_LEVELS && cfg->n_levels < cfg->factor, IPK_ERR_INVALID, "flow train: bad level count");
  IPK_CHECK(cfg->kernel_h == 2 && cfg->kernel_w == 3, IPK_ERR_UNSUPPORTED, "flow train: only kernel_size (2,3) is supported");
  IPK_CHECK(cfg->precision == IPK_PREC_FP32_SIMT || cfg->precision == IPK_PREC_FP32_SPLIT, IPK_ERR_UNSUPPORTED,
            "flow train: precision must be fp32 (bf16x3 tensor cores) or fp32_simt");
  IPK_CHECK(cfg->max_batch > 0 && cfg->h_channels % 8 == 0 && cfg->flow_mid_channels % 64 == 0, IPK_ERR_UNSUPPORTED, "flow train: bad sizes");
  levels_of(*cfg);
  ipk_flowtrain* f = new ipk_flowtrain();
  f->cfg = *cfg;
  f->C0 = cfg->flow_in_channels; f->Hd = cfg->flow_mid_channels; f->hch = cfg->h_channels; f->eng = cfg->precision;
  f->omode = cfg->precision == IPK_PREC_FP32_SIMT ? OUT_F32_NHWC : OUT_BF16_SPLIT;
  {
    const char* e = getenv("IPK_TRAIN_GRAPH");
    f->use_graph = !(e && e[0] == '0');
    const char* r = getenv("IPK_TRAIN_RECOMPUTE");
    f->keep_acts = !(r && r[0] == '1');
  }
  *out = f;
  IPK_CATCH
}

extern "C" int ipk_flowtrain_set_tensor(ipk_flowtrain* f, const char* name, const void* param, float* grad, int64_t numel, int dtype) {
  IPK_TRY
  IPK_CHECK(f && name && param, IPK_ERR_INVALID, "ipk_flowtrain_set_tensor: null argument");
  IPK_CHECK(!f->finalized, IPK_ERR_STATE, "ipk_flowtrain_set_tensor after finalize");
  f->tensors[name] = TrainRef{param, grad, numel, dtype};
  IPK_CATCH
}

extern "C" int ipk_flowtrain_finalize(ipk_flowtrain* f, void* stream) {
  IPK_TRY
  IPK_CHECK(f && !f->finalized, IPK_ERR_STATE, "flow train: null or already finalized");
  cudaStream_t st = (cudaStream_t)stream;
  const int Hd = f->Hd, hch = f->hch, eng = f->eng;
  int Cmax = 0, K1max = 0, N3max = 0, colmax = 0;
  for (const LogicalOp& o : logical_program(f->cfg, true)) {
    TrainOp t;
    t.kind = o.kind; t.C = o.C;
    switch (o.kind) {
      case L_ACTNORM: {
        const TrainRef& ls = tneed(f, o.prefix + "log_scale", o.cnt, IPK_F32);
        const TrainRef& b = tneed(f, o.prefix + "bias", o.cnt, IPK_F32);
        t.coff = o.coff; t.cnt = o.cnt; t.ls = (const float*)ls.p; t.bias = (const float*)b.p; t.g_ls = ls.g; t.g_bias = b.g;
        break;
      }
      case L_SHUFFLE: {
        const TrainRef& ix = tneed(f, o.prefix + "forward_shuffle_idx", o.C, IPK_I64);
        t.idx = f->pool.alloc<int>(o.C);
        i64_to_i32((const long long*)ix.p, t.idx, o.C, st);
        break;
      }
      case L_MCF: {
        McfTrain m;
        m.p = o.prefix; m.C = o.C; m.Cp = round_up(o.C, 8); m.order = o.order; m.hid = 4 * o.C; m.K1 = m.hid + hch; m.C2 = 2 * o.C;
        const int kh = o.order < 2 ? f->cfg.kernel_h : f->cfg.kernel_w, kw = o.order < 2 ? f->cfg.kernel_w : f->cfg.kernel_h;
        m.taps = mcf_taps(o.order, kh, kw);
        const TrainRef& ws = tneed(f, o.prefix + "net.shift_conv.weight", (int64_t)m.hid * o.C * kh * kw, IPK_F32);
        const TrainRef& v = tneed(f, o.prefix + "net.conv1x1.conv.weight_v", (int64_t)m.C2 * m.K1, IPK_F32);
        const TrainRef& g = tneed(f, o.prefix + "net.conv1x1.conv.weight_g", m.C2, IPK_F32);
        const TrainRef& b = tneed(f, o.prefix + "net.conv1x1.conv.bias", m.C2, IPK_F32);
        m.v_ws = (const float*)ws.p; m.v1 = (const float*)v.p; m.g1 = (const float*)g.p; m.b1 = (const float*)b.p;
        m.g_ws = ws.g; m.g_v1 = v.g; m.g_g1 = g.g; m.g_b1 = b.g;
        m.weff = f->pool.alloc<float>((size_t)m.C2 * m.K1);     // the MCF contractions run on sgemm: no packed operands
        Cmax = std::max(Cmax, o.C);
        t.a = (int)f->mcfs.size();
        f->mcfs.push_back(m);
        break;
      }
      default: {
        NiceTrain n;
        std::vector<int> iz, ip;
        nice_indices(o.C, o.factor, o.skip, o.up, iz, ip);
        n.p = o.prefix; n.n_z = (int)iz.size(); n.n_p = (int)ip.size(); n.N3 = 2 * n.n_p; n.N3p = round_up(n.N3, 8);
        n.K1 = 9 * n.n_z; n.K1p = round_up(n.K1, 8);
        n.d_iz = up_ints(f, iz, st); n.d_ip = up_ints(f, ip, st);
        const TrainRef& w1 = tneed(f, o.prefix + "net.conv1.weight", (int64_t)Hd * n.K1, IPK_F32);
        const TrainRef& w2 = tneed(f, o.prefix + "net.conv2.weight", (int64_t)Hd * Hd, IPK_F32);
        const TrainRef& v3 = tneed(f, o.prefix + "net.conv3.conv.weight_v", (int64_t)n.N3 * Hd * 9, IPK_F32);
        const TrainRef& g3 = tneed(f, o.prefix + "net.conv3.conv.weight_g", n.N3, IPK_F32);
        const TrainRef& b3 = tneed(f, o.prefix + "net.conv3.conv.bias", n.N3, IPK_F32);
        n.w1 = (const float*)w1.p; n.w2 = (const float*)w2.p; n.v3 = (const float*)v3.p; n.g3 = (const float*)g3.p; n.b3 = (const float*)b3.p;
        n.g_w1 = w1.g; n.g_w2 = w2.g; n.g_v3 = v3.g; n.g_g3 = g3.g; n.g_b3 = b3.g;
        n.c1 = conv_alloc(f->pool, eng, 1, n.K1, Hd, false);
        n.c1T = conv_alloc(f->pool, eng, 1, Hd, n.K1, false);
        n.c2 = conv_alloc(f->pool, eng, 1, Hd, Hd, false);
        n.c2T = conv_alloc(f->pool, eng, 1, Hd, Hd, false);
        n.c3 = conv_alloc(f->pool, eng, 9, Hd, n.N3, true);
        n.c3T = conv_alloc(f->pool, eng, 9, n.N3, Hd, false);
        n.weff3 = f->pool.alloc<float>((size_t)n.N3 * Hd * 9);
        K1max = std::max(K1max, n.K1p); N3max = std::max(N3max, n.N3p);
        t.a = (int)f->nices.size();
        f->nices.push_back(n);
        break;
      }
    }
    f->ops.push_back(t);
  }
  // job tables of the per-step re-packing (all pointers are fixed from here on)
  {
    std::vector<WnJob> wn;
    std::vector<char> jobs;
    const size_t jb = conv_pack_job_bytes();
    auto add = [&](ConvW& dst, const float* w, int N, int Ksrc, int ntaps, bool transposed) {
      PackSrc s2; s2.w = w; s2.N = N; s2.Ksrc = Ksrc; s2.kh = ntaps; s2.kw = 1; s2.transposed = transposed;
      jobs.resize(jobs.size() + jb);
      conv_pack_job(dst, 0, s2, iota(ntaps), jobs.data() + jobs.size() - jb);
    };
    for (McfTrain& m : f->mcfs) {
      wn.push_back(WnJob{m.v1, m.g1, m.weff, m.C2, m.K1});
      f->wn_maxrows = std::max(f->wn_maxrows, m.C2);
    }
    for (NiceTrain& n : f->nices) {
      add(n.c1, n.w1, Hd, n.K1, 1, false);
      add(n.c1T, n.w1, n.K1, Hd, 1, true);
      add(n.c2, n.w2, Hd, Hd, 1, false);
      add(n.c2T, n.w2, Hd, Hd, 1, true);
      add(n.c3, n.weff3, n.N3, Hd, 9, false);
      add(n.c3T, n.weff3, Hd, n.N3, 9, true);
      wn.push_back(WnJob{n.v3, n.g3, n.weff3, n.N3, Hd * 9});
      f->wn_maxrows = std::max(f->wn_maxrows, n.N3);
    }
    f->n_pack_jobs = (int)(jobs.size() / jb);
    f->n_wn_jobs = (int)wn.size();
    f->d_pack_jobs = f->pool.alloc<char>(std::max<size_t>(jobs.size(), 16));
    f->d_wn_jobs = f->pool.alloc<WnJob>(std::max<size_t>(wn.size(), 1));
    IPK_CUDA(cudaMemcpyAsync(f->d_pack_jobs, jobs.data(), jobs.size(), cudaMemcpyHostToDevice, st));
    IPK_CUDA(cudaMemcpyAsync(f->d_wn_jobs, wn.data(), wn.size() * sizeof(WnJob), cudaMemcpyHostToDevice, st));
    IPK_CUDA(cudaStreamSynchronize(st));
  }
  // workspace
  const size_t M = (size_t)f->cfg.max_batch * 64;
  const int hidmax = 4 * Cmax, K1m = hidmax + hch;
  f->ldc1 = round_up(std::max(hidmax, 16), 16);
  f->ldE = round_up(K1m, 16);
  f->ldP = round_up(std::max(2 * Cmax, N3max), 16);
  colmax = std::max(K1max, round_up(Cmax, 16));
  f->ldcol = round_up(colmax, 16);
  f->ldstack = round_up(std::max(9 * N3max, 6 * Cmax), 16);
  f->ldwout = round_up(std::max(std::max(Hd, K1m), f->ldcol), 16);
  const size_t wout_rows = std::max<size_t>(std::max<size_t>(Hd, f->ldstack), 2 * Cmax);
  f->opA_elems = M * std::max<size_t>(std::max<size_t>(Hd, f->ldE), std::max<size_t>(f->ldcol, f->ldP));
  f->opT_elems = std::max<size_t>(Hd, f->ldstack) * M;
  f->opW_elems = std::max<size_t>(std::max<size_t>(Hd, f->ldE), f->ldcol) * M;
  auto rb = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t slot = M * f->C0;
  size_t bytes = rb((f->ops.size() + 1) * slot * 4) + 5 * rb(slot * 4) + rb(f->cfg.max_batch * 4) + rb(f->cfg.max_batch * 4) + 3 * rb(M * hch * 4) + 4096 + rb(M * f->ldc1 * 4) + rb(M * f->ldE * 4) +
                 2 * rb(M * f->ldP * 4) + 3 * rb(M * Hd * 4) + 2 * rb(M * f->ldcol * 4) + 2 * rb(M * f->ldstack * 4) + rb(wout_rows * f->ldwout * 4) + rb(9 * M * f->ldP * 4) +
                 2 * rb(f->opA_elems * 4) + rb(f->opT_elems * 4) + rb(f->opW_elems * 4) + (1 << 16);
  f->ws.init(bytes);
  f->tape = f->ws.alloc<float>((f->ops.size() + 1) * slot);
  f->G = f->ws.alloc<float>(slot); f->Gtmp = f->ws.alloc<float>(slot);
  f->logdet = f->ws.alloc<float>(f->cfg.max_batch);
  f->dld = f->ws.alloc<float>(f->cfg.max_batch); f->dz_in = f->ws.alloc<float>(slot);
  f->cond_nhwc = f->ws.alloc<float>(M * hch); f->Ecache = f->ws.alloc<float>(M * hch);
  f->x_in = f->ws.alloc<float>(slot); f->cond_in = f->ws.alloc<float>(M * hch); f->z_dev = f->ws.alloc<float>(slot); f->loss_dev = f->ws.alloc<float>(64);
  f->c1 = f->ws.alloc<float>(M * f->ldc1); f->E = f->ws.alloc<float>(M * f->ldE);
  f->P = f->ws.alloc<float>(M * f->ldP); f->dP = f->ws.alloc<float>(M * f->ldP);
  f->a1 = f->ws.alloc<float>(M * Hd); f->a2 = f->ws.alloc<float>(M * Hd); f->da = f->ws.alloc<float>(M * Hd);
  f->col = f->ws.alloc<float>(M * f->ldcol); f->dcol = f->ws.alloc<float>(M * f->ldcol);
  f->stack = f->ws.alloc<float>(M * f->ldstack);
  f->dstack = f->ws.alloc<float>(M * f->ldstack);
  f->wout = f->ws.alloc<float>(wout_rows * f->ldwout);
  f->slices = f->ws.alloc<float>(9 * M * f->ldP);
  // operand scratch: fp32 rows (SIMT) or two bf16 planes (tensor cores) share one allocation of 4 bytes per element
  auto planes = [&](size_t elems, void** hi, void** lo) {
    char* p = (char*)f->ws.alloc<float>(elems);
    *hi = p;
    *lo = f->omode == OUT_BF16_SPLIT ? p + elems * 2 : nullptr;
  };
  planes(f->opA_elems, &f->opA, &f->opA_lo);
  planes(f->opA_elems, &f->opB, &f->opB_lo);
  planes(f->opT_elems, &f->opT, &f->opT_lo);
  planes(f->opW_elems, &f->opW, &f->opW_lo);
  if (f->keep_acts) {
    for (size_t i = 0; i < f->mcfs.size(); ++i) { f->sv_mc1.push_back(f->pool.alloc<float>(M * f->ldc1)); f->sv_mP.push_back(f->pool.alloc<float>(M * f->ldP)); }
    for (size_t i = 0; i < f->nices.size(); ++i) {
      f->sv_ncol.push_back(f->pool.alloc<float>(M * f->ldcol)); f->sv_na1.push_back(f->pool.alloc<float>(M * Hd));
      f->sv_na2.push_back(f->pool.alloc<float>(M * Hd)); f->sv_nP.push_back(f->pool.alloc<float>(M * f->ldP));
    }
  }
  IPK_CUDA(cudaMemsetAsync(f->ws.base, 0, f->ws.off, st));
  IPK_CUDA(cudaStreamSynchronize(st));
  f->finalized = true;
  IPK_CATCH
}


namespace ipk {
// re-pack every layer from the master parameters, then the density direction keeping the tape (inputs staged in x_in / cond_in)
static void train_repack_and_forward(ipk_flowtrain* f, int B, cudaStream_t st) {
  const long long M = (long long)B * 64;
  {
    ProfScope ps("train.repack", st);
    // every layer of the flow in three launches: effective weight-normed weights, all packings, then the (tiny) bias copies
    wn_apply_multi_kernel<<<dim3(f->wn_maxrows, f->n_wn_jobs), 128, 0, st>>>(f->d_wn_jobs);
    IPK_LAUNCH_CHECK();
    conv_pack_run_jobs(f->d_pack_jobs, f->n_pack_jobs, st);
    for (NiceTrain& n : f->nices) conv_pack_bias(n.c3, 0, n.b3, n.N3, 0.f, st);
  }
  nchw_to_nhwc(f->x_in, f->tape, B, f->C0, 64, f->C0, st);
  nchw_to_nhwc(f->cond_in, f->cond_nhwc, B, f->hch, 64, f->hch, st);
  elu_kernel<<<gridn(M * f->hch), 256, 0, st>>>(f->cond_nhwc, f->Ecache, M * f->hch);
  IPK_LAUNCH_CHECK();
  IPK_CUDA(cudaMemsetAsync(f->logdet, 0, B * sizeof(float), st));
  {
    ProfScope ps("train.forward", st);
    train_forward(f, B, st);
  }
}
}  // namespace ipk

// One training step without the optimizer: forward (z, logdet, loss) and the gradient of the loss w.r.t. every registered fp32
// tensor, written (not accumulated) into its gradient buffer.  x: [B][C0][8][8], cond: [B][h][8][8] (NCHW, device).
extern "C" int ipk_flowtrain_step(ipk_flowtrain* f, const float* x, const float* cond, float* loss_out, float* z_out, float* logdet_out,
                                  int32_t B, void* stream) {
  IPK_TRY
  IPK_CHECK(f && f->finalized, IPK_ERR_STATE, "flow train: not finalized");
  IPK_CHECK(x && cond && loss_out, IPK_ERR_INVALID, "ipk_flowtrain_step: null buffer");
  IPK_CHECK(B > 0 && B <= f->cfg.max_batch, IPK_ERR_INVALID, "flow train: batch %d outside (0, %d]", B, f->cfg.max_batch);
  cudaStream_t stream_ = (cudaStream_t)stream;
  const long long M = (long long)B * 64;
  // inputs are staged into plan-owned buffers so that the step is a fixed launch sequence: ~30 000 small launches whose host-side
  // cost dominates when issued one by one -> replayed as ONE CUDA graph from the third call on (IPK_TRAIN_GRAPH=0 disables)
  IPK_CUDA(cudaMemcpyAsync(f->x_in, x, (size_t)B * f->C0 * 64 * sizeof(float), cudaMemcpyDeviceToDevice, stream_));
  IPK_CUDA(cudaMemcpyAsync(f->cond_in, cond, (size_t)B * f->hch * 64 * sizeof(float), cudaMemcpyDeviceToDevice, stream_));
  const size_t slot = (size_t)f->cfg.max_batch * 64 * f->C0;
  const float* zS = f->tape + f->ops.size() * slot;
  run_graphed_step(f, B, 1, stream_, [&](cudaStream_t st) {
    train_repack_and_forward(f, B, st);
    IPK_CUDA(cudaMemsetAsync(f->loss_dev, 0, sizeof(float), st));
    loss_kernel<<<64, 256, 0, st>>>(zS, f->G, M * f->C0, f->logdet, B, f->loss_dev, f->dld);
    IPK_LAUNCH_CHECK();
    nhwc_to_nchw(zS, f->z_dev, B, f->C0, 64, f->C0, st);
    {
      ProfScope ps("train.backward", st);
      train_backward(f, B, st);
    }
  }, f->use_graph);
  IPK_CUDA(cudaMemcpyAsync(loss_out, f->loss_dev, sizeof(float), cudaMemcpyDeviceToDevice, stream_));
  if (z_out) IPK_CUDA(cudaMemcpyAsync(z_out, f->z_dev, (size_t)B * f->C0 * 64 * sizeof(float), cudaMemcpyDeviceToDevice, stream_));
  if (logdet_out) IPK_CUDA(cudaMemcpyAsync(logdet_out, f->logdet, B * sizeof(float), cudaMemcpyDeviceToDevice, stream_));
  IPK_CATCH
}


// The same step split at the loss, for callers that compute their own loss (torch.autograd.Function around the flow module:
// `out, logdet = self.flow(x, cond)` ... `loss.backward()`, models/second_stage_video.py:409-415):
//   ipk_flowtrain_forward  : re-pack, density direction with the tape kept -> z [B][C0][8][8], logdet [B]
//   ipk_flowtrain_backward : upstream gradients dz [B][C0][8][8] and dlogdet [B] (either may be null = zeros) -> every registered
//                            gradient buffer is written (not accumulated); dx_out (optional) receives dL/dx [B][C0][8][8].
//                            Must follow a forward of the same B on the same plan.
extern "C" int ipk_flowtrain_forward(ipk_flowtrain* f, const float* x, const float* cond, float* z_out, float* logdet_out, int32_t B, void* stream) {
  IPK_TRY
  IPK_CHECK(f && f->finalized, IPK_ERR_STATE, "flow train: not finalized");
  IPK_CHECK(x && cond && z_out && logdet_out, IPK_ERR_INVALID, "ipk_flowtrain_forward: null buffer");
  IPK_CHECK(B > 0 && B <= f->cfg.max_batch, IPK_ERR_INVALID, "flow train: batch %d outside (0, %d]", B, f->cfg.max_batch);
  cudaStream_t stream_ = (cudaStream_t)stream;
  IPK_CUDA(cudaMemcpyAsync(f->x_in, x, (size_t)B * f->C0 * 64 * sizeof(float), cudaMemcpyDeviceToDevice, stream_));
  IPK_CUDA(cudaMemcpyAsync(f->cond_in, cond, (size_t)B * f->hch * 64 * sizeof(float), cudaMemcpyDeviceToDevice, stream_));
  const size_t slot = (size_t)f->cfg.max_batch * 64 * f->C0;
  const float* zS = f->tape + f->ops.size() * slot;
  run_graphed_step(f, B, 2, stream_, [&](cudaStream_t st) {
    train_repack_and_forward(f, B, st);
    nhwc_to_nchw(zS, f->z_dev, B, f->C0, 64, f->C0, st);
  }, f->use_graph);
  f->fwd_batch = B;
  IPK_CUDA(cudaMemcpyAsync(z_out, f->z_dev, (size_t)B * f->C0 * 64 * sizeof(float), cudaMemcpyDeviceToDevice, stream_));
  IPK_CUDA(cudaMemcpyAsync(logdet_out, f->logdet, B * sizeof(float), cudaMemcpyDeviceToDevice, stream_));
  IPK_CATCH
}

extern "C" int ipk_flowtrain_backward(ipk_flowtrain* f, const float* dz, const float* dlogdet, float* dx_out, int32_t B, void* stream) {
  IPK_TRY
  IPK_CHECK(f && f->finalized, IPK_ERR_STATE, "flow train: not finalized");
  IPK_CHECK(B > 0 && B == f->fwd_batch, IPK_ERR_STATE, "ipk_flowtrain_backward: batch %d does not match the preceding forward (%d)", B, f->fwd_batch);
  cudaStream_t stream_ = (cudaStream_t)stream;
  const size_t n = (size_t)B * f->C0 * 64;
  if (dz) IPK_CUDA(cudaMemcpyAsync(f->dz_in, dz, n * sizeof(float), cudaMemcpyDeviceToDevice, stream_));
  else IPK_CUDA(cudaMemsetAsync(f->dz_in, 0, n * sizeof(float), stream_));
  if (dlogdet) IPK_CUDA(cudaMemcpyAsync(f->dld, dlogdet, B * sizeof(float), cudaMemcpyDeviceToDevice, stream_));
  else IPK_CUDA(cudaMemsetAsync(f->dld, 0, B * sizeof(float), stream_));
  run_graphed_step(f, B, 3, stream_, [&](cudaStream_t st) {
    nchw_to_nhwc(f->dz_in, f->G, B, f->C0, 64, f->C0, st);
    {
      ProfScope ps("train.backward", st);
      train_backward(f, B, st);
    }
    nhwc_to_nchw(f->G, f->dz_in, B, f->C0, 64, f->C0, st);      // dL/dx (the staging buffer is free again)
  }, f->use_graph);
  if (dx_out) IPK_CUDA(cudaMemcpyAsync(dx_out, f->dz_in, n * sizeof(float), cudaMemcpyDeviceToDevice, stream_));
  f->fwd_batch = 0;
  IPK_CATCH
}

void ipk_graphs_drop(const void* handle);   // capi.cu
extern "C" int ipk_flowtrain_destroy(ipk_flowtrain* f) {
  if (!f) return IPK_OK;
  ipk_graphs_drop(f);
  f->pool.release();
  f->ws.release();
  delete f;
  return IPK_OK;
}

// torch.optim.Adam(amsgrad) on a contiguous fp32 shard (second_stage_video.py:633-636); step counts from 1; vmax may be null (plain Adam).
// grad_scale multiplies the gradient first (1 / world_size after a summing reduce-scatter).
extern "C" int ipk_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq, int64_t n, double lr, double beta1,
                             double beta2, double eps, double weight_decay, int32_t step, float grad_scale, void* stream) {
  IPK_TRY
  IPK_CHECK(param && grad && exp_avg && exp_avg_sq && n >= 0 && step >= 1, IPK_ERR_INVALID, "ipk_adam_step: bad argument");
  if (n == 0) return IPK_OK;
  // torch.optim.adam._single_tensor_adam: step_size = lr / (1 - beta1^t) and sqrt(1 - beta2^t) are Python doubles, rounded to fp32 only
  // when they meet the tensors (1 - powf(0.999f, 1) would already be off by 1e-5 relative)
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  adam_kernel<<<gridn(n), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, max_exp_avg_sq, n, (float)(lr / bc1), (float)beta1, (float)beta2,
                                                          (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps, (float)weight_decay, (float)sqrt(bc2), grad_scale);
  IPK_LAUNCH_CHECK();
  IPK_CATCH
}
