// Convolution / GEMM layer abstraction shared by the fp32 SIMT engine and the tcgen05 engine.
#pragma once
#include <vector>
#include <algorithm>
#include "common.cuh"

namespace ipk {

constexpr int MAX_TAPS = 9;

// Output mapping + epilogue of one contraction launch.
struct ConvOut {
  void* p = nullptr;          // fp32* or bf16* (hi plane)
  void* p_lo = nullptr;       // bf16 lo plane (OUT_BF16_SPLIT)
  int mode = OUT_F32_NHWC;
  int cstride = 0, coff = 0;  // channels per output pixel in memory, first channel written
  int Ho = 0, Wo = 0;         // output grid
  int ymul = 1, yadd = 0, xmul = 1, xadd = 0;   // out pixel = (f, y*ymul+yadd, x*xmul+xadd)
  int act = ACT_NONE;
  const float* bias = nullptr;        // [>=N] or null
  long long split_stride = 0;         // elements between split-K partial slices (fp32 NHWC only)
  // optional second destination (tensor-core engine): output columns [split_col, N) go to `second` (its p / p_lo / mode /
  // cstride / coff / act; column index rebased to n - split_col); pixel mapping and bias array are shared
  const ConvOut* second = nullptr;
  int split_col = 0;                  // multiple of 32
  // optional fused residual branch + output statistics (tensor-core engine, fp32 NHWC output, N >= 64):
  //   out = act(conv + bias) + res_act((res - mean) * rstd)      with res[pix][res_cstride] fp32 in the output's pixel grid and
  //   (mean, rstd) = res_mr[frame][N][2]; and stats[frame][N][2] += (sum, sum of squares) of `out` over the frame's pixels
  const float* res = nullptr;
  int res_cstride = 0, res_act = ACT_NONE;
  const float* res_mr = nullptr;
  double* stats = nullptr;            // accumulated with atomics; zero it first
};

// Activation operand of one launch.
struct ConvIn {
  const void* p = nullptr;    // fp32* (SIMT) or bf16* hi plane (TC)
  const void* p_lo = nullptr; // bf16 lo plane
  int cstride = 0, coff = 0;
  int F = 0, H = 0, W = 0;
};

// Packed weights of one layer.  Taps are stored in the order given at pack time.
struct ConvW {
  int engine = IPK_PREC_FP32_SIMT;   // ipk_precision
  int ntaps = 1, K = 0, Kpad = 0, N = 0, Npad = 0;
  float* w_f32 = nullptr;            // SIMT: [ntaps][K][Npad]
  __nv_bfloat16* w_hi = nullptr;     // TC: [ntaps][Npad][Kpad]  (K-major rows, zero padded)
  __nv_bfloat16* w_lo = nullptr;     // TC split only
  float* bias = nullptr;             // [Npad] (zero padded) or null
};

struct TapList {
  int n = 1;
  int dy[MAX_TAPS] = {0}, dx[MAX_TAPS] = {0};
  int widx[MAX_TAPS] = {0};          // index of the packed weight slice used by each tap
};

// Run all taps of `taps`.  nsplit > 1 asks for split-K into fp32 partial slices (out.split_stride apart): the SIMT engine
// splits the tap range, the tensor-core engine the (tap, 64-wide k-block) iteration range.  Both return the number of
// slices actually written.
int conv_simt_run(const ConvW& w, const ConvIn& in, const ConvOut& out, const TapList& taps, int nsplit, cudaStream_t st);
int conv_tc_run(const ConvW& w, const ConvIn& in, const ConvOut& out, const TapList& taps, int nsplit, cudaStream_t st);

// Tensor-core engine only: up to 4 tap lists ("sub-convolutions") over the same input and packed weights in ONE launch, each
// writing its own output phase (yadd, xadd) -- the four parity classes of a stride-2 ConvTranspose2d.
struct ConvSub { TapList taps; int yadd = 0, xadd = 0; };
void conv_tc_run_multi(const ConvW& w, const ConvIn& in, const ConvOut& out, const ConvSub* subs, int nsub, cudaStream_t st);

inline int conv_run(const ConvW& w, const ConvIn& in, const ConvOut& out, const TapList& taps, int nsplit, cudaStream_t st) {
  if (w.engine == IPK_PREC_FP32_SIMT) return conv_simt_run(w, in, out, taps, nsplit, st);
  return conv_tc_run(w, in, out, taps, nsplit, st);
}
inline int conv_split_count(const ConvW& w, const TapList& taps, int nsplit) {
  const int total = w.engine == IPK_PREC_FP32_SIMT ? taps.n : taps.n * (w.Kpad / 64);
  nsplit = std::max(1, std::min(nsplit, total));
  const int per = (total + nsplit - 1) / nsplit;
  return (total + per - 1) / per;
}

// ---- packing (device kernels; sources are the reference's parameter tensors, fp32 on device) ----
struct PackSrc {
  const float* w = nullptr;       // OIHW [N][Ksrc][kh][kw]; IOHW [Ksrc][N][kh][kw] when `transposed` (ConvTranspose2d)
  int N = 0, Ksrc = 0, kh = 1, kw = 1;
  bool transposed = false;
  const float* oscale = nullptr;  // per-output-channel multiplier [N] (weight-norm g/||v||), may be null
  const float* gscale = nullptr;  // device scalar; weights are divided by *gscale (spectral-norm sigma), may be null
  int k_off = 0;                  // first source input channel used
  const int* k_map = nullptr;     // optional device map packed-k -> source channel (before k_off), length K
};
// allocate zeroed packed buffers for `engine`
ConvW conv_alloc(DevPool& pool, int engine, int ntaps, int K, int N, bool with_bias);
// dst[t][n_off + n][k] = src.w(n, k_src, ky, kx) * oscale[n] / gscale   for t in tap_src (tap_src[t] = ky*kw+kx)
void conv_pack_into(ConvW& dst, int n_off, const PackSrc& src, const std::vector<int>& tap_src, cudaStream_t st);
// the same as a job record (conv_pack_job_bytes() bytes, host memory) for conv_pack_run_jobs: many packings in ONE launch from a
// device array of job records (training re-packs every layer of the flow every step)
size_t conv_pack_job_bytes();
void conv_pack_job(ConvW& dst, int n_off, const PackSrc& src, const std::vector<int>& tap_src, void* job_out);
void conv_pack_run_jobs(const void* d_jobs, int njobs, cudaStream_t st);
// dst.bias[n_off + i] = bias_src[i] + add
void conv_pack_bias(ConvW& dst, int n_off, const float* bias_src, int n, float add, cudaStream_t st);

inline TapList taps_3x3() {
  TapList t;
  t.n = 9;
  for (int i = 0; i < 9; ++i) { t.dy[i] = i / 3 - 1; t.dx[i] = i % 3 - 1; t.widx[i] = i; }
  return t;
}
inline TapList taps_1x1() { TapList t; t.n = 1; return t; }

// ConvTranspose2d(k3, s2, p1, op1) parity class (a, b): out[2y+a][2x+b] = sum over the listed taps of in[y+dy][x+dx] * W[ky][kx].
// oy = 2*iy - 1 + ky  ->  a=0: ky=1 (dy=0);  a=1: ky=0 (dy=+1), ky=2 (dy=0).   Same for columns.
inline TapList taps_convT(int a, int b) {
  TapList t;
  t.n = 0;
  int kys[2], dys[2], nky, kxs[2], dxs[2], nkx;
  if (a == 0) { nky = 1; kys[0] = 1; dys[0] = 0; } else { nky = 2; kys[0] = 0; dys[0] = 1; kys[1] = 2; dys[1] = 0; }
  if (b == 0) { nkx = 1; kxs[0] = 1; dxs[0] = 0; } else { nkx = 2; kxs[0] = 0; dxs[0] = 1; kxs[1] = 2; dxs[1] = 0; }
  for (int i = 0; i < nky; ++i)
    for (int j = 0; j < nkx; ++j) {
      t.dy[t.n] = dys[i]; t.dx[t.n] = dxs[j]; t.widx[t.n] = kys[i] * 3 + kxs[j]; t.n++;
    }
  return t;
}

}  // namespace ipk
