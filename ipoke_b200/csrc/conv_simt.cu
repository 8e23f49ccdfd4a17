// fp32 SIMT implicit-GEMM convolution (NHWC) -- the validation engine and the engine of the small layers
// (ConvGRU gates, SPADE 3->128 conv, out_conv 64->3) -- plus the weight packing kernels of both engines.
#include "conv.cuh"

namespace ipk {

struct SimtArgs {
  const float* in; int in_cstride, in_coff, K; int F, H, W;
  const float* w; int Npad, N;
  const float* bias;
  int ntaps, taps_per_split;
  int dy[MAX_TAPS], dx[MAX_TAPS], widx[MAX_TAPS];
  int act;
  float* out; __nv_bfloat16* out_lo; int out_mode, out_cstride, out_coff, Ho, Wo, ymul, yadd, xmul, xadd;
  long long split_stride;
};

constexpr int SIMT_BK = 16;

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256) conv_simt_kernel(const SimtArgs a) {
  constexpr int NT = 256;
  static_assert((BM / TM) * (BN / TN) == NT, "thread tiling");
  constexpr int A_LD = BM * SIMT_BK / NT;                                 // A elements loaded per thread
  constexpr int W_ELEMS = SIMT_BK * BN;
  constexpr int W_LD = (W_ELEMS + NT - 1) / NT;
  __shared__ __align__(16) float As[SIMT_BK][BM + 4];
  __shared__ __align__(16) float Ws[SIMT_BK][BN];

  const int tid = threadIdx.x;
  const long long M = (long long)a.F * a.H * a.W;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tap_begin = blockIdx.z * a.taps_per_split;
  const int tap_end = min(a.ntaps, tap_begin + a.taps_per_split);

  // A-load bookkeeping: element e = tid + j*NT -> (row = e / BK, k = e % BK)
  const int lk = tid % SIMT_BK;
  int pf[A_LD], py[A_LD], px[A_LD];
#pragma unroll
  for (int j = 0; j < A_LD; ++j) {
    long long m = m0 + (tid / SIMT_BK) + j * (NT / SIMT_BK);
    if (m < M) {
      int hw = a.H * a.W;
      pf[j] = (int)(m / hw);
      int r = (int)(m % hw);
      py[j] = r / a.W;
      px[j] = r % a.W;
    } else {
      pf[j] = -1; py[j] = 0; px[j] = 0;
    }
  }

  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int t = tap_begin; t < tap_end; ++t) {
    const int dy = a.dy[t], dx = a.dx[t];
    const float* wt = a.w + (size_t)a.widx[t] * a.K * a.Npad;
    for (int k0 = 0; k0 < a.K; k0 += SIMT_BK) {
      // ---- stage A (gather, zero fill) ----
#pragma unroll
      for (int j = 0; j < A_LD; ++j) {
        float v = 0.f;
        int k = k0 + lk;
        int yy = py[j] + dy, xx = px[j] + dx;
        if (pf[j] >= 0 && k < a.K && yy >= 0 && yy < a.H && xx >= 0 && xx < a.W)
          v = __ldg(a.in + ((size_t)((size_t)pf[j] * a.H + yy) * a.W + xx) * a.in_cstride + a.in_coff + k);
        As[lk][(tid / SIMT_BK) + j * (NT / SIMT_BK)] = v;
      }
      // ---- stage W ----
#pragma unroll
      for (int j = 0; j < W_LD; ++j) {
        int e = tid + j * NT;
        if (e < W_ELEMS) {
          int k = e / BN, n = e % BN;
          float v = 0.f;
          if (k0 + k < a.K && n0 + n < a.Npad) v = __ldg(wt + (size_t)(k0 + k) * a.Npad + n0 + n);
          Ws[k][n] = v;
        }
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < SIMT_BK; ++kk) {
        float av[TM], wv[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) av[i] = As[kk][ty * TM + i];
#pragma unroll
        for (int j = 0; j < TN; ++j) wv[j] = Ws[kk][tx * TN + j];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---- epilogue ----
  float* outp = a.out + (size_t)blockIdx.z * a.split_stride;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    long long m = m0 + ty * TM + i;
    if (m >= M) continue;
    int hw = a.H * a.W;
    int f = (int)(m / hw);
    int r = (int)(m % hw);
    int oy = (r / a.W) * a.ymul + a.yadd, ox = (r % a.W) * a.xmul + a.xadd;
    size_t opix = ((size_t)f * a.Ho + oy) * a.Wo + ox;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n >= a.N) continue;
      float v = acc[i][j];
      if (a.bias) v += __ldg(a.bias + n);
      v = act_apply(v, a.act);
      if (a.out_mode == OUT_F32_NHWC) {
        outp[opix * a.out_cstride + a.out_coff + n] = v;
      } else if (a.out_mode == OUT_F32_NCHW) {
        outp[(((size_t)f * a.N + n) * a.Ho + oy) * a.Wo + ox] = v;
      } else if (a.out_mode == OUT_BF16_SPLIT) {
        __nv_bfloat16 hi, lo;
        split_bf16(v, hi, lo);
        ((__nv_bfloat16*)a.out)[opix * a.out_cstride + a.out_coff + n] = hi;
        a.out_lo[opix * a.out_cstride + a.out_coff + n] = lo;
      } else {
        ((__nv_bfloat16*)a.out)[opix * a.out_cstride + a.out_coff + n] = __float2bfloat16_rn(v);
      }
    }
  }
}

int conv_simt_run(const ConvW& w, const ConvIn& in, const ConvOut& out, const TapList& taps, int nsplit, cudaStream_t st) {
  IPK_CHECK(w.w_f32 != nullptr, IPK_ERR_STATE, "conv_simt_run: layer was not packed for the SIMT engine");
  SimtArgs a;
  a.in = (const float*)in.p; a.in_cstride = in.cstride; a.in_coff = in.coff; a.K = w.K;
  a.F = in.F; a.H = in.H; a.W = in.W;
  a.w = w.w_f32; a.Npad = w.Npad; a.N = w.N;
  a.bias = out.bias;
  a.ntaps = taps.n;
  nsplit = std::max(1, std::min(nsplit, taps.n));
  a.taps_per_split = cdiv(taps.n, nsplit);
  nsplit = cdiv(taps.n, a.taps_per_split);
  for (int i = 0; i < MAX_TAPS; ++i) { a.dy[i] = taps.dy[i]; a.dx[i] = taps.dx[i]; a.widx[i] = taps.widx[i]; }
  a.act = out.act;
  a.out = (float*)out.p; a.out_lo = (__nv_bfloat16*)out.p_lo; a.out_mode = out.mode;
  a.out_cstride = out.cstride; a.out_coff = out.coff; a.Ho = out.Ho; a.Wo = out.Wo;
  a.ymul = out.ymul; a.yadd = out.yadd; a.xmul = out.xmul; a.xadd = out.xadd;
  a.split_stride = out.split_stride;
  IPK_CHECK(nsplit == 1 || (out.mode == OUT_F32_NHWC && out.split_stride > 0), IPK_ERR_INVALID, "split-K needs fp32 partial slices");
  long long M = (long long)in.F * in.H * in.W;
  if (M == 0) return 0;
  if (w.N <= 4) {
    dim3 g((unsigned)((M + 255) / 256), cdiv(w.N, 4), nsplit);
    conv_simt_kernel<256, 4, 1, 4><<<g, 256, 0, st>>>(a);
  } else if (w.N <= 32) {
    dim3 g((unsigned)((M + 127) / 128), cdiv(w.N, 32), nsplit);
    conv_simt_kernel<128, 32, 4, 4><<<g, 256, 0, st>>>(a);
  } else {
    dim3 g((unsigned)((M + 127) / 128), cdiv(w.N, 64), nsplit);
    conv_simt_kernel<128, 64, 8, 4><<<g, 256, 0, st>>>(a);
  }
  IPK_LAUNCH_CHECK();
  return nsplit;
}

// ------------------------------------------------------------------------------------------ packing
struct PackArgs {
  const float* w; int N, Ksrc, kh, kw, transposed;
  const float* oscale; const float* gscale; int k_off; const int* k_map;
  int ntaps; int tap_src[MAX_TAPS];
  int K, Kpad, Npad, n_off;
  float* dst_f32; __nv_bfloat16* dst_hi; __nv_bfloat16* dst_lo;
};

__device__ __forceinline__ void pack_elements(const PackArgs& a, long long first, long long step) {
  long long total = (long long)a.ntaps * a.N * a.K;
  for (long long e = first; e < total; e += step) {
    int k = (int)(e % a.K);
    int n = (int)((e / a.K) % a.N);
    int t = (int)(e / ((long long)a.K * a.N));
    int ks = (a.k_map ? a.k_map[k] : k) + a.k_off;
    int tap = a.tap_src[t];
    size_t si = a.transposed ? (((size_t)ks * a.N + n) * (a.kh * a.kw) + tap) : (((size_t)n * a.Ksrc + ks) * (a.kh * a.kw) + tap);
    float v = a.w[si];
    if (a.oscale) v *= a.oscale[n];
    if (a.gscale) v = v / a.gscale[0];
    if (a.dst_f32) {
      a.dst_f32[((size_t)t * a.K + k) * a.Npad + a.n_off + n] = v;
    } else {
      size_t di = ((size_t)t * a.Npad + a.n_off + n) * a.Kpad + k;
      __nv_bfloat16 hi = __float2bfloat16_rn(v);
      a.dst_hi[di] = hi;
      if (a.dst_lo) a.dst_lo[di] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
  }
}
__global__ void pack_kernel(const PackArgs a) {
  pack_elements(a, blockIdx.x * (long long)blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}
// many packing jobs in one launch (training re-packs every layer every step): blockIdx.y = job
__global__ void pack_multi_kernel(const PackArgs* __restrict__ jobs) {
  const PackArgs a = jobs[blockIdx.y];
  pack_elements(a, blockIdx.x * (long long)blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}

__global__ void pack_bias_kernel(float* dst, const float* src, int n, float add) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (src ? src[i] : 0.f) + add;
}

ConvW conv_alloc(DevPool& pool, int engine, int ntaps, int K, int N, bool with_bias) {
  ConvW w;
  w.engine = engine;
  w.ntaps = ntaps;
  w.K = K;
  w.N = N;
  if (engine == IPK_PREC_FP32_SIMT) {
    w.Kpad = K;
    w.Npad = round_up(N, 4);
    w.w_f32 = pool.alloc<float>((size_t)ntaps * K * w.Npad, true);
  } else {
    w.Kpad = round_up(K, 64);
    w.Npad = round_up(N, 16);
    w.w_hi = pool.alloc<__nv_bfloat16>((size_t)ntaps * w.Npad * w.Kpad, true);
    if (engine == IPK_PREC_FP32_SPLIT) w.w_lo = pool.alloc<__nv_bfloat16>((size_t)ntaps * w.Npad * w.Kpad, true);
  }
  if (with_bias) w.bias = pool.alloc<float>(w.Npad, true);
  return w;
}

static PackArgs make_pack_args(ConvW& dst, int n_off, const PackSrc& src, const std::vector<int>& tap_src) {
  IPK_CHECK((int)tap_src.size() == dst.ntaps && dst.ntaps <= MAX_TAPS, IPK_ERR_INVALID, "pack: tap count mismatch");
  IPK_CHECK(n_off + src.N <= dst.Npad, IPK_ERR_INVALID, "pack: N overflow");
  PackArgs a;
  memset(&a, 0, sizeof(a));
  a.w = src.w; a.N = src.N; a.Ksrc = src.Ksrc; a.kh = src.kh; a.kw = src.kw; a.transposed = src.transposed ? 1 : 0;
  a.oscale = src.oscale; a.gscale = src.gscale; a.k_off = src.k_off; a.k_map = src.k_map;
  a.ntaps = dst.ntaps;
  for (int i = 0; i < dst.ntaps; ++i) a.tap_src[i] = tap_src[i];
  a.K = dst.K; a.Kpad = dst.Kpad; a.Npad = dst.Npad; a.n_off = n_off;
  a.dst_f32 = dst.w_f32; a.dst_hi = dst.w_hi; a.dst_lo = dst.w_lo;
  return a;
}
size_t conv_pack_job_bytes() { return sizeof(PackArgs); }
void conv_pack_job(ConvW& dst, int n_off, const PackSrc& src, const std::vector<int>& tap_src, void* job_out) {
  const PackArgs a = make_pack_args(dst, n_off, src, tap_src);
  memcpy(job_out, &a, sizeof(a));
}
void conv_pack_run_jobs(const void* d_jobs, int njobs, cudaStream_t st) {
  for (int j0 = 0; j0 < njobs; j0 += 65535) {
    const int nj = std::min(65535, njobs - j0);
    pack_multi_kernel<<<dim3(48, nj), 256, 0, st>>>((const PackArgs*)d_jobs + j0);
    IPK_LAUNCH_CHECK();
  }
}

void conv_pack_into(ConvW& dst, int n_off, const PackSrc& src, const std::vector<int>& tap_src, cudaStream_t st) {
  IPK_CHECK((int)tap_src.size() == dst.ntaps && dst.ntaps <= MAX_TAPS, IPK_ERR_INVALID, "pack: tap count mismatch");
  IPK_CHECK(n_off + src.N <= dst.Npad, IPK_ERR_INVALID, "pack: N overflow");
  PackArgs a;
  a.w = src.w; a.N = src.N; a.Ksrc = src.Ksrc; a.kh = src.kh; a.kw = src.kw; a.transposed = src.transposed ? 1 : 0;
  a.oscale = src.oscale; a.gscale = src.gscale; a.k_off = src.k_off; a.k_map = src.k_map;
  a.ntaps = dst.ntaps;
  for (int i = 0; i < dst.ntaps; ++i) a.tap_src[i] = tap_src[i];
  a.K = dst.K; a.Kpad = dst.Kpad; a.Npad = dst.Npad; a.n_off = n_off;
  a.dst_f32 = dst.w_f32; a.dst_hi = dst.w_hi; a.dst_lo = dst.w_lo;
  long long total = (long long)a.ntaps * a.N * a.K;
  int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  pack_kernel<<<blocks, 256, 0, st>>>(a);
  IPK_LAUNCH_CHECK();
}

void conv_pack_bias(ConvW& dst, int n_off, const float* bias_src, int n, float add, cudaStream_t st) {
  IPK_CHECK(dst.bias != nullptr && n_off + n <= dst.Npad, IPK_ERR_INVALID, "pack bias: no bias buffer / overflow");
  pack_bias_kernel<<<cdiv(n, 128), 128, 0, st>>>(dst.bias + n_off, bias_src, n, add);
  IPK_LAUNCH_CHECK();
}

}  // namespace ipk
