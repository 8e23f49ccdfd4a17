// First-stage 3-D conv video encoder (training path of the second stage):
//   PokeMotionModel.encode_first_stage      models/second_stage_video.py:352-359
//   ResNetMotionEncoder.forward / __init__  models/modules/motion_models/motion_encoder.py:150-241
//   BasicBlock                              models/modules/motion_models/motion_encoder.py:45-74
//
// Data layout: NDHWC fp32 [B][T][H][W][C] (the 3 input channels padded to 4).  Every Conv3d is an implicit GEMM over
// (tap, channel) with the taps outermost, so one k-chunk of a tile row is a contiguous run of channels of one input voxel.
// GroupNorm(16) statistics run over the whole (T, H, W) volume of a sample: the shared norm pass with F = B, P = T*H*W.
// Tensor-core precisions (fp32 = bf16x3, bf16): every Conv3d with >= 64 input channels runs on the tcgen05 engine of conv3d_tc.cu
// (5-D TMA boxes, strides as element strides, temporal-padding taps skipped, GroupNorm statistics fused into the epilogue); the
// norm pass then writes the next conv's bf16 operand planes next to the fp32 activation.  The 3-channel stem and the fp32_simt
// validation mode use the FFMA kernel below.
#include <map>
#include <string>
#include <cmath>
#include "conv.cuh"
#include "conv3d_tc.cuh"
#include "elementwise.cuh"

namespace ipk {

struct TensorRefE { const void* p; int64_t numel; };

struct Conv3dDesc {
  int Cin, Cout;          // Cin as stored (multiple of 4)
  int kt, ky, kx, st, sy, sx, pt, py, px;
  float* w = nullptr;     // [kt][ky][kx][Cin][Cout]
  float* bias = nullptr;  // [Cout] or null
  ConvW tcw;              // tensor-core packing [tap][Cout][Cin] (hi [, lo]) when `tc`
  bool tc = false;
};

// an activation volume: fp32 NDHWC, plus (tensor-core precisions) the bf16 operand planes of the same values
struct ActVol { float* f = nullptr; __nv_bfloat16* hi = nullptr; __nv_bfloat16* lo = nullptr; };

struct Conv3dArgs {
  const float* in; float* out; const float* w; const float* bias;
  int B, Ti, Hi, Wi, Cin, To, Ho, Wo, Cout;
  int kt, ky, kx, st, sy, sx, pt, py, px;
};

constexpr int C3_BM = 64, C3_BN = 64, C3_BK = 16;

// 256 threads: 64 output voxels x 64 output channels per CTA, 4 x 4 outputs per thread
__global__ void __launch_bounds__(256) conv3d_simt_kernel(const Conv3dArgs a) {
  __shared__ __align__(16) float As[C3_BK][C3_BM + 4];
  __shared__ __align__(16) float Ws[C3_BK][C3_BN];
  __shared__ int vb[C3_BM], vt[C3_BM], vy[C3_BM], vx[C3_BM];
  pdl_wait();
  pdl_trigger();
  const int tid = threadIdx.x;
  const long long M = (long long)a.B * a.To * a.Ho * a.Wo;
  const long long m0 = (long long)blockIdx.x * C3_BM;
  const int n0 = blockIdx.y * C3_BN;
  if (tid < C3_BM) {
    long long m = m0 + tid;
    if (m < M) {
      int x = (int)(m % a.Wo); m /= a.Wo;
      int y = (int)(m % a.Ho); m /= a.Ho;
      int t = (int)(m % a.To);
      vb[tid] = (int)(m / a.To); vt[tid] = t * a.st - a.pt; vy[tid] = y * a.sy - a.py; vx[tid] = x * a.sx - a.px;
    } else {
      vb[tid] = -1; vt[tid] = 0; vy[tid] = 0; vx[tid] = 0;
    }
  }
  __syncthreads();
  const int kc = a.Cin < C3_BK ? a.Cin : C3_BK;          // channels per k-chunk (Cin is 4 or a multiple of 16)
  const int arow = tid >> 2, aq = tid & 3;               // A staging: row, float4 index inside the chunk
  const int tx = tid & 15, ty = tid >> 4;                // compute: 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ntaps = a.kt * a.ky * a.kx;
  for (int tap = 0; tap < ntaps; ++tap) {
    const int dt = tap / (a.ky * a.kx), dy = (tap / a.kx) % a.ky, dx = tap % a.kx;
    const int ti = vt[arow] + dt, yi = vy[arow] + dy, xi = vx[arow] + dx;
    const bool ok = vb[arow] >= 0 && ti >= 0 && ti < a.Ti && yi >= 0 && yi < a.Hi && xi >= 0 && xi < a.Wi;
    const float* ip = a.in + ((((size_t)max(vb[arow], 0) * a.Ti + max(ti, 0)) * a.Hi + max(yi, 0)) * a.Wi + max(xi, 0)) * a.Cin;
    const float* wt = a.w + (size_t)tap * a.Cin * a.Cout;
    for (int c0 = 0; c0 < a.Cin; c0 += kc) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok && aq * 4 < kc) v = __ldg((const float4*)(ip + c0) + aq);
      As[aq * 4 + 0][arow] = v.x; As[aq * 4 + 1][arow] = v.y; As[aq * 4 + 2][arow] = v.z; As[aq * 4 + 3][arow] = v.w;
      for (int e = tid; e < C3_BK * C3_BN; e += 256) {
        const int k = e / C3_BN, n = e % C3_BN;
        Ws[k][n] = (k < kc && n0 + n < a.Cout) ? __ldg(wt + (size_t)(c0 + k) * a.Cout + n0 + n) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < C3_BK; ++kk) {
        float av[4], wv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) wv[j] = Ws[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < a.Cout) a.out[(size_t)m * a.Cout + n] = acc[i][j] + (a.bias ? __ldg(a.bias + n) : 0.f);
    }
  }
}

// w: OIDHW [Cout][CinSrc][kt][ky][kx] -> [tap][Cin (zero padded)][Cout]
__global__ void pack_conv3d_kernel(const float* __restrict__ w, float* __restrict__ dst, int Cout, int CinSrc, int Cin, int ntaps,
                                   const float* __restrict__ sigma) {
  const long long total = (long long)ntaps * Cin * Cout;
  const float inv = sigma ? 1.0f / sigma[0] : 1.0f;     // eval-mode spectral norm: W / sigma
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(e % Cout), c = (int)((e / Cout) % Cin), tap = (int)(e / ((long long)Cout * Cin));
    dst[e] = c < CinSrc ? (sigma ? w[((size_t)n * CinSrc + c) * ntaps + tap] / sigma[0] : w[((size_t)n * CinSrc + c) * ntaps + tap]) : 0.f;
    (void)inv;
  }
}

// X [B][3][T][H][W] (NCDHW) -> [B][T][H][W][4] (4th channel zero)
__global__ void ncdhw_to_ndhwc4_kernel(const float* __restrict__ in, float* __restrict__ out, int B, long long V) {
  const long long total = (long long)B * V;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / V, v = e % V;
    const float* ip = in + (size_t)b * 3 * V + v;
    ((float4*)out)[e] = make_float4(ip[0], ip[V], ip[2 * V], 0.f);
  }
}

// Tensor-core stem (Conv3d 3 -> C0, (3,7,7), stride 2, pad (1,3,3)): the 7 x-taps are folded into K.  The input is stored as bf16
// planes [B][T][H][Wp][8] (3 channels + 5 zeros per pixel, 3 zero pixels of left padding, Wp = W + 8), so the 8 pixels x 8 channels
// starting at padded pixel 2*ox are 64 CONTIGUOUS elements = one 128-byte K block of output column ox; the TMA map views the rows with
// an overlapped x stride of 2 pixels (32 B).  The conv then has 3 x 7 (dt, dy) taps of K = 64 (56 used), y / t padding by OOB fill.
constexpr int STEM_CP = 8, STEM_XPAD = 3;
__global__ void stem_input_planes_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int B, int T,
                                         int H, int W, int Wp) {
  const long long rows = (long long)B * T * H, total = rows * Wp;
  const long long V = (long long)T * H * W;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int xp = (int)(e % Wp);
    const long long row = e / Wp;                      // (b, t, y)
    const long long b = row / ((long long)T * H), ty = row % ((long long)T * H);
    const int x = xp - STEM_XPAD;
    float v[3] = {0.f, 0.f, 0.f};
    if (x >= 0 && x < W) {
      const float* ip = in + (size_t)b * 3 * V + (size_t)ty * W + x;
      v[0] = ip[0]; v[1] = ip[V]; v[2] = ip[2 * V];
    }
    uint32_t h[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0};
    __nv_bfloat16 hb[3], lb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { hb[c] = __float2bfloat16_rn(v[c]); lb[c] = __float2bfloat16_rn(v[c] - __bfloat162float(hb[c])); }
    h[0] = (uint32_t)__bfloat16_as_ushort(hb[0]) | ((uint32_t)__bfloat16_as_ushort(hb[1]) << 16); h[1] = (uint32_t)__bfloat16_as_ushort(hb[2]);
    l[0] = (uint32_t)__bfloat16_as_ushort(lb[0]) | ((uint32_t)__bfloat16_as_ushort(lb[1]) << 16); l[1] = (uint32_t)__bfloat16_as_ushort(lb[2]);
    ((uint4*)hi)[e] = make_uint4(h[0], h[1], h[2], h[3]);
    if (lo) ((uint4*)lo)[e] = make_uint4(l[0], l[1], l[2], l[3]);
  }
}
// w: OIDHW [Cout][3][3][7][7] -> planes [tap = dt*7 + dy][Npad][64] with k = dx*8 + c
__global__ void pack_stem_tc_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int Cout, int Npad) {
  const int total = 21 * Cout * 21;                    // (tap, n, dx*3 + c)
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int j = e % 21, n = (e / 21) % Cout, tap = e / (21 * Cout);
    const int dx = j / 3, c = j % 3, dt = tap / 7, dy = tap % 7;
    const float v = w[((((size_t)n * 3 + c) * 3 + dt) * 7 + dy) * 7 + dx];
    const size_t di = ((size_t)tap * Npad + n) * 64 + dx * STEM_CP + c;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[di] = h;
    if (lo) lo[di] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// x [B][C][P] (NCHW, C <= 4) -> [B][P][4] (missing channels zero)
__global__ void nchw_to_nhwc4_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int C, long long P) {
  const long long total = (long long)B * P;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / P, p = e % P;
    const float* ip = in + (size_t)b * C * P + p;
    ((float4*)out)[e] = make_float4(ip[0], C > 1 ? ip[P] : 0.f, C > 2 ? ip[2 * P] : 0.f, C > 3 ? ip[3 * P] : 0.f);
  }
}

// x [B][C][P] (NCHW, C <= Cp) -> [B][P][Cp] (missing channels zero), Cp a multiple of 4
__global__ void nchw_to_nhwc_pad_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int C, int Cp, long long P) {
  const long long total = (long long)B * P * Cp;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % Cp);
    const long long bp = e / Cp, b = bp / P, p = bp % P;
    out[e] = c < C ? in[((size_t)b * C + c) * P + p] : 0.f;
  }
}

// heads [B*64][2z] (mu | logvar) + eps [B][z][64] -> z, mu, logvar NCHW [B][z][64]     (reparameterize, motion_encoder.py:218-222)
__global__ void reparam_kernel(const float* __restrict__ heads, const float* __restrict__ eps, float* __restrict__ zo, float* __restrict__ mu,
                               float* __restrict__ lv, int B, int z) {
  const int total = B * z * 64;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int p = e % 64, c = (e / 64) % z, b = e / (64 * z);
    const float m = heads[((size_t)b * 64 + p) * 2 * z + c], l = heads[((size_t)b * 64 + p) * 2 * z + z + c];
    mu[e] = m;
    lv[e] = l;
    zo[e] = eps[e] * expf(0.5f * l) + m;
  }
}

struct EncBlock {
  Conv3dDesc c1, c2, ds;
  bool has_ds = false;
  float *g1w, *g1b, *g2w, *g2b, *gdw, *gdb;
};

}  // namespace ipk

using namespace ipk;

struct ipk_enc {
  ipk_enc_config cfg;
  std::map<std::string, TensorRefE> tensors;
  bool finalized = false;
  DevPool pool;
  Arena ws;
  Conv3dDesc stem;
  ConvW stem_tc;             // tensor-core stem: 21 (dt, dy) taps x K = 64 (x taps folded into K), see stem_input_planes_kernel
  bool stem_on_tc = false;
  __nv_bfloat16 *Xp_hi = nullptr, *Xp_lo = nullptr;
  float *stem_gw = nullptr, *stem_gb = nullptr;
  std::vector<EncBlock> blocks;
  ConvW heads;               // conv_mu | conv_var fused along N (2-D 3x3, fp32 FFMA engine)
  int last_C = 0;
  // workspace
  float *X4 = nullptr, *bufC = nullptr, *bufD = nullptr, *hbuf = nullptr;
  ActVol actA, actB, actX;      // block input / output ping-pong (actA, actX) and the mid-block activation (actB)
  double* sums = nullptr; float* mr = nullptr;
  size_t act_elems = 0;
};

namespace ipk {

static const TensorRefE& eneed(ipk_enc* e, const std::string& name, int64_t numel) {
  auto it = e->tensors.find(name);
  IPK_CHECK(it != e->tensors.end(), IPK_ERR_MISSING, "encoder: missing tensor '%s'", name.c_str());
  IPK_CHECK(it->second.numel == numel, IPK_ERR_SHAPE, "encoder: tensor '%s' has %lld elements, expected %lld", name.c_str(),
            (long long)it->second.numel, (long long)numel);
  return it->second;
}
static float* ecopy(ipk_enc* e, const std::string& name, int n, cudaStream_t st) {
  float* p = e->pool.alloc<float>(n);
  IPK_CUDA(cudaMemcpyAsync(p, eneed(e, name, n).p, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return p;
}
static Conv3dDesc build_conv3d(ipk_enc* e, const std::string& name, int Cout, int CinSrc, int kt, int ky, int kx, int st_, int sy, int sx,
                               int pt, int py, int px, cudaStream_t st) {
  Conv3dDesc d;
  d.Cin = round_up(CinSrc, 4); d.Cout = Cout;
  d.kt = kt; d.ky = ky; d.kx = kx; d.st = st_; d.sy = sy; d.sx = sx; d.pt = pt; d.py = py; d.px = px;
  const int ntaps = kt * ky * kx;
  const TensorRefE& w = eneed(e, name, (int64_t)Cout * CinSrc * ntaps);
  d.w = e->pool.alloc<float>((size_t)ntaps * d.Cin * Cout);
  const long long total = (long long)ntaps * d.Cin * Cout;
  pack_conv3d_kernel<<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, st>>>((const float*)w.p, d.w, Cout, CinSrc, d.Cin, ntaps, nullptr);
  IPK_LAUNCH_CHECK();
  if (e->cfg.precision != IPK_PREC_FP32_SIMT && CinSrc % 64 == 0 && Cout % 64 == 0) {
    d.tcw = conv_alloc(e->pool, e->cfg.precision, ntaps, CinSrc, Cout, false);
    conv3d_tc_pack(d.tcw, (const float*)w.p, Cout, CinSrc, st);
    d.tc = true;
  }
  return d;
}
static inline int out_extent(int n, int k, int s, int p) { return (n + 2 * p - k) / s + 1; }

struct Vol { int T, H, W, C; size_t voxels() const { return (size_t)T * H * W; } };

static Vol conv3d_out(const Conv3dDesc& d, const Vol& v) {
  return Vol{out_extent(v.T, d.kt, d.st, d.pt), out_extent(v.H, d.ky, d.sy, d.py), out_extent(v.W, d.kx, d.sx, d.px), d.Cout};
}

static Conv3dShape shape_of(const Conv3dDesc& d, const Vol& v) {
  return Conv3dShape{d.Cin, d.Cout, v.T, v.H, v.W, d.kt, d.ky, d.kx, d.st, d.sy, d.sx, d.pt, d.py, d.px};
}

// Conv3d of `in` into the fp32 volume `out`.  Returns true when the GroupNorm statistics of the output were accumulated into
// `sums` by the conv's own epilogue (tensor-core path); the caller zeroes `sums` first.
static bool run_conv3d(const Conv3dDesc& d, const ActVol& in, const Vol& v, float* out, int B, double* sums, Vol* ov, cudaStream_t st) {
  IPK_CHECK(v.C == d.Cin, IPK_ERR_STATE, "encoder: conv input has %d channels, layer expects %d", v.C, d.Cin);
  const Vol o = conv3d_out(d, v);
  *ov = o;
  if (d.tc && in.hi != nullptr && conv3d_tc_supported(shape_of(d, v))) {
    ProfScope ps("enc.conv3d.tc", st);
    conv3d_tc_run(d.tcw, shape_of(d, v), in.hi, in.lo, B, out, sums, st);
    return sums != nullptr;
  }
  ProfScope ps("enc.conv3d.simt", st);
  Conv3dArgs a{in.f, out, d.w, d.bias, B, v.T, v.H, v.W, d.Cin, o.T, o.H, o.W, d.Cout, d.kt, d.ky, d.kx, d.st, d.sy, d.sx, d.pt, d.py, d.px};
  const long long M = (long long)B * o.voxels();
  dim3 g((unsigned)((M + C3_BM - 1) / C3_BM), (unsigned)cdiv(d.Cout, C3_BN));
  launch_k(conv3d_simt_kernel, g, dim3(256), 0, st, a);
  return false;
}

// conv -> GroupNorm(16) over the (T, H, W) volume of each sample (+ affine, optional residual, ReLU) -> `out` (fp32 and, when
// its plane pointers are set, the bf16 operand planes of the next conv).  `tmp` receives the raw conv output.
static Vol conv_gn(ipk_enc* e, const Conv3dDesc& d, const ActVol& in, const Vol& v, float* tmp, int B, const float* gw, const float* gb,
                   const float* add, bool relu, const ActVol& out, cudaStream_t st) {
  Vol o;
  IPK_CUDA(cudaMemsetAsync(e->sums, 0, (size_t)B * d.Cout * 2 * sizeof(double), st));
  const bool have_stats = run_conv3d(d, in, v, tmp, B, e->sums, &o, st);
  ProfScope ps("enc.group_norm", st);
  const long long P = (long long)o.voxels();
  if (!have_stats) {
    NormApply s; s.x = tmp; s.F = B; s.C = o.C; s.P = P; s.stats_out = e->sums;
    norm_apply(s, st);
  }
  finalize_stats(e->sums, e->mr, B, P, o.C, 16, 1e-5f, st);
  NormApply n; n.x = tmp; n.F = B; n.C = o.C; n.P = P; n.mr = e->mr; n.w = gw; n.b = gb; n.add = add; n.act = relu ? ACT_RELU : ACT_NONE;
  n.act_last = add != nullptr; n.out_f32 = out.f; n.out_hi = out.hi; n.out_lo = out.lo;
  norm_apply(n, st);
  return o;
}

// ResNetMotionEncoder.__init__ layer plan (motion_encoder.py:161-190)
struct StagePlan { int inplanes, planes, st, sy, sx; };
static std::vector<StagePlan> stage_plan(const ipk_enc_config& c) {
  std::vector<int> ch(c.channels, c.channels + c.n_channels);
  const bool first_down = ((int)ch.size() - 1 < (int)std::ceil(std::log2((double)c.max_frames))) || c.full_seq;
  std::vector<StagePlan> v;
  v.push_back({ch[0], ch[1], first_down ? 2 : 1, 1, 1});
  v.push_back({ch[1], ch[2], 2, 2, 2});
  v.push_back({ch[2], ch[3], 2, 2, 2});
  bool has4 = c.full_seq && c.max_frames >= 16;
  int s4t = 2, s4s = 1;
  if (c.img_size / 8 > c.min_spatial_size) { has4 = true; s4s = 2; }
  if (has4) {
    if (ch.size() < 5) ch.push_back(ch.back());
    v.push_back({ch[3], ch[4], s4t, s4s, s4s});
  }
  if (c.img_size / 16 > c.min_spatial_size) {
    IPK_CHECK(ch.size() >= 6, IPK_ERR_INVALID, "encoder: layer5 needs ENC_M_channels[5]");
    v.push_back({ch[4], ch[5], 2, 2, 2});
  }
  return v;
}

}  // namespace ipk

extern "C" int ipk_enc_create(const ipk_enc_config* cfg, ipk_enc** out) {
  IPK_TRY
  IPK_CHECK(cfg && out, IPK_ERR_INVALID, "ipk_enc_create: null argument");
  IPK_CHECK(cfg->n_channels >= 4 && cfg->n_channels <= IPK_MAX_DEC, IPK_ERR_INVALID, "encoder: ENC_M_channels needs 4..8 entries");
  IPK_CHECK(cfg->z_dim > 0 && cfg->z_dim % 2 == 0 && cfg->max_batch > 0 && cfg->max_frames > 1 && cfg->img_size >= 32, IPK_ERR_INVALID, "encoder: bad sizes");
  for (int i = 0; i < cfg->n_channels; ++i)
    IPK_CHECK(cfg->channels[i] % 16 == 0, IPK_ERR_UNSUPPORTED, "encoder: channel counts must be multiples of 16 (GroupNorm(16))");
  IPK_CHECK(cfg->precision >= 0 && cfg->precision <= 2, IPK_ERR_INVALID, "encoder: bad precision");
  stage_plan(*cfg);
  ipk_enc* e = new ipk_enc();
  e->cfg = *cfg;
  *out = e;
  IPK_CATCH
}

extern "C" int ipk_enc_set_tensor(ipk_enc* e, const char* name, const void* dev_ptr, int64_t numel, int dtype) {
  IPK_TRY
  IPK_CHECK(e && name && dev_ptr, IPK_ERR_INVALID, "ipk_enc_set_tensor: null argument");
  IPK_CHECK(!e->finalized, IPK_ERR_STATE, "ipk_enc_set_tensor after finalize");
  IPK_CHECK(dtype == IPK_F32, IPK_ERR_SHAPE, "encoder: tensor '%s' must be fp32", name);
  e->tensors[name] = TensorRefE{dev_ptr, numel};
  IPK_CATCH
}

extern "C" int ipk_enc_finalize(ipk_enc* e, void* stream) {
  IPK_TRY
  IPK_CHECK(e && !e->finalized, IPK_ERR_STATE, "encoder: null or already finalized");
  cudaStream_t st = (cudaStream_t)stream;
  const ipk_enc_config& c = e->cfg;
  e->stem = build_conv3d(e, "conv1.weight", c.channels[0], 3, 3, 7, 7, 2, 2, 2, 1, 3, 3, st);
  if (c.precision != IPK_PREC_FP32_SIMT && c.channels[0] % 64 == 0) {
    const int S = c.img_size;
    Conv3dShape ss{64, c.channels[0], c.max_frames + 1, S, S / 2, 3, 7, 1, 2, 2, 1, 1, 3, 0, 2 * STEM_CP * 2, (long long)(S + 8) * STEM_CP * 2};
    if (conv3d_tc_supported(ss)) {
      e->stem_tc = conv_alloc(e->pool, c.precision, 21, 64, c.channels[0], false);
      pack_stem_tc_kernel<<<cdiv(21 * c.channels[0] * 21, 256), 256, 0, st>>>((const float*)eneed(e, "conv1.weight", (int64_t)c.channels[0] * 3 * 147).p,
                                                                              e->stem_tc.w_hi, e->stem_tc.w_lo, c.channels[0], e->stem_tc.Npad);
      IPK_LAUNCH_CHECK();
      e->stem_on_tc = true;
    }
  }
  e->stem_gw = ecopy(e, "bn1.weight", c.channels[0], st);
  e->stem_gb = ecopy(e, "bn1.bias", c.channels[0], st);
  auto plan = stage_plan(c);
  // walk the shapes for the workspace while building
  Vol v{c.max_frames + 1, c.img_size, c.img_size, 4};          // the second stage feeds max_frames + 1 frames (full_seq)
  size_t maxe = v.voxels() * 4;
  v = conv3d_out(e->stem, v);
  maxe = std::max(maxe, v.voxels() * v.C);
  for (size_t li = 0; li < plan.size(); ++li) {
    const StagePlan& sp = plan[li];
    int inp = sp.inplanes;
    for (int b = 0; b < 2; ++b) {
      const std::string p = "layer" + std::to_string(li + 1) + "." + std::to_string(b) + ".";
      EncBlock blk;
      const int s_t = b == 0 ? sp.st : 1, s_y = b == 0 ? sp.sy : 1, s_x = b == 0 ? sp.sx : 1;
      blk.c1 = build_conv3d(e, p + "conv1.weight", sp.planes, inp, 3, 3, 3, s_t, s_y, s_x, 1, 1, 1, st);
      blk.g1w = ecopy(e, p + "bn1.weight", sp.planes, st); blk.g1b = ecopy(e, p + "bn1.bias", sp.planes, st);
      blk.c2 = build_conv3d(e, p + "conv2.weight", sp.planes, sp.planes, 3, 3, 3, 1, 1, 1, 1, 1, 1, st);
      blk.g2w = ecopy(e, p + "bn2.weight", sp.planes, st); blk.g2b = ecopy(e, p + "bn2.bias", sp.planes, st);
      blk.has_ds = b == 0 && (s_t != 1 || s_y != 1 || s_x != 1 || inp != sp.planes);
      if (blk.has_ds) {
        blk.ds = build_conv3d(e, p + "downsample.0.weight", sp.planes, inp, 1, 1, 1, s_t, s_y, s_x, 0, 0, 0, st);
        blk.gdw = ecopy(e, p + "downsample.1.weight", sp.planes, st); blk.gdb = ecopy(e, p + "downsample.1.bias", sp.planes, st);
      }
      v = conv3d_out(blk.c1, v);
      maxe = std::max(maxe, v.voxels() * v.C);
      e->blocks.push_back(blk);
      inp = sp.planes;
    }
  }
  IPK_CHECK(v.H == 8 && v.W == 8, IPK_ERR_INVALID, "encoder: final grid is %dx%d, expected 8x8", v.H, v.W);
  e->last_C = v.C;
  // heads: conv_mu | conv_var, 2-D 3x3, pad 1
  const int z = c.z_dim;
  e->heads = conv_alloc(e->pool, IPK_PREC_FP32_SIMT, 9, v.C, 2 * z, true);
  {
    PackSrc s; s.N = z; s.Ksrc = v.C; s.kh = 3; s.kw = 3;
    const std::vector<int> all9 = {0, 1, 2, 3, 4, 5, 6, 7, 8};
    s.w = (const float*)eneed(e, "conv_mu.weight", (int64_t)z * v.C * 9).p;
    conv_pack_into(e->heads, 0, s, all9, st);
    conv_pack_bias(e->heads, 0, (const float*)eneed(e, "conv_mu.bias", z).p, z, 0.f, st);
    s.w = (const float*)eneed(e, "conv_var.weight", (int64_t)z * v.C * 9).p;
    conv_pack_into(e->heads, z, s, all9, st);
    conv_pack_bias(e->heads, z, (const float*)eneed(e, "conv_var.bias", z).p, z, 0.f, st);
  }
  e->act_elems = maxe;
  const size_t B = c.max_batch;
  auto rb = [](size_t b) { return (b + 255) / 256 * 256; };
  const bool planes = c.precision != IPK_PREC_FP32_SIMT;
  const bool lo = c.precision == IPK_PREC_FP32_SPLIT;
  const size_t xp_elems = (size_t)(c.max_frames + 1) * c.img_size * (c.img_size + 8) * STEM_CP;
  e->ws.init((6 + (planes ? 3 : 0)) * rb(B * maxe * 4) + (e->stem_on_tc ? 2 * rb(B * xp_elems * 2) : 0) + rb(B * 64 * 2 * z * 4) + rb(B * 1024 * 2 * 8) + rb(B * 1024 * 2 * 4) + 65536);
  e->X4 = e->ws.alloc<float>(B * maxe);
  e->bufC = e->ws.alloc<float>(B * maxe);
  e->bufD = e->ws.alloc<float>(B * maxe);
  if (e->stem_on_tc) {
    e->Xp_hi = e->ws.alloc<__nv_bfloat16>(B * xp_elems);
    if (lo) e->Xp_lo = e->ws.alloc<__nv_bfloat16>(B * xp_elems);
  }
  for (ActVol* a : {&e->actA, &e->actB, &e->actX}) {
    a->f = e->ws.alloc<float>(B * maxe);
    if (planes) {
      a->hi = e->ws.alloc<__nv_bfloat16>(B * maxe);
      if (lo) a->lo = e->ws.alloc<__nv_bfloat16>(B * maxe);
    }
  }
  e->hbuf = e->ws.alloc<float>(B * 64 * 2 * z);
  e->sums = e->ws.alloc<double>(B * 1024 * 2);
  e->mr = e->ws.alloc<float>(B * 1024 * 2);
  IPK_CUDA(cudaStreamSynchronize(st));
  e->tensors.clear();
  e->finalized = true;
  IPK_CATCH
}

extern "C" int ipk_enc_forward(ipk_enc* e, const float* X, const float* eps, float* z_out, float* mu, float* logvar, int32_t B, int32_t T, void* stream) {
  IPK_TRY
  IPK_CHECK(e && e->finalized, IPK_ERR_STATE, "encoder not finalized");
  IPK_CHECK(X && eps && z_out && mu && logvar, IPK_ERR_INVALID, "ipk_enc_forward: null buffer");
  IPK_CHECK(B > 0 && B <= e->cfg.max_batch, IPK_ERR_INVALID, "encoder: batch %d outside (0, %d]", B, e->cfg.max_batch);
  IPK_CHECK(T > 0 && T <= e->cfg.max_frames + 1, IPK_ERR_INVALID, "encoder: %d frames outside (0, %d]", T, e->cfg.max_frames + 1);
  cudaStream_t st = (cudaStream_t)stream;
  const int S = e->cfg.img_size;
  Vol v{T, S, S, 4};
  ActVol x = e->actA;
  if (e->stem_on_tc) {
    // stem on tcgen05: padded bf16 planes of X, x taps folded into K (see stem_input_planes_kernel), GroupNorm statistics fused
    const int Wp = S + 8;
    {
      ProfScope ps("enc.stem.planes", st);
      const long long tot = (long long)B * T * S * Wp;
      stem_input_planes_kernel<<<(int)std::min<long long>((tot + 255) / 256, 148 * 32), 256, 0, st>>>(X, e->Xp_hi, e->Xp_lo, B, T, S, S, Wp);
      IPK_LAUNCH_CHECK();
    }
    const Conv3dShape ss{64, e->stem.Cout, T, S, S / 2, 3, 7, 1, 2, 2, 1, 1, 3, 0, 2 * STEM_CP * 2, (long long)Wp * STEM_CP * 2};
    const Vol o = conv3d_out(e->stem, v);
    IPK_CUDA(cudaMemsetAsync(e->sums, 0, (size_t)B * o.C * 2 * sizeof(double), st));
    {
      ProfScope ps("enc.conv3d.tc", st);
      conv3d_tc_run(e->stem_tc, ss, e->Xp_hi, e->Xp_lo, B, e->bufC, e->sums, st);
    }
    ProfScope ps("enc.group_norm", st);
    const long long P = (long long)o.voxels();
    finalize_stats(e->sums, e->mr, B, P, o.C, 16, 1e-5f, st);
    NormApply n; n.x = e->bufC; n.F = B; n.C = o.C; n.P = P; n.mr = e->mr; n.w = e->stem_gw; n.b = e->stem_gb; n.act = ACT_RELU;
    n.out_f32 = x.f; n.out_hi = x.hi; n.out_lo = x.lo;
    norm_apply(n, st);
    v = o;
  } else {
    // stem: Conv3d(3 -> C0, (3,7,7), stride 2, pad (1,3,3)) + GroupNorm(16) + ReLU on the FFMA kernel
    const long long V = (long long)v.voxels();
    ncdhw_to_ndhwc4_kernel<<<(int)std::min<long long>(((long long)B * V + 255) / 256, 148 * 32), 256, 0, st>>>(X, e->X4, B, V);
    IPK_LAUNCH_CHECK();
    ActVol in; in.f = e->X4;
    v = conv_gn(e, e->stem, in, v, e->bufC, B, e->stem_gw, e->stem_gb, nullptr, true, x, st);
  }
  // BasicBlocks (motion_encoder.py:56-74): out = relu(gn2(conv2(relu(gn1(conv1(x))))) + residual)
  for (const EncBlock& blk : e->blocks) {
    ActVol y = e->actB;
    {   // the mid-block activation feeds conv2 only: when that conv runs on tcgen05 it reads the bf16 planes, so skip the fp32 copy
      const Vol o1 = conv3d_out(blk.c1, v);
      if (blk.c2.tc && y.hi != nullptr && conv3d_tc_supported(shape_of(blk.c2, o1))) y.f = nullptr;
    }
    const Vol o = conv_gn(e, blk.c1, x, v, e->bufC, B, blk.g1w, blk.g1b, nullptr, true, y, st);
    const float* res = x.f;
    if (blk.has_ds) {
      ActVol r; r.f = e->bufD;
      conv_gn(e, blk.ds, x, v, e->bufC, B, blk.gdw, blk.gdb, nullptr, false, r, st);
      res = r.f;
    }
    const ActVol nx = (x.f == e->actA.f) ? e->actX : e->actA;          // ping-pong the block output
    conv_gn(e, blk.c2, y, o, e->bufC, B, blk.g2w, blk.g2b, res, true, nx, st);
    x = nx;
    v = o;
  }
  IPK_CHECK(v.T == 1 && v.H == 8 && v.W == 8, IPK_ERR_INVALID, "encoder: %d frames leave a %dx%dx%d volume; the temporal extent must collapse to 1 "
            "(motion_encoder.py:241 squeezes it)", T, v.T, v.H, v.W);
  // conv_mu | conv_var (2-D 3x3) and the reparameterisation with host-supplied eps
  {
    ConvIn in; in.p = x.f; in.cstride = v.C; in.F = B; in.H = 8; in.W = 8;
    ConvOut o; o.p = e->hbuf; o.cstride = 2 * e->cfg.z_dim; o.Ho = 8; o.Wo = 8; o.bias = e->heads.bias;
    conv_run(e->heads, in, o, taps_3x3(), 1, st);
    reparam_kernel<<<cdiv(B * e->cfg.z_dim * 64, 256), 256, 0, st>>>(e->hbuf, eps, z_out, mu, logvar, B, e->cfg.z_dim);
    IPK_LAUNCH_CHECK();
  }
  IPK_CATCH
}

extern "C" int ipk_enc_destroy(ipk_enc* e) {
  if (!e) return IPK_OK;
  e->pool.release();
  e->ws.release();
  delete e;
  return IPK_OK;
}


// ================================================================================================ conditioning encoders
// ConvEncoder (models/modules/autoencoders/fully_conv_models.py:28-94), deterministic: the frozen poke embedder and image
// conditioner of make_flow_input (models/second_stage_video.py:268-287).  Stride-2 3x3 convs run on the same implicit-GEMM
// kernel as the video encoder (T = 1); Group/InstanceNorm + ELU + residual through the shared norm pass.
namespace ipk {
struct CBlock {                 // Conv2dBlock: conv (+bias) -> norm -> act
  Conv3dDesc conv;
  float *gw = nullptr, *gb = nullptr;   // GroupNorm affine (null: InstanceNorm, no affine)
};
}  // namespace ipk

struct ipk_cenc {
  ipk_cenc_config cfg;
  std::map<std::string, TensorRefE> tensors;
  bool finalized = false;
  DevPool pool;
  Arena ws;
  CBlock stem;
  struct Res { CBlock c1, c2, rc; bool has_rc; int stride; };
  std::vector<Res> blocks;     // stride-2 ResBlocks, then the bottleneck ResBlock (stride 1)
  float *X4 = nullptr, *bufA = nullptr, *bufB = nullptr, *bufC = nullptr, *bufD = nullptr;
  double* sums = nullptr; float* mr = nullptr;
};

namespace ipk {

static const TensorRefE& cneed(ipk_cenc* e, const std::string& name, int64_t numel) {
  auto it = e->tensors.find(name);
  IPK_CHECK(it != e->tensors.end(), IPK_ERR_MISSING, "cond encoder: missing tensor '%s'", name.c_str());
  IPK_CHECK(it->second.numel == numel, IPK_ERR_SHAPE, "cond encoder: tensor '%s' has %lld elements, expected %lld", name.c_str(),
            (long long)it->second.numel, (long long)numel);
  return it->second;
}
static bool chas(ipk_cenc* e, const std::string& name) { return e->tensors.find(name) != e->tensors.end(); }

static CBlock build_cblock(ipk_cenc* e, const std::string& p, int Cout, int CinSrc, int stride, bool group_norm_affine, cudaStream_t st) {
  CBlock b;
  Conv3dDesc& d = b.conv;
  d.Cin = round_up(CinSrc, 4); d.Cout = Cout;
  d.kt = 1; d.ky = 3; d.kx = 3; d.st = 1; d.sy = stride; d.sx = stride; d.pt = 0; d.py = 1; d.px = 1;
  const int64_t numel = (int64_t)Cout * CinSrc * 9;
  const float* w; const float* sigma = nullptr;
  if (chas(e, p + "conv.weight_orig")) {
    w = (const float*)cneed(e, p + "conv.weight_orig", numel).p;
    float* sg = e->pool.alloc<float>(1);
    spectral_sigma(w, (const float*)cneed(e, p + "conv.weight_u", Cout).p, (const float*)cneed(e, p + "conv.weight_v", (int64_t)CinSrc * 9).p, sg,
                   Cout, CinSrc, 9, false, st);
    sigma = sg;
  } else {
    w = (const float*)cneed(e, p + "conv.weight", numel).p;
  }
  d.w = e->pool.alloc<float>((size_t)9 * d.Cin * Cout);
  const long long total = 9LL * d.Cin * Cout;
  pack_conv3d_kernel<<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, st>>>(w, d.w, Cout, CinSrc, d.Cin, 9, sigma);
  IPK_LAUNCH_CHECK();
  d.bias = e->pool.alloc<float>(Cout);
  IPK_CUDA(cudaMemcpyAsync(d.bias, cneed(e, p + "conv.bias", Cout).p, Cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (group_norm_affine) {
    b.gw = e->pool.alloc<float>(Cout); b.gb = e->pool.alloc<float>(Cout);
    IPK_CUDA(cudaMemcpyAsync(b.gw, cneed(e, p + "norm.weight", Cout).p, Cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
    IPK_CUDA(cudaMemcpyAsync(b.gb, cneed(e, p + "norm.bias", Cout).p, Cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  return b;
}

// Conv2dBlock.forward (util.py:256-273): conv -> GroupNorm(16) | InstanceNorm -> activation [-> + residual]
static Vol run_cblock(ipk_cenc* e, const CBlock& b, const float* in, const Vol& v, float* tmp, float* out, int B, int act, const float* add,
                      cudaStream_t st) {
  Vol o;
  ActVol ain; ain.f = const_cast<float*>(in);
  run_conv3d(b.conv, ain, v, tmp, B, nullptr, &o, st);
  const long long P = (long long)o.voxels();
  IPK_CUDA(cudaMemsetAsync(e->sums, 0, (size_t)B * o.C * 2 * sizeof(double), st));
  NormApply s; s.x = tmp; s.F = B; s.C = o.C; s.P = P; s.stats_out = e->sums;
  norm_apply(s, st);
  finalize_stats(e->sums, e->mr, B, P, o.C, b.gw ? 16 : 0, 1e-5f, st);
  NormApply n; n.x = tmp; n.F = B; n.C = o.C; n.P = P; n.mr = e->mr; n.w = b.gw; n.b = b.gb; n.act = act; n.add = add; n.out_f32 = out;
  norm_apply(n, st);
  return o;
}

static std::vector<int> cenc_widths(const ipk_cenc_config& c) {
  std::vector<int> w{32};
  for (int i = 1; i < c.n_stages; ++i) w.push_back(std::min(w.back() * 2, c.nf_max));
  return w;
}

}  // namespace ipk

extern "C" int ipk_cenc_create(const ipk_cenc_config* cfg, ipk_cenc** out) {
  IPK_TRY
  IPK_CHECK(cfg && out, IPK_ERR_INVALID, "ipk_cenc_create: null argument");
  IPK_CHECK(cfg->nf_in >= 1 && cfg->nf_in <= 8, IPK_ERR_UNSUPPORTED, "cond encoder: nf_in must be 1..8 (got %d)", cfg->nf_in);
  IPK_CHECK(cfg->nf_max % 16 == 0 && cfg->nf_max >= 32, IPK_ERR_UNSUPPORTED, "cond encoder: nf_max must be a multiple of 16 >= 32");
  IPK_CHECK(cfg->n_stages >= 1 && cfg->n_stages <= 8 && cfg->spatial == (cfg->min_spatial_size << cfg->n_stages), IPK_ERR_INVALID,
            "cond encoder: spatial %d != min_spatial_size %d << n_stages %d", cfg->spatial, cfg->min_spatial_size, cfg->n_stages);
  IPK_CHECK(cfg->max_batch > 0, IPK_ERR_INVALID, "cond encoder: max_batch must be positive");
  ipk_cenc* e = new ipk_cenc();
  e->cfg = *cfg;
  *out = e;
  IPK_CATCH
}

extern "C" int ipk_cenc_set_tensor(ipk_cenc* e, const char* name, const void* dev_ptr, int64_t numel, int dtype) {
  IPK_TRY
  IPK_CHECK(e && name && dev_ptr, IPK_ERR_INVALID, "ipk_cenc_set_tensor: null argument");
  IPK_CHECK(!e->finalized, IPK_ERR_STATE, "ipk_cenc_set_tensor after finalize");
  IPK_CHECK(dtype == IPK_F32, IPK_ERR_SHAPE, "cond encoder: tensor '%s' must be fp32", name);
  e->tensors[name] = TensorRefE{dev_ptr, numel};
  IPK_CATCH
}

extern "C" int ipk_cenc_finalize(ipk_cenc* e, void* stream) {
  IPK_TRY
  IPK_CHECK(e && !e->finalized, IPK_ERR_STATE, "cond encoder: null or already finalized");
  cudaStream_t st = (cudaStream_t)stream;
  const ipk_cenc_config& c = e->cfg;
  const std::vector<int> w = cenc_widths(c);
  e->stem = build_cblock(e, "model.0.", w[0], c.nf_in, 2, true, st);
  for (size_t i = 1; i < w.size(); ++i) {
    const std::string p = "model." + std::to_string(i) + ".";
    ipk_cenc::Res r;
    r.c1 = build_cblock(e, p + "conv1.", w[i], w[i - 1], 2, true, st);
    r.c2 = build_cblock(e, p + "conv2.", w[i], w[i], 1, true, st);
    r.rc = build_cblock(e, p + "res_conv.", w[i], w[i - 1], 2, false, st);
    r.has_rc = true; r.stride = 2;
    e->blocks.push_back(r);
  }
  {
    const std::string p = "bottleneck.0.";
    ipk_cenc::Res r;
    r.c1 = build_cblock(e, p + "conv1.", c.nf_max, w.back(), 1, true, st);
    r.c2 = build_cblock(e, p + "conv2.", c.nf_max, c.nf_max, 1, true, st);
    r.has_rc = w.back() != c.nf_max; r.stride = 1;
    if (r.has_rc) r.rc = build_cblock(e, p + "res_conv.", c.nf_max, w.back(), 1, false, st);
    e->blocks.push_back(r);
  }
  const size_t B = c.max_batch;
  const size_t maxe = std::max<size_t>((size_t)c.spatial * c.spatial * round_up(c.nf_in, 4), (size_t)(c.spatial / 2) * (c.spatial / 2) * std::max(w[0], c.nf_max));
  auto rb = [](size_t b) { return (b + 255) / 256 * 256; };
  e->ws.init(5 * rb(B * maxe * 4) + rb(B * 1024 * 2 * 8) + rb(B * 1024 * 2 * 4) + 65536);
  e->X4 = e->ws.alloc<float>(B * maxe);
  e->bufA = e->ws.alloc<float>(B * maxe);
  e->bufB = e->ws.alloc<float>(B * maxe);
  e->bufC = e->ws.alloc<float>(B * maxe);
  e->bufD = e->ws.alloc<float>(B * maxe);
  e->sums = e->ws.alloc<double>(B * 1024 * 2);
  e->mr = e->ws.alloc<float>(B * 1024 * 2);
  IPK_CUDA(cudaStreamSynchronize(st));
  e->tensors.clear();
  e->finalized = true;
  IPK_CATCH
}

extern "C" int ipk_cenc_forward(ipk_cenc* e, const float* x, float* out, float* mean, int32_t B, void* stream) {
  IPK_TRY
  IPK_CHECK(e && e->finalized, IPK_ERR_STATE, "cond encoder not finalized");
  IPK_CHECK(x && out, IPK_ERR_INVALID, "ipk_cenc_forward: null buffer");
  IPK_CHECK(B > 0 && B <= e->cfg.max_batch, IPK_ERR_INVALID, "cond encoder: batch %d outside (0, %d]", B, e->cfg.max_batch);
  cudaStream_t st = (cudaStream_t)stream;
  const int S = e->cfg.spatial;
  {
    const long long P = (long long)S * S;
    if (e->cfg.nf_in <= 4) nchw_to_nhwc4_kernel<<<(int)std::min<long long>(((long long)B * P + 255) / 256, 148 * 32), 256, 0, st>>>(x, e->X4, B, e->cfg.nf_in, P);
    else   // poke map + start frame (embed_poke_and_image, second_stage_video.py:265-266): 5 channels, stored padded to 8
      nchw_to_nhwc_pad_kernel<<<(int)std::min<long long>(((long long)B * P * 8 + 255) / 256, 148 * 32), 256, 0, st>>>(x, e->X4, B, e->cfg.nf_in, 8, P);
    IPK_LAUNCH_CHECK();
  }
  Vol v{1, S, S, round_up(e->cfg.nf_in, 4)};
  float* cur = e->bufA;
  v = run_cblock(e, e->stem, e->X4, v, e->bufB, cur, B, ACT_ELU, nullptr, st);
  for (size_t i = 0; i < e->blocks.size(); ++i) {
    const ipk_cenc::Res& r = e->blocks[i];
    if (i + 1 == e->blocks.size() && mean) nhwc_to_nchw(cur, mean, B, v.C, v.H * v.W, v.C, st);     // `mean` = input of the bottleneck
    // ResBlock.forward (util.py:185-192): out = conv2(conv1(x)) + res_conv(x)
    const float* res = cur;
    if (r.has_rc) { run_cblock(e, r.rc, cur, v, e->bufB, e->bufD, B, ACT_ELU, nullptr, st); res = e->bufD; }
    Vol o = run_cblock(e, r.c1, cur, v, e->bufB, e->bufC, B, ACT_ELU, nullptr, st);
    float* nxt = (cur == e->bufA) ? e->X4 : e->bufA;
    run_cblock(e, r.c2, e->bufC, o, e->bufB, nxt, B, ACT_NONE, res, st);
    cur = nxt;
    v = o;
  }
  IPK_CHECK(v.H == e->cfg.min_spatial_size && v.C == e->cfg.nf_max, IPK_ERR_STATE, "cond encoder: unexpected output volume %dx%dx%d", v.H, v.W, v.C);
  nhwc_to_nchw(cur, out, B, v.C, v.H * v.W, v.C, st);
  IPK_CATCH
}

extern "C" int ipk_cenc_destroy(ipk_cenc* e) {
  if (!e) return IPK_OK;
  e->pool.release();
  e->ws.release();
  delete e;
  return IPK_OK;
}
