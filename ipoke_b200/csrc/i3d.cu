// Native I3D (Inception-v1 inflated to 3-D) forward: the feature extractor of the in-repo Frechet-video-distance used by every
// validation epoch of the second stage.
//   I3D.forward                      utils/metrics.py:1079-1105          (logits = mean over time of conv3d_0c_1x1(avg_pool(...)))
//   Unit3Dpy (TF-SAME Conv3d + BatchNorm3d(eval, eps 1e-3) + ReLU)        utils/metrics.py:857-937, get_padding_shape :814-842
//   MaxPool3dTFPadding (zero pad by T mod stride, MaxPool3d ceil_mode)    utils/metrics.py:940-960
//   Mixed (4 branches, channel concat)                                    utils/metrics.py:963-997
//
// Layout: NDHWC bf16 operand planes (hi [, lo]) between layers, exactly what conv3d_tc.cu consumes.  Every Unit3Dpy is ONE launch of the
// tcgen05 Conv3d engine: TF-SAME padding = low-side offset of the TMA box + out-of-bounds zero fill, BatchNorm folded to a per-channel
// scale / shift at finalize and applied with the ReLU in the epilogue, which writes the next layer's operand planes straight into the
// channel slice of the Inception concat.  The 7x7x7 stride-2 stem folds its 7 x-taps into K (overlapped tensor-map view, as the video
// encoder's stem does).  Max-pools run on the planes (the max of hi + lo pairs is the pair of the winning element).
#include <map>
#include <string>
#include "conv.cuh"
#include "conv3d_tc.cuh"
#include "elementwise.cuh"

namespace ipk {

struct I3dTensor { const void* p; int64_t numel; int dtype; };

struct I3dUnit {
  std::string name;
  int cin = 0, cout = 0, k = 1, stride = 1;
  bool bn = true;
  ConvW w;
  float* scale = nullptr;   // folded BatchNorm (null for the logits layer)
  float* shift = nullptr;   // folded BatchNorm shift, or the conv bias
};

struct Planes { __nv_bfloat16* hi = nullptr; __nv_bfloat16* lo = nullptr; };

constexpr int I3D_STEM_CP = 8;       // channels per pixel of the stem input planes (3 used)

// x: [B][3][T][S][S] fp32 (NCDHW, the reference's input layout) -> stem planes [B][T][S][Wp][8] with `xpad` zero pixels on the left
__global__ void i3d_stem_planes_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int B, int T, int H, int W,
                                       int Wp, int xpad) {
  const long long rows = (long long)B * T * H, total = rows * Wp;
  const long long V = (long long)T * H * W;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int xp = (int)(e % Wp);
    const long long row = e / Wp;
    const long long b = row / ((long long)T * H), ty = row % ((long long)T * H);
    const int x = xp - xpad;
    float v[3] = {0.f, 0.f, 0.f};
    if (x >= 0 && x < W) {
      const float* ip = in + (size_t)b * 3 * V + (size_t)ty * W + x;
      v[0] = ip[0]; v[1] = ip[V]; v[2] = ip[2 * V];
    }
    __nv_bfloat16 hb[3], lb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { hb[c] = __float2bfloat16_rn(v[c]); lb[c] = __float2bfloat16_rn(v[c] - __bfloat162float(hb[c])); }
    const uint32_t h0 = (uint32_t)__bfloat16_as_ushort(hb[0]) | ((uint32_t)__bfloat16_as_ushort(hb[1]) << 16), h1 = (uint32_t)__bfloat16_as_ushort(hb[2]);
    const uint32_t l0 = (uint32_t)__bfloat16_as_ushort(lb[0]) | ((uint32_t)__bfloat16_as_ushort(lb[1]) << 16), l1 = (uint32_t)__bfloat16_as_ushort(lb[2]);
    ((uint4*)hi)[e] = make_uint4(h0, h1, 0u, 0u);
    if (lo) ((uint4*)lo)[e] = make_uint4(l0, l1, 0u, 0u);
  }
}
// w: OIDHW [Cout][3][7][7][7] -> planes [tap = dt*7 + dy][Npad][64] with k = dx*8 + c
__global__ void i3d_pack_stem_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int Cout, int Npad) {
  const int total = 49 * Cout * 21;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int j = e % 21, n = (e / 21) % Cout, tap = e / (21 * Cout);
    const int dx = j / 3, c = j % 3, dt = tap / 7, dy = tap % 7;
    const float v = w[((((size_t)n * 3 + c) * 7 + dt) * 7 + dy) * 7 + dx];
    const size_t di = ((size_t)tap * Npad + n) * 64 + dx * I3D_STEM_CP + c;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[di] = h;
    if (lo) lo[di] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}
// BatchNorm3d(eval): y = (x - mean) / sqrt(var + eps) * gamma + beta  ->  y = x * scale + shift
__global__ void i3d_fold_bn_kernel(const float* __restrict__ g, const float* __restrict__ b, const float* __restrict__ m, const float* __restrict__ v, float eps,
                                   float* __restrict__ scale, float* __restrict__ shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const float s = g[c] / sqrtf(v[c] + eps);
    scale[c] = s;
    shift[c] = b[c] - m[c] * s;
  }
}

// MaxPool3dTFPadding on operand planes: explicit ZERO padding (lo sides pt / py / px; the zeros take part in the max) followed by
// MaxPool3d(ceil_mode=True), whose windows ignore whatever lies beyond the padded extent (Tp, Hp, Wp).
struct PoolArgs {
  int B, C, Ti, Hi, Wi, To, Ho, Wo;
  int kt, ky, kx, st, sy, sx, pt, py, px, Tp, Hp, Wp;
};
__global__ void i3d_maxpool_kernel(const __nv_bfloat16* __restrict__ ihi, const __nv_bfloat16* __restrict__ ilo, __nv_bfloat16* __restrict__ ohi,
                                   __nv_bfloat16* __restrict__ olo, const PoolArgs a) {
  const int C2 = a.C >> 1;                         // two channels per thread (bf16x2)
  const long long total = (long long)a.B * a.To * a.Ho * a.Wo * C2;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c2 = (int)(e % C2);
    long long r = e / C2;
    const int xo = (int)(r % a.Wo); r /= a.Wo;
    const int yo = (int)(r % a.Ho); r /= a.Ho;
    const int to = (int)(r % a.To);
    const int b = (int)(r / a.To);
    float m0 = -INFINITY, m1 = -INFINITY;
    for (int dt = 0; dt < a.kt; ++dt) {
      const int tp = to * a.st + dt;               // coordinate in the zero-padded volume
      if (tp >= a.Tp) break;
      const int t = tp - a.pt;
      for (int dy = 0; dy < a.ky; ++dy) {
        const int yp = yo * a.sy + dy;
        if (yp >= a.Hp) break;
        const int y = yp - a.py;
        for (int dx = 0; dx < a.kx; ++dx) {
          const int xp = xo * a.sx + dx;
          if (xp >= a.Wp) break;
          const int x = xp - a.px;
          float v0 = 0.f, v1 = 0.f;                // padding value
          if (t >= 0 && t < a.Ti && y >= 0 && y < a.Hi && x >= 0 && x < a.Wi) {
            const size_t i = ((((size_t)b * a.Ti + t) * a.Hi + y) * a.Wi + x) * C2 + c2;
            const float2 h = __bfloat1622float2(((const __nv_bfloat162*)ihi)[i]);
            v0 = h.x; v1 = h.y;
            if (ilo) { const float2 l = __bfloat1622float2(((const __nv_bfloat162*)ilo)[i]); v0 += l.x; v1 += l.y; }
          }
          m0 = fmaxf(m0, v0); m1 = fmaxf(m1, v1);
        }
      }
    }
    const __nv_bfloat162 h = __floats2bfloat162_rn(m0, m1);
    ((__nv_bfloat162*)ohi)[e] = h;
    if (olo) {
      const float2 hf = __bfloat1622float2(h);
      ((__nv_bfloat162*)olo)[e] = __floats2bfloat162_rn(m0 - hf.x, m1 - hf.y);
    }
  }
}

// avg_pool (2,7,7)/(1,1,1) -> conv3d_0c_1x1 (bias, no BN, no ReLU) -> mean over the remaining time steps (utils/metrics.py:1096-1102).
// One block per sample; x planes [B][T][7][7][C]; wl [N][C] fp32; out [B][N].
__global__ void __launch_bounds__(256) i3d_head_kernel(const __nv_bfloat16* __restrict__ xhi, const __nv_bfloat16* __restrict__ xlo, int T, int HW, int C,
                                                       const float* __restrict__ wl, const float* __restrict__ bias, int N, float* __restrict__ out) {
  extern __shared__ float avg[];                   // [T - 1][C]
  const int b = blockIdx.x, To = T - 1;
  const float inv = 1.0f / (float)(2 * HW);
  for (int i = threadIdx.x; i < To * C; i += blockDim.x) {
    const int t = i / C, c = i - t * C;
    float s = 0.f;
    for (int dt = 0; dt < 2; ++dt)
      for (int p = 0; p < HW; ++p) {
        const size_t j = (((size_t)b * T + t + dt) * HW + p) * C + c;
        float v = __bfloat162float(xhi[j]);
        if (xlo) v += __bfloat162float(xlo[j]);
        s += v;
      }
    avg[i] = s * inv;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int n = warp; n < N; n += nw) {
    float acc = 0.f;
    for (int t = 0; t < To; ++t) {
      float d = 0.f;
      for (int c = lane; c < C; c += 32) d = fmaf(wl[(size_t)n * C + c], avg[t * C + c], d);
      acc += d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[(size_t)b * N + n] = acc / (float)To + bias[n];
  }
}

// ---- preprocess (utils/metrics.py:786-802): bilinear resize (align_corners) to 224 x 224, then [-1, 1] -> [0, 1] when the SET has a negative value
__global__ void i3d_resize_kernel(const float* __restrict__ in, float* __restrict__ out, long long planes, int S, int R, int* __restrict__ min_bits) {
  const long long total = planes * R * R;
  const float sc = R > 1 ? (float)(S - 1) / (float)(R - 1) : 0.f;
  float lmin = INFINITY;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(e % R), y = (int)((e / R) % R);
    const long long pl = e / ((long long)R * R);
    const float fy = y * sc, fx = x * sc;
    const int y0 = min((int)fy, S - 1), x0 = min((int)fx, S - 1);
    const int y1 = min(y0 + 1, S - 1), x1 = min(x0 + 1, S - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float* p = in + pl * S * S;
    const float v = (1.f - ly) * ((1.f - lx) * p[y0 * S + x0] + lx * p[y0 * S + x1]) + ly * ((1.f - lx) * p[y1 * S + x0] + lx * p[y1 * S + x1]);
    out[e] = v;
    lmin = fminf(lmin, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
  if ((threadIdx.x & 31) == 0 && lmin < 0.f) atomicOr(min_bits, 1);       // all the caller needs: "is any value negative"
}
__global__ void i3d_denorm_kernel(float* __restrict__ x, long long n, const int* __restrict__ neg_flag) {
  if (*neg_flag == 0) return;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) x[e] = (x[e] + 1.0f) * 0.5f;
}

}  // namespace ipk

using namespace ipk;

struct ipk_i3d {
  ipk_i3d_config cfg;
  std::map<std::string, I3dTensor> tensors;
  bool finalized = false;
  DevPool pool;
  Arena ws;
  int eng = IPK_PREC_FP32_SPLIT;
  std::map<std::string, I3dUnit> units;
  Planes xp, buf[4];
  size_t buf_elems = 0, xp_elems = 0;
  float* head_w = nullptr;   // [N][1024]
  float* head_b = nullptr;
  float* logits = nullptr;
};

namespace ipk {

static const I3dTensor& ineed(ipk_i3d* m, const std::string& name, int64_t numel) {
  auto it = m->tensors.find(name);
  IPK_CHECK(it != m->tensors.end(), IPK_ERR_MISSING, "i3d: missing tensor '%s'", name.c_str());
  IPK_CHECK(it->second.numel == numel && it->second.dtype == IPK_F32, IPK_ERR_SHAPE, "i3d: tensor '%s' has %lld elements / dtype %d, expected %lld fp32", name.c_str(),
            (long long)it->second.numel, it->second.dtype, (long long)numel);
  return it->second;
}

struct MixedSpec { const char* name; int cin; int o[6]; };
static const MixedSpec MIXED[9] = {        // utils/metrics.py:1047-1066
    {"mixed_3b", 192, {64, 96, 128, 16, 32, 32}},   {"mixed_3c", 256, {128, 128, 192, 32, 96, 64}},  {"mixed_4b", 480, {192, 96, 208, 16, 48, 64}},
    {"mixed_4c", 512, {160, 112, 224, 24, 64, 64}}, {"mixed_4d", 512, {128, 128, 256, 24, 64, 64}},  {"mixed_4e", 512, {112, 144, 288, 32, 64, 64}},
    {"mixed_4f", 528, {256, 160, 320, 32, 128, 128}}, {"mixed_5b", 832, {256, 160, 320, 32, 128, 128}}, {"mixed_5c", 832, {384, 192, 384, 48, 128, 128}}};

static void build_unit(ipk_i3d* m, const std::string& name, int cin, int cout, int k, int stride, cudaStream_t st) {
  I3dUnit u;
  u.name = name; u.cin = cin; u.cout = cout; u.k = k; u.stride = stride;
  const int k3 = k * k * k;
  const I3dTensor& w = ineed(m, name + ".conv3d.weight", (int64_t)cout * cin * k3);
  if (k == 7) {        // stem: (dt, dy) taps of K = 64 = 7 x-taps x 8 channel slots
    IPK_CHECK(cin == 3 && stride == 2, IPK_ERR_UNSUPPORTED, "i3d: only the rgb 7x7x7 stride-2 stem is supported");
    u.w = conv_alloc(m->pool, m->eng, 49, 64, cout, false);
    i3d_pack_stem_kernel<<<cdiv(49 * cout * 21, 256), 256, 0, st>>>((const float*)w.p, u.w.w_hi, u.w.w_lo, cout, u.w.Npad);
    IPK_LAUNCH_CHECK();
  } else {
    u.w = conv_alloc(m->pool, m->eng, k3, cin, cout, false);
    conv3d_tc_pack(u.w, (const float*)w.p, cout, cin, st);
  }
  u.scale = m->pool.alloc<float>(cout);
  u.shift = m->pool.alloc<float>(cout);
  i3d_fold_bn_kernel<<<cdiv(cout, 128), 128, 0, st>>>((const float*)ineed(m, name + ".batch3d.weight", cout).p, (const float*)ineed(m, name + ".batch3d.bias", cout).p,
                                                      (const float*)ineed(m, name + ".batch3d.running_mean", cout).p,
                                                      (const float*)ineed(m, name + ".batch3d.running_var", cout).p, 1e-3f, u.scale, u.shift, cout);
  IPK_LAUNCH_CHECK();
  m->units[name] = u;
}

// get_padding_shape (utils/metrics.py:814-842) for one dimension: (low, high) zero padding
static void tf_pad(int k, int s, int mod, int& lo, int& hi) {
  const int along = mod ? std::max(k - mod, 0) : std::max(k - s, 0);
  lo = along / 2;
  hi = along - lo;
}

struct Vol5 { int T, H, W, C; size_t voxels(int B) const { return (size_t)B * T * H * W; } };

// Unit3Dpy.forward on planes: in [B][T][H][W][cin] -> out channel slice [coff, coff + cout) of [B][To][Ho][Wo][cstride]
static Vol5 run_unit(ipk_i3d* m, const I3dUnit& u, const Planes& in, Vol5 v, const Planes& out, int cstride, int coff, int B, cudaStream_t st) {
  IPK_CHECK(v.C == u.cin, IPK_ERR_STATE, "i3d: %s expects %d input channels, got %d", u.name.c_str(), u.cin, v.C);
  int lo_t, hi_t, lo_s, hi_s;
  tf_pad(u.k, u.stride, u.stride > 1 ? v.T % u.stride : 0, lo_t, hi_t);      // runtime depth padding by T mod stride (:927-931)
  tf_pad(u.k, u.stride, 0, lo_s, hi_s);
  Conv3dShape s{u.cin, u.cout, v.T, v.H, v.W, u.k, u.k, u.k, u.stride, u.stride, u.stride, lo_t, lo_s, lo_s};
  s.To = (v.T + lo_t + hi_t - u.k) / u.stride + 1;
  s.Ho = (v.H + lo_s + hi_s - u.k) / u.stride + 1;
  s.Wo = (v.W + lo_s + hi_s - u.k) / u.stride + 1;
  IPK_CHECK((size_t)B * s.To * s.Ho * s.Wo * cstride <= m->buf_elems, IPK_ERR_STATE, "i3d: activation buffer too small for %s", u.name.c_str());
  Conv3dEpi e;
  e.out_hi = out.hi; e.out_lo = out.lo; e.cstride = cstride; e.coff = coff; e.scale = u.scale; e.shift = u.shift; e.relu = 1;
  conv3d_tc_run_ex(u.w, s, in.hi, in.lo, B, e, st);
  return Vol5{s.To, s.Ho, s.Wo, cstride};
}

static Vol5 run_pool(ipk_i3d* m, const Planes& in, Vol5 v, const Planes& out, int kt, int ks, int st_t, int st_s, int B, cudaStream_t st) {
  PoolArgs a;
  a.B = B; a.C = v.C; a.Ti = v.T; a.Hi = v.H; a.Wi = v.W;
  a.kt = kt; a.ky = ks; a.kx = ks; a.st = st_t; a.sy = st_s; a.sx = st_s;
  int hi;
  tf_pad(kt, st_t, st_t > 1 ? v.T % st_t : 0, a.pt, hi); a.Tp = v.T + a.pt + hi;
  tf_pad(ks, st_s, 0, a.py, hi); a.Hp = v.H + a.py + hi;
  a.px = a.py; a.Wp = v.W + a.px + hi;
  auto outlen = [](int Lp, int k, int s) {           // MaxPool3d(ceil_mode=True) on the padded length
    int o = (Lp - k + s - 1) / s + 1;
    if ((o - 1) * s >= Lp) --o;
    return o;
  };
  a.To = outlen(a.Tp, kt, st_t); a.Ho = outlen(a.Hp, ks, st_s); a.Wo = outlen(a.Wp, ks, st_s);
  const long long total = (long long)B * a.To * a.Ho * a.Wo * (v.C / 2);
  IPK_CHECK((size_t)total * 2 <= m->buf_elems, IPK_ERR_STATE, "i3d: activation buffer too small for a pooling output");
  i3d_maxpool_kernel<<<(int)std::min<long long>((total + 255) / 256, 148LL * 32), 256, 0, st>>>(in.hi, in.lo, out.hi, out.lo, a);
  IPK_LAUNCH_CHECK();
  return Vol5{a.To, a.Ho, a.Wo, v.C};
}

// Mixed.forward: x (buffer bi) -> concat (buffer bo); buffers t1, t2 hold the 1x1 bottlenecks / the pooled input
static Vol5 run_mixed(ipk_i3d* m, const MixedSpec& ms, int bi, int bo, int t1, int t2, Vol5 v, int B, cudaStream_t st) {
  const std::string p = ms.name;
  const int ctot = ms.o[0] + ms.o[2] + ms.o[4] + ms.o[5];
  ProfScope ps(("i3d." + p).c_str(), st);
  Vol5 o = run_unit(m, m->units[p + ".branch_0"], m->buf[bi], v, m->buf[bo], ctot, 0, B, st);
  Vol5 b1 = run_unit(m, m->units[p + ".branch_1.0"], m->buf[bi], v, m->buf[t1], ms.o[1], 0, B, st);
  run_unit(m, m->units[p + ".branch_1.1"], m->buf[t1], b1, m->buf[bo], ctot, ms.o[0], B, st);
  Vol5 b2 = run_unit(m, m->units[p + ".branch_2.0"], m->buf[bi], v, m->buf[t1], ms.o[3], 0, B, st);
  run_unit(m, m->units[p + ".branch_2.1"], m->buf[t1], b2, m->buf[bo], ctot, ms.o[0] + ms.o[2], B, st);
  Vol5 pv = run_pool(m, m->buf[bi], v, m->buf[t2], 3, 3, 1, 1, B, st);
  run_unit(m, m->units[p + ".branch_3.1"], m->buf[t2], pv, m->buf[bo], ctot, ms.o[0] + ms.o[2] + ms.o[4], B, st);
  return o;
}

}  // namespace ipk

// ----------------------------------------------------------------------------------------------- C ABI
extern "C" int ipk_i3d_create(const ipk_i3d_config* cfg, ipk_i3d** out) {
  IPK_TRY
  IPK_CHECK(cfg && out, IPK_ERR_INVALID, "ipk_i3d_create: null argument");
  IPK_CHECK(cfg->max_batch > 0 && cfg->max_frames >= 9, IPK_ERR_INVALID, "i3d: max_batch must be positive and max_frames >= 9 (the (2,7,7) average pool needs 2 time steps after three halvings)");
  IPK_CHECK(cfg->num_classes > 0 && cfg->num_classes % 8 == 0, IPK_ERR_UNSUPPORTED, "i3d: num_classes must be a positive multiple of 8");
  IPK_CHECK(cfg->precision == IPK_PREC_FP32_SPLIT || cfg->precision == IPK_PREC_BF16, IPK_ERR_UNSUPPORTED, "i3d: precision must be fp32 (bf16x3) or bf16");
  ipk_i3d* m = new ipk_i3d();
  m->cfg = *cfg;
  m->eng = cfg->precision;
  *out = m;
  IPK_CATCH
}

extern "C" int ipk_i3d_set_tensor(ipk_i3d* m, const char* name, const void* dev_ptr, int64_t numel, int dtype) {
  IPK_TRY
  IPK_CHECK(m && name && dev_ptr, IPK_ERR_INVALID, "ipk_i3d_set_tensor: null argument");
  IPK_CHECK(!m->finalized, IPK_ERR_STATE, "ipk_i3d_set_tensor after finalize");
  m->tensors[name] = I3dTensor{dev_ptr, numel, dtype};
  IPK_CATCH
}

extern "C" int ipk_i3d_finalize(ipk_i3d* m, void* stream) {
  IPK_TRY
  IPK_CHECK(m && !m->finalized, IPK_ERR_STATE, "i3d: null or already finalized");
  cudaStream_t st = (cudaStream_t)stream;
  build_unit(m, "conv3d_1a_7x7", 3, 64, 7, 2, st);
  build_unit(m, "conv3d_2b_1x1", 64, 64, 1, 1, st);
  build_unit(m, "conv3d_2c_3x3", 64, 192, 3, 1, st);
  for (const MixedSpec& ms : MIXED) {
    const std::string p = ms.name;
    build_unit(m, p + ".branch_0", ms.cin, ms.o[0], 1, 1, st);
    build_unit(m, p + ".branch_1.0", ms.cin, ms.o[1], 1, 1, st);
    build_unit(m, p + ".branch_1.1", ms.o[1], ms.o[2], 3, 1, st);
    build_unit(m, p + ".branch_2.0", ms.cin, ms.o[3], 1, 1, st);
    build_unit(m, p + ".branch_2.1", ms.o[3], ms.o[4], 3, 1, st);
    build_unit(m, p + ".branch_3.1", ms.cin, ms.o[5], 1, 1, st);
  }
  const int N = m->cfg.num_classes;
  m->head_w = m->pool.alloc<float>((size_t)N * 1024);
  m->head_b = m->pool.alloc<float>(N);
  IPK_CUDA(cudaMemcpyAsync(m->head_w, ineed(m, "conv3d_0c_1x1.conv3d.weight", (int64_t)N * 1024).p, (size_t)N * 1024 * 4, cudaMemcpyDeviceToDevice, st));
  IPK_CUDA(cudaMemcpyAsync(m->head_b, ineed(m, "conv3d_0c_1x1.conv3d.bias", N).p, (size_t)N * 4, cudaMemcpyDeviceToDevice, st));
  // workspace: the largest activation is the stem output [B][T/2][112][112][64]; the 3c concat ([B][T/2][28][28][480]) is smaller
  const size_t B = m->cfg.max_batch, T = m->cfg.max_frames, T1 = (T + 1) / 2;
  m->buf_elems = B * T1 * 112 * 112 * 64;
  m->buf_elems = std::max(m->buf_elems, B * T1 * 56 * 56 * 192);
  m->xp_elems = B * T * 224 * 232 * I3D_STEM_CP;
  const bool split = m->eng == IPK_PREC_FP32_SPLIT;
  auto rb = [](size_t b) { return (b + 255) / 256 * 256; };
  m->ws.init((split ? 2 : 1) * (4 * rb(m->buf_elems * 2) + rb(m->xp_elems * 2)) + rb(B * N * 4) + 65536);
  auto planes = [&](size_t elems) {
    Planes p;
    p.hi = m->ws.alloc<__nv_bfloat16>(elems);
    p.lo = split ? m->ws.alloc<__nv_bfloat16>(elems) : nullptr;
    return p;
  };
  m->xp = planes(m->xp_elems);
  for (int i = 0; i < 4; ++i) m->buf[i] = planes(m->buf_elems);
  m->logits = m->ws.alloc<float>(B * N);
  IPK_CUDA(cudaStreamSynchronize(st));
  m->tensors.clear();
  m->finalized = true;
  IPK_CATCH
}

// x: [B][3][T][224][224] fp32 in [0, 1] (what the reference feeds: model(batch.permute(0, 2, 1, 3, 4)), utils/metrics.py:726) -> logits [B][num_classes]
extern "C" int ipk_i3d_forward(ipk_i3d* m, const float* x, float* logits, int32_t B, int32_t T, void* stream) {
  IPK_TRY
  IPK_CHECK(m && m->finalized, IPK_ERR_STATE, "i3d: not finalized");
  IPK_CHECK(x && logits, IPK_ERR_INVALID, "ipk_i3d_forward: null buffer");
  IPK_CHECK(B > 0 && B <= m->cfg.max_batch && T >= 9 && T <= m->cfg.max_frames, IPK_ERR_INVALID, "i3d: batch %d / frames %d outside (0, %d] / [9, %d]", B, T,
            m->cfg.max_batch, m->cfg.max_frames);
  cudaStream_t st = (cudaStream_t)stream;
  const int S = 224, Wp = 232;
  // ---- stem: conv3d_1a_7x7, stride 2, TF-SAME: low-side padding 2 (H, W) and 2 or 3 (T even / odd)
  int lo_t, hi_t, lo_s, hi_s;
  tf_pad(7, 2, T % 2, lo_t, hi_t);
  tf_pad(7, 2, 0, lo_s, hi_s);
  Vol5 v;
  {
    ProfScope ps("i3d.stem", st);
    const long long tot = (long long)B * T * S * Wp;
    i3d_stem_planes_kernel<<<(int)std::min<long long>((tot + 255) / 256, 148LL * 32), 256, 0, st>>>(x, m->xp.hi, m->xp.lo, B, T, S, S, Wp, lo_s);
    IPK_LAUNCH_CHECK();
    const I3dUnit& u = m->units["conv3d_1a_7x7"];
    // view: one "column" per output x (stride 2 pixels = 32 bytes), its 8 pixels x 8 channels = the K block of the 7 x-taps
    Conv3dShape ss{64, 64, T, S, S / 2, 7, 7, 1, 2, 2, 1, lo_t, lo_s, 0, 2 * I3D_STEM_CP * 2, (long long)Wp * I3D_STEM_CP * 2};
    ss.To = (T + lo_t + hi_t - 7) / 2 + 1; ss.Ho = (S + lo_s + hi_s - 7) / 2 + 1; ss.Wo = S / 2;
    Conv3dEpi e;
    e.out_hi = m->buf[0].hi; e.out_lo = m->buf[0].lo; e.cstride = 64; e.coff = 0; e.scale = u.scale; e.shift = u.shift; e.relu = 1;
    conv3d_tc_run_ex(u.w, ss, m->xp.hi, m->xp.lo, B, e, st);
    v = Vol5{ss.To, ss.Ho, ss.Wo, 64};
  }
  {
    ProfScope ps("i3d.block2", st);
    v = run_pool(m, m->buf[0], v, m->buf[1], 1, 3, 1, 2, B, st);                                   // maxPool3d_2a_3x3
    v = run_unit(m, m->units["conv3d_2b_1x1"], m->buf[1], v, m->buf[0], 64, 0, B, st);
    v = run_unit(m, m->units["conv3d_2c_3x3"], m->buf[0], v, m->buf[1], 192, 0, B, st);
    v = run_pool(m, m->buf[1], v, m->buf[0], 1, 3, 1, 2, B, st);                                   // maxPool3d_3a_3x3
  }
  int cur = 0;                                     // buffer holding the current activation; 2 and 3 are scratch
  auto mixed = [&](int idx) {
    const int nxt = cur ^ 1;
    v = run_mixed(m, MIXED[idx], cur, nxt, 2, 3, v, B, st);
    cur = nxt;
  };
  mixed(0); mixed(1);
  { v = run_pool(m, m->buf[cur], v, m->buf[cur ^ 1], 3, 3, 2, 2, B, st); cur ^= 1; }              // maxPool3d_4a_3x3
  for (int i = 2; i < 7; ++i) mixed(i);
  { v = run_pool(m, m->buf[cur], v, m->buf[cur ^ 1], 2, 2, 2, 2, B, st); cur ^= 1; }              // maxPool3d_5a_2x2
  mixed(7); mixed(8);
  IPK_CHECK(v.C == 1024 && v.H == 7 && v.W == 7 && v.T >= 2, IPK_ERR_STATE, "i3d: unexpected final volume %d x %d x %d x %d", v.T, v.H, v.W, v.C);
  {
    ProfScope ps("i3d.head", st);
    i3d_head_kernel<<<B, 256, (size_t)(v.T - 1) * 1024 * sizeof(float), st>>>(m->buf[cur].hi, m->buf[cur].lo, v.T, 49, 1024, m->head_w, m->head_b, m->cfg.num_classes,
                                                                               m->logits);
    IPK_LAUNCH_CHECK();
  }
  IPK_CUDA(cudaMemcpyAsync(logits, m->logits, (size_t)B * m->cfg.num_classes * sizeof(float), cudaMemcpyDeviceToDevice, st));
  IPK_CATCH
}

// preprocess (utils/metrics.py:786-802) of ONE set: videos [n_frames_total = N*T][3][S][S] fp32 -> out [N*T][3][224][224]; the [-1,1] -> [0,1]
// map is applied when any resized value of the set is negative (decided on the device, no host round trip)
extern "C" int ipk_i3d_preprocess(const float* videos, float* out, int64_t n_frames, int32_t S, void* stream) {
  IPK_TRY
  IPK_CHECK(videos && out && n_frames > 0 && S > 1, IPK_ERR_INVALID, "ipk_i3d_preprocess: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int* flag = nullptr;
  IPK_CUDA(cudaMallocAsync((void**)&flag, sizeof(int), st));
  IPK_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
  const long long planes = (long long)n_frames * 3, total = planes * 224 * 224;
  const int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
  i3d_resize_kernel<<<grid, 256, 0, st>>>(videos, out, planes, S, 224, flag);
  IPK_LAUNCH_CHECK();
  i3d_denorm_kernel<<<grid, 256, 0, st>>>(out, total, flag);
  IPK_LAUNCH_CHECK();
  IPK_CUDA(cudaFreeAsync(flag, st));
  IPK_CATCH
}

extern "C" int ipk_i3d_destroy(ipk_i3d* m) {
  if (!m) return IPK_OK;
  m->pool.release();
  m->ws.release();
  delete m;
  return IPK_OK;
}
