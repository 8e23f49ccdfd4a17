// First-stage decoder plan: latent ConvGRU + SPADE-conditioned decoder.
//   PokeMotionModel.decode_first_stage      models/second_stage_video.py:361-382
//   ConvGRU / ConvGRUCell                   models/modules/motion_models/rnn.py:4-133
//   SpadeCondConvDecoder.forward            models/modules/autoencoders/fully_conv_models.py:135-177
//   ResBlock / Conv2dBlock / Conv2dTransposeBlock / Spade   models/modules/autoencoders/util.py:7-73,106-273,473-500
//
// Data layout: NDHWC -- every frame of every video of a chunk is one "image" of an NHWC tensor [F = videos*T][H][W][C]
// (kT = 1: the decoder has no temporal kernel), so all T frames go through each conv together.
// The x0-only SPADE gamma/beta maps are computed once per video and broadcast over its T frames.
#include <map>
#include <string>
#include "conv.cuh"
#include "elementwise.cuh"

namespace ipk {
struct FsTensor { const void* p; int64_t numel; int dtype; };

struct UpBlock {
  int Cin, Cout, s_in;       // input grid s_in x s_in -> output 2 s_in
  ConvW ct1, ctr, c2;        // conv1 (ConvT), res_conv (ConvT), conv2 (3x3)
  ConvW ctf;                 // tensor-core engines: conv1 | res_conv fused along N (one read of the input per tap)
  ConvW sp1, spgb;           // SPADE: 3->128, 128->(1+gamma | beta)
  float* SP = nullptr;       // [max_batch][(2 s_in)^2][2 Cout]
  int groups = 16;
};
}  // namespace ipk

void ipk_graphs_drop(const void* handle);   // capi.cu
using namespace ipk;

struct ipk_fs {
  ipk_fs_config cfg;
  std::map<std::string, FsTensor> tensors;
  bool finalized = false;
  DevPool pool;
  Arena ws;
  int z = 0, S = 0, L = 0, nd = 0, eng = 0, act_mode = OUT_F32_NHWC;
  int chunk_videos = 1;
  // GRU
  std::vector<ConvW> gru_ur, gru_o;
  float* motion_bias = nullptr;  // [z][64]
  // decoder
  ConvW in_c1, in_c2, in_res;
  bool has_in_res = false;
  float *gn1_w = nullptr, *gn1_b = nullptr, *gn2_w = nullptr, *gn2_b = nullptr;
  std::vector<UpBlock> blocks;
  ConvW out_conv;
  cudaEvent_t chunk_done = nullptr;
  OutConvPlan* out_direct = nullptr;   // fp32 halo-tile kernel when the last decoder width is 64 (out_conv.cu)
  // workspace
  std::vector<char*> XH, XRH;    // [L] conv-input operands (x | h), (x | r*h): [Mmax][2z] in the engine's storage mode
  std::vector<float*> Hf;        // [L] fp32 hidden state [Mmax][z]
  float* xtmp = nullptr;         // [Mmax][z] fp32 staging of the layer-0 input
  void* spI = nullptr; void* spI_lo = nullptr;   // SPADE 3->128 conv input: 3x3 im2col rows [B*S*S][32]
  float* U = nullptr; float* graw = nullptr; float* Hseq = nullptr;
  float* x0r = nullptr; void* spY = nullptr; void* spY_lo = nullptr;
  void* bufA = nullptr; void* bufA_lo = nullptr; void* bufY1 = nullptr; void* bufY1_lo = nullptr;
  float *bufR = nullptr, *bufY2 = nullptr, *bufS = nullptr;
  double* sums = nullptr; float* mr = nullptr;
  float* hin = nullptr; float* x0_dev_unused = nullptr;
  size_t frame_elems = 0;  // max elements of one activation tensor per frame
  int Fmax = 0;
};

namespace ipk {

static const FsTensor& fneed(ipk_fs* d, const std::string& name, int64_t numel) {
  auto it = d->tensors.find(name);
  IPK_CHECK(it != d->tensors.end(), IPK_ERR_MISSING, "first stage: missing tensor '%s'", name.c_str());
  IPK_CHECK(it->second.numel == numel, IPK_ERR_SHAPE, "first stage: tensor '%s' has %lld elements, expected %lld", name.c_str(),
            (long long)it->second.numel, (long long)numel);
  IPK_CHECK(it->second.dtype == IPK_F32, IPK_ERR_SHAPE, "first stage: tensor '%s' must be fp32", name.c_str());
  return it->second;
}
static bool fhas(ipk_fs* d, const std::string& name) { return d->tensors.find(name) != d->tensors.end(); }

static float* dev_copy(ipk_fs* d, const void* src, size_t n, cudaStream_t st) {
  float* p = d->pool.alloc<float>(n);
  IPK_CUDA(cudaMemcpyAsync(p, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return p;
}

static const std::vector<int> ALL9 = {0, 1, 2, 3, 4, 5, 6, 7, 8};

// pack a 3x3 conv (or ConvTranspose) whose weight may carry legacy spectral norm (weight_orig / weight_u / weight_v) into
// output columns [n_off, n_off + Cout) of `w`
static void pack_conv3_into(ipk_fs* d, ConvW& w, int n_off, const std::string& p, int Cout, int Cin, bool transposed, float bias_add, cudaStream_t st) {
  PackSrc s;
  s.N = Cout; s.Ksrc = Cin; s.kh = 3; s.kw = 3; s.transposed = transposed;
  const int64_t numel = (int64_t)Cout * Cin * 9;
  if (fhas(d, p + "weight_orig")) {
    const FsTensor& wo = fneed(d, p + "weight_orig", numel);
    const FsTensor& u = fneed(d, p + "weight_u", Cout);
    const FsTensor& v = fneed(d, p + "weight_v", (int64_t)Cin * 9);
    float* sigma = d->pool.alloc<float>(1);
    // eval-mode spectral norm: sigma = u^T W_mat v from the stored vectors; dim = 1 for ConvTranspose2d
    if (transposed) spectral_sigma((const float*)wo.p, (const float*)u.p, (const float*)v.p, sigma, Cin, Cout, 9, true, st);
    else spectral_sigma((const float*)wo.p, (const float*)u.p, (const float*)v.p, sigma, Cout, Cin, 9, false, st);
    s.w = (const float*)wo.p;
    s.gscale = sigma;
  } else {
    s.w = (const float*)fneed(d, p + "weight", numel).p;
  }
  conv_pack_into(w, n_off, s, ALL9, st);
  conv_pack_bias(w, n_off, (const float*)fneed(d, p + "bias", Cout).p, Cout, bias_add, st);
}
static ConvW build_conv3(ipk_fs* d, const std::string& p, int engine, int Cout, int Cin, bool transposed, float bias_add, cudaStream_t st) {
  ConvW w = conv_alloc(d->pool, engine, 9, Cin, Cout, true);
  pack_conv3_into(d, w, 0, p, Cout, Cin, transposed, bias_add, st);
  return w;
}

static int gn_groups(int C) {
  int g = 16;
  while (C % g != 0) --g;   // Spade.__init__ (util.py:477-478)
  return g;
}

static void stats_norm(ipk_fs* d, const float* x, int F, long long P, int C, int groups, cudaStream_t st) {
  IPK_CUDA(cudaMemsetAsync(d->sums, 0, (size_t)F * C * 2 * sizeof(double), st));
  NormApply n; n.x = x; n.F = F; n.C = C; n.P = P; n.stats_out = d->sums;     // pure statistics pass
  norm_apply(n, st);
  finalize_stats(d->sums, d->mr, F, P, C, groups, 1e-5f, st);
}

// write `v` (already final fp32 values in x) as the conv-input operand of the big engine
static void to_operand(ipk_fs* d, NormApply& n, void* dst, void* dst_lo) {
  if (d->act_mode == OUT_F32_NHWC) n.out_f32 = (float*)dst;
  else { n.out_hi = (__nv_bfloat16*)dst; n.out_lo = d->act_mode == OUT_BF16_SPLIT ? (__nv_bfloat16*)dst_lo : nullptr; }
}

static void run_convT(const ConvW& w, const ConvIn& in, ConvOut out, cudaStream_t st) {
  // ConvTranspose2d(3, s2, p1, op1) as four stride-1 sub-convolutions, one per output parity class
  out.ymul = 2; out.xmul = 2;
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      out.yadd = a; out.xadd = b;
      conv_run(w, in, out, taps_convT(a, b), 1, st);
    }
}

// SPADE gamma/beta maps of every video (x0-only, frame-invariant): Spade.forward (util.py:494-499)
static void spade_maps(ipk_fs* d, const float* x0, int B, cudaStream_t st) {
  for (UpBlock& ub : d->blocks) {
    const int s = 2 * ub.s_in;
    bilinear_nchw_to_nhwc(x0, d->x0r, B, 3, d->S, s, st);
    ConvOut o; o.p = d->spY; o.p_lo = d->spY_lo; o.mode = d->act_mode; o.cstride = 128; o.Ho = s; o.Wo = s; o.act = ACT_LRELU02; o.bias = ub.sp1.bias;
    if (d->eng == IPK_PREC_FP32_SIMT) {
      ConvIn in; in.p = d->x0r; in.cstride = 3; in.F = B; in.H = s; in.W = s;
      conv_run(ub.sp1, in, o, taps_3x3(), 1, st);
    } else {
      OperandDst im; im.p = d->spI; im.p_lo = d->spI_lo; im.mode = d->act_mode; im.cstride = 32; im.coff = 0;
      im2col3x3_small(d->x0r, B, s, 3, im, 32, st);
      ConvIn in; in.p = d->spI; in.p_lo = d->spI_lo; in.cstride = 32; in.F = B; in.H = s; in.W = s;
      conv_run(ub.sp1, in, o, taps_1x1(), 1, st);
    }
    ConvIn in2; in2.p = d->spY; in2.p_lo = d->spY_lo; in2.cstride = 128; in2.F = B; in2.H = s; in2.W = s;
    ConvOut o2; o2.p = ub.SP; o2.mode = OUT_F32_NHWC; o2.cstride = 2 * ub.Cout; o2.Ho = s; o2.Wo = s; o2.bias = ub.spgb.bias;
    conv_run(ub.spgb, in2, o2, taps_3x3(), 1, st);
  }
}

// decoder for F = nv*T frames whose latents are h [F][64][z] (fp32 NHWC) -> frames [F][3][S][S]; SPADE maps of the nv
// videos start at video offset v0.
static void decode_frames(ipk_fs* d, const float* h, int nv, int T, int v0, float* frames, cudaStream_t st) {
  const int F = nv * T, z = d->z, C0 = d->cfg.dec_channels[0];
  // last SPADE block fused into the final conv (IPK_OUTCONV_FUSED=0: separate normalisation pass)
  static const bool fused_out_env = []() { const char* e = getenv("IPK_OUTCONV_FUSED"); return !(e && e[0] == '0'); }();
  const bool fused_out = fused_out_env && d->eng != IPK_PREC_FP32_SIMT;
  IPK_CHECK(F <= d->Fmax, IPK_ERR_INVALID, "decoder chunk of %d frames exceeds workspace (%d)", F, d->Fmax);
  // ---- in_block: ResBlock(z -> C0, norm='group') at 8x8 (util.py:140-192)
  {
    ProfScope psi("dec.in_block", st);
    NormApply cp; cp.x = h; cp.F = F; cp.C = z; cp.P = 64;
    to_operand(d, cp, d->bufA, d->bufA_lo);
    if (d->act_mode != OUT_F32_NHWC) norm_apply(cp, st);
    ConvIn in; in.p = d->act_mode == OUT_F32_NHWC ? (const void*)h : d->bufA; in.p_lo = d->bufA_lo; in.cstride = z; in.F = F; in.H = 8; in.W = 8;
    const float* resid = h;
    if (d->has_in_res) {
      ConvOut o; o.p = d->bufR; o.cstride = C0; o.Ho = 8; o.Wo = 8; o.bias = d->in_res.bias;
      conv_run(d->in_res, in, o, taps_3x3(), 1, st);
      stats_norm(d, d->bufR, F, 64, C0, 0, st);
      NormApply n; n.x = d->bufR; n.F = F; n.C = C0; n.P = 64; n.mr = d->mr; n.act = ACT_ELU; n.out_f32 = d->bufR;
      norm_apply(n, st);
      resid = d->bufR;
    }
    ConvOut o1; o1.p = d->bufY2; o1.cstride = C0; o1.Ho = 8; o1.Wo = 8; o1.bias = d->in_c1.bias;
    conv_run(d->in_c1, in, o1, taps_3x3(), 1, st);
    stats_norm(d, d->bufY2, F, 64, C0, 16, st);
    NormApply n1; n1.x = d->bufY2; n1.F = F; n1.C = C0; n1.P = 64; n1.mr = d->mr; n1.w = d->gn1_w; n1.b = d->gn1_b; n1.act = ACT_ELU;
    to_operand(d, n1, d->bufY1, d->bufY1_lo);
    norm_apply(n1, st);
    ConvIn in2; in2.p = d->bufY1; in2.p_lo = d->bufY1_lo; in2.cstride = C0; in2.F = F; in2.H = 8; in2.W = 8;
    ConvOut o2; o2.p = d->bufY2; o2.cstride = C0; o2.Ho = 8; o2.Wo = 8; o2.bias = d->in_c2.bias;
    conv_run(d->in_c2, in2, o2, taps_3x3(), 1, st);
    stats_norm(d, d->bufY2, F, 64, C0, 16, st);
    NormApply n2; n2.x = d->bufY2; n2.F = F; n2.C = C0; n2.P = 64; n2.mr = d->mr; n2.w = d->gn2_w; n2.b = d->gn2_b; n2.add = resid;
    to_operand(d, n2, d->bufA, d->bufA_lo);
    norm_apply(n2, st);
  }
  // ---- up blocks + SPADE
  for (size_t i = 0; i < d->blocks.size(); ++i) {
    UpBlock& ub = d->blocks[i];
    const bool last = i + 1 == d->blocks.size();
    const int s = ub.s_in, so = 2 * s;
    const long long P = (long long)so * so;
    ConvIn in; in.p = d->bufA; in.p_lo = d->bufA_lo; in.cstride = ub.Cin; in.F = F; in.H = s; in.W = s;
    // conv1: ConvT + ReLU ("elu" maps to nn.ReLU in Conv2dTransposeBlock, util.py:41-42)
    // res_conv: ConvT (+ InstanceNorm + ReLU applied below)
    const std::string bi = std::to_string(i);
    ConvOut o1; o1.p = d->bufY1; o1.p_lo = d->bufY1_lo; o1.mode = d->act_mode; o1.cstride = ub.Cout; o1.Ho = so; o1.Wo = so; o1.act = ACT_RELU;
    ConvOut orr; orr.p = d->bufR; orr.cstride = ub.Cout; orr.Ho = so; orr.Wo = so;
    if (d->eng == IPK_PREC_FP32_SIMT) {
      o1.bias = ub.ct1.bias; orr.bias = ub.ctr.bias;
      { ProfScope ps(("dec.up.convT1.b" + bi).c_str(), st); run_convT(ub.ct1, in, o1, st); }
      { ProfScope ps(("dec.up.convTres.b" + bi).c_str(), st); run_convT(ub.ctr, in, orr, st); }
    } else {
      // both transposed convs and all four output parity classes in ONE launch: N = conv1 | res_conv, dual destination
      ProfScope ps(("dec.up.convT.b" + bi).c_str(), st);
      IPK_CUDA(cudaMemsetAsync(d->sums, 0, (size_t)F * ub.Cout * 2 * sizeof(double), st));
      orr.stats = d->sums;          // InstanceNorm statistics of res_conv's output, accumulated by the epilogue
      o1.bias = ub.ctf.bias; o1.second = &orr; o1.split_col = ub.Cout; o1.ymul = 2; o1.xmul = 2;
      ConvSub subs[4];
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) { subs[a * 2 + b].taps = taps_convT(a, b); subs[a * 2 + b].yadd = a; subs[a * 2 + b].xadd = b; }
      std::swap(subs[0], subs[3]);      // heaviest class (4 taps) first
      conv_tc_run_multi(ub.ctf, in, o1, subs, 4, st);
    }
    // conv2: ZeroPad(1) + 3x3, no norm, no activation;  out = conv2 + ReLU(IN(res))      (ResBlock.forward, util.py:185-192)
    ConvIn in2; in2.p = d->bufY1; in2.p_lo = d->bufY1_lo; in2.cstride = ub.Cout; in2.F = F; in2.H = so; in2.W = so;
    if (d->eng == IPK_PREC_FP32_SIMT) {
      ConvOut o2; o2.p = d->bufY2; o2.cstride = ub.Cout; o2.Ho = so; o2.Wo = so; o2.bias = ub.c2.bias;
      { ProfScope ps(("dec.up.conv2.b" + bi).c_str(), st); conv_run(ub.c2, in2, o2, taps_3x3(), 1, st); }
      ProfScope pse(("dec.up.norm_spade.b" + bi).c_str(), st);
      stats_norm(d, d->bufR, F, P, ub.Cout, 0, st);
      // (the GroupNorm statistics of `out` needed by SPADE are accumulated by the same pass)
      IPK_CUDA(cudaMemsetAsync(d->sums, 0, (size_t)F * ub.Cout * 2 * sizeof(double), st));
      NormApply n; n.x = d->bufR; n.F = F; n.C = ub.Cout; n.P = P; n.mr = d->mr; n.act = ACT_RELU; n.add = d->bufY2; n.out_f32 = d->bufS;
      n.stats_out = d->sums;
      norm_apply(n, st);
    } else {
      // tensor-core engines: InstanceNorm statistics of res first, then conv2's epilogue adds ReLU(IN(res)), writes `out`
      // and accumulates the GroupNorm statistics SPADE needs -- no separate residual / statistics passes over the tensor
      { ProfScope pse(("dec.up.norm_spade.b" + bi).c_str(), st); finalize_stats(d->sums, d->mr, F, P, ub.Cout, 0, 1e-5f, st); }
      IPK_CUDA(cudaMemsetAsync(d->sums, 0, (size_t)F * ub.Cout * 2 * sizeof(double), st));
      ConvOut o2; o2.p = d->bufS; o2.cstride = ub.Cout; o2.Ho = so; o2.Wo = so; o2.bias = ub.c2.bias;
      o2.res = d->bufR; o2.res_cstride = ub.Cout; o2.res_act = ACT_RELU; o2.res_mr = d->mr; o2.stats = d->sums;
      ProfScope ps(("dec.up.conv2.b" + bi).c_str(), st);
      conv_run(ub.c2, in2, o2, taps_3x3(), 1, st);
    }
    ProfScope pse(("dec.up.norm_spade.b" + bi).c_str(), st);
    // SPADE: GroupNorm(16, affine=False)(out) * (1 + gamma) + beta
    finalize_stats(d->sums, d->mr, F, P, ub.Cout, ub.groups, 1e-5f, st);
    if (last && d->out_direct && fused_out) break;      // the final conv normalises its input tiles itself (out_conv.cu)
    NormApply sp; sp.x = d->bufS; sp.F = F; sp.C = ub.Cout; sp.P = P; sp.mr = d->mr; sp.spade = ub.SP + (size_t)v0 * P * 2 * ub.Cout; sp.T = T;
    if (last && d->out_direct) sp.out_f32 = (float*)d->bufA;      // the final conv reads fp32
    else to_operand(d, sp, d->bufA, d->bufA_lo);
    norm_apply(sp, st);
  }
  // ---- out_conv: 3x3 -> 3 channels + tanh, written straight into the NCHW frame tensor
  {
    ProfScope ps("dec.out_conv", st);
    if (d->out_direct && fused_out) {
      const UpBlock& ub = d->blocks.back();
      out_conv_run(d->out_direct, d->bufS, frames, F, d->S, st, d->mr, ub.SP + (size_t)v0 * d->S * d->S * 2 * ub.Cout, T);
    } else if (d->out_direct) {
      out_conv_run(d->out_direct, (const float*)d->bufA, frames, F, d->S, st);
    } else {
      const int Cl = d->cfg.dec_channels[d->nd - 1];
      ConvIn in; in.p = d->bufA; in.p_lo = d->bufA_lo; in.cstride = Cl; in.F = F; in.H = d->S; in.W = d->S;
      ConvOut o; o.p = frames; o.mode = OUT_F32_NCHW; o.Ho = d->S; o.Wo = d->S; o.act = ACT_TANH; o.bias = d->out_conv.bias;
      conv_run(d->out_conv, in, o, taps_3x3(), 1, st);
    }
  }
}

static OperandDst gru_operand(ipk_fs* d, char* base, int coff) {
  const size_t Mmax = (size_t)d->cfg.max_batch * 64;
  OperandDst o;
  o.p = base; o.mode = d->act_mode; o.cstride = 2 * d->z; o.coff = coff;
  o.p_lo = d->act_mode == OUT_BF16_SPLIT ? base + Mmax * 2 * d->z * 2 : nullptr;
  return o;
}

// (x | h) and (x | r*h) operands of layer l from fp32 sources: x [M][z] (may be null = leave), h = Hf[l]
static void gru_load_layer(ipk_fs* d, int l, const float* x, int B, cudaStream_t st) {
  const int z = d->z;
  const long long M = (long long)B * 64;
  operand_copy(d->Hf[l], z, 0, gru_operand(d, d->XH[l], z), M, z, st);
  if (x) {
    operand_copy(x, z, 0, gru_operand(d, d->XH[l], 0), M, z, st);
    operand_copy(x, z, 0, gru_operand(d, d->XRH[l], 0), M, z, st);
  }
}

// one ConvGRU step over all layers (ConvGRU.forward, rnn.py:104-133); XH[l] = (x | h), XRH[l] = (x | r*h)
static void gru_step(ipk_fs* d, int B, float* seq_out, int T, int t, cudaStream_t st) {
  const int z = d->z;
  const long long M = (long long)B * 64;
  for (int l = 0; l < d->L; ++l) {
    const OperandDst xh = gru_operand(d, d->XH[l], 0), xrh = gru_operand(d, d->XRH[l], 0);
    ConvIn in; in.p = xh.p; in.p_lo = xh.p_lo; in.cstride = 2 * z; in.F = B; in.H = 8; in.W = 8;
    ConvOut o; o.p = d->graw; o.cstride = d->gru_ur[l].Npad; o.Ho = 8; o.Wo = 8; o.bias = d->gru_ur[l].bias;
    conv_run(d->gru_ur[l], in, o, taps_3x3(), 1, st);
    gru_gate1(d->graw, d->Hf[l], d->U, gru_operand(d, d->XRH[l], z), M, z, st);
    ConvIn in2; in2.p = xrh.p; in2.p_lo = xrh.p_lo; in2.cstride = 2 * z; in2.F = B; in2.H = 8; in2.W = 8;
    ConvOut o2; o2.p = d->graw; o2.cstride = d->gru_o[l].Npad; o2.Ho = 8; o2.Wo = 8; o2.bias = d->gru_o[l].bias;
    conv_run(d->gru_o[l], in2, o2, taps_3x3(), 1, st);
    OperandDst dst[3];
    int nd = 0;
    dst[nd++] = gru_operand(d, d->XH[l], z);
    if (l + 1 < d->L) {
      dst[nd++] = gru_operand(d, d->XH[l + 1], 0);
      dst[nd++] = gru_operand(d, d->XRH[l + 1], 0);
    }
    gru_gate2(d->graw, d->U, d->Hf[l], M, z, dst, nd, (l + 1 == d->L) ? seq_out : nullptr, T, t, st);
  }
}

}  // namespace ipk

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" int ipk_fs_create(const ipk_fs_config* cfg, ipk_fs** out) {
  IPK_TRY
  IPK_CHECK(cfg && out, IPK_ERR_INVALID, "ipk_fs_create: null argument");
  IPK_CHECK(cfg->n_dec >= 2 && cfg->n_dec <= IPK_MAX_DEC, IPK_ERR_INVALID, "first stage: bad n_dec");
  IPK_CHECK(cfg->spatial == (8 << (cfg->n_dec - 1)), IPK_ERR_INVALID, "first stage: spatial %d does not match %d up blocks from 8x8", cfg->spatial, cfg->n_dec - 1);
  IPK_CHECK(cfg->z_dim % 4 == 0 && cfg->z_dim > 0, IPK_ERR_UNSUPPORTED, "first stage: z_dim must be a multiple of 4");
  IPK_CHECK(cfg->precision >= 0 && cfg->precision <= 2, IPK_ERR_INVALID, "first stage: bad precision");
  IPK_CHECK(cfg->max_batch > 0 && cfg->max_frames > 0 && cfg->n_gru_layers > 0, IPK_ERR_INVALID, "first stage: bad sizes");
  for (int i = 0; i < cfg->n_dec; ++i) {
    IPK_CHECK(cfg->dec_channels[i] % 16 == 0, IPK_ERR_UNSUPPORTED, "first stage: dec_channels must be multiples of 16");
    if (cfg->precision != IPK_PREC_FP32_SIMT)
      IPK_CHECK(cfg->dec_channels[i] % 64 == 0, IPK_ERR_UNSUPPORTED, "first stage: tensor-core engine needs dec_channels %% 64 == 0");
  }
  ipk_fs* d = new ipk_fs();
  d->cfg = *cfg;
  d->z = cfg->z_dim; d->S = cfg->spatial; d->L = cfg->n_gru_layers; d->nd = cfg->n_dec; d->eng = cfg->precision;
  d->act_mode = d->eng == IPK_PREC_FP32_SIMT ? OUT_F32_NHWC : (d->eng == IPK_PREC_FP32_SPLIT ? OUT_BF16_SPLIT : OUT_BF16);
  d->chunk_videos = cfg->chunk_videos > 0 ? cfg->chunk_videos : std::max(1, 256 / cfg->max_frames);
  d->chunk_videos = std::min(d->chunk_videos, cfg->max_batch);
  *out = d;
  IPK_CATCH
}

extern "C" int ipk_fs_set_tensor(ipk_fs* d, const char* name, const void* dev_ptr, int64_t numel, int dtype) {
  IPK_TRY
  IPK_CHECK(d && name && dev_ptr, IPK_ERR_INVALID, "ipk_fs_set_tensor: null argument");
  IPK_CHECK(!d->finalized, IPK_ERR_STATE, "ipk_fs_set_tensor after finalize");
  d->tensors[name] = FsTensor{dev_ptr, numel, dtype};
  IPK_CATCH
}

extern "C" int ipk_fs_finalize(ipk_fs* d, void* stream) {
  IPK_TRY
  IPK_CHECK(d && !d->finalized, IPK_ERR_STATE, "first stage: null or already finalized");
  cudaStream_t st = (cudaStream_t)stream;
  const int z = d->z, eng = d->eng;
  const int* dc = d->cfg.dec_channels;
  // ---- ConvGRU: update|reset gates share one contraction (N = 2z)
  const int geng = (eng != IPK_PREC_FP32_SIMT && z % 16 == 0) ? eng : IPK_PREC_FP32_SIMT;
  IPK_CHECK(geng == eng, IPK_ERR_UNSUPPORTED, "first stage: tensor-core engine needs z_dim %% 16 == 0");
  for (int l = 0; l < d->L; ++l) {
    std::string p = "rnn.cells." + std::to_string(l) + ".";
    ConvW ur = conv_alloc(d->pool, geng, 9, 2 * z, 2 * z, true);
    PackSrc s; s.N = z; s.Ksrc = 2 * z; s.kh = 3; s.kw = 3;
    s.w = (const float*)fneed(d, p + "update_gate.weight", (int64_t)z * 2 * z * 9).p;
    conv_pack_into(ur, 0, s, ALL9, st);
    conv_pack_bias(ur, 0, (const float*)fneed(d, p + "update_gate.bias", z).p, z, 0.f, st);
    s.w = (const float*)fneed(d, p + "reset_gate.weight", (int64_t)z * 2 * z * 9).p;
    conv_pack_into(ur, z, s, ALL9, st);
    conv_pack_bias(ur, z, (const float*)fneed(d, p + "reset_gate.bias", z).p, z, 0.f, st);
    d->gru_ur.push_back(ur);
    ConvW og = conv_alloc(d->pool, geng, 9, 2 * z, z, true);
    s.w = (const float*)fneed(d, p + "out_gate.weight", (int64_t)z * 2 * z * 9).p;
    conv_pack_into(og, 0, s, ALL9, st);
    conv_pack_bias(og, 0, (const float*)fneed(d, p + "out_gate.bias", z).p, z, 0.f, st);
    d->gru_o.push_back(og);
  }
  d->motion_bias = dev_copy(d, fneed(d, "motion_bias", (int64_t)z * 64).p, (size_t)z * 64, st);
  // ---- in_block
  d->in_c1 = build_conv3(d, "gen.in_block.conv1.conv.", eng, dc[0], z, false, 0.f, st);
  d->in_c2 = build_conv3(d, "gen.in_block.conv2.conv.", eng, dc[0], dc[0], false, 0.f, st);
  d->gn1_w = dev_copy(d, fneed(d, "gen.in_block.conv1.norm.weight", dc[0]).p, dc[0], st);
  d->gn1_b = dev_copy(d, fneed(d, "gen.in_block.conv1.norm.bias", dc[0]).p, dc[0], st);
  d->gn2_w = dev_copy(d, fneed(d, "gen.in_block.conv2.norm.weight", dc[0]).p, dc[0], st);
  d->gn2_b = dev_copy(d, fneed(d, "gen.in_block.conv2.norm.bias", dc[0]).p, dc[0], st);
  d->has_in_res = z != dc[0];
  if (d->has_in_res) d->in_res = build_conv3(d, "gen.in_block.res_conv.conv.", eng, dc[0], z, false, 0.f, st);
  // ---- up blocks
  size_t fe = (size_t)64 * std::max(dc[0], z);
  for (int i = 0; i + 1 < d->nd; ++i) {
    UpBlock ub;
    ub.Cin = dc[i]; ub.Cout = dc[i + 1]; ub.s_in = 8 << i;
    std::string p = "gen.blocks." + std::to_string(i) + ".";
    if (eng == IPK_PREC_FP32_SIMT) {
      ub.ct1 = build_conv3(d, p + "conv1.conv.", eng, ub.Cout, ub.Cin, true, 0.f, st);
      ub.ctr = build_conv3(d, p + "res_conv.conv.", eng, ub.Cout, ub.Cin, true, 0.f, st);
    } else {
      ub.ctf = conv_alloc(d->pool, eng, 9, ub.Cin, 2 * ub.Cout, true);
      pack_conv3_into(d, ub.ctf, 0, p + "conv1.conv.", ub.Cout, ub.Cin, true, 0.f, st);
      pack_conv3_into(d, ub.ctf, ub.Cout, p + "res_conv.conv.", ub.Cout, ub.Cin, true, 0.f, st);
    }
    ub.c2 = build_conv3(d, p + "conv2.conv.", eng, ub.Cout, ub.Cout, false, 0.f, st);
    std::string sp = "gen.spade_blocks." + std::to_string(i) + ".";
    if (eng == IPK_PREC_FP32_SIMT) {
      ub.sp1 = build_conv3(d, sp + "conv.", IPK_PREC_FP32_SIMT, 128, 3, false, 0.f, st);
    } else {
      // one-tap GEMM over 3x3 im2col rows: k = tap*3 + c  <-  OIHW column c*9 + tap
      ub.sp1 = conv_alloc(d->pool, eng, 1, 27, 128, true);
      std::vector<int> kmap(27);
      for (int t = 0; t < 9; ++t)
        for (int c = 0; c < 3; ++c) kmap[t * 3 + c] = c * 9 + t;
      int* d_kmap = d->pool.alloc<int>(27);
      IPK_CUDA(cudaMemcpyAsync(d_kmap, kmap.data(), 27 * sizeof(int), cudaMemcpyHostToDevice, st));
      IPK_CUDA(cudaStreamSynchronize(st));
      PackSrc ps; ps.w = (const float*)fneed(d, sp + "conv.weight", 128 * 27).p; ps.N = 128; ps.Ksrc = 27; ps.k_map = d_kmap;
      conv_pack_into(ub.sp1, 0, ps, {0}, st);
      conv_pack_bias(ub.sp1, 0, (const float*)fneed(d, sp + "conv.bias", 128).p, 128, 0.f, st);
    }
    // gamma and beta convolutions fused along N; "+1" of (1 + gamma) folded into the gamma bias
    ub.spgb = conv_alloc(d->pool, eng, 9, 128, 2 * ub.Cout, true);
    PackSrc s; s.N = ub.Cout; s.Ksrc = 128; s.kh = 3; s.kw = 3;
    s.w = (const float*)fneed(d, sp + "conv_gamma.weight", (int64_t)ub.Cout * 128 * 9).p;
    conv_pack_into(ub.spgb, 0, s, ALL9, st);
    conv_pack_bias(ub.spgb, 0, (const float*)fneed(d, sp + "conv_gamma.bias", ub.Cout).p, ub.Cout, 1.0f, st);
    s.w = (const float*)fneed(d, sp + "conv_beta.weight", (int64_t)ub.Cout * 128 * 9).p;
    conv_pack_into(ub.spgb, ub.Cout, s, ALL9, st);
    conv_pack_bias(ub.spgb, ub.Cout, (const float*)fneed(d, sp + "conv_beta.bias", ub.Cout).p, ub.Cout, 0.f, st);
    ub.groups = gn_groups(ub.Cout);
    const size_t so = 2 * (size_t)ub.s_in;
    ub.SP = d->pool.alloc<float>((size_t)d->cfg.max_batch * so * so * 2 * ub.Cout);
    fe = std::max(fe, so * so * (size_t)ub.Cout);
    fe = std::max(fe, (size_t)ub.s_in * ub.s_in * ub.Cin);
    d->blocks.push_back(ub);
  }
  // Final 3x3 64 -> 3 conv: the fp32 FFMA halo-tile kernel (out_conv.cu).  IPK_OUTCONV_TC=1 runs it on the tensor-core engine in halo mode
  // instead (BN = 32); measured on B200 that is SLOWER (3.48 vs 2.53 ms per 1 024 frames): one 128-pixel row per tile leaves ~4 us of
  // TMA round-trip latency per tile that a 3-stage ring cannot hide, while the MMA work per tile is ~0.1 us.
  static const bool outconv_tc = []() { const char* e = getenv("IPK_OUTCONV_TC"); return e && e[0] == '1'; }();
  if (dc[d->nd - 1] == OUT_CONV_CIN && !(outconv_tc && eng != IPK_PREC_FP32_SIMT && d->S == 128)) {
    const std::string p = "gen.out_conv.conv.";
    float* packed = d->pool.alloc<float>(OUT_CONV_PACKED_FLOATS);
    const float* bias = (const float*)fneed(d, p + "bias", 3).p;
    if (fhas(d, p + "weight_orig")) {
      const FsTensor& wo = fneed(d, p + "weight_orig", 3 * 64 * 9);
      float* sigma = d->pool.alloc<float>(1);
      spectral_sigma((const float*)wo.p, (const float*)fneed(d, p + "weight_u", 3).p, (const float*)fneed(d, p + "weight_v", 64 * 9).p, sigma, 3, 64, 9, false, st);
      out_conv_pack((const float*)wo.p, sigma, bias, packed, st);
    } else {
      out_conv_pack((const float*)fneed(d, p + "weight", 3 * 64 * 9).p, nullptr, bias, packed, st);
    }
    d->out_direct = out_conv_plan_create(packed, st);
  } else {
    d->out_conv = build_conv3(d, "gen.out_conv.conv.", eng, 3, dc[d->nd - 1], false, 0.f, st);
  }
  d->frame_elems = fe;
  // ---- workspace
  const size_t Mmax = (size_t)d->cfg.max_batch * 64;
  d->Fmax = d->chunk_videos * d->cfg.max_frames;
  const size_t Fm = std::max<size_t>(d->Fmax, 1);
  auto rb = [](size_t b) { return (b + 255) / 256 * 256; };
  size_t bytes = 0;
  bytes += (size_t)d->L * (2 * rb(Mmax * 2 * z * 4) + rb(Mmax * z * 4)) + 2 * rb(Mmax * z * 4) + rb(Mmax * 2 * z * 4);
  bytes += rb((size_t)d->cfg.max_batch * d->S * d->S * 32 * 4);
  bytes += rb(Mmax * d->cfg.max_frames * z * 4);                               // Hseq
  bytes += rb((size_t)d->cfg.max_batch * d->S * d->S * 3 * 4) + rb((size_t)d->cfg.max_batch * d->S * d->S * 128 * 4);
  bytes += 5 * rb(Fm * fe * 4);
  bytes += rb(Fm * 512 * 2 * 8) + rb(Fm * 512 * 2 * 4) + 65536;
  d->ws.init(bytes);
  for (int l = 0; l < d->L; ++l) {
    d->XH.push_back((char*)d->ws.alloc<float>(Mmax * 2 * z));
    d->XRH.push_back((char*)d->ws.alloc<float>(Mmax * 2 * z));
    d->Hf.push_back(d->ws.alloc<float>(Mmax * z));
  }
  d->xtmp = d->ws.alloc<float>(Mmax * z);
  {
    const size_t n = (size_t)d->cfg.max_batch * d->S * d->S * 32;
    char* p = (char*)d->ws.alloc<float>(n);
    d->spI = p;
    d->spI_lo = d->act_mode == OUT_BF16_SPLIT ? p + n * 2 : nullptr;
  }
  d->U = d->ws.alloc<float>(Mmax * z);
  d->graw = d->ws.alloc<float>(Mmax * 2 * z);
  d->Hseq = d->ws.alloc<float>(Mmax * d->cfg.max_frames * z);
  d->x0r = d->ws.alloc<float>((size_t)d->cfg.max_batch * d->S * d->S * 3);
  const size_t spn = (size_t)d->cfg.max_batch * d->S * d->S * 128;
  char* spy = (char*)d->ws.alloc<float>(spn);
  d->spY = spy;
  d->spY_lo = d->act_mode == OUT_BF16_SPLIT ? spy + spn * 2 : nullptr;
  char* a = (char*)d->ws.alloc<float>(Fm * fe);
  char* y1 = (char*)d->ws.alloc<float>(Fm * fe);
  d->bufA = a; d->bufY1 = y1;
  d->bufA_lo = d->act_mode == OUT_BF16_SPLIT ? a + Fm * fe * 2 : nullptr;
  d->bufY1_lo = d->act_mode == OUT_BF16_SPLIT ? y1 + Fm * fe * 2 : nullptr;
  d->bufR = d->ws.alloc<float>(Fm * fe);
  d->bufY2 = d->ws.alloc<float>(Fm * fe);
  d->bufS = d->ws.alloc<float>(Fm * fe);
  d->sums = d->ws.alloc<double>(Fm * 512 * 2);
  d->mr = d->ws.alloc<float>(Fm * 512 * 2);
  IPK_CUDA(cudaStreamSynchronize(st));
  d->tensors.clear();
  d->finalized = true;
  IPK_CATCH
}

// frames_host / copy_st (optional): every finished chunk of frames is copied to the host on a second stream while the next
// chunk is being decoded; the caller synchronises copy_st.  u8_dev / u8_host (optional, same chunking): the chunk is also
// converted to uint8 NTHWC (frames_to_u8) and THAT leaves for the host instead of the fp32 frames.
static void fs_decode_impl(ipk_fs* d, const float* motion, bool motion_is_nhwc, const float* x0, float* frames, int B, int T, cudaStream_t st,
                           float* frames_host = nullptr, cudaStream_t copy_st = nullptr, uint8_t* u8_dev = nullptr, uint8_t* u8_host = nullptr) {
  IPK_CHECK(d && d->finalized, IPK_ERR_STATE, "first stage not finalized");
  IPK_CHECK(B > 0 && B <= d->cfg.max_batch, IPK_ERR_INVALID, "first stage: batch %d outside (0, %d]", B, d->cfg.max_batch);
  IPK_CHECK(T > 0 && T <= d->cfg.max_frames, IPK_ERR_INVALID, "first stage: T %d outside (0, %d]", T, d->cfg.max_frames);
  const int z = d->z;
  const long long M = (long long)B * 64;
  // hidden = [motion] * n_layers ; in_rnn = motion_bias repeated over the batch (second_stage_video.py:365-372)
  for (int l = 0; l < d->L; ++l) {
    if (motion_is_nhwc) copy_channels(motion, z, 0, d->Hf[l], z, 0, M, z, st);
    else nchw_to_nhwc(motion, d->Hf[l], B, z, 64, z, st);
  }
  broadcast_chw_to_nhwc(d->motion_bias, d->xtmp, B, z, 64, z, 0, st);
  for (int l = 0; l < d->L; ++l) gru_load_layer(d, l, l == 0 ? d->xtmp : nullptr, B, st);
  {
    ProfScope ps("gru", st);
    for (int t = 0; t < T; ++t) gru_step(d, B, d->Hseq, T, t, st);
  }
  {
    ProfScope ps("spade_maps", st);
    spade_maps(d, x0, B, st);
  }
  for (int v0 = 0; v0 < B; v0 += d->chunk_videos) {
    int nv = std::min(d->chunk_videos, B - v0);
    const size_t off = (size_t)v0 * T * 3 * d->S * d->S, cnt = (size_t)nv * T * 3 * d->S * d->S;
    decode_frames(d, d->Hseq + (size_t)v0 * T * 64 * z, nv, T, v0, frames + off, st);
    if (u8_dev) frames_to_u8(frames + off, u8_dev + off, (long long)nv * T, d->S * d->S, st);
    if (frames_host || u8_host) {
      if (!d->chunk_done) IPK_CUDA(cudaEventCreateWithFlags(&d->chunk_done, cudaEventDisableTiming));
      IPK_CUDA(cudaEventRecord(d->chunk_done, st));
      IPK_CUDA(cudaStreamWaitEvent(copy_st, d->chunk_done, 0));
      if (u8_host) IPK_CUDA(cudaMemcpyAsync(u8_host + off, u8_dev + off, cnt, cudaMemcpyDeviceToHost, copy_st));
      else IPK_CUDA(cudaMemcpyAsync(frames_host + off, frames + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, copy_st));
    }
  }
}

extern "C" int ipk_fs_decode(ipk_fs* d, const float* motion, const float* x0, float* frames, int32_t B, int32_t T, void* stream) {
  IPK_TRY
  IPK_CHECK(motion && x0 && frames, IPK_ERR_INVALID, "ipk_fs_decode: null buffer");
  fs_decode_impl(d, motion, false, x0, frames, B, T, (cudaStream_t)stream);
  IPK_CATCH
}

// internal entry used by ipk_sample: motion already NHWC [B][64][z] on device
int ipk_fs_decode_nhwc(ipk_fs* d, const float* motion_nhwc, const float* x0, float* frames, int B, int T, cudaStream_t st,
                       float* frames_host, cudaStream_t copy_st, uint8_t* u8_dev, uint8_t* u8_host) {
  fs_decode_impl(d, motion_nhwc, true, x0, frames, B, T, st, frames_host, copy_st, u8_dev, u8_host);
  return 0;
}

extern "C" int ipk_fs_gru_step(ipk_fs* d, const float* x, const float* hidden, float* new_hidden, int32_t B, void* stream) {
  IPK_TRY
  IPK_CHECK(d && d->finalized, IPK_ERR_STATE, "first stage not finalized");
  IPK_CHECK(B > 0 && B <= d->cfg.max_batch && x && hidden && new_hidden, IPK_ERR_INVALID, "ipk_fs_gru_step: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int z = d->z;
  const size_t per = (size_t)B * z * 64;
  for (int l = 0; l < d->L; ++l) nchw_to_nhwc(hidden + l * per, d->Hf[l], B, z, 64, z, st);
  nchw_to_nhwc(x, d->xtmp, B, z, 64, z, st);
  for (int l = 0; l < d->L; ++l) gru_load_layer(d, l, l == 0 ? d->xtmp : nullptr, B, st);
  gru_step(d, B, nullptr, 1, 0, st);
  for (int l = 0; l < d->L; ++l) nhwc_to_nchw(d->Hf[l], new_hidden + l * per, B, z, 64, z, st);
  IPK_CATCH
}

extern "C" int ipk_fs_gen(ipk_fs* d, const float* h, const float* x0, float* frame, int32_t B, void* stream) {
  IPK_TRY
  IPK_CHECK(d && d->finalized, IPK_ERR_STATE, "first stage not finalized");
  IPK_CHECK(B > 0 && B <= d->cfg.max_batch && h && x0 && frame, IPK_ERR_INVALID, "ipk_fs_gen: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int z = d->z;
  nchw_to_nhwc(h, d->Hseq, B, z, 64, z, st);
  spade_maps(d, x0, B, st);
  const int per = std::max(1, std::min(d->Fmax, B));
  for (int v0 = 0; v0 < B; v0 += per) {
    int nv = std::min(per, B - v0);
    decode_frames(d, d->Hseq + (size_t)v0 * 64 * z, nv, 1, v0, frame + (size_t)v0 * 3 * d->S * d->S, st);
  }
  IPK_CATCH
}

extern "C" int ipk_fs_destroy(ipk_fs* d) {
  if (!d) return IPK_OK;
  ipk_graphs_drop(d);
  d->pool.release();
  d->ws.release();
  if (d->out_direct) out_conv_plan_destroy(d->out_direct);
  if (d->chunk_done) cudaEventDestroy(d->chunk_done);
  delete d;
  return IPK_OK;
}

int ipk_fs_spatial(ipk_fs* d) { return d->S; }
