// Conditional MaCow flow plan: SupervisedMacowTransformer.forward / reverse (models/modules/INN/INN.py:469-481),
// MultiScaleInternal.forward (models/modules/INN/macow2.py:873-920), MaCowStep (macow2.py:1066-1117),
// MaCowUnit (macow2.py:957-995), MultiScalePrior (macow2.py:569-593), NICE2d (macow2.py:395-448),
// NICEConvBlock (models/modules/INN/macow_utils.py:313-337).
//
// Data layout: the whole flow runs IN PLACE on one fp32 NHWC state buffer [B][8*8][C0].  Level L works on the first
// C_L channels of every pixel; the channels split off to the prior by earlier levels stay where they are, which is
// exactly the reference's output order z = [final | L14 slice | ... | L0 slice] (macow2.py:893-899,904-907).
// NICE split/unsplit never moves data: each coupling carries two channel index lists (network input z, transformed zp).
#include <map>
#include <string>
#include <cstdarg>
#include "conv.cuh"
#include "elementwise.cuh"
#include "flow_segment.cuh"
#include "flow_plan.cuh"

namespace ipk {

struct TensorRef { const void* p; int64_t numel; int dtype; };

struct NiceLayer {
  ConvW c1, c2, c3;          // c3: one GEMM producing all 9 tap responses, N = 9 * N3p (tap-major columns)
  int n_z = 0, n_p = 0, K1pad = 0, N3p = 0, nsplit3 = 1;
  int* d_iz = nullptr;
  int* d_ip = nullptr;
  float* bias3 = nullptr;  // [2*n_p]
};

struct Stage {          // one segment launch followed (optionally) by one NICE network
  std::vector<MicroOp> host_ops;
  MicroOp* d_ops = nullptr;
  int C = 0;
  bool has_mcf = false;
  int nice_id = -1;     // NICE whose network runs after this segment (the segment ends with its IM2COL)
};

struct McfPacked { float* Wc; float* W1x; int C, Cp, hid; int hcol; const float* v_src; const float* os; const float* b_src; };

}  // namespace ipk

void ipk_graphs_drop(const void* handle);   // capi.cu
using namespace ipk;

struct ipk_flow {
  ipk_flow_config cfg;
  std::map<std::string, TensorRef> tensors;
  bool finalized = false;
  DevPool pool;
  Arena ws;
  int C0 = 0, Hd = 0, hch = 0;
  std::vector<NiceLayer> nices;
  std::map<std::string, int> nice_by_prefix;
  std::map<std::string, McfPacked> mcfs;
  std::map<std::string, int*> shuffles;     // key = prefix + ("f"|"b")
  std::vector<Stage> prog_fwd, prog_inv;
  // workspace
  float* state = nullptr;
  float* cond = nullptr;
  void* A1 = nullptr; void* A1_lo = nullptr;
  void* H1 = nullptr; void* H1_lo = nullptr;
  void* H2 = nullptr; void* H2_lo = nullptr;
  float* partials = nullptr;
  float* logdet_ws = nullptr;
  // conditioning terms of all MCFs: Hterm[M][hstride] = bias + W1h * ELU(cond), one GEMM per flow pass
  ConvW hterm_w;
  int hterm_cols = 0, hstride = 0;
  void* E = nullptr; void* E_lo = nullptr;   // ELU(cond) operand [M][hch]
  float* Hterm = nullptr;
  // two half-batches on two streams (see run_program)
  bool dual_stream = false;
  cudaStream_t st2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int K1pad_max = 0, Npad3_max = 0;
  int act_mode = OUT_F32_NHWC;
};

namespace ipk {

static const TensorRef& need(ipk_flow* f, const std::string& name, int64_t numel, int dtype) {
  auto it = f->tensors.find(name);
  IPK_CHECK(it != f->tensors.end(), IPK_ERR_MISSING, "flow: missing tensor '%s'", name.c_str());
  IPK_CHECK(it->second.numel == numel, IPK_ERR_SHAPE, "flow: tensor '%s' has %lld elements, expected %lld", name.c_str(),
            (long long)it->second.numel, (long long)numel);
  IPK_CHECK(it->second.dtype == dtype, IPK_ERR_SHAPE, "flow: tensor '%s' has dtype %d, expected %d", name.c_str(), it->second.dtype, dtype);
  return it->second;
}

std::vector<LevelInfo> levels_of(const ipk_flow_config& c) {
  // MultiScaleInternal.__init__ channel bookkeeping (macow2.py:825-871)
  std::vector<LevelInfo> v;
  int C = c.flow_in_channels, factor = c.factor, step = C / c.factor;
  for (int L = 0; L < c.n_levels; ++L) {
    LevelInfo li{L, C, c.num_steps[L], factor, C / factor, C - C / factor};
    v.push_back(li);
    C -= step;
    IPK_CHECK(C == li.z1, IPK_ERR_INVALID, "flow: channel bookkeeping mismatch at level %d (macow2.py:868)", L);
    factor -= 1;
  }
  return v;
}

// channel index lists of NICE2d.split (macow2.py:301-317,364-377)
void nice_indices(int C, int factor, bool skip, bool up, std::vector<int>& iz, std::vector<int>& ip) {
  if (skip && (C % 2 == 1)) skip = false;
  int cout = C / factor, cin = C - cout;
  int z1c = up ? cin : cout;
  std::vector<int> i1, i2;
  if (!skip) {
    for (int i = 0; i < z1c; ++i) i1.push_back(i);
    for (int i = z1c; i < C; ++i) i2.push_back(i);
  } else {
    for (int i = 0; i < C; i += 2) i1.push_back(i);
    for (int i = 1; i < C; i += 2) i2.push_back(i);
  }
  if (up) { iz = i1; ip = i2; } else { iz = i2; ip = i1; }
}

static int* upload_ints(ipk_flow* f, const std::vector<int>& v, cudaStream_t st) {
  int* d = f->pool.alloc<int>(v.size());
  IPK_CUDA(cudaMemcpyAsync(d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  IPK_CUDA(cudaStreamSynchronize(st));
  return d;
}

static int build_nice(ipk_flow* f, const std::string& p, int C, int factor, bool skip, bool up, cudaStream_t st) {
  NiceLayer n;
  std::vector<int> iz, ip;
  nice_indices(C, factor, skip, up, iz, ip);
  n.n_z = (int)iz.size();
  n.n_p = (int)ip.size();
  const int Hd = f->Hd, eng = f->cfg.precision;
  n.K1pad = round_up(9 * n.n_z, 64);
  n.d_iz = upload_ints(f, iz, st);
  n.d_ip = upload_ints(f, ip, st);
  // conv1: 3x3, no bias; packed as a 1-tap GEMM over the im2col'd rows: k = tap * n_z + j  (tap = ky*3+kx)
  const TensorRef& w1 = need(f, p + "net.conv1.weight", (int64_t)Hd * n.n_z * 9, IPK_F32);
  n.c1 = conv_alloc(f->pool, eng, 1, 9 * n.n_z, Hd, false);
  {
    // view OIHW [Hd][n_z][3][3] as a [Hd][n_z*9] matrix whose column (j*9 + tap) must land at k = tap*n_z + j
    std::vector<int> kmap(9 * n.n_z);
    for (int t = 0; t < 9; ++t)
      for (int j = 0; j < n.n_z; ++j) kmap[t * n.n_z + j] = j * 9 + t;
    int* d_kmap = upload_ints(f, kmap, st);
    PackSrc s;
    s.w = (const float*)w1.p; s.N = Hd; s.Ksrc = 9 * n.n_z; s.kh = 1; s.kw = 1; s.k_map = d_kmap;
    conv_pack_into(n.c1, 0, s, {0}, st);
  }
  // conv2: 1x1, no bias
  const TensorRef& w2 = need(f, p + "net.conv2.weight", (int64_t)Hd * Hd, IPK_F32);
  n.c2 = conv_alloc(f->pool, eng, 1, Hd, Hd, false);
  {
    PackSrc s;
    s.w = (const float*)w2.p; s.N = Hd; s.Ksrc = Hd;
    conv_pack_into(n.c2, 0, s, {0}, st);
  }
  // conv3: weight-normed 3x3 with bias (Conv2dWeightNorm, macow_utils.py:211-251): w = g * v / ||v||
  const int N3 = 2 * n.n_p;
  const TensorRef& v3 = need(f, p + "net.conv3.conv.weight_v", (int64_t)N3 * Hd * 9, IPK_F32);
  const TensorRef& g3 = need(f, p + "net.conv3.conv.weight_g", N3, IPK_F32);
  const TensorRef& b3 = need(f, p + "net.conv3.conv.bias", N3, IPK_F32);
  float* os = f->pool.alloc<float>(N3);
  weight_norm_scale((const float*)v3.p, (const float*)g3.p, os, N3, Hd * 9, st);
  // conv3 runs as ONE 1x1 GEMM T[q][t*N3p + j] = W_t[j,:] . H2[q,:] over all pixels q (the hidden activations are read
  // once instead of once per tap); the 3x3 gather  out[p] = sum_t T[p + delta_t][t]  happens in the affine micro-op.
  n.N3p = round_up(N3, 4);
  n.c3 = conv_alloc(f->pool, eng, 1, Hd, 9 * n.N3p, false);
  for (int t = 0; t < 9; ++t) {
    PackSrc s;
    s.w = (const float*)v3.p; s.N = N3; s.Ksrc = Hd; s.kh = 3; s.kw = 3; s.oscale = os;
    conv_pack_into(n.c3, t * n.N3p, s, {t}, st);
  }
  {
    // split-K (tensor-core engine only): the slice count that minimises (waves of the persistent grid) x (k-blocks per slice), with a
    // small charge per slice for the partial sums the next segment gathers.  The engine runs layers wider than 128 columns as
    // cdiv(Npad, 256) N tiles on CTA pairs (74 slots), narrower ones on single CTAs (148 slots).  (The previous rule assumed 128-column
    // tiles: the 288-column conv3 of the two widest levels got 3 slices = 96 pair-units = two waves, the second 30 % full.)
    const int mt = cdiv(f->cfg.max_batch * 64, 128);
    const int nkb = n.c3.Kpad / 64;
    const bool pairs = n.c3.Npad > 128 && mt >= 2;
    const long long per_slice = (long long)(pairs ? cdiv(mt, 2) : mt) * cdiv(n.c3.Npad, 256);
    const int slots = pairs ? 74 : 148;
    int want = 1;
    double best = 1e30;
    for (int ns = 1; ns < MAX_NSPLIT; ++ns) {
      const double cost = (double)cdiv((int)(per_slice * ns), slots) * cdiv(nkb, ns) + 0.25 * ns;
      if (cost < best - 1e-9) { best = cost; want = ns; }
    }
    n.nsplit3 = conv_split_count(n.c3, taps_1x1(), want);
  }
  n.bias3 = f->pool.alloc<float>(N3);
  IPK_CUDA(cudaMemcpyAsync(n.bias3, b3.p, N3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  f->K1pad_max = std::max(f->K1pad_max, n.K1pad);
  f->Npad3_max = std::max(f->Npad3_max, n.c3.Npad);
  f->nices.push_back(n);
  return (int)f->nices.size() - 1;
}

static const McfPacked& build_mcf(ipk_flow* f, const std::string& p, int C, int order, cudaStream_t st) {
  auto it = f->mcfs.find(p);
  if (it != f->mcfs.end()) return it->second;
  IPK_CHECK(C <= 96, IPK_ERR_UNSUPPORTED, "flow: MCF with more than 96 channels is not supported (got %d)", C);
  McfPacked m;
  m.C = C; m.Cp = round_up(C, 4); m.hid = 4 * C;   // macow2.py:36-40
  const int kh = (order < 2) ? f->cfg.kernel_h : f->cfg.kernel_w;
  const int kw = (order < 2) ? f->cfg.kernel_w : f->cfg.kernel_h;
  const int hch = f->hch, row = m.hid + hch, C2 = 2 * C;
  const TensorRef& ws = need(f, p + "net.shift_conv.weight", (int64_t)m.hid * C * kh * kw, IPK_F32);
  const TensorRef& v = need(f, p + "net.conv1x1.conv.weight_v", (int64_t)C2 * row, IPK_F32);
  const TensorRef& g = need(f, p + "net.conv1x1.conv.weight_g", C2, IPK_F32);
  const TensorRef& b = need(f, p + "net.conv1x1.conv.bias", C2, IPK_F32);
  float* os = f->pool.alloc<float>(C2);
  weight_norm_scale((const float*)v.p, (const float*)g.p, os, C2, row, st);
  if (f->cfg.precision != IPK_PREC_FP32_SIMT && C <= MCF_MMA_MAXC) {
    // tensor-core precisions: mma.sync bf16x3 fragments (flow_segment.cu: mcf_mma)
    uint32_t* wa = f->pool.alloc<uint32_t>(mcf_mma_conv_words(C));
    uint32_t* w1 = f->pool.alloc<uint32_t>(mcf_mma_1x1_words(C));
    pack_mcf_mma((const float*)ws.p, (const float*)v.p, os, wa, w1, m.hid, C, kh, kw, order, row, st);
    m.Wc = (float*)wa;
    m.W1x = (float*)w1;
  } else {
    m.Wc = f->pool.alloc<float>((size_t)6 * m.Cp * m.hid);
    pack_mcf_shift((const float*)ws.p, m.Wc, m.hid, C, m.Cp, kh, kw, order, st);
    m.W1x = f->pool.alloc<float>((size_t)m.hid * C2);
    pack_rows4((const float*)v.p, os, m.W1x, C2, row, 0, m.hid, st);
  }
  // the x-independent part of the 1x1 (columns hid.. of weight_v, macow_utils.py:429-432) goes into the shared Hterm GEMM
  m.hcol = f->hterm_cols;
  f->hterm_cols += round_up(C2, 4);
  m.v_src = (const float*)v.p; m.os = os; m.b_src = (const float*)b.p;
  f->mcfs[p] = m;
  return f->mcfs[p];
}

static int* build_shuffle(ipk_flow* f, const std::string& p, int C, bool fwd, cudaStream_t st) {
  std::string key = p + (fwd ? "f" : "b");
  auto it = f->shuffles.find(key);
  if (it != f->shuffles.end()) return it->second;
  const TensorRef& t = need(f, p + (fwd ? "forward_shuffle_idx" : "backward_shuffle_idx"), C, IPK_I64);
  int* d = f->pool.alloc<int>(C);
  i64_to_i32((const long long*)t.p, d, C, st);
  f->shuffles[key] = d;
  return d;
}

// ---- logical programs (exact mirror of the reference's module order) ----
static void unit_ops(std::vector<LogicalOp>& v, const std::string& p, int C, bool fwd) {
  // MaCowUnit.forward (macow2.py:962-995)
  std::vector<LogicalOp> u;
  auto mcf = [&](const char* n, int order) { LogicalOp o; o.kind = L_MCF; o.C = C; o.prefix = p + n; o.order = order; u.push_back(o); };
  auto act = [&](const char* n) { LogicalOp o; o.kind = L_ACTNORM; o.C = C; o.prefix = p + n; o.coff = 0; o.cnt = C; u.push_back(o); };
  mcf("conv1.", 0); mcf("conv2.", 1); act("actnorm1."); mcf("conv3.", 2); mcf("conv4.", 3); act("actnorm2.");
  if (!fwd) std::reverse(u.begin(), u.end());
  v.insert(v.end(), u.begin(), u.end());
}

std::vector<LogicalOp> logical_program(const ipk_flow_config& cfg, bool fwd) {
  std::vector<LogicalOp> prog;
  auto lv = levels_of(cfg);
  auto add_level = [&](const LevelInfo& li) {
    std::vector<LogicalOp> ops;  // forward order; reversed afterwards for the inverse
    const int C = li.C;
    auto act = [&](const std::string& p, int coff, int cnt) { LogicalOp o; o.kind = L_ACTNORM; o.C = C; o.prefix = p; o.coff = coff; o.cnt = cnt; ops.push_back(o); };
    auto shf = [&](const std::string& p) { LogicalOp o; o.kind = L_SHUFFLE; o.C = C; o.prefix = p; o.fwd_idx = fwd; ops.push_back(o); };
    auto nice = [&](const std::string& p, int factor, bool skip, bool up) {
      LogicalOp o; o.kind = L_NICE; o.C = C; o.prefix = p; o.factor = factor; o.skip = skip; o.up = up; ops.push_back(o);
    };
    for (int s = 0; s < li.steps; ++s) {
      // MaCowStep.forward (macow2.py:1066-1091)
      std::string p = "flow.layers." + std::to_string(li.L) + "." + std::to_string(s) + ".";
      act(p + "actnorm1.", 0, C);
      shf(p + "conv1x1.");
      for (int u = 0; u < 2; ++u) { std::vector<LogicalOp> t; unit_ops(t, p + "units1." + std::to_string(u) + ".", C, true); ops.insert(ops.end(), t.begin(), t.end()); }
      nice(p + "coupling1_up.", 2, false, true);
      nice(p + "coupling1_dn.", 2, false, false);
      act(p + "actnorm2.", 0, C);
      for (int u = 0; u < 2; ++u) { std::vector<LogicalOp> t; unit_ops(t, p + "units2." + std::to_string(u) + ".", C, true); ops.insert(ops.end(), t.begin(), t.end()); }
      nice(p + "coupling2_up.", 2, true, true);
      nice(p + "coupling2_dn.", 2, true, false);
    }
    // MultiScalePrior.forward (macow2.py:569-581) then the level shuffle (macow2.py:889)
    std::string pp = "flow.priors." + std::to_string(li.L) + ".";
    shf(pp + "conv1x1.");
    nice(pp + "coupling.", li.prior_factor, false, true);
    act(pp + "actnorm.", li.z1, li.prior_out);
    shf("flow.shuffle_layers." + std::to_string(li.L) + ".");
    if (!fwd) std::reverse(ops.begin(), ops.end());
    prog.insert(prog.end(), ops.begin(), ops.end());
  };
  if (fwd) for (size_t i = 0; i < lv.size(); ++i) add_level(lv[i]);
  else for (size_t i = lv.size(); i-- > 0;) add_level(lv[i]);
  return prog;
}

static std::vector<Stage> compile_program(ipk_flow* f, const std::vector<LogicalOp>& prog, cudaStream_t st) {
  std::vector<Stage> stages;
  Stage cur;
  bool open = false;
  auto flush = [&]() {
    if (open && (!cur.host_ops.empty())) stages.push_back(cur);
    cur = Stage();
    open = false;
  };
  auto ensure = [&](int C) {
    if (open && cur.C != C) flush();
    if (!open) { cur = Stage(); cur.C = C; open = true; }
  };
  const int mode = f->act_mode;
  for (const LogicalOp& o : prog) {
    ensure(o.C);
    MicroOp m;
    memset(&m, 0, sizeof(m));
    switch (o.kind) {
      case L_ACTNORM: {
        const TensorRef& ls = need(f, o.prefix + "log_scale", o.cnt, IPK_F32);
        const TensorRef& b = need(f, o.prefix + "bias", o.cnt, IPK_F32);
        float* d = f->pool.alloc<float>(2 * (size_t)o.cnt);
        IPK_CUDA(cudaMemcpyAsync(d, ls.p, o.cnt * sizeof(float), cudaMemcpyDeviceToDevice, st));
        IPK_CUDA(cudaMemcpyAsync(d + o.cnt, b.p, o.cnt * sizeof(float), cudaMemcpyDeviceToDevice, st));
        m.kind = MK_ACTNORM; m.i0 = o.coff; m.i1 = o.cnt; m.p0 = d; m.p1 = d + o.cnt;
        cur.host_ops.push_back(m);
        break;
      }
      case L_SHUFFLE: {
        m.kind = MK_SHUFFLE; m.i0 = o.C; m.idx = build_shuffle(f, o.prefix, o.C, o.fwd_idx, st);
        cur.host_ops.push_back(m);
        break;
      }
      case L_MCF: {
        const McfPacked& p = build_mcf(f, o.prefix, o.C, o.order, st);
        m.kind = MK_MCF; m.i0 = o.order; m.i1 = p.C; m.i2 = p.Cp; m.i3 = p.hid;
        m.p0 = p.Wc; m.p1 = p.W1x; m.l0 = p.hcol;   // p2 / l0 patched after the Hterm buffer exists
        cur.host_ops.push_back(m);
        cur.has_mcf = true;
        break;
      }
      case L_NICE: {
        int id;
        auto it = f->nice_by_prefix.find(o.prefix);
        if (it == f->nice_by_prefix.end()) {
          id = build_nice(f, o.prefix, o.C, o.factor, o.skip, o.up, st);
          f->nice_by_prefix[o.prefix] = id;
        } else id = it->second;
        const NiceLayer& n = f->nices[id];
        // end the current segment with the operand build of this coupling ...
        m.kind = MK_IM2COL; m.i0 = n.n_z; m.i1 = n.K1pad; m.i2 = mode; m.idx = n.d_iz;   // out0/out1 patched after workspace alloc
        cur.host_ops.push_back(m);
        cur.nice_id = id;
        int C = cur.C;
        flush();
        // ... and open the next one with its affine update
        ensure(C);
        MicroOp a;
        memset(&a, 0, sizeof(a));
        a.kind = MK_AFFINE; a.i0 = n.nsplit3; a.i1 = n.c3.Npad; a.i2 = n.n_p; a.i3 = n.N3p; a.p1 = n.bias3; a.idx = n.d_ip;   // p0 / l0 patched later
        cur.host_ops.push_back(a);
        break;
      }
    }
  }
  flush();
  return stages;
}

static void upload_programs(ipk_flow* f, std::vector<Stage>& stages, cudaStream_t st) {
  const long long Mmax = (long long)f->cfg.max_batch * 64;
  for (Stage& s : stages) {
    for (MicroOp& m : s.host_ops) {
      if (m.kind == MK_IM2COL) { m.out0 = f->A1; m.out1 = f->A1_lo; }
      if (m.kind == MK_AFFINE) { m.p0 = f->partials; m.l0 = Mmax * f->Npad3_max; }
      if (m.kind == MK_MCF) { m.p2 = f->Hterm + m.l0; m.l0 = f->hstride; }
    }
    s.d_ops = f->pool.alloc<MicroOp>(s.host_ops.size());
    IPK_CUDA(cudaMemcpyAsync(s.d_ops, s.host_ops.data(), s.host_ops.size() * sizeof(MicroOp), cudaMemcpyHostToDevice, st));
  }
  IPK_CUDA(cudaStreamSynchronize(st));
}

// element size of one operand plane row entry (fp32 rows, or bf16 planes)
static inline size_t act_esz(const ipk_flow* f) { return f->act_mode == OUT_F32_NHWC ? 4 : 2; }
static inline void* poff(void* p, size_t bytes) { return p ? (char*)p + bytes : nullptr; }

// the coupling network of samples [b0, b0 + nb): every workspace buffer is indexed by pixel row, so a batch slice is a
// pointer offset
static void run_nice_net(ipk_flow* f, const NiceLayer& n, int b0, int nb, cudaStream_t st) {
  const int M = nb * 64, Hd = f->Hd;
  const size_t r0 = (size_t)b0 * 64, es = act_esz(f);
  const long long Mmax = (long long)f->cfg.max_batch * 64;
  // conv1 (im2col GEMM) + ELU
  {
    ConvIn in; in.p = poff(f->A1, r0 * n.K1pad * es); in.p_lo = poff(f->A1_lo, r0 * n.K1pad * es); in.cstride = n.K1pad; in.F = M; in.H = 1; in.W = 1;
    ConvOut out; out.p = poff(f->H1, r0 * Hd * es); out.p_lo = poff(f->H1_lo, r0 * Hd * es); out.mode = f->act_mode; out.cstride = Hd; out.Ho = 1; out.Wo = 1; out.act = ACT_ELU;
    ProfScope ps("flow.nice.conv1", st);
    conv_run(n.c1, in, out, taps_1x1(), 1, st);
  }
  // conv2 (1x1) + ELU
  {
    ConvIn in; in.p = poff(f->H1, r0 * Hd * es); in.p_lo = poff(f->H1_lo, r0 * Hd * es); in.cstride = Hd; in.F = M; in.H = 1; in.W = 1;
    ConvOut out; out.p = poff(f->H2, r0 * Hd * es); out.p_lo = poff(f->H2_lo, r0 * Hd * es); out.mode = f->act_mode; out.cstride = Hd; out.Ho = 1; out.Wo = 1; out.act = ACT_ELU;
    ProfScope ps("flow.nice.conv2", st);
    conv_run(n.c2, in, out, taps_1x1(), 1, st);
  }
  // conv3: all nine tap responses in one GEMM, split-K into fp32 partial slices; bias, the 3x3 gather and the affine
  // transform happen in the next segment
  {
    ConvIn in; in.p = poff(f->H2, r0 * Hd * es); in.p_lo = poff(f->H2_lo, r0 * Hd * es); in.cstride = Hd; in.F = M; in.H = 1; in.W = 1;
    ConvOut out; out.p = f->partials + r0 * n.c3.Npad; out.mode = OUT_F32_NHWC; out.cstride = n.c3.Npad; out.Ho = 1; out.Wo = 1;
    out.split_stride = Mmax * f->Npad3_max;
    ProfScope ps("flow.nice.conv3", st);
    const int ns = conv_run(n.c3, in, out, taps_1x1(), n.nsplit3, st);
    IPK_CHECK(ns == n.nsplit3, IPK_ERR_STATE, "flow: conv3 split-K produced %d slices, plan expected %d", ns, n.nsplit3);
  }
}

// Hterm[B*64][hstride] = bias + W1h_all * ELU(cond): the conditioning input of every MCF's 1x1 (MCFBlock.forward,
// macow_utils.py:429-432: cat -> ELU -> 1x1 splits into a state part and this state-independent part)
static void run_hterm(ipk_flow* f, int b0, int nb, cudaStream_t st) {
  if (f->hterm_cols == 0) return;
  ProfScope ps("flow.hterm_gemm", st);
  const long long M = (long long)nb * 64;
  const size_t r0 = (size_t)b0 * 64;
  const bool simt = f->hterm_w.engine == IPK_PREC_FP32_SIMT;
  const size_t es = simt ? 4 : 2;
  NormApply e; e.x = f->cond + r0 * f->hch; e.F = 1; e.P = M; e.C = f->hch; e.act = ACT_ELU;
  if (simt) e.out_f32 = (float*)poff(f->E, r0 * f->hch * es);
  else { e.out_hi = (__nv_bfloat16*)poff(f->E, r0 * f->hch * es); e.out_lo = (__nv_bfloat16*)poff(f->E_lo, r0 * f->hch * es); }
  norm_apply(e, st);
  ConvIn in; in.p = poff(f->E, r0 * f->hch * es); in.p_lo = poff(f->E_lo, r0 * f->hch * es); in.cstride = f->hch; in.F = (int)M; in.H = 1; in.W = 1;
  ConvOut out; out.p = f->Hterm + r0 * f->hstride; out.mode = OUT_F32_NHWC; out.cstride = f->hstride; out.Ho = 1; out.Wo = 1; out.bias = f->hterm_w.bias;
  conv_run(f->hterm_w, in, out, taps_1x1(), 1, st);
}

static void run_program_slice(ipk_flow* f, std::vector<Stage>& stages, bool fwd, int b0, int nb, cudaStream_t st) {
  run_hterm(f, b0, nb, st);
  for (Stage& s : stages) {
    SegmentLaunch sl{s.d_ops, (int)s.host_ops.size(), s.C, s.has_mcf, f->cfg.precision != IPK_PREC_FP32_SIMT};
    {
      ProfScope ps(s.has_mcf ? "flow.segment.mcf" : "flow.segment.light", st);
      flow_segment_run(sl, fwd, f->state, f->C0, f->logdet_ws, b0, nb, st);
    }
    if (s.nice_id >= 0) run_nice_net(f, f->nices[s.nice_id], b0, nb, st);
  }
}

// Samples are independent, and the flow alternates between kernels that fill the machine (coupling GEMMs) and kernels that
// cannot (one CTA per sample on the 6 400-step MCF chain).  IPK_FLOW_DUAL_STREAM=1 runs large batches as two half-batches on
// two streams so one half's GEMMs can take the SMs the other half's MCF lines leave idle.  It is OFF by default: the MCF
// chain is latency- not throughput-bound, so every half still pays the full chain, and the persistent 225 KB-smem GEMM CTAs
// do not co-reside with segment CTAs.
static void run_program(ipk_flow* f, std::vector<Stage>& stages, bool fwd, int B, cudaStream_t st) {
  const bool split = f->dual_stream && B >= 16 && !Prof::enabled();
  if (!split) {
    run_program_slice(f, stages, fwd, 0, B, st);
    return;
  }
  const int nbA = (B + 1) / 2;
  IPK_CUDA(cudaEventRecord(f->ev_fork, st));
  IPK_CUDA(cudaStreamWaitEvent(f->st2, f->ev_fork, 0));
  run_program_slice(f, stages, fwd, 0, nbA, st);
  run_program_slice(f, stages, fwd, nbA, B - nbA, f->st2);
  IPK_CUDA(cudaEventRecord(f->ev_join, f->st2));
  IPK_CUDA(cudaStreamWaitEvent(st, f->ev_join, 0));
}

}  // namespace ipk

// ----------------------------------------------------------------------------------------------- C ABI
extern "C" int ipk_flow_create(const ipk_flow_config* cfg, ipk_flow** out) {
  IPK_TRY
  IPK_CHECK(cfg && out, IPK_ERR_INVALID, "ipk_flow_create: null argument");
  IPK_CHECK(cfg->n_levels > 0 && cfg->n_levels <= IPK_MAX_LEVELS, IPK_ERR_INVALID, "flow: bad n_levels %d", cfg->n_levels);
  IPK_CHECK(cfg->n_levels < cfg->factor, IPK_ERR_INVALID, "flow: num_layers must be < factor (macow2.py:834)");
  IPK_CHECK(cfg->kernel_h == 2 && cfg->kernel_w == 3, IPK_ERR_UNSUPPORTED, "flow: only kernel_size (2,3) is supported");
  IPK_CHECK(cfg->flow_in_channels >= cfg->factor && cfg->flow_in_channels <= 96, IPK_ERR_UNSUPPORTED, "flow: flow_in_channels out of range");
  IPK_CHECK(cfg->h_channels % 4 == 0 && cfg->h_channels > 0, IPK_ERR_UNSUPPORTED, "flow: h_channels must be a positive multiple of 4");
  IPK_CHECK(cfg->precision >= 0 && cfg->precision <= 2, IPK_ERR_INVALID, "flow: bad precision");
  IPK_CHECK(cfg->max_batch > 0, IPK_ERR_INVALID, "flow: max_batch must be positive");
  if (cfg->precision != IPK_PREC_FP32_SIMT)
    IPK_CHECK(cfg->h_channels % 8 == 0, IPK_ERR_UNSUPPORTED, "flow: tensor-core engine needs h_channels %% 8 == 0");
  if (cfg->precision != IPK_PREC_FP32_SIMT)
    IPK_CHECK(cfg->flow_mid_channels % 64 == 0, IPK_ERR_UNSUPPORTED, "flow: tensor-core engine needs flow_mid_channels %% 64 == 0");
  levels_of(*cfg);
  ipk_flow* f = new ipk_flow();
  f->cfg = *cfg;
  f->C0 = cfg->flow_in_channels; f->Hd = cfg->flow_mid_channels; f->hch = cfg->h_channels;
  f->act_mode = cfg->precision == IPK_PREC_FP32_SIMT ? OUT_F32_NHWC : (cfg->precision == IPK_PREC_FP32_SPLIT ? OUT_BF16_SPLIT : OUT_BF16);
  *out = f;
  IPK_CATCH
}

extern "C" int ipk_flow_set_tensor(ipk_flow* f, const char* name, const void* dev_ptr, int64_t numel, int dtype) {
  IPK_TRY
  IPK_CHECK(f && name && dev_ptr, IPK_ERR_INVALID, "ipk_flow_set_tensor: null argument");
  IPK_CHECK(!f->finalized, IPK_ERR_STATE, "ipk_flow_set_tensor after finalize");
  f->tensors[name] = TensorRef{dev_ptr, numel, dtype};
  IPK_CATCH
}

// ---------------------------------------------------------------------------------- data-dependent initialisation
// First forward of a freshly constructed flow (every `initialized` buffer 0): ActNorm2dFlow.init (macow2.py:503-505,526-539)
// and Conv2dWeightNorm.init (macow_utils.py:231-250).  Both weight-normed convs of the flow are built with zero_init=True
// (macow_utils.py:281,423), so their init is g <- init_scale / (std + 1e-6) = 0 and bias <- -mean * 0 = 0: every MCF and every
// coupling is the identity during (and right after) the init pass, and the pass reduces to the ActNorms and Shuffles applied in
// forward order, each ActNorm taking its statistics from the state that reaches it:
//     out = x * exp(ls0) + b0;  mean, unbiased std over (B, H, W) per channel;  ls <- log(1 / (std + 1e-6));  b <- -mean / (std + 1e-6)
// One CTA walks the whole op list on the [B*64][C0] state in global memory (one-off, ~600 ops on <= 1 MB of state).
namespace ipk {
struct InitOp { int kind, C, coff, cnt; float* ls; float* bias; const long long* idx; };

__global__ void __launch_bounds__(1024, 1)
flow_data_init_kernel(float* __restrict__ S, int C0, int M, const InitOp* __restrict__ ops, int nops) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int i = 0; i < nops; ++i) {
    const InitOp op = ops[i];
    if (op.kind == L_ACTNORM) {
      for (int c = warp; c < op.cnt; c += nwarps) {
        const float e0 = expf(op.ls[c]), b0 = op.bias[c];
        double s = 0.0;
        for (int m = lane; m < M; m += 32) s += (double)(S[(size_t)m * C0 + op.coff + c] * e0 + b0);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const double mean = s / (double)M;
        double q = 0.0;
        for (int m = lane; m < M; m += 32) {
          const double d = (double)(S[(size_t)m * C0 + op.coff + c] * e0 + b0) - mean;
          q += d * d;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float stdv = (float)sqrt(q / (double)(M > 1 ? M - 1 : 1));     // torch.std: unbiased
        const float inv = 1.0f / (stdv + 1e-6f);
        const float ls = logf(inv), b = -(float)mean * inv;
        const float e1 = expf(ls);
        for (int m = lane; m < M; m += 32) {
          const size_t o = (size_t)m * C0 + op.coff + c;
          S[o] = S[o] * e1 + b;                                               // forward with the fresh parameters (macow2.py:511)
        }
        if (lane == 0) { op.ls[c] = ls; op.bias[c] = b; }
      }
    } else if (op.kind == L_SHUFFLE) {
      float tmp[96];
      for (int m = threadIdx.x; m < M; m += blockDim.x) {
        float* row = S + (size_t)m * C0;
        for (int c = 0; c < op.C; ++c) tmp[c] = row[c];
        for (int c = 0; c < op.C; ++c) row[c] = tmp[(int)op.idx[c]];
      }
    }
    __syncthreads();
  }
}
}  // namespace ipk

// Runs on a created, NOT yet finalized plan whose tensors were registered with ipk_flow_set_tensor: the registered ActNorm
// log_scale / bias and weight-norm weight_g / bias tensors are OVERWRITTEN in place (they are the caller's parameters); the caller
// then sets its `initialized` buffers to 1 and finalizes the plan.
extern "C" int ipk_flow_data_init(ipk_flow* f, const float* x, int32_t B, void* stream) {
  IPK_TRY
  IPK_CHECK(f && x, IPK_ERR_INVALID, "ipk_flow_data_init: null argument");
  IPK_CHECK(!f->finalized, IPK_ERR_STATE, "ipk_flow_data_init: call before ipk_flow_finalize");
  IPK_CHECK(B > 1, IPK_ERR_INVALID, "ipk_flow_data_init: the unbiased standard deviation needs more than one sample");
  IPK_CHECK(f->C0 <= 96, IPK_ERR_UNSUPPORTED, "ipk_flow_data_init: more than 96 channels");
  cudaStream_t st = (cudaStream_t)stream;
  const int M = B * 64;
  std::vector<InitOp> ops;
  for (const LogicalOp& o : logical_program(f->cfg, true)) {
    InitOp io;
    memset(&io, 0, sizeof(io));
    io.kind = o.kind; io.C = o.C;
    if (o.kind == L_ACTNORM) {
      io.coff = o.coff; io.cnt = o.cnt;
      io.ls = (float*)const_cast<void*>(need(f, o.prefix + "log_scale", o.cnt, IPK_F32).p);
      io.bias = (float*)const_cast<void*>(need(f, o.prefix + "bias", o.cnt, IPK_F32).p);
      ops.push_back(io);
    } else if (o.kind == L_SHUFFLE) {
      io.idx = (const long long*)need(f, o.prefix + "forward_shuffle_idx", o.C, IPK_I64).p;
      ops.push_back(io);
    } else {
      // zero_init weight-normed conv of this MCF / coupling: g <- 0, bias <- 0
      const std::string p = o.prefix + (o.kind == L_MCF ? "net.conv1x1.conv." : "net.conv3.conv.");
      int n2;
      if (o.kind == L_MCF) n2 = 2 * o.C;
      else {
        std::vector<int> iz, ip;
        nice_indices(o.C, o.factor, o.skip, o.up, iz, ip);
        n2 = 2 * (int)ip.size();
      }
      IPK_CUDA(cudaMemsetAsync(const_cast<void*>(need(f, p + "weight_g", n2, IPK_F32).p), 0, n2 * sizeof(float), st));
      IPK_CUDA(cudaMemsetAsync(const_cast<void*>(need(f, p + "bias", n2, IPK_F32).p), 0, n2 * sizeof(float), st));
    }
  }
  float* S = nullptr;
  InitOp* d_ops = nullptr;
  IPK_CUDA(cudaMalloc((void**)&S, (size_t)M * f->C0 * sizeof(float)));
  IPK_CUDA(cudaMalloc((void**)&d_ops, ops.size() * sizeof(InitOp)));
  IPK_CUDA(cudaMemcpyAsync(d_ops, ops.data(), ops.size() * sizeof(InitOp), cudaMemcpyHostToDevice, st));
  nchw_to_nhwc(x, S, B, f->C0, 64, f->C0, st);
  flow_data_init_kernel<<<1, 1024, 0, st>>>(S, f->C0, M, d_ops, (int)ops.size());
  IPK_LAUNCH_CHECK();
  IPK_CUDA(cudaStreamSynchronize(st));
  cudaFree(S);
  cudaFree(d_ops);
  IPK_CATCH
}

extern "C" int ipk_flow_finalize(ipk_flow* f, void* stream) {
  IPK_TRY
  IPK_CHECK(f, IPK_ERR_INVALID, "null flow");
  IPK_CHECK(!f->finalized, IPK_ERR_STATE, "flow already finalized");
  cudaStream_t st = (cudaStream_t)stream;
  flow_segment_init();
  auto pf = logical_program(f->cfg, true);
  auto pi = logical_program(f->cfg, false);
  f->prog_fwd = compile_program(f, pf, st);
  f->prog_inv = compile_program(f, pi, st);
  // one GEMM for the conditioning terms of all MCFs (always error-compensated on the tensor-core engines: K = h_channels only)
  {
    const int heng = f->cfg.precision == IPK_PREC_FP32_SIMT ? IPK_PREC_FP32_SIMT : IPK_PREC_FP32_SPLIT;
    f->hterm_w = conv_alloc(f->pool, heng, 1, f->hch, std::max(f->hterm_cols, 4), true);
    for (auto& kv : f->mcfs) {
      const McfPacked& m = kv.second;
      PackSrc s;
      s.w = m.v_src; s.N = 2 * m.C; s.Ksrc = m.hid + f->hch; s.k_off = m.hid; s.oscale = m.os;
      conv_pack_into(f->hterm_w, m.hcol, s, {0}, st);
      conv_pack_bias(f->hterm_w, m.hcol, m.b_src, 2 * m.C, 0.f, st);
    }
    f->hstride = f->hterm_w.Npad;
  }
  // workspace
  const size_t Mmax = (size_t)f->cfg.max_batch * 64;
  const size_t esz = 4;  // fp32, or two bf16 planes
  size_t bytes = 0;
  auto rb = [](size_t b) { return (b + 255) / 256 * 256; };
  bytes += rb(Mmax * f->C0 * 4) + rb(Mmax * f->hch * 4) + rb(Mmax * f->K1pad_max * esz) + 2 * rb(Mmax * f->Hd * esz) +
           rb((size_t)MAX_NSPLIT * Mmax * f->Npad3_max * 4) + rb(f->cfg.max_batch * 4) + rb(Mmax * f->hch * 4) + rb(Mmax * (size_t)f->hstride * 4) + 4096;
  f->ws.init(bytes);
  f->state = f->ws.alloc<float>(Mmax * f->C0);
  f->cond = f->ws.alloc<float>(Mmax * f->hch);
  char* a1 = (char*)f->ws.alloc<float>(Mmax * f->K1pad_max);
  char* h1 = (char*)f->ws.alloc<float>(Mmax * f->Hd);
  char* h2 = (char*)f->ws.alloc<float>(Mmax * f->Hd);
  f->A1 = a1; f->H1 = h1; f->H2 = h2;
  if (f->act_mode == OUT_BF16_SPLIT) {
    f->A1_lo = a1 + Mmax * f->K1pad_max * 2;
    f->H1_lo = h1 + Mmax * f->Hd * 2;
    f->H2_lo = h2 + Mmax * f->Hd * 2;
  }
  f->partials = f->ws.alloc<float>((size_t)MAX_NSPLIT * Mmax * f->Npad3_max);
  f->logdet_ws = f->ws.alloc<float>(f->cfg.max_batch);
  {
    char* e = (char*)f->ws.alloc<float>(Mmax * f->hch);
    f->E = e;
    f->E_lo = f->hterm_w.engine == IPK_PREC_FP32_SPLIT ? e + Mmax * f->hch * 2 : nullptr;
    f->Hterm = f->ws.alloc<float>(Mmax * (size_t)f->hstride);
  }
  upload_programs(f, f->prog_fwd, st);
  upload_programs(f, f->prog_inv, st);
  IPK_CUDA(cudaStreamSynchronize(st));
  {
    const char* e = getenv("IPK_FLOW_DUAL_STREAM");
    f->dual_stream = e && e[0] == '1';     // measured on B200 at B=64: 85.6 ms/step with, 82.2 without -> off by default
    IPK_CUDA(cudaStreamCreateWithFlags(&f->st2, cudaStreamNonBlocking));
    IPK_CUDA(cudaEventCreateWithFlags(&f->ev_fork, cudaEventDisableTiming));
    IPK_CUDA(cudaEventCreateWithFlags(&f->ev_join, cudaEventDisableTiming));
  }
  f->tensors.clear();
  f->finalized = true;
  IPK_CATCH
}

static void flow_run(ipk_flow* f, bool fwd, const float* in, const float* cond, float* out, float* logdet, int B, cudaStream_t st) {
  IPK_CHECK(f && f->finalized, IPK_ERR_STATE, "flow not finalized");
  IPK_CHECK(B > 0 && B <= f->cfg.max_batch, IPK_ERR_INVALID, "flow: batch %d outside (0, max_batch=%d]", B, f->cfg.max_batch);
  IPK_CHECK(in && cond && out, IPK_ERR_INVALID, "flow: null buffer");
  nchw_to_nhwc(in, f->state, B, f->C0, 64, f->C0, st);
  nchw_to_nhwc(cond, f->cond, B, f->hch, 64, f->hch, st);
  if (fwd) IPK_CUDA(cudaMemsetAsync(f->logdet_ws, 0, B * sizeof(float), st));
  run_program(f, fwd ? f->prog_fwd : f->prog_inv, fwd, B, st);
  nhwc_to_nchw(f->state, out, B, f->C0, 64, f->C0, st);
  if (fwd && logdet) IPK_CUDA(cudaMemcpyAsync(logdet, f->logdet_ws, B * sizeof(float), cudaMemcpyDeviceToDevice, st));
}

extern "C" int ipk_flow_reverse(ipk_flow* f, const float* z, const float* cond, float* out, int32_t B, void* stream) {
  IPK_TRY
  flow_run(f, false, z, cond, out, nullptr, B, (cudaStream_t)stream);
  IPK_CATCH
}

extern "C" int ipk_flow_forward(ipk_flow* f, const float* x, const float* cond, float* z, float* logdet, int32_t B, void* stream) {
  IPK_TRY
  flow_run(f, true, x, cond, z, logdet, B, (cudaStream_t)stream);
  IPK_CATCH
}

extern "C" int ipk_flow_destroy(ipk_flow* f) {
  if (!f) return IPK_OK;
  ipk_graphs_drop(f);
  f->pool.release();
  f->ws.release();
  if (f->st2) cudaStreamDestroy(f->st2);
  if (f->ev_fork) cudaEventDestroy(f->ev_fork);
  if (f->ev_join) cudaEventDestroy(f->ev_join);
  delete f;
  return IPK_OK;
}

// internal (capi.cu): reverse pass that leaves the result in the NHWC state buffer and returns its address
int ipk_flow_reverse_nhwc(ipk_flow* f, const float* z, const float* cond, const float** state_nhwc, int B, cudaStream_t st) {
  IPK_CHECK(f && f->finalized, IPK_ERR_STATE, "flow not finalized");
  IPK_CHECK(B > 0 && B <= f->cfg.max_batch, IPK_ERR_INVALID, "flow: batch %d outside (0, max_batch=%d]", B, f->cfg.max_batch);
  nchw_to_nhwc(z, f->state, B, f->C0, 64, f->C0, st);
  nchw_to_nhwc(cond, f->cond, B, f->hch, 64, f->hch, st);
  run_program(f, f->prog_inv, false, B, st);
  *state_nhwc = f->state;
  return 0;
}
struct FlowDims { int C0, hch; };
FlowDims ipk_flow_dims(ipk_flow* f) { return FlowDims{f->C0, f->hch}; }
