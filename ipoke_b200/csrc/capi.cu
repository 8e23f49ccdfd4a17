// C ABI glue: error state, launch counter, whole-step entry points and kernel-level test hooks.
#include <cstdarg>
#include <mutex>
#include <map>
#include <tuple>
#include <nvtx3/nvToolsExt.h>
#include "conv.cuh"
#include "conv3d_tc.cuh"
#include "elementwise.cuh"

namespace ipk {

bool nvtx_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("IPK_NVTX");
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on == 1;
}
void nvtx_push(const char* tag) { nvtxRangePushA(tag); }
void nvtx_pop() { nvtxRangePop(); }

thread_local std::string g_last_error;
thread_local int64_t g_launches = 0;

void set_last_error(const std::string& m) { g_last_error = m; }

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("IPK_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

[[noreturn]] void fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw Error(code, buf);
}

// ---- profiler
namespace {
struct ProfRec { std::string tag; cudaEvent_t e0, e1; };
std::vector<ProfRec> g_prof;
std::vector<size_t> g_prof_open;
bool g_prof_on = false;
}
bool& Prof::enabled() { return g_prof_on; }
void Prof::begin(const char* tag, cudaStream_t st) {
  ProfRec r;
  r.tag = tag;
  cudaEventCreate(&r.e0);
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, st);
  g_prof.push_back(r);
  g_prof_open.push_back(g_prof.size() - 1);
}
void Prof::end(cudaStream_t st) {
  if (g_prof_open.empty()) return;
  cudaEventRecord(g_prof[g_prof_open.back()].e1, st);
  g_prof_open.pop_back();
}

}  // namespace ipk

using namespace ipk;

extern "C" void ipk_prof_enable(int on) {
  for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  g_prof.clear();
  g_prof_open.clear();
  g_prof_on = on != 0;
}
// writes "tag count total_ms\n" lines into buf; returns the number of bytes needed
extern "C" int ipk_prof_report(char* buf, int cap) {
  cudaDeviceSynchronize();
  std::map<std::string, std::pair<int, double>> acc;
  std::vector<std::string> order;
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) { cudaGetLastError(); continue; }
    if (!acc.count(r.tag)) order.push_back(r.tag);
    acc[r.tag].first += 1;
    acc[r.tag].second += ms;
  }
  std::string out;
  char line[256];
  for (auto& t : order) {
    snprintf(line, sizeof(line), "%s %d %.6f\n", t.c_str(), acc[t].first, acc[t].second);
    out += line;
  }
  if (buf && cap > 0) { snprintf(buf, cap, "%s", out.c_str()); }
  return (int)out.size() + 1;
}

// defined in flow.cu / decoder.cu
struct ipk_flow;
struct ipk_fs;
int ipk_fs_decode_nhwc(ipk_fs* d, const float* motion_nhwc, const float* x0, float* frames, int B, int T, cudaStream_t st,
                       float* frames_host = nullptr, cudaStream_t copy_st = nullptr, uint8_t* u8_dev = nullptr, uint8_t* u8_host = nullptr);
int ipk_flow_reverse_nhwc(ipk_flow* f, const float* z, const float* cond, const float** state_nhwc, int B, cudaStream_t st);

extern "C" int ipk_version(void) { return IPK_VERSION; }
extern "C" const char* ipk_last_error(void) { return g_last_error.c_str(); }
extern "C" int64_t ipk_launch_count(void) { return g_launches; }
extern "C" void ipk_launch_count_reset(void) { g_launches = 0; }

// ---- CUDA-graph replay of the whole sampling step.  A step is ~6 500 dependent launches; replaying them as one graph removes the
// per-launch driver work from the critical path (the programmatic-dependent-launch edges are kept by stream capture).  A graph is
// keyed by (plans, buffers, B, T): the first call with a key runs eagerly (lazy one-time setup: function attributes, tensor maps),
// the second is captured on an internal stream -- torch's default stream is the legacy stream, which cannot capture -- and from
// then on the instantiated graph is launched into the caller's stream.  Opt-in with IPK_GRAPH=1 (the step is 1 298 launches of ~60 us
// each with programmatic dependent launch already hiding the prologues: measured 827.9 vs 813.6 videos/s); never used while profiling.
namespace {
struct GraphKey {
  const void *f, *d, *a0, *a1, *a2, *a3; int B, T, kind;
  bool operator<(const GraphKey& o) const {
    return std::tie(f, d, a0, a1, a2, a3, B, T, kind) < std::tie(o.f, o.d, o.a0, o.a1, o.a2, o.a3, o.B, o.T, o.kind);
  }
};
struct GraphEntry { cudaGraphExec_t exec = nullptr; int64_t launches = 0; int seen = 0; bool failed = false; };
std::map<GraphKey, GraphEntry> g_graphs;
std::mutex g_graphs_mu;
cudaStream_t g_cap_sts[IPK_MAX_DEVICES] = {nullptr};     // capture stream per device
bool graph_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("IPK_GRAPH");
    on = (e && e[0] == '1') ? 1 : 0;     // opt-in: measured +1.7 % on the iper_128 step (1 298 launches of ~60 us each)
  }
  return on == 1 && !Prof::enabled();
}
// runs body(stream) eagerly, or captures / replays it; body must only enqueue work on the stream it is given
template <typename Body>
void run_graphed(const GraphKey& key, cudaStream_t st, Body&& body, bool enabled) {
  if (!enabled || Prof::enabled()) { body(st); return; }
  std::lock_guard<std::mutex> lk(g_graphs_mu);
  if (g_graphs.size() > 32 && !g_graphs.count(key)) {     // buffers keep changing: stop hoarding instantiated graphs
    for (auto& kv : g_graphs)
      if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    g_graphs.clear();
  }
  GraphEntry& e = g_graphs[key];
  if (e.exec) {
    IPK_CUDA(cudaGraphLaunch(e.exec, st));
    g_launches += e.launches;
    return;
  }
  if (e.failed || e.seen++ == 0) { body(st); return; }
  cudaStream_t& g_cap_st = g_cap_sts[current_device_slot()];
  if (!g_cap_st) IPK_CUDA(cudaStreamCreateWithFlags(&g_cap_st, cudaStreamNonBlocking));
  const int64_t before = g_launches;
  cudaGraph_t graph = nullptr;
  IPK_CUDA(cudaStreamBeginCapture(g_cap_st, cudaStreamCaptureModeThreadLocal));
  bool ok = true;
  std::string err;
  try { body(g_cap_st); } catch (const Error& ex) { ok = false; err = ex.what(); }
  cudaError_t ce = cudaStreamEndCapture(g_cap_st, &graph);
  if (ok && ce == cudaSuccess && graph) {
    cudaGraphExec_t exec = nullptr;
    ce = cudaGraphInstantiate(&exec, graph, 0);
    if (ce == cudaSuccess) { e.exec = exec; e.launches = g_launches - before; }
  }
  if (graph) cudaGraphDestroy(graph);
  g_launches = before;
  if (!e.exec) {            // capture is an optimisation: fall back to eager launches for this key
    cudaGetLastError();
    e.failed = true;
    body(st);
    return;
  }
  IPK_CUDA(cudaGraphLaunch(e.exec, st));
  g_launches += e.launches;
}
}  // namespace

namespace ipk {
void run_graphed_step(const void* owner, int B, int kind, cudaStream_t st, const std::function<void(cudaStream_t)>& body, bool enabled) {
  run_graphed(GraphKey{owner, nullptr, nullptr, nullptr, nullptr, nullptr, B, 0, kind}, st, body, enabled);
}
}  // namespace ipk

// plans call this from their destroy entry points: a recycled handle must not replay a stale graph
void ipk_graphs_drop(const void* handle) {
  std::lock_guard<std::mutex> lk(g_graphs_mu);
  for (auto it = g_graphs.begin(); it != g_graphs.end();) {
    if (it->first.f == handle || it->first.d == handle) {
      if (it->second.exec) cudaGraphExecDestroy(it->second.exec);
      it = g_graphs.erase(it);
    } else ++it;
  }
}

extern "C" int ipk_sample(ipk_flow* f, ipk_fs* d, const float* z, const float* cond, const float* x0, float* frames,
                          int32_t B, int32_t T, void* stream) {
  IPK_TRY
  IPK_CHECK(f && d && z && cond && x0 && frames, IPK_ERR_INVALID, "ipk_sample: null argument");
  run_graphed(GraphKey{f, d, z, cond, x0, frames, B, T, 0}, (cudaStream_t)stream, [&](cudaStream_t st) {
    const float* motion = nullptr;   // flow state stays NHWC on device and feeds the GRU directly
    ipk_flow_reverse_nhwc(f, z, cond, &motion, B, st);
    ipk_fs_decode_nhwc(d, motion, x0, frames, B, T, st);
  }, graph_enabled());
  IPK_CATCH
}

namespace {
struct HostStage {
  float *z = nullptr, *cond = nullptr, *x0 = nullptr, *frames = nullptr, *u8 = nullptr;   // u8: byte buffer, capacity counted in floats
  size_t nz = 0, nc = 0, nx = 0, nf = 0, nu = 0;
  cudaStream_t copy_st = nullptr;
  void ensure(float** p, size_t* cap, size_t n) {
    if (*cap >= n) return;
    if (*p) cudaFree(*p);
    IPK_CUDA(cudaMalloc((void**)p, n * sizeof(float)));
    *cap = n;
  }
};
HostStage g_stages[IPK_MAX_DEVICES];      // staging buffers / copy stream of ipk_sample_host, one set per device
std::mutex g_stage_mu;
}  // namespace

struct FlowDims { int C0, hch; };
FlowDims ipk_flow_dims(ipk_flow* f);
int ipk_fs_spatial(ipk_fs* d);

static void sample_host_impl(ipk_flow* f, ipk_fs* d, const float* z_host, const float* cond_host, const float* x0_host,
                             float* frames_host, uint8_t* u8_host, int B, int T, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_stage_mu);
  HostStage& g_stage = g_stages[current_device_slot()];
  FlowDims fd = ipk_flow_dims(f);
  const int S = ipk_fs_spatial(d);
  const size_t nz = (size_t)B * fd.C0 * 64, nc = (size_t)B * fd.hch * 64, nx = (size_t)B * 3 * S * S, nf = (size_t)B * T * 3 * S * S;
  g_stage.ensure(&g_stage.z, &g_stage.nz, nz);
  g_stage.ensure(&g_stage.cond, &g_stage.nc, nc);
  g_stage.ensure(&g_stage.x0, &g_stage.nx, nx);
  g_stage.ensure(&g_stage.frames, &g_stage.nf, nf);
  if (u8_host) g_stage.ensure(&g_stage.u8, &g_stage.nu, (nf + 3) / 4);
  IPK_CUDA(cudaMemcpyAsync(g_stage.z, z_host, nz * 4, cudaMemcpyHostToDevice, st));
  IPK_CUDA(cudaMemcpyAsync(g_stage.cond, cond_host, nc * 4, cudaMemcpyHostToDevice, st));
  IPK_CUDA(cudaMemcpyAsync(g_stage.x0, x0_host, nx * 4, cudaMemcpyHostToDevice, st));
  const float* motion = nullptr;
  ipk_flow_reverse_nhwc(f, g_stage.z, g_stage.cond, &motion, B, st);
  // frames leave chunk by chunk on a second stream while the next chunk is decoded
  if (!g_stage.copy_st) IPK_CUDA(cudaStreamCreateWithFlags(&g_stage.copy_st, cudaStreamNonBlocking));
  ipk_fs_decode_nhwc(d, motion, g_stage.x0, g_stage.frames, B, T, st, frames_host, g_stage.copy_st,
                     u8_host ? (uint8_t*)g_stage.u8 : nullptr, u8_host);
  IPK_CUDA(cudaStreamSynchronize(g_stage.copy_st));
  IPK_CUDA(cudaStreamSynchronize(st));
}

extern "C" int ipk_sample_host(ipk_flow* f, ipk_fs* d, const float* z_host, const float* cond_host, const float* x0_host,
                               float* frames_host, int32_t B, int32_t T, void* stream) {
  IPK_TRY
  IPK_CHECK(f && d && z_host && cond_host && x0_host && frames_host, IPK_ERR_INVALID, "ipk_sample_host: null argument");
  sample_host_impl(f, d, z_host, cond_host, x0_host, frames_host, nullptr, B, T, (cudaStream_t)stream);
  IPK_CATCH
}

extern "C" int ipk_sample_host_u8(ipk_flow* f, ipk_fs* d, const float* z_host, const float* cond_host, const float* x0_host,
                                  uint8_t* frames_u8_host, int32_t B, int32_t T, void* stream) {
  IPK_TRY
  IPK_CHECK(f && d && z_host && cond_host && x0_host && frames_u8_host, IPK_ERR_INVALID, "ipk_sample_host_u8: null argument");
  sample_host_impl(f, d, z_host, cond_host, x0_host, nullptr, frames_u8_host, B, T, (cudaStream_t)stream);
  IPK_CATCH
}

extern "C" int ipk_frames_to_u8(const float* frames, uint8_t* out, int64_t n_frames, int32_t spatial, void* stream) {
  IPK_TRY
  IPK_CHECK(frames && out && n_frames > 0 && spatial > 0, IPK_ERR_INVALID, "ipk_frames_to_u8: bad argument");
  frames_to_u8(frames, out, n_frames, spatial * spatial, (cudaStream_t)stream);
  IPK_CATCH
}

// ------------------------------------------------------------------------------------------ test hooks
namespace {
struct Tmp {
  std::vector<void*> v;
  template <typename T> T* alloc(size_t n) { void* p; IPK_CUDA(cudaMalloc(&p, std::max<size_t>(n * sizeof(T), 16))); v.push_back(p); return (T*)p; }
  ~Tmp() { for (void* p : v) cudaFree(p); }
};

// fp32 NHWC tensor -> operand of `engine` (returns hi pointer; lo through *lo)
void* make_operand(Tmp& tmp, const float* x, long long rows, int C, int engine, void** lo, cudaStream_t st) {
  *lo = nullptr;
  if (engine == IPK_PREC_FP32_SIMT) return (void*)x;
  __nv_bfloat16* hi = tmp.alloc<__nv_bfloat16>((size_t)rows * C);
  NormApply n; n.x = x; n.F = 1; n.P = rows; n.C = C; n.out_hi = hi;
  if (engine == IPK_PREC_FP32_SPLIT) { n.out_lo = tmp.alloc<__nv_bfloat16>((size_t)rows * C); *lo = n.out_lo; }
  norm_apply(n, st);
  return hi;
}
}  // namespace

extern "C" int ipk_test_gemm(const float* A, const float* W, float* out, int32_t M, int32_t N, int32_t K, int32_t precision, void* stream) {
  IPK_TRY
  cudaStream_t st = (cudaStream_t)stream;
  IPK_CHECK(K % 4 == 0, IPK_ERR_UNSUPPORTED, "ipk_test_gemm: K must be a multiple of 4");
  if (precision != IPK_PREC_FP32_SIMT) IPK_CHECK(K % 8 == 0, IPK_ERR_UNSUPPORTED, "ipk_test_gemm: tensor-core engine needs K %% 8 == 0");
  DevPool pool;
  Tmp tmp;
  ConvW w = conv_alloc(pool, precision, 1, K, N, false);
  PackSrc s; s.w = W; s.N = N; s.Ksrc = K;
  conv_pack_into(w, 0, s, {0}, st);
  ConvIn in; in.cstride = K; in.F = M; in.H = 1; in.W = 1;
  in.p = make_operand(tmp, A, M, K, precision, (void**)&in.p_lo, st);
  ConvOut o; o.p = out; o.cstride = N; o.Ho = 1; o.Wo = 1;
  conv_run(w, in, o, taps_1x1(), 1, st);
  IPK_CUDA(cudaStreamSynchronize(st));
  pool.release();
  IPK_CATCH
}

static int test_conv_impl(const float* in_, const float* w_, const float* bias, float* out, int F, int H, int W, int Cin, int Cout,
                          int precision, bool transposed, cudaStream_t st) {
  IPK_TRY
  IPK_CHECK(Cin % 4 == 0, IPK_ERR_UNSUPPORTED, "test conv: Cin must be a multiple of 4");
  if (precision != IPK_PREC_FP32_SIMT) IPK_CHECK(Cin % 8 == 0, IPK_ERR_UNSUPPORTED, "test conv: tensor-core engine needs Cin %% 8 == 0");
  DevPool pool;
  Tmp tmp;
  ConvW w = conv_alloc(pool, precision, 9, Cin, Cout, bias != nullptr);
  PackSrc s; s.w = w_; s.N = Cout; s.Ksrc = Cin; s.kh = 3; s.kw = 3; s.transposed = transposed;
  conv_pack_into(w, 0, s, {0, 1, 2, 3, 4, 5, 6, 7, 8}, st);
  if (bias) conv_pack_bias(w, 0, bias, Cout, 0.f, st);
  ConvIn in; in.cstride = Cin; in.F = F; in.H = H; in.W = W;
  in.p = make_operand(tmp, in_, (long long)F * H * W, Cin, precision, (void**)&in.p_lo, st);
  ConvOut o; o.p = out; o.cstride = Cout; o.bias = w.bias;
  if (!transposed) {
    o.Ho = H; o.Wo = W;
    conv_run(w, in, o, taps_3x3(), 1, st);
  } else {
    o.Ho = 2 * H; o.Wo = 2 * W; o.ymul = 2; o.xmul = 2;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        o.yadd = a; o.xadd = b;
        conv_run(w, in, o, taps_convT(a, b), 1, st);
      }
  }
  IPK_CUDA(cudaStreamSynchronize(st));
  pool.release();
  IPK_CATCH
}

extern "C" int ipk_test_conv3x3(const float* in, const float* w, const float* bias, float* out, int32_t F, int32_t H, int32_t W,
                                int32_t Cin, int32_t Cout, int32_t precision, void* stream) {
  return test_conv_impl(in, w, bias, out, F, H, W, Cin, Cout, precision, false, (cudaStream_t)stream);
}
extern "C" int ipk_test_convT3x3(const float* in, const float* w, const float* bias, float* out, int32_t F, int32_t H, int32_t W,
                                 int32_t Cin, int32_t Cout, int32_t precision, void* stream) {
  return test_conv_impl(in, w, bias, out, F, H, W, Cin, Cout, precision, true, (cudaStream_t)stream);
}

// Conv3d on the tcgen05 engine (conv3d_tc.cu): NDHWC in[B,T,H,W,Cin] fp32, OIDHW w[Cout,Cin,kt,ky,kx] -> NDHWC out[B,To,Ho,Wo,Cout] and
// the fused per-(sample, channel) statistics stats[B,Cout,2] (sum, sum of squares; fp64).  dims = {B,T,H,W,Cin,Cout, kt,ky,kx, st,sy,sx, pt,py,px}
extern "C" int ipk_test_conv3d(const float* in, const float* w, float* out, double* stats, const int32_t* dims, int32_t precision, void* stream) {
  IPK_TRY
  cudaStream_t st = (cudaStream_t)stream;
  IPK_CHECK(in && w && out && dims, IPK_ERR_INVALID, "ipk_test_conv3d: null argument");
  IPK_CHECK(precision == IPK_PREC_FP32_SPLIT || precision == IPK_PREC_BF16, IPK_ERR_UNSUPPORTED, "ipk_test_conv3d: tensor-core precisions only");
  const int B = dims[0], T = dims[1], H = dims[2], W = dims[3], Cin = dims[4], Cout = dims[5];
  Conv3dShape sh{Cin, Cout, T, H, W, dims[6], dims[7], dims[8], dims[9], dims[10], dims[11], dims[12], dims[13], dims[14]};
  IPK_CHECK(conv3d_tc_supported(sh), IPK_ERR_UNSUPPORTED, "ipk_test_conv3d: shape not supported by the tensor-core engine");
  DevPool pool;
  Tmp tmp;
  ConvW cw = conv_alloc(pool, precision, dims[6] * dims[7] * dims[8], Cin, Cout, false);
  conv3d_tc_pack(cw, w, Cout, Cin, st);
  void* lo = nullptr;
  void* hi = make_operand(tmp, in, (long long)B * T * H * W, Cin, precision, &lo, st);
  if (stats) IPK_CUDA(cudaMemsetAsync(stats, 0, (size_t)B * Cout * 2 * sizeof(double), st));
  conv3d_tc_run(cw, sh, hi, lo, B, out, stats, st);
  IPK_CUDA(cudaStreamSynchronize(st));
  pool.release();
  IPK_CATCH
}
