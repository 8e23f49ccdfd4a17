// Shared helpers for the ipoke_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdexcept>
#include <vector>
#include <algorithm>
#include <utility>
#include <cstring>
#include <cstdlib>
#include <functional>

#include "../../include/ipoke_b200.h"

namespace ipk {

// ---------------------------------------------------------------- errors
struct Error : public std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& m);
[[noreturn]] void fail(int code, const char* fmt, ...);

#define IPK_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      ::ipk::fail(IPK_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define IPK_CHECK(cond, code, ...)                  \
  do {                                              \
    if (!(cond)) ::ipk::fail((code), __VA_ARGS__);  \
  } while (0)

// launch bookkeeping (gpu_launches in bench.py)
extern thread_local int64_t g_launches;
inline void count_launch(int n = 1) { g_launches += n; }
#define IPK_LAUNCH_CHECK()                \
  do {                                    \
    ::ipk::count_launch();                \
    IPK_CUDA(cudaGetLastError());         \
  } while (0)

// ---------------------------------------------------------------- programmatic dependent launch
// Every hot kernel is launched with programmatic stream serialization: its CTAs may be scheduled while the previous kernel in
// the stream drains, run their prologue (barrier init, TMEM allocation, index setup) and then block in pdl_wait() until the
// previous grid has completed and its writes are visible.  pdl_trigger() lets the NEXT kernel do the same with this one.
// Kernels launched this way MUST call pdl_wait() before their first global-memory access.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();   // IPK_PDL=0 disables (plain stream-ordered launches)

// cluster_x > 1 launches thread-block clusters of that many CTAs along x (gridDim.x must be a multiple of it)
template <typename... KArgs, typename... Args>
inline void launch_kc(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster_x > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)cluster_x; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
  if (e != cudaSuccess) ::ipk::fail(IPK_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
  count_launch();
}
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  launch_kc(kernel, grid, block, smem, st, 1, std::forward<Args>(args)...);
}

// ---------------------------------------------------------------- optional per-phase device timing
// When enabled (ipk_prof_enable), every ProfScope brackets the launches issued inside it with CUDA events on the launch
// stream; ipk_prof_report() sums the elapsed time per tag.  Disabled: zero cost besides one branch.
struct Prof {
  static bool& enabled();
  static void begin(const char* tag, cudaStream_t st);
  static void end(cudaStream_t st);
};
// Every ProfScope is also an NVTX range (domain-less push/pop, header-only nvtx3) when IPK_NVTX=1, so that nsys / ncu --nvtx timelines
// carry the stage names (flow.nice.conv2, dec.up.convT.b3, train.backward, ...); without the variable the cost is one branch.
bool nvtx_enabled();
void nvtx_push(const char* tag);
void nvtx_pop();
struct ProfScope {
  cudaStream_t st;
  bool on, nv;
  ProfScope(const char* tag, cudaStream_t s) : st(s), on(Prof::enabled()), nv(nvtx_enabled()) {
    if (nv) nvtx_push(tag);
    if (on) Prof::begin(tag, st);
  }
  ~ProfScope() {
    if (on) Prof::end(st);
    if (nv) nvtx_pop();
  }
};

// ---------------------------------------------------------------- activations
enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_ELU = 2, ACT_LRELU02 = 3, ACT_TANH = 4, ACT_SIGMOID = 5 };

__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_ELU: return v > 0.f ? v : expm1f(v);
    case ACT_LRELU02: return v > 0.f ? v : 0.2f * v;
    case ACT_TANH: return tanhf(v);
    case ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}

// Activation-operand storage modes (what a producing kernel writes for the next contraction)
enum OutMode : int {
  OUT_F32_NHWC = 0,     // fp32 [pixels][cstride]
  OUT_F32_NCHW = 1,     // fp32 [frame][C][Ho][Wo]   (frames at the ABI edge)
  OUT_BF16_SPLIT = 2,   // two bf16 planes (hi, lo) [pixels][cstride]; lo plane at +plane_stride elements
  OUT_BF16 = 3          // one bf16 plane
};

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// ordinal of the calling thread's current device, folded into [0, 64): per-device one-time setup (function attributes, SM counts,
// staging buffers) is keyed by it, so that one process may drive several GPUs through this library
constexpr int IPK_MAX_DEVICES = 64;
inline int current_device_slot() {
  int d = 0;
  IPK_CUDA(cudaGetDevice(&d));
  return d & (IPK_MAX_DEVICES - 1);
}

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
inline int64_t round_up64(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- device arena (bump allocator)
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0;
  void init(size_t bytes) {
    release();
    IPK_CUDA(cudaMalloc((void**)&base, bytes));
    cap = bytes;
    off = 0;
  }
  void release() {
    if (base) cudaFree(base);
    base = nullptr;
    cap = off = 0;
  }
  template <typename T>
  T* alloc(size_t n) {
    size_t bytes = (n * sizeof(T) + 255) / 256 * 256;
    IPK_CHECK(off + bytes <= cap, IPK_ERR_INVALID, "arena overflow: need %zu more bytes (cap %zu)", bytes, cap);
    T* p = (T*)(base + off);
    off += bytes;
    return p;
  }
};

// owned device allocations (packed weights): chunked bump allocator, 256-byte aligned
struct DevPool {
  std::vector<void*> chunks;
  char* cur = nullptr;
  size_t cur_cap = 0, cur_off = 0;
  size_t total = 0;
  static constexpr size_t kChunk = 256ull << 20;
  template <typename T>
  T* alloc(size_t n, bool zero = false) {
    size_t bytes = (std::max<size_t>(n * sizeof(T), 16) + 255) / 256 * 256;
    char* p;
    if (bytes > kChunk / 4) {
      void* q = nullptr;
      IPK_CUDA(cudaMalloc(&q, bytes));
      chunks.push_back(q);
      p = (char*)q;
    } else {
      if (cur == nullptr || cur_off + bytes > cur_cap) {
        void* q = nullptr;
        IPK_CUDA(cudaMalloc(&q, kChunk));
        chunks.push_back(q);
        cur = (char*)q;
        cur_cap = kChunk;
        cur_off = 0;
      }
      p = cur + cur_off;
      cur_off += bytes;
    }
    total += bytes;
    if (zero) {
      IPK_CUDA(cudaMemset(p, 0, bytes));
      IPK_CUDA(cudaDeviceSynchronize());   // plan-build time only: order the memset before packing kernels on any stream
    }
    return (T*)p;
  }
  void release() {
    for (void* p : chunks) cudaFree(p);
    chunks.clear();
    cur = nullptr;
    cur_cap = cur_off = total = 0;
  }
};

// CUDA-graph replay of a fixed launch sequence (capi.cu).  body(stream) must only enqueue work whose device pointers and shapes are
// fully determined by (owner, B, kind).  The first call with a key runs eagerly (lazy one-time setup), the second is captured on an
// internal stream and instantiated, later calls replay the graph into `st`.  enabled = false, or an active profiler, runs eagerly.
void run_graphed_step(const void* owner, int B, int kind, cudaStream_t st, const std::function<void(cudaStream_t)>& body, bool enabled);

#define IPK_TRY try {
#define IPK_CATCH                                                                                \
  }                                                                                              \
  catch (const ::ipk::Error& e) { ::ipk::set_last_error(e.what()); return e.code; }              \
  catch (const std::exception& e) { ::ipk::set_last_error(e.what()); return IPK_ERR_INVALID; }   \
  return IPK_OK;

}  // namespace ipk
