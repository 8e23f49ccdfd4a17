// tcgen05 implicit-GEMM convolution engine (sm_100a).
//
//   D[128 pixels x BN] (fp32, TMEM) += sum over taps, K-blocks of  A_tap[128 x 64] (bf16, smem) * W_tap[BN x 64]^T (bf16, smem)
//
// * A operand: NHWC activation planes [F][H][W][C] read by TMA as 4-D boxes (64 channels x bw x bh x bf pixels = 128
//   rows) whose start coordinate is shifted by the tap offset (dy, dx); out-of-bounds rows/columns are zero-filled by
//   the TMA unit, which IS the zero padding of the convolution -- no im2col buffer exists anywhere.
// * B operand: packed weights [tap][Npad][Kpad] (K-major), 2-D TMA boxes of 64 x BN.
// * 128-byte swizzled K-major tiles, UMMA M=128, N=BN, K=16 (kind::f16, bf16 inputs, fp32 accumulation in TMEM).
// * fp32 fidelity mode (NSPLIT=3): every operand is a pair of bf16 planes (hi, lo = bf16(x - hi)) and each K-step issues
//   three MMAs  hi*hi + lo*hi + hi*lo  into the same accumulator (error-compensated "bf16x3", ~2^-16 relative).
// * warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane), warps 2..5 = epilogue
//   (TMEM -> registers -> bias / activation / operand split -> global).
#include <cuda.h>
#include <map>
#include <mutex>
#include <tuple>
#include "conv.cuh"

namespace ipk {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_THREADS = 192;
constexpr size_t TC_SMEM_BUDGET = 200 * 1024;

struct TcArgs {
  int F, H, W;
  int bw, bh, bf;                 // pixel box (bw*bh*bf == 128)
  int tiles_x, tiles_y;           // tiles along x and y (tiles along f = gridDim.x / (tiles_x*tiles_y))
  int ntaps, taps_per_split, nkb; // nkb = Kpad / 64
  int dy[MAX_TAPS], dx[MAX_TAPS], widx[MAX_TAPS];
  int Npad;
  int stages;
  // epilogue
  const float* bias;
  int act, out_mode, out_cstride, out_coff, Ho, Wo, ymul, yadd, xmul, xadd;
  void* out;
  void* out_lo;
  long long split_stride;
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a pipeline bug must surface as a trap, never as a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("ipoke_b200 conv_tc: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO=1<<16 |
// SBO=(8 rows * 128 B)>>4 <<32 | version=1<<46 | layout SWIZZLE_128B=2<<61
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ kernel
template <int BN, int NSPLIT>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo, const TcArgs a) {
  constexpr int A_BYTES = TC_BM * TC_BK * 2;         // 16 KB
  constexpr int W_BYTES = BN * TC_BK * 2;
  constexpr int NPLANES = NSPLIT == 3 ? 2 : 1;
  constexpr int STAGE_BYTES = NPLANES * (A_BYTES + W_BYTES);
  constexpr uint32_t IDESC = umma_idesc_bf16(TC_BM, BN);
  constexpr int MAX_STAGES = 8;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B needs 1024-byte alignment
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stages = a.stages;

  // tile coordinates
  const int txy = a.tiles_x * a.tiles_y;
  const int tf = blockIdx.x / txy;
  const int ty = (blockIdx.x % txy) / a.tiles_x;
  const int tx = blockIdx.x % a.tiles_x;
  const int f0 = tf * a.bf, y0 = ty * a.bh, x0 = tx * a.bw;
  const int n0 = blockIdx.y * BN;
  const int tap_begin = blockIdx.z * a.taps_per_split;
  const int tap_end = min(a.ntaps, tap_begin + a.taps_per_split);
  const int iters = (tap_end - tap_begin) * a.nkb;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation by one full warp; the same warp deallocates
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"((uint32_t)BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = tap_begin; t < tap_end; ++t) {
        const int dy = a.dy[t], dx = a.dx[t], wrow = a.widx[t] * a.Npad + n0;
        for (int kb = 0; kb < a.nkb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* st = smem + (size_t)s * STAGE_BYTES;
          mbar_expect_tx(&full_bar[s], STAGE_BYTES);
          tma_load_4d(st, &tmA_hi, &full_bar[s], kb * TC_BK, x0 + dx, y0 + dy, f0);
          tma_load_2d(st + NPLANES * A_BYTES, &tmW_hi, &full_bar[s], kb * TC_BK, wrow);
          if (NSPLIT == 3) {
            tma_load_4d(st + A_BYTES, &tmA_lo, &full_bar[s], kb * TC_BK, x0 + dx, y0 + dy, f0);
            tma_load_2d(st + NPLANES * A_BYTES + W_BYTES, &tmW_lo, &full_bar[s], kb * TC_BK, wrow);
          }
          if (++s == stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint32_t sw = sa + NPLANES * A_BYTES;
        const uint64_t da_hi = umma_desc_sw128(sa), dw_hi = umma_desc_sw128(sw);
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          const uint64_t koff = (uint64_t)((k * 32) >> 4);   // 16 bf16 = 32 bytes along K inside the swizzle atom
          umma_bf16(tmem_base, da_hi + koff, dw_hi + koff, IDESC, (it > 0 || k > 0) ? 1u : 0u);
          if (NSPLIT == 3) {
            const uint64_t da_lo = umma_desc_sw128(sa + A_BYTES), dw_lo = umma_desc_sw128(sw + W_BYTES);
            umma_bf16(tmem_base, da_lo + koff, dw_hi + koff, IDESC, 1u);
            umma_bf16(tmem_base, da_hi + koff, dw_lo + koff, IDESC, 1u);
          }
        }
        umma_commit(&empty_bar[s]);                 // frees the stage once the MMAs above have read it
        if (it == iters - 1) umma_commit(&tmem_full_bar);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====================
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int r = q * 32 + lane;                    // accumulator row == pixel index inside the box
    const int xl = r % a.bw, yl = (r / a.bw) % a.bh, fl = r / (a.bw * a.bh);
    const int f = f0 + fl, y = y0 + yl, x = x0 + xl;
    const bool valid = (f < a.F) && (y < a.H) && (x < a.W);
    const size_t opix = ((size_t)f * a.Ho + (size_t)(y * a.ymul + a.yadd)) * a.Wo + (size_t)(x * a.xmul + a.xadd);
    const size_t obase = opix * a.out_cstride + a.out_coff;
    float* outf = (float*)a.out + (size_t)blockIdx.z * a.split_stride;
#pragma unroll 1
    for (int c = 0; c < BN; c += 16) {
      if (n0 + c >= a.Npad) break;                  // warp-uniform
      uint32_t rr[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, rr);
      tmem_ld_wait();
      if (!valid) continue;
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float t = __uint_as_float(rr[j]);
        if (a.bias) t += __ldg(a.bias + n0 + c + j);
        v[j] = act_apply(t, a.act);
      }
      const size_t o = obase + n0 + c;
      if (a.out_mode == OUT_F32_NHWC) {
        float4* p = (float4*)(outf + o);
#pragma unroll
        for (int j = 0; j < 4; ++j) p[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      } else {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          __nv_bfloat16 h0, l0, h1, l1;
          split_bf16(v[2 * j], h0, l0);
          split_bf16(v[2 * j + 1], h1, l1);
          __nv_bfloat162 hh(h0, h1), ll(l0, l1);
          hi[j] = *(uint32_t*)&hh;
          lo[j] = *(uint32_t*)&ll;
        }
        uint4* ph = (uint4*)((__nv_bfloat16*)a.out + o);
        ph[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        ph[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        if (a.out_mode == OUT_BF16_SPLIT) {
          uint4* pl = (uint4*)((__nv_bfloat16*)a.out_lo + o);
          pl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          pl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
  });
  IPK_CHECK(fn != nullptr, IPK_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  return fn;
}

struct MapKey {
  const void* p; long long d0, d1, d2, d3, s1, s2, s3; int b0, b1, b2, b3, rank;
  bool operator<(const MapKey& o) const {
    return std::tie(p, d0, d1, d2, d3, s1, s2, s3, b0, b1, b2, b3, rank) <
           std::tie(o.p, o.d0, o.d1, o.d2, o.d3, o.s1, o.s2, o.s3, o.b0, o.b1, o.b2, o.b3, o.rank);
  }
};
static std::map<MapKey, CUtensorMap> g_maps;
static std::mutex g_maps_mu;

static CUtensorMap make_map(const void* base, int rank, const long long* dims, const long long* strides_bytes, const int* box) {
  MapKey k{base, dims[0], dims[1], rank > 2 ? dims[2] : 1, rank > 3 ? dims[3] : 1, strides_bytes[0], rank > 2 ? strides_bytes[1] : 0,
           rank > 3 ? strides_bytes[2] : 0, box[0], box[1], rank > 2 ? box[2] : 1, rank > 3 ? box[3] : 1, rank};
  std::lock_guard<std::mutex> lk(g_maps_mu);
  auto it = g_maps.find(k);
  if (it != g_maps.end()) return it->second;
  CUtensorMap m;
  cuuint64_t gd[4]; cuuint64_t gs[3]; cuuint32_t bx[4]; cuuint32_t es[4] = {1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) { gd[i] = (cuuint64_t)dims[i]; bx[i] = (cuuint32_t)box[i]; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = (cuuint64_t)strides_bytes[i];
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IPK_CHECK(r == CUDA_SUCCESS, IPK_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): rank %d dims [%lld,%lld,%lld,%lld] box [%d,%d,%d,%d] stride1 %lld",
            (int)r, rank, k.d0, k.d1, k.d2, k.d3, k.b0, k.b1, k.b2, k.b3, k.s1);
  if (g_maps.size() > 8192) g_maps.clear();
  g_maps[k] = m;
  return m;
}

template <int BN, int NSPLIT>
static void launch_tc(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi, const CUtensorMap& w_lo, TcArgs& a,
                      dim3 grid, cudaStream_t st) {
  constexpr int STAGE_BYTES = (NSPLIT == 3 ? 2 : 1) * (TC_BM * TC_BK * 2 + BN * TC_BK * 2);
  int stages = (int)std::min<size_t>(8, TC_SMEM_BUDGET / STAGE_BYTES);
  a.stages = stages;
  size_t smem = (size_t)stages * STAGE_BYTES + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    IPK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TC_SMEM_BUDGET + 1024)));
    attr_set = true;
  }
  conv_tc_kernel<BN, NSPLIT><<<grid, TC_THREADS, smem, st>>>(a_hi, a_lo, w_hi, w_lo, a);
  IPK_LAUNCH_CHECK();
}

void conv_tc_run(const ConvW& w, const ConvIn& in, const ConvOut& out, const TapList& taps, int nsplit, cudaStream_t st) {
  IPK_CHECK(w.w_hi != nullptr, IPK_ERR_STATE, "conv_tc_run: layer was not packed for the tensor-core engine");
  const bool split = w.engine == IPK_PREC_FP32_SPLIT;
  IPK_CHECK(!split || (in.p_lo && w.w_lo), IPK_ERR_STATE, "conv_tc_run: split precision needs hi and lo operand planes");
  IPK_CHECK(in.cstride % 8 == 0 && in.coff % 8 == 0, IPK_ERR_UNSUPPORTED, "conv_tc_run: activation rows must be 16-byte aligned (cstride %d, coff %d)", in.cstride, in.coff);
  IPK_CHECK(out.mode != OUT_F32_NCHW, IPK_ERR_UNSUPPORTED, "conv_tc_run: NCHW output is served by the SIMT engine");
  IPK_CHECK((out.cstride % 4 == 0) && (out.coff % 4 == 0) && out.coff + w.Npad <= out.cstride, IPK_ERR_UNSUPPORTED,
            "conv_tc_run: output row (cstride %d, coff %d) cannot hold Npad %d", out.cstride, out.coff, w.Npad);
  if (out.mode != OUT_F32_NHWC) IPK_CHECK(out.cstride % 8 == 0 && out.coff % 8 == 0, IPK_ERR_UNSUPPORTED, "conv_tc_run: bf16 output rows must be 16-byte aligned");
  long long M = (long long)in.F * in.H * in.W;
  if (M == 0) return;
  TcArgs a;
  a.F = in.F; a.H = in.H; a.W = in.W;
  if ((long long)in.H * in.W <= TC_BM) {
    IPK_CHECK(TC_BM % (in.H * in.W) == 0, IPK_ERR_UNSUPPORTED, "conv_tc_run: grid %dx%d does not tile 128 rows", in.H, in.W);
    a.bw = in.W; a.bh = in.H; a.bf = TC_BM / (in.H * in.W);
  } else if (in.W >= TC_BM) {
    a.bw = TC_BM; a.bh = 1; a.bf = 1;
  } else {
    IPK_CHECK(TC_BM % in.W == 0, IPK_ERR_UNSUPPORTED, "conv_tc_run: width %d does not divide 128", in.W);
    a.bw = in.W; a.bh = TC_BM / in.W; a.bf = 1;
  }
  a.tiles_x = cdiv(in.W, a.bw); a.tiles_y = cdiv(in.H, a.bh);
  const int tiles_f = cdiv(in.F, a.bf);
  a.ntaps = taps.n;
  nsplit = std::max(1, std::min(nsplit, taps.n));
  a.taps_per_split = cdiv(taps.n, nsplit);
  nsplit = cdiv(taps.n, a.taps_per_split);
  IPK_CHECK(nsplit == 1 || (out.mode == OUT_F32_NHWC && out.split_stride > 0), IPK_ERR_INVALID, "split-K needs fp32 partial slices");
  a.nkb = w.Kpad / TC_BK;
  for (int i = 0; i < MAX_TAPS; ++i) { a.dy[i] = taps.dy[i]; a.dx[i] = taps.dx[i]; a.widx[i] = taps.widx[i]; }
  a.Npad = w.Npad;
  a.bias = out.bias; a.act = out.act; a.out_mode = out.mode; a.out_cstride = out.cstride; a.out_coff = out.coff;
  a.Ho = out.Ho; a.Wo = out.Wo; a.ymul = out.ymul; a.yadd = out.yadd; a.xmul = out.xmul; a.xadd = out.xadd;
  a.out = out.p; a.out_lo = out.p_lo; a.split_stride = out.split_stride;
  a.stages = 2;

  // activation maps: dims (C, W, H, F); the C extent is the true channel count so the K tail is zero-filled
  long long ad[4] = {w.K, in.W, in.H, in.F};
  long long as[3] = {(long long)in.cstride * 2, (long long)in.W * in.cstride * 2, (long long)in.H * in.W * in.cstride * 2};
  int ab[4] = {TC_BK, a.bw, a.bh, a.bf};
  const __nv_bfloat16* ahi = (const __nv_bfloat16*)in.p + in.coff;
  CUtensorMap mA_hi = make_map(ahi, 4, ad, as, ab);
  CUtensorMap mA_lo = split ? make_map((const __nv_bfloat16*)in.p_lo + in.coff, 4, ad, as, ab) : mA_hi;
  int BN = w.Npad >= 256 ? 256 : (w.Npad > 64 ? 128 : (w.Npad > 32 ? 64 : 32));
  long long wd[2] = {w.Kpad, (long long)w.ntaps * w.Npad};
  long long wsb[1] = {(long long)w.Kpad * 2};
  int wb[2] = {TC_BK, BN};
  CUtensorMap mW_hi = make_map(w.w_hi, 2, wd, wsb, wb);
  CUtensorMap mW_lo = split ? make_map(w.w_lo, 2, wd, wsb, wb) : mW_hi;

  dim3 grid((unsigned)(a.tiles_x * a.tiles_y * tiles_f), (unsigned)cdiv(w.Npad, BN), (unsigned)nsplit);
#define IPK_TC_CASE(bn)                                                             \
  case bn:                                                                          \
    if (split) launch_tc<bn, 3>(mA_hi, mA_lo, mW_hi, mW_lo, a, grid, st);           \
    else launch_tc<bn, 1>(mA_hi, mA_lo, mW_hi, mW_lo, a, grid, st);                 \
    break;
  switch (BN) {
    IPK_TC_CASE(32)
    IPK_TC_CASE(64)
    IPK_TC_CASE(128)
    IPK_TC_CASE(256)
  }
#undef IPK_TC_CASE
}

}  // namespace ipk
