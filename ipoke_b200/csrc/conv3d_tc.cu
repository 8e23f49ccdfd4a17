// tcgen05 implicit-GEMM Conv3d (sm_100a): the 3-D convolutions of the first-stage video encoder
// (ResNetMotionEncoder / BasicBlock, models/modules/motion_models/motion_encoder.py:45-74,150-241).
//
//   D[128 output voxels x BN] (fp32, TMEM) += sum over taps (dt,dy,dx), K-blocks of  A_tap[128 x 64] * W_tap[BN x 64]^T
//
// * A operand: NDHWC activation planes [B][T][H][W][C] (bf16 hi [, lo]) read by TMA as 5-D boxes
//   (64 channels x bw x bh x 1 x bb voxels = 128 rows).  The box start is the tap-shifted INPUT coordinate
//   (x0*sx - px + dx, y0*sy - py + dy, t*st - pt + dt, b0); spatial strides are the tensor map's element strides (the box
//   spans (bw-1)*sx+1 input columns and lands as bw dense rows), out-of-bounds voxels are zero-filled by the TMA unit, which IS
//   the conv padding.  No im2col buffer.
// * A tile covers ONE output time step, so a tap whose input time t*st - pt + dt falls outside [0, Ti) contributes only
//   padding: producer and MMA issuer both skip it (for the encoder's late layers, T_out = 1..3, this removes 22..67 % of the
//   MMA work instead of multiplying zeros).
// * B operand: packed weights [tap][Npad][Kpad] (K-major, ConvW of conv.cuh), 128-byte swizzle, UMMA M128 x BN x K16,
//   kind::f16 bf16 -> fp32 in TMEM; fp32 fidelity = bf16x3 (hi*hi + lo*hi + hi*lo), as in conv_tc.cu.
// * persistent, warp-specialised (TMA / MMA / 8 epilogue warps), two TMEM accumulator stages; epilogue: fp32 NDHWC rows through
//   an xor-swizzled smem transpose (coalesced 128-byte segments) and, optionally, the per-(sample, channel) sum / sum of squares
//   of the output for the GroupNorm(16) that follows every Conv3d (fp64 atomics), so no separate statistics pass reads it back.
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include "conv.cuh"
#include "conv3d_tc.cuh"
#include "tc_ptx.cuh"

namespace ipk {

constexpr int C3T_BM = 128;
constexpr int C3T_BK = 64;
constexpr int C3T_EPI_WARPS = 8;
constexpr int C3T_THREADS = 64 + 32 * C3T_EPI_WARPS;
constexpr size_t C3T_SMEM_BUDGET = 192 * 1024;
constexpr size_t C3T_EPI_STAGE_BYTES = 4096;

struct C3TArgs {
  int B, Ti, Hi, Wi, To, Ho, Wo;
  int st, sy, sx, pt, py, px, kt, ky, kx;
  int bw, bh, bb;                       // output box, bw * bh * bb == 128
  int tiles_x, tiles_y, tiles_b, tiles_n;
  int nkb, Npad, N, stages;
  float* out; int cstride;              // fp32 rows (may be null when only planes are written); already offset to the first column
  double* stats;                        // [B][N][2] or null
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;   // operand planes (optional), same [voxel][cstride] geometry
  const float* scale; const float* shift; int relu;
};

template <int BN, int NSPLIT>
__global__ void __launch_bounds__(C3T_THREADS, 1)
conv3d_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo, const C3TArgs a) {
  constexpr int A_BYTES = C3T_BM * C3T_BK * 2;
  constexpr int W_BYTES = BN * C3T_BK * 2;
  constexpr int NPLANES = NSPLIT == 3 ? 2 : 1;
  constexpr int STAGE_BYTES = NPLANES * (A_BYTES + W_BYTES);
  constexpr uint32_t IDESC = umma_idesc_bf16(C3T_BM, BN);
  constexpr int MAX_STAGES = 8;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  constexpr int HALF_COLS = BN / 2;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* epi_stage = smem + (size_t)a.stages * STAGE_BYTES;
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stages = a.stages;
  const int txy = a.tiles_x * a.tiles_y;
  const int tiles_m = a.tiles_b * a.To * txy;
  const int total_tiles = tiles_m * a.tiles_n;
  const int ntaps = a.kt * a.ky * a.kx, kyx = a.ky * a.kx;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], C3T_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();
  pdl_trigger();

  // tile id -> (m tile, n tile), n fastest; m tile -> (batch tile, output time step, y tile, x tile)
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile / a.tiles_n, nt = tile - mt * a.tiles_n;
        const int tb = mt / (a.To * txy), r1 = mt - tb * (a.To * txy);
        const int t = r1 / txy, r2 = r1 - t * txy;
        const int ty = r2 / a.tiles_x, tx = r2 - ty * a.tiles_x;
        const int b0 = tb * a.bb, n0 = nt * BN;
        const int xi0 = tx * a.bw * a.sx - a.px, yi0 = ty * a.bh * a.sy - a.py, ti0 = t * a.st - a.pt;
        for (int tap = 0; tap < ntaps; ++tap) {
          const int dt = tap / kyx, rr = tap - dt * kyx, dy = rr / a.kx, dx = rr - dy * a.kx;
          const int ti = ti0 + dt;
          if (ti < 0 || ti >= a.Ti) continue;                  // this tap only sees temporal padding for the whole tile
          const int wrow = tap * a.Npad + n0;
          for (int kb = 0; kb < a.nkb; ++kb) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* stp = smem + (size_t)s * STAGE_BYTES;
            mbar_expect_tx(&full_bar[s], STAGE_BYTES);
            tma_load_5d(stp, &tmA_hi, &full_bar[s], kb * C3T_BK, xi0 + dx, yi0 + dy, ti, b0);
            tma_load_2d(stp + NPLANES * A_BYTES, &tmW_hi, &full_bar[s], kb * C3T_BK, wrow);
            if (NSPLIT == 3) {
              tma_load_5d(stp + A_BYTES, &tmA_lo, &full_bar[s], kb * C3T_BK, xi0 + dx, yi0 + dy, ti, b0);
              tma_load_2d(stp + NPLANES * A_BYTES + W_BYTES, &tmW_lo, &full_bar[s], kb * C3T_BK, wrow);
            }
            if (++s == stages) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile / a.tiles_n;
        const int t = (mt % (a.To * txy)) / txy;
        const int ti0 = t * a.st - a.pt;
        // taps with an in-range input time step: dt in [dt_lo, dt_hi)
        const int dt_lo = max(0, -ti0), dt_hi = min(a.kt, a.Ti - ti0);
        const int iters = max(0, dt_hi - dt_lo) * kyx * a.nkb;
        mbar_wait(&tmem_empty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(as * BN);
        for (int it = 0; it < iters; ++it) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
          const uint32_t sw = sa + NPLANES * A_BYTES;
          const uint64_t da_hi = umma_desc_sw128(sa), dw_hi = umma_desc_sw128(sw);
#pragma unroll
          for (int k = 0; k < C3T_BK / 16; ++k) {
            const uint64_t koff = (uint64_t)((k * 32) >> 4);
            umma_bf16(tacc, da_hi + koff, dw_hi + koff, IDESC, (it > 0 || k > 0) ? 1u : 0u);
            if (NSPLIT == 3) {
              const uint64_t da_lo = umma_desc_sw128(sa + A_BYTES), dw_lo = umma_desc_sw128(sw + W_BYTES);
              umma_bf16(tacc, da_lo + koff, dw_hi + koff, IDESC, 1u);
              umma_bf16(tacc, da_hi + koff, dw_lo + koff, IDESC, 1u);
            }
          }
          umma_commit(&empty_bar[s]);
          if (it == iters - 1) umma_commit(&tmem_full_bar[as]);
          if (++s == stages) { s = 0; ph ^= 1; }
        }
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: warps 2..9; TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 ============
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int xl = r % a.bw, yl = (r / a.bw) % a.bh, bl = r / (a.bw * a.bh);
    int as = 0;
    uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = tile / a.tiles_n, nt = tile - mt * a.tiles_n;
      const int tb = mt / (a.To * txy), r1 = mt - tb * (a.To * txy);
      const int t = r1 / txy, r2 = r1 - t * txy;
      const int ty = r2 / a.tiles_x, tx = r2 - ty * a.tiles_x;
      const int b = tb * a.bb + bl, y = ty * a.bh + yl, x = tx * a.bw + xl;
      const int n0 = nt * BN;
      const bool valid = (b < a.B) && (y < a.Ho) && (x < a.Wo);
      const size_t opix = (((size_t)b * a.To + t) * a.Ho + (size_t)y) * a.Wo + (size_t)x;
      unsigned long long trow[8];
      {
        const unsigned long long mine = valid ? (unsigned long long)opix : ~0ull;
#pragma unroll
        for (int i = 0; i < 8; ++i) trow[i] = __shfl_sync(0xffffffffu, mine, (lane >> 3) + 4 * i);
      }
      mbar_wait(&tmem_full_bar[as], aph);
      tc_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
      for (int c = half * HALF_COLS; c < (half + 1) * HALF_COLS; c += 32) {
        if (n0 + c >= a.N) break;
        const int ncols = min(32, a.N - (n0 + c));             // N is a multiple of 4: whole 16-byte segments
        float v[32];
        {
          uint32_t rr[32];
          tmem_ld32(tacc + (uint32_t)c, rr);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = valid ? __uint_as_float(rr[j]) : 0.f;
        }
        if (a.scale != nullptr || a.shift != nullptr) {        // folded BatchNorm / bias, warp-uniform branch
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = min(n0 + c + j, a.N - 1);
            const float sc = a.scale ? __ldg(a.scale + n) : 1.0f, sh = a.shift ? __ldg(a.shift + n) : 0.0f;
            v[j] = fmaf(v[j], sc, sh);
          }
        }
        if (a.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        uint8_t* stg = epi_stage + (size_t)(warp - 2) * C3T_EPI_STAGE_BYTES + lane * 128;
        const int sw = lane & 7;
#pragma unroll
        for (int j = 0; j < 8; ++j) *(float4*)(stg + ((j ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        const uint8_t* rd = epi_stage + (size_t)(warp - 2) * C3T_EPI_STAGE_BYTES;
        const int seg = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = (lane >> 3) + 4 * i;
          if (trow[i] == ~0ull) continue;
          const uint4 dv = *(const uint4*)(rd + row * 128 + ((seg ^ (row & 7)) << 4));
          if (seg * 4 < ncols) {
            const size_t o = (size_t)trow[i] * a.cstride + n0 + c + seg * 4;
            if (a.out) *(uint4*)(a.out + o) = dv;
            if (a.out_hi) {
              const float f0 = __uint_as_float(dv.x), f1 = __uint_as_float(dv.y), f2 = __uint_as_float(dv.z), f3 = __uint_as_float(dv.w);
              const __nv_bfloat162 h01 = __floats2bfloat162_rn(f0, f1), h23 = __floats2bfloat162_rn(f2, f3);
              *(uint2*)(a.out_hi + o) = make_uint2(*(const uint32_t*)&h01, *(const uint32_t*)&h23);
              if (a.out_lo) {
                const float2 g01 = __bfloat1622float2(h01), g23 = __bfloat1622float2(h23);
                const __nv_bfloat162 l01 = __floats2bfloat162_rn(f0 - g01.x, f1 - g01.y), l23 = __floats2bfloat162_rn(f2 - g23.x, f3 - g23.y);
                *(uint2*)(a.out_lo + o) = make_uint2(*(const uint32_t*)&l01, *(const uint32_t*)&l23);
              }
            }
          }
        }
        if (a.stats != nullptr) {
          // the 32 rows of this warp belong to one sample (bw*bh >= 32): lane = column, sums straight from the staging rows
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int row = 0; row < 32; ++row) {
            const float tv = *(const float*)(rd + row * 128 + (((lane >> 2) ^ (row & 7)) << 4) + (lane & 3) * 4);
            s1 += tv;
            s2 = fmaf(tv, tv, s2);
          }
          const int bw0 = __shfl_sync(0xffffffffu, b, 0);
          if (lane < ncols && n0 + c + lane < a.N && bw0 < a.B) {
            double* sp = a.stats + ((size_t)bw0 * a.N + n0 + c + lane) * 2;
            atomicAdd(sp, (double)s1);
            atomicAdd(sp + 1, (double)s2);
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
      if (++as == 2) { as = 0; aph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn3 get_encode3() {
  static EncodeTiledFn3 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn3)p;
  });
  IPK_CHECK(fn != nullptr, IPK_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  return fn;
}

static CUtensorMap encode_map(const void* base, int rank, const cuuint64_t* gd, const cuuint64_t* gs, const cuuint32_t* bx, const cuuint32_t* es) {
  CUtensorMap m;
  CUresult r = get_encode3()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IPK_CHECK(r == CUDA_SUCCESS, IPK_ERR_CUDA, "conv3d_tc: cuTensorMapEncodeTiled failed (%d), rank %d box [%u,%u,%u,%u,%u] strides [%u,%u,%u]", (int)r, rank,
            bx[0], bx[1], rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0, rank > 4 ? bx[4] : 0, es[1], rank > 2 ? es[2] : 0, rank > 3 ? es[3] : 0);
  return m;
}

static int sm_count3() {
  static int n[IPK_MAX_DEVICES] = {0};
  const int slot = current_device_slot();
  if (n[slot] == 0) {
    int dev = 0;
    IPK_CUDA(cudaGetDevice(&dev));
    IPK_CUDA(cudaDeviceGetAttribute(&n[slot], cudaDevAttrMultiProcessorCount, dev));
  }
  return n[slot];
}

template <int BN, int NSPLIT>
static void launch_c3t(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi, const CUtensorMap& w_lo, C3TArgs& a, cudaStream_t st) {
  constexpr int STAGE_BYTES = (NSPLIT == 3 ? 2 : 1) * (C3T_BM * C3T_BK * 2 + BN * C3T_BK * 2);
  a.stages = (int)std::min<size_t>(8, C3T_SMEM_BUDGET / STAGE_BYTES);
  const size_t smem = (size_t)a.stages * STAGE_BYTES + 1024 + C3T_EPI_WARPS * C3T_EPI_STAGE_BYTES;
  static bool attr_set[IPK_MAX_DEVICES] = {false};
  const int slot = current_device_slot();
  if (!attr_set[slot]) {
    IPK_CUDA(cudaFuncSetAttribute(conv3d_tc_kernel<BN, NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(C3T_SMEM_BUDGET + 1024 + C3T_EPI_WARPS * C3T_EPI_STAGE_BYTES)));
    attr_set[slot] = true;
  }
  const long long total = (long long)a.tiles_b * a.To * a.tiles_y * a.tiles_x * a.tiles_n;
  const unsigned grid = (unsigned)std::min<long long>(total, sm_count3());
  launch_k(conv3d_tc_kernel<BN, NSPLIT>, dim3(grid), dim3(C3T_THREADS), smem, st, a_hi, a_lo, w_hi, w_lo, a);
}

// w: OIDHW [Cout][CinSrc][ntaps] fp32 -> dst planes [tap][Npad][Kpad] (hi [, lo]); the buffers come zeroed from conv_alloc
__global__ void pack_conv3d_tc_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int Cout, int CinSrc,
                                      int ntaps, int Npad, int Kpad) {
  const long long total = (long long)ntaps * Cout * CinSrc;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e % CinSrc), n = (int)((e / CinSrc) % Cout), tap = (int)(e / ((long long)CinSrc * Cout));
    const float v = w[((size_t)n * CinSrc + k) * ntaps + tap];
    const size_t di = ((size_t)tap * Npad + n) * Kpad + k;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[di] = h;
    if (lo) lo[di] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}
void conv3d_tc_pack(ConvW& dst, const float* w_oidhw, int Cout, int CinSrc, cudaStream_t st) {
  IPK_CHECK(dst.w_hi && dst.N == Cout && dst.K >= CinSrc, IPK_ERR_INVALID, "conv3d_tc_pack: destination does not match");
  const long long total = (long long)dst.ntaps * Cout * CinSrc;
  pack_conv3d_tc_kernel<<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, st>>>(w_oidhw, dst.w_hi, dst.w_lo, Cout, CinSrc, dst.ntaps,
                                                                                                  dst.Npad, dst.Kpad);
  IPK_LAUNCH_CHECK();
}

static void out_extent(const Conv3dShape& s, int& To, int& Ho, int& Wo) {
  To = s.To > 0 ? s.To : (s.Ti + 2 * s.pt - s.kt) / s.st + 1;
  Ho = s.Ho > 0 ? s.Ho : (s.Hi + 2 * s.py - s.ky) / s.sy + 1;
  Wo = s.Wo > 0 ? s.Wo : (s.Wi + 2 * s.px - s.kx) / s.sx + 1;
}

bool conv3d_tc_supported(const Conv3dShape& s) {
  int To, Ho, Wo;
  out_extent(s, To, Ho, Wo);
  if (s.Cin % 64 != 0 || s.Cout % 64 != 0) return false;
  if (!(C3T_BM % Wo == 0 || Wo % C3T_BM == 0)) return false;
  const int bw = std::min(Wo, C3T_BM), rest = C3T_BM / bw;
  if (rest > 1 && !(rest % Ho == 0 || Ho % rest == 0)) return false;
  const int bh = std::min(Ho, rest);
  if (bw * bh < 32) return false;                          // the fused statistics need a warp's 32 rows inside one sample
  if ((bw - 1) * s.sx + 1 > 256 || (bh - 1) * s.sy + 1 > 256 || s.sx > 8 || s.sy > 8) return false;
  return true;
}
bool conv3d_tc_supported_ex(const Conv3dShape& s) {
  return s.Cin % 8 == 0 && s.Cout % 8 == 0 && s.sx >= 1 && s.sx <= 8 && s.sy >= 1 && s.sy <= 8;
}

// 128-voxel output box (bw x bh x bb, powers of two) covering [Wo][Ho][B] with the fewest out-of-range voxels; ties -> widest rows
static void choose_box(int Wo, int Ho, int B, int sx, int sy, bool need_stats, int& bw, int& bh, int& bb) {
  double best = 1e300;
  bw = 0;
  for (int w = C3T_BM; w >= 1; w >>= 1)
    for (int h = C3T_BM / w; h >= 1; h >>= 1) {
      const int b = C3T_BM / (w * h);
      if ((w - 1) * sx + 1 > 256 || (h - 1) * sy + 1 > 256 || b > 256) continue;
      if (need_stats && w * h < 32) continue;
      const double cover = (double)cdiv(Wo, w) * w * cdiv(Ho, h) * h * cdiv(B, b) * b;
      if (cover < best) { best = cover; bw = w; bh = h; bb = b; }
    }
  IPK_CHECK(bw > 0, IPK_ERR_UNSUPPORTED, "conv3d_tc: no 128-voxel box fits output %d x %d (strides %d, %d)", Wo, Ho, sx, sy);
}

void conv3d_tc_run_ex(const ConvW& w, const Conv3dShape& s, const void* in_hi, const void* in_lo, int B, const Conv3dEpi& epi, cudaStream_t st) {
  IPK_CHECK(conv3d_tc_supported_ex(s), IPK_ERR_UNSUPPORTED, "conv3d_tc_run: shape not supported by the tensor-core engine");
  IPK_CHECK(w.w_hi != nullptr && w.ntaps == s.kt * s.ky * s.kx && w.K == s.Cin && w.N == s.Cout, IPK_ERR_STATE, "conv3d_tc_run: packed weights do not match the layer");
  const bool split = w.engine == IPK_PREC_FP32_SPLIT;
  IPK_CHECK(!split || (in_lo && w.w_lo), IPK_ERR_STATE, "conv3d_tc_run: split precision needs hi and lo operand planes");
  IPK_CHECK(epi.out_f32 || epi.out_hi, IPK_ERR_INVALID, "conv3d_tc_run: no output");
  IPK_CHECK(epi.cstride % 4 == 0 && epi.coff % 4 == 0 && epi.coff + s.Cout <= epi.cstride, IPK_ERR_INVALID, "conv3d_tc_run: output slice (%d + %d of %d)", epi.coff, s.Cout, epi.cstride);
  C3TArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.Ti = s.Ti; a.Hi = s.Hi; a.Wi = s.Wi;
  out_extent(s, a.To, a.Ho, a.Wo);
  a.st = s.st; a.sy = s.sy; a.sx = s.sx; a.pt = s.pt; a.py = s.py; a.px = s.px; a.kt = s.kt; a.ky = s.ky; a.kx = s.kx;
  if (conv3d_tc_supported(s)) {          // the encoder's shapes keep their boxes (a warp's 32 rows inside one sample: fused statistics)
    a.bw = std::min(a.Wo, C3T_BM);
    a.bh = std::min(a.Ho, C3T_BM / a.bw);
    a.bb = C3T_BM / (a.bw * a.bh);
  } else {
    IPK_CHECK(epi.stats == nullptr, IPK_ERR_UNSUPPORTED, "conv3d_tc_run: fused statistics need a power-of-two output geometry");
    choose_box(a.Wo, a.Ho, B, s.sx, s.sy, false, a.bw, a.bh, a.bb);
  }
  a.tiles_x = cdiv(a.Wo, a.bw); a.tiles_y = cdiv(a.Ho, a.bh); a.tiles_b = cdiv(B, a.bb);
  a.nkb = w.Kpad / C3T_BK; a.Npad = w.Npad; a.N = w.N;
  a.out = epi.out_f32 ? epi.out_f32 + epi.coff : nullptr; a.cstride = epi.cstride; a.stats = epi.stats;
  a.out_hi = epi.out_hi ? epi.out_hi + epi.coff : nullptr; a.out_lo = epi.out_lo ? epi.out_lo + epi.coff : nullptr;
  a.scale = epi.scale; a.shift = epi.shift; a.relu = epi.relu;
  // activation map: dims (C, W, H, T, B); element strides carry the spatial conv stride
  const cuuint64_t cs = s.stride_x > 0 ? (cuuint64_t)s.stride_x : (cuuint64_t)s.Cin * 2;
  const cuuint64_t rs = s.stride_y > 0 ? (cuuint64_t)s.stride_y : cs * s.Wi;
  cuuint64_t gd[5] = {(cuuint64_t)s.Cin, (cuuint64_t)s.Wi, (cuuint64_t)s.Hi, (cuuint64_t)s.Ti, (cuuint64_t)B};
  cuuint64_t gs[4] = {cs, rs, rs * s.Hi, rs * s.Hi * s.Ti};
  cuuint32_t bx[5] = {(cuuint32_t)C3T_BK, (cuuint32_t)((a.bw - 1) * s.sx + 1), (cuuint32_t)((a.bh - 1) * s.sy + 1), 1u, (cuuint32_t)a.bb};
  cuuint32_t es[5] = {1u, (cuuint32_t)s.sx, (cuuint32_t)s.sy, 1u, 1u};
  const CUtensorMap mA_hi = encode_map(in_hi, 5, gd, gs, bx, es);
  const CUtensorMap mA_lo = split ? encode_map(in_lo, 5, gd, gs, bx, es) : mA_hi;
  // N tile: the widest that still gives every SM a tile; padded columns cost MMA work, every extra N tile one more read of the A tile
  const long long tiles_m = (long long)a.tiles_b * a.To * a.tiles_y * a.tiles_x;
  int BN = 64;
  {
    double best = 1e300;
    for (int bn : {256, 128, 64}) {
      const int tn = cdiv(w.Npad, bn);
      double cost = (double)tn * bn * (1.0 + 32.0 / bn);
      if (tiles_m * tn < sm_count3()) cost *= 2.0;          // too few tiles to fill the machine
      if (cost < best) { best = cost; BN = bn; }
    }
  }
  a.tiles_n = cdiv(w.Npad, BN);
  cuuint64_t wd[2] = {(cuuint64_t)w.Kpad, (cuuint64_t)w.ntaps * w.Npad};
  cuuint64_t wsb[1] = {(cuuint64_t)w.Kpad * 2};
  cuuint32_t wb[2] = {(cuuint32_t)C3T_BK, (cuuint32_t)BN};
  cuuint32_t we[2] = {1u, 1u};
  const CUtensorMap mW_hi = encode_map(w.w_hi, 2, wd, wsb, wb, we);
  const CUtensorMap mW_lo = split ? encode_map(w.w_lo, 2, wd, wsb, wb, we) : mW_hi;
  switch (BN) {
    case 64: if (split) launch_c3t<64, 3>(mA_hi, mA_lo, mW_hi, mW_lo, a, st); else launch_c3t<64, 1>(mA_hi, mA_lo, mW_hi, mW_lo, a, st); break;
    case 128: if (split) launch_c3t<128, 3>(mA_hi, mA_lo, mW_hi, mW_lo, a, st); else launch_c3t<128, 1>(mA_hi, mA_lo, mW_hi, mW_lo, a, st); break;
    default: if (split) launch_c3t<256, 3>(mA_hi, mA_lo, mW_hi, mW_lo, a, st); else launch_c3t<256, 1>(mA_hi, mA_lo, mW_hi, mW_lo, a, st); break;
  }
}

void conv3d_tc_run(const ConvW& w, const Conv3dShape& s, const void* in_hi, const void* in_lo, int B, float* out, double* stats, cudaStream_t st) {
  IPK_CHECK(conv3d_tc_supported(s), IPK_ERR_UNSUPPORTED, "conv3d_tc_run: shape not supported by the tensor-core engine");
  Conv3dEpi e;
  e.out_f32 = out; e.cstride = s.Cout; e.stats = stats;
  conv3d_tc_run_ex(w, s, in_hi, in_lo, B, e, st);
}

}  // namespace ipk
