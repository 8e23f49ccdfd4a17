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
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include "conv.cuh"
#include "tc_ptx.cuh"

namespace ipk {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr size_t TC_SMEM_BUDGET = 192 * 1024;      // pipeline stages
constexpr size_t TC_EPI_STAGE_BYTES = 4096;          // per epilogue warp: 32 rows x 128 B transpose buffer (xor-swizzled)

constexpr int TC_MAX_SUB = 4;
struct TcSub {                    // one tap list + output phase (a parity class of a transposed conv; plain convs have one)
  int ntaps, yadd, xadd;
  int dy[MAX_TAPS], dx[MAX_TAPS], widx[MAX_TAPS];
};
struct TcOut {                    // destination of a column range
  void* out; void* out_lo;
  int act, mode, cstride, coff;
};
struct TcArgs {
  int F, H, W;
  int bw, bh, bf;                 // pixel box (bw*bh*bf == 128)
  int tiles_x, tiles_y;           // tiles along x and y
  int tiles_m, tiles_n, nsplit;   // tiles_m = tiles_x * tiles_y * tiles_f
  int nsub;                       // sub-convolutions in this launch (nsplit == 1 when > 1)
  int iters_per_split, nkb;       // nkb = Kpad / 64; a split covers iters_per_split consecutive (tap, k-block) iterations
  TcSub sub[TC_MAX_SUB];
  int Npad, N;
  int stages;
  // epilogue
  const float* bias;
  int Ho, Wo, ymul, xmul;
  int split_col;                  // columns >= split_col (when > 0) go to o[1]
  TcOut o[2];
  long long split_stride;
  // fused residual branch + output statistics (see ConvOut)
  const float* res; int res_cstride, res_act; const float* res_mr; double* stats;
  int stats_oi, stats_C;          // statistics cover the columns of destination o[stats_oi], stats_C channels per frame
  int halo_variant;               // HALO kernels: 1 = row-shifted descriptors carry the swizzle base offset, 2 = they do not
};

// PTX wrappers (mbarrier, TMA, UMMA descriptors, TMEM loads): tc_ptx.cuh

// ------------------------------------------------------------------------------------------------ kernel
// activation applied to a register tile; the branch is warp-uniform.  ELU uses ex2.approx (abs error ~1e-7, far inside
// the engine's own operand rounding); the fp32 SIMT validation engine keeps expm1f.
template <int NV>
__device__ __forceinline__ void act_tile(float (&v)[NV], int act) {
  if (act == ACT_ELU) {
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = v[j] > 0.f ? v[j] : (exp2f(v[j] * 1.4426950408889634f) - 1.0f);
  } else if (act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (act == ACT_LRELU02) {
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
  } else if (act != ACT_NONE) {
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = act_apply(v[j], act);
  }
}

// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue.
// Two TMEM accumulator stages (2 x BN columns): the epilogue of tile i overlaps the main loop of tile i+1.
// Tile order: linear id -> (split z, m tile, n tile) with n fastest, so CTAs running side by side share the A tile in L2.
// FUSED: residual branch / output statistics in the epilogue (see ConvOut).
// HALO (3x3 convs on 128-wide images, one image row per tile): the producer loads each input row ONCE as a 130-pixel box
// (x = -1 .. 128, zero-filled outside) and the three dx taps are row-shifted views of that box (UMMA descriptor start advanced by
// dx * 128 bytes) instead of three separate TMA loads: A traffic / 3.
constexpr int TC_HALO_A_BYTES = 17 * 1024;      // 130 rows x 128 B = 16 640 B, padded to keep the regions 1024-byte aligned
// CG = 2 (CTA pair, launched as clusters of 2): the two CTAs own two consecutive M tiles of the same N tile; each loads its own A
// tile and HALF of the W tile (BN / 2 rows), and the pair's leader issues cta_group::2 MMAs (M = 256, N = BN) that read both CTAs'
// shared memory -- every W byte is fetched from L2 once per pair instead of once per CTA, which is what bounds these kernels
// (measured: NICE conv2 sits at the chip's TMA throughput, not at the tensor pipe).  Each CTA's TMEM holds its own 128 rows x BN
// columns, so the epilogue is the single-CTA one.
template <int BN, int NSPLIT, bool FUSED, bool HALO, int CG>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo, const TcArgs a) {
  constexpr int A_BYTES = TC_BM * TC_BK * 2;         // 16 KB
  constexpr int WROWS = BN / CG;                     // W rows this CTA stages
  constexpr int W_BYTES = WROWS * TC_BK * 2;
  constexpr int NPLANES = NSPLIT == 3 ? 2 : 1;
  constexpr int STAGE_BYTES = HALO ? NPLANES * (TC_HALO_A_BYTES + 3 * W_BYTES) : NPLANES * (A_BYTES + W_BYTES);
  constexpr uint32_t IDESC = umma_idesc_bf16(TC_BM * CG, BN);
  constexpr int MAX_STAGES = 8;
  constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  constexpr int EPI_CHUNK = BN >= 64 ? 32 : 16;      // columns per TMEM load
  constexpr int HALF_COLS = BN / 2;                  // columns owned by one of the two epilogue warps of a lane quarter

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B needs 1024-byte alignment
  uint8_t* epi_stage = smem + (size_t)a.stages * STAGE_BYTES;                      // 8 x 4 KB epilogue transpose buffers
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stages = a.stages;
  const int txy = a.tiles_x * a.tiles_y;
  // work units: (split / sub-convolution z, M-tile group, N tile); a group is CG consecutive M tiles, one per CTA of the pair
  const int tiles_mg = (a.tiles_m + CG - 1) / CG;
  const int tiles_mn = tiles_mg * a.tiles_n;
  const int total_tiles = tiles_mn * (a.nsub > 1 ? a.nsub : a.nsplit);
  const uint32_t crank = CG == 2 ? cluster_ctarank() : 0u;
  const int unit0 = (int)blockIdx.x / CG, unit_step = (int)gridDim.x / CG;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], CG * TC_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation by one full warp (of each CTA of a pair); the same warp deallocates
    if constexpr (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();     // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();        // prologue above overlapped the previous kernel's tail; its outputs are visible from here on
  pdl_trigger();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      // loads of either CTA of a pair complete on the LEADER's full barrier, which expects the bytes of both
      auto lda = [&](void* dst, const CUtensorMap* tm, int st_i, int c0, int c1, int c2, int c3) {
        if constexpr (CG == 2) tma_load_4d_2cta(dst, tm, mapa_u32(smem_u32(&full_bar[st_i]), 0), c0, c1, c2, c3);
        else tma_load_4d(dst, tm, &full_bar[st_i], c0, c1, c2, c3);
      };
      auto ldw = [&](void* dst, const CUtensorMap* tm, int st_i, int c0, int c1) {
        if constexpr (CG == 2) tma_load_2d_2cta(dst, tm, mapa_u32(smem_u32(&full_bar[st_i]), 0), c0, c1);
        else tma_load_2d(dst, tm, &full_bar[st_i], c0, c1);
      };
      for (int tile = unit0; tile < total_tiles; tile += unit_step) {
        const int z = tile / tiles_mn, rem = tile - z * tiles_mn;      // z: split-K slice, or sub-convolution when nsub > 1
        const int mg = rem / a.tiles_n, nt = rem - mg * a.tiles_n;
        const int mt = mg * CG + (int)crank;                            // beyond tiles_m: every box is out of bounds -> zero fill
        const int tf = mt / txy, r2 = mt - tf * txy;
        const int ty = r2 / a.tiles_x, tx = r2 - ty * a.tiles_x;
        const int f0 = tf * a.bf, y0 = ty * a.bh, x0 = tx * a.bw, n0 = nt * BN + (int)crank * WROWS;
        const TcSub& sb = a.sub[a.nsub > 1 ? z : 0];
        if constexpr (HALO) {
          // one stage per (input row dy, k-block): the 130-pixel row box of both planes + the weights of its three dx taps
          for (int dyi = 0; dyi < 3; ++dyi)
            for (int kb = 0; kb < a.nkb; ++kb) {
              mbar_wait(&empty_bar[s], ph ^ 1);
              uint8_t* st = smem + (size_t)s * STAGE_BYTES;
              if (crank == 0) mbar_expect_tx(&full_bar[s], CG * NPLANES * (130 * 128 + 3 * W_BYTES));
              lda(st, &tmA_hi, s, kb * TC_BK, -1, y0 + dyi - 1, f0);
              if (NSPLIT == 3) lda(st + TC_HALO_A_BYTES, &tmA_lo, s, kb * TC_BK, -1, y0 + dyi - 1, f0);
              uint8_t* wst = st + NPLANES * TC_HALO_A_BYTES;
              for (int dxi = 0; dxi < 3; ++dxi) {
                const int wrow = (dyi * 3 + dxi) * a.Npad + n0;
                ldw(wst + dxi * W_BYTES, &tmW_hi, s, kb * TC_BK, wrow);
                if (NSPLIT == 3) ldw(wst + (3 + dxi) * W_BYTES, &tmW_lo, s, kb * TC_BK, wrow);
              }
              if (++s == stages) { s = 0; ph ^= 1; }
            }
          continue;
        }
        const int it_begin = a.nsub > 1 ? 0 : z * a.iters_per_split;
        const int it_end = min(sb.ntaps * a.nkb, it_begin + a.iters_per_split);
        int t = it_begin / a.nkb, kb = it_begin - t * a.nkb;
        for (int it = it_begin; it < it_end; ++it) {
          const int dy = sb.dy[t], dx = sb.dx[t], wrow = sb.widx[t] * a.Npad + n0;
          {
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* st = smem + (size_t)s * STAGE_BYTES;
            if (crank == 0) mbar_expect_tx(&full_bar[s], CG * STAGE_BYTES);
            lda(st, &tmA_hi, s, kb * TC_BK, x0 + dx, y0 + dy, f0);
            ldw(st + NPLANES * A_BYTES, &tmW_hi, s, kb * TC_BK, wrow);
            if (NSPLIT == 3) {
              lda(st + A_BYTES, &tmA_lo, s, kb * TC_BK, x0 + dx, y0 + dy, f0);
              ldw(st + NPLANES * A_BYTES + W_BYTES, &tmW_lo, s, kb * TC_BK, wrow);
            }
            if (++s == stages) { s = 0; ph ^= 1; }
          }
          if (++kb == a.nkb) { kb = 0; ++t; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (the pair's leader only when CG = 2) =====================
    if (lane == 0 && crank == 0) {
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t acc) {
        if constexpr (CG == 2) umma_bf16_2cta(d, da, db, IDESC, acc);
        else umma_bf16(d, da, db, IDESC, acc);
      };
      auto commit = [&](uint64_t* bar) {
        if constexpr (CG == 2) umma_commit_2cta(bar);
        else umma_commit(bar);
      };
      for (int tile = unit0; tile < total_tiles; tile += unit_step) {
        const int z = tile / tiles_mn;
        const int it_begin = a.nsub > 1 ? 0 : z * a.iters_per_split;
        const int iters = HALO ? 3 * a.nkb : min(a.sub[a.nsub > 1 ? z : 0].ntaps * a.nkb, it_begin + a.iters_per_split) - it_begin;
        mbar_wait(&tmem_empty_bar[as], aph ^ 1);      // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(as * BN);
        if constexpr (HALO) {
          for (int it = 0; it < iters; ++it) {
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
            const uint32_t sw = sa + NPLANES * TC_HALO_A_BYTES;
#pragma unroll
            for (int dxi = 0; dxi < 3; ++dxi) {
              // rows dxi .. dxi+127 of the 130-row box: output pixel x reads input x + dx = box row x + dxi
              const uint32_t boff = a.halo_variant == 1 ? (uint32_t)dxi : 0u;
              const uint64_t da_hi = umma_desc_sw128_off(sa + dxi * 128, boff);
              const uint64_t dw_hi = umma_desc_sw128(sw + dxi * W_BYTES);
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k) {
                const uint64_t koff = (uint64_t)((k * 32) >> 4);
                mma(tacc, da_hi + koff, dw_hi + koff, (it > 0 || dxi > 0 || k > 0) ? 1u : 0u);
                if (NSPLIT == 3) {
                  const uint64_t da_lo = umma_desc_sw128_off(sa + TC_HALO_A_BYTES + dxi * 128, boff);
                  const uint64_t dw_lo = umma_desc_sw128(sw + (3 + dxi) * W_BYTES);
                  mma(tacc, da_lo + koff, dw_hi + koff, 1u);
                  mma(tacc, da_hi + koff, dw_lo + koff, 1u);
                }
              }
            }
            commit(&empty_bar[s]);
            if (it == iters - 1) commit(&tmem_full_bar[as]);
            if (++s == stages) { s = 0; ph ^= 1; }
          }
          if (++as == 2) { as = 0; aph ^= 1; }
          continue;
        }
        for (int it = 0; it < iters; ++it) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
          const uint32_t sw = sa + NPLANES * A_BYTES;
          const uint64_t da_hi = umma_desc_sw128(sa), dw_hi = umma_desc_sw128(sw);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t koff = (uint64_t)((k * 32) >> 4);   // 16 bf16 = 32 bytes along K inside the swizzle atom
            mma(tacc, da_hi + koff, dw_hi + koff, (it > 0 || k > 0) ? 1u : 0u);
            if (NSPLIT == 3) {
              const uint64_t da_lo = umma_desc_sw128(sa + A_BYTES), dw_lo = umma_desc_sw128(sw + W_BYTES);
              mma(tacc, da_lo + koff, dw_hi + koff, 1u);
              mma(tacc, da_hi + koff, dw_lo + koff, 1u);
            }
          }
          commit(&empty_bar[s]);                      // frees the stage (in both CTAs of a pair) once the MMAs above have read it
          if (it == iters - 1) commit(&tmem_full_bar[as]);
          if (++s == stages) { s = 0; ph ^= 1; }
        }
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: warps 2..9; TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 ============
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;                    // accumulator row == pixel index inside the box
    const int xl = r % a.bw, yl = (r / a.bw) % a.bh, fl = r / (a.bw * a.bh);
    int as = 0;
    uint32_t aph = 0;
    // the accumulator stage goes back to the MMA issuer: the leader's barrier collects the epilogue warps of both CTAs of a pair
    const uint32_t tmem_empty_addr0 = CG == 2 ? mapa_u32(smem_u32(&tmem_empty_bar[0]), 0) : 0u;
    for (int tile = unit0; tile < total_tiles; tile += unit_step) {
      const int z = tile / tiles_mn, rem = tile - z * tiles_mn;
      const int mg = rem / a.tiles_n, nt = rem - mg * a.tiles_n;
      const int mt = mg * CG + (int)crank;
      const int tf = mt / txy, r2 = mt - tf * txy;
      const int ty = r2 / a.tiles_x, tx = r2 - ty * a.tiles_x;
      const int f = tf * a.bf + fl, y = ty * a.bh + yl, x = tx * a.bw + xl;
      const int n0 = nt * BN;
      const bool valid = (mt < a.tiles_m) && (f < a.F) && (y < a.H) && (x < a.W);
      const TcSub& sb = a.sub[a.nsub > 1 ? z : 0];
      const int oy = y * a.ymul + sb.yadd, ox = x * a.xmul + sb.xadd;
      const size_t opix = ((size_t)f * a.Ho + (size_t)oy) * a.Wo + (size_t)ox;
      const size_t zoff = a.nsub > 1 ? 0 : (size_t)z * a.split_stride;

      // rows this lane stores in the transposed (coalesced) write-out: row_i = lane/8 + 4*i; their output pixels come from
      // the lanes that own them (all-ones = row outside the image)
      unsigned long long trow[8];
      if constexpr (EPI_CHUNK == 32) {
        const unsigned long long mine = valid ? (unsigned long long)opix : ~0ull;
#pragma unroll
        for (int i = 0; i < 8; ++i) trow[i] = __shfl_sync(0xffffffffu, mine, (lane >> 3) + 4 * i);
      }

      mbar_wait(&tmem_full_bar[as], aph);
      tc_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
      for (int c = half * HALF_COLS; c < (half + 1) * HALF_COLS; c += EPI_CHUNK) {
        if (n0 + c >= a.Npad) break;                  // warp-uniform
        const int ncols = min(EPI_CHUNK, a.Npad - (n0 + c));   // Npad is a multiple of 16: a 32-column chunk may be half valid
        float v[EPI_CHUNK];
        {
          uint32_t rr[EPI_CHUNK];
          if constexpr (EPI_CHUNK == 32) tmem_ld32(tacc + (uint32_t)c, rr);
          else tmem_ld16(tacc + (uint32_t)c, rr);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < EPI_CHUNK; ++j) v[j] = __uint_as_float(rr[j]);
        }
        const int oi = (a.split_col > 0 && n0 + c >= a.split_col) ? 1 : 0;     // warp-uniform: destination of this chunk
        const TcOut& od = a.o[oi];
        const size_t colbase = (size_t)od.coff + (size_t)(n0 + c - (oi ? a.split_col : 0));
        if (a.bias) {
          const float4* b4 = (const float4*)(a.bias + n0 + c);
#pragma unroll
          for (int j = 0; j < EPI_CHUNK / 4; ++j) {
            if (4 * j >= ncols) break;
            const float4 b = __ldg(b4 + j);
            v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
          }
        }
        act_tile<EPI_CHUNK>(v, od.act);
        if (FUSED && EPI_CHUNK == 32 && a.res != nullptr) {
          // residual branch of a ResBlock: + res_act((res - mean) * rstd), per-(frame, channel) statistics from res_mr
          float rv[EPI_CHUNK];
          if (valid) {
            const float4* rp = (const float4*)(a.res + opix * a.res_cstride + n0 + c);
#pragma unroll
            for (int j = 0; j < EPI_CHUNK / 4; ++j) {
              const float4 t4 = __ldg(rp + j);
              rv[4 * j] = t4.x; rv[4 * j + 1] = t4.y; rv[4 * j + 2] = t4.z; rv[4 * j + 3] = t4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < EPI_CHUNK; ++j) rv[j] = 0.f;
          }
          if (a.res_mr) {
            const float4* mp = (const float4*)(a.res_mr + ((size_t)min(f, a.F - 1) * a.N + n0 + c) * 2);     // (mean, rstd) pairs
#pragma unroll
            for (int j = 0; j < EPI_CHUNK / 2; ++j) {
              const float4 m4 = __ldg(mp + j);
              rv[2 * j] = (rv[2 * j] - m4.x) * m4.y;
              rv[2 * j + 1] = (rv[2 * j + 1] - m4.z) * m4.w;
            }
          }
          act_tile<EPI_CHUNK>(rv, a.res_act);
#pragma unroll
          for (int j = 0; j < EPI_CHUNK; ++j) v[j] = valid ? v[j] + rv[j] : 0.f;      // rows outside the image contribute 0 to the stats
        }
        if (FUSED && a.stats != nullptr && a.res == nullptr && !valid) {
#pragma unroll
          for (int j = 0; j < EPI_CHUNK; ++j) v[j] = 0.f;      // rows outside the image contribute 0 to the statistics
        }
        if (od.mode == OUT_F32_NCHW) {
          // frames at the ABI edge: [f][N][Ho][Wo]; consecutive lanes are consecutive x -> coalesced per channel
          if (valid) {
#pragma unroll
            for (int j = 0; j < EPI_CHUNK; ++j) {
              const int n = n0 + c + j;
              if (n < a.N) ((float*)od.out)[(((size_t)f * a.N + n) * a.Ho + oy) * a.Wo + ox] = v[j];
            }
          }
        } else if (EPI_CHUNK == 32 && od.mode == OUT_F32_NHWC) {
          // ---- fp32 rows: transpose through shared memory: lane = row on the way in, 8 lanes = one 128-byte row segment on the
          //      way out, so every global store instruction writes full, contiguous sectors (4 rows x 128 B).  (Measured: a win
          //      for fp32 outputs; for the bf16 operand planes the extra instructions cost more than the scattered 32-byte
          //      stores, so those keep the direct path below.)
          uint8_t* stg = epi_stage + (size_t)(warp - 2) * TC_EPI_STAGE_BYTES + lane * 128;
          const int sw = lane & 7;
#pragma unroll
          for (int j = 0; j < 8; ++j) *(float4*)(stg + ((j ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          __syncwarp();
          const uint8_t* rd = epi_stage + (size_t)(warp - 2) * TC_EPI_STAGE_BYTES;
          const int seg = lane & 7;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = (lane >> 3) + 4 * i;
            if (trow[i] == ~0ull) continue;
            const uint4 dv = *(const uint4*)(rd + row * 128 + ((seg ^ (row & 7)) << 4));
            const size_t o = (size_t)trow[i] * od.cstride + colbase;
            if (seg * 4 < ncols) *(uint4*)((float*)od.out + zoff + o + seg * 4) = dv;
          }
          if (FUSED && a.stats != nullptr && oi == a.stats_oi) {
            // per-channel sum / sum of squares of this warp's 32 rows (one frame): lane = column, straight from the staging rows
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int row = 0; row < 32; ++row) {
              const float t = *(const float*)(rd + row * 128 + (((lane >> 2) ^ (row & 7)) << 4) + (lane & 3) * 4);
              s1 += t;
              s2 = fmaf(t, t, s2);
            }
            const int fw = __shfl_sync(0xffffffffu, f, 0);          // frame of this warp's rows
            if (lane < ncols && fw < a.F) {
              double* sp = a.stats + ((size_t)fw * a.stats_C + (n0 + c - (oi ? a.split_col : 0)) + lane) * 2;
              atomicAdd(sp, (double)s1);
              atomicAdd(sp + 1, (double)s2);
            }
          }
          __syncwarp();       // staging rows are rewritten by the next chunk
        } else {
          // bf16 operand planes, and narrow tiles (BN = 32): direct row-per-lane stores
          if (valid) {
            const size_t ocol = opix * od.cstride + colbase;
            if (od.mode == OUT_F32_NHWC) {
              float4* p = (float4*)((float*)od.out + zoff + ocol);
#pragma unroll
              for (int j = 0; j < EPI_CHUNK / 4; ++j)
                if (4 * j < ncols) p[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
              uint32_t hi[EPI_CHUNK / 2], lo[EPI_CHUNK / 2];
#pragma unroll
              for (int j = 0; j < EPI_CHUNK / 2; ++j) {
                const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                const float2 hf = __bfloat1622float2(hh);
                const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
                hi[j] = *(const uint32_t*)&hh;
                lo[j] = *(const uint32_t*)&ll;
              }
              // a lane owns 64 contiguous bytes of its row per plane: 256-bit stores (STG.256) write whole 32-byte sectors, so L2 sees no
              // partially written sectors (the 16-byte version left every sector to be completed by a second store instruction)
              __nv_bfloat16* ph = (__nv_bfloat16*)od.out + ocol;
              __nv_bfloat16* pl = od.mode == OUT_BF16_SPLIT ? (__nv_bfloat16*)od.out_lo + ocol : nullptr;
              const bool wide = EPI_CHUNK == 32 && ncols == 32 && ((((uintptr_t)ph) | ((uintptr_t)pl)) & 31) == 0;
              if (wide) {
#pragma unroll
                for (int j = 0; j < EPI_CHUNK / 16; ++j) {
                  st_global_v8(ph + 16 * j, hi + 8 * j);
                  if (pl) st_global_v8(pl + 16 * j, lo + 8 * j);
                }
              } else {
                uint4* ph4 = (uint4*)ph;
#pragma unroll
                for (int j = 0; j < EPI_CHUNK / 8; ++j)
                  if (8 * j < ncols) ph4[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                if (pl) {
                  uint4* pl4 = (uint4*)pl;
#pragma unroll
                  for (int j = 0; j < EPI_CHUNK / 8; ++j)
                    if (8 * j < ncols) pl4[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                }
              }
            }
          }
        }
      }
      // all TMEM reads of this warp are complete (wait::ld above): hand the accumulator stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(tmem_empty_addr0 + (uint32_t)(as * sizeof(uint64_t)));
        else mbar_arrive(&tmem_empty_bar[as]);
      }
      if (++as == 2) { as = 0; aph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();      // the leader's MMAs read the peer's shared memory: nobody leaves before both are done
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
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

static int sm_count() {
  static int n[IPK_MAX_DEVICES] = {0};
  const int slot = current_device_slot();
  if (n[slot] == 0) {
    int dev = 0;
    IPK_CUDA(cudaGetDevice(&dev));
    IPK_CUDA(cudaDeviceGetAttribute(&n[slot], cudaDevAttrMultiProcessorCount, dev));
  }
  return n[slot];
}

template <int BN, int NSPLIT, bool FUSED, bool HALO, int CG>
static void launch_tc(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi, const CUtensorMap& w_lo, TcArgs& a,
                      cudaStream_t st) {
  constexpr int STAGE_BYTES = HALO ? (NSPLIT == 3 ? 2 : 1) * (TC_HALO_A_BYTES + 3 * (BN / CG) * TC_BK * 2)
                                   : (NSPLIT == 3 ? 2 : 1) * (TC_BM * TC_BK * 2 + (BN / CG) * TC_BK * 2);
  int stages = (int)std::min<size_t>(8, TC_SMEM_BUDGET / STAGE_BYTES);
  IPK_CHECK(stages >= 2, IPK_ERR_UNSUPPORTED, "conv_tc: pipeline needs at least two stages (stage %d bytes)", STAGE_BYTES);
  a.stages = stages;
  size_t smem = (size_t)stages * STAGE_BYTES + 1024 + TC_EPI_WARPS * TC_EPI_STAGE_BYTES;
  static bool attr_set[IPK_MAX_DEVICES] = {false};      // function attributes are per device
  const int slot = current_device_slot();
  if (!attr_set[slot]) {
    IPK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, NSPLIT, FUSED, HALO, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TC_SMEM_BUDGET + 1024 + TC_EPI_WARPS * TC_EPI_STAGE_BYTES)));
    attr_set[slot] = true;
  }
  const long long units = (long long)cdiv(a.tiles_m, CG) * a.tiles_n * (a.nsub > 1 ? a.nsub : a.nsplit);
  // persistent: one CTA per SM (CG = 2: one CTA pair per SM pair, grid a multiple of the cluster size)
  const unsigned grid = (unsigned)std::min<long long>(units * CG, (sm_count() / CG) * CG);
  launch_kc(conv_tc_kernel<BN, NSPLIT, FUSED, HALO, CG>, dim3(grid), dim3(TC_THREADS), smem, st, CG, a_hi, a_lo, w_hi, w_lo, a);
}

static int conv_tc_impl(const ConvW& w, const ConvIn& in, const ConvOut& out, const ConvSub* subs, int nsub, int nsplit, cudaStream_t st) {
  IPK_CHECK(w.w_hi != nullptr, IPK_ERR_STATE, "conv_tc_run: layer was not packed for the tensor-core engine");
  IPK_CHECK(nsub >= 1 && nsub <= TC_MAX_SUB, IPK_ERR_INVALID, "conv_tc_run: bad sub-convolution count %d", nsub);
  const bool split = w.engine == IPK_PREC_FP32_SPLIT;
  IPK_CHECK(!split || (in.p_lo && w.w_lo), IPK_ERR_STATE, "conv_tc_run: split precision needs hi and lo operand planes");
  IPK_CHECK(in.cstride % 8 == 0 && in.coff % 8 == 0, IPK_ERR_UNSUPPORTED, "conv_tc_run: activation rows must be 16-byte aligned (cstride %d, coff %d)", in.cstride, in.coff);
  const ConvOut* outs[2] = {&out, out.second};
  const int ncol[2] = {out.second ? out.split_col : w.Npad, out.second ? w.Npad - out.split_col : 0};
  if (out.second) IPK_CHECK(out.split_col > 0 && out.split_col % 32 == 0 && out.split_col < w.Npad, IPK_ERR_INVALID, "conv_tc_run: bad split_col %d", out.split_col);
  for (int i = 0; i < 2; ++i) {
    const ConvOut* o = outs[i];
    if (!o) continue;
    if (o->mode != OUT_F32_NCHW)
      IPK_CHECK((o->cstride % 4 == 0) && (o->coff % 4 == 0) && o->coff + ncol[i] <= o->cstride, IPK_ERR_UNSUPPORTED,
                "conv_tc_run: output row (cstride %d, coff %d) cannot hold %d columns", o->cstride, o->coff, ncol[i]);
    if (o->mode == OUT_BF16_SPLIT || o->mode == OUT_BF16)
      IPK_CHECK(o->cstride % 8 == 0 && o->coff % 8 == 0, IPK_ERR_UNSUPPORTED, "conv_tc_run: bf16 output rows must be 16-byte aligned");
  }
  long long M = (long long)in.F * in.H * in.W;
  if (M == 0) return 0;
  TcArgs a;
  memset(&a, 0, sizeof(a));
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
  a.nkb = w.Kpad / TC_BK;
  a.nsub = nsub;
  int max_taps = 0;
  for (int i = 0; i < nsub; ++i) {
    const TapList& t = subs[i].taps;
    a.sub[i].ntaps = t.n; a.sub[i].yadd = subs[i].yadd; a.sub[i].xadd = subs[i].xadd;
    for (int j = 0; j < MAX_TAPS; ++j) { a.sub[i].dy[j] = t.dy[j]; a.sub[i].dx[j] = t.dx[j]; a.sub[i].widx[j] = t.widx[j]; }
    max_taps = std::max(max_taps, t.n);
  }
  const int total_iters = max_taps * a.nkb;
  if (nsub > 1) nsplit = 1;
  nsplit = std::max(1, std::min(nsplit, total_iters));
  a.iters_per_split = cdiv(total_iters, nsplit);
  nsplit = cdiv(total_iters, a.iters_per_split);
  IPK_CHECK(nsplit == 1 || (out.mode == OUT_F32_NHWC && out.split_stride > 0 && !out.second), IPK_ERR_INVALID, "split-K needs fp32 partial slices");
  a.Npad = w.Npad; a.N = w.N;
  a.bias = out.bias;
  a.Ho = out.Ho; a.Wo = out.Wo; a.ymul = out.ymul; a.xmul = out.xmul;
  a.split_col = out.second ? out.split_col : 0;
  for (int i = 0; i < 2; ++i) {
    const ConvOut* o = outs[i] ? outs[i] : &out;
    a.o[i].out = o->p; a.o[i].out_lo = o->p_lo; a.o[i].act = o->act; a.o[i].mode = o->mode; a.o[i].cstride = o->cstride; a.o[i].coff = o->coff;
  }
  a.split_stride = out.split_stride;
  a.res = out.res; a.res_cstride = out.res_cstride; a.res_act = out.res_act; a.res_mr = out.res_mr; a.stats = out.stats;
  a.stats_oi = 0; a.stats_C = w.N;
  if (out.res)
    IPK_CHECK(out.mode == OUT_F32_NHWC && !out.second && nsplit == 1 && w.Npad == w.N && w.N % 32 == 0 && w.N >= 64 &&
                  (long long)in.H * in.W >= 32 && out.ymul == 1 && out.xmul == 1,
              IPK_ERR_UNSUPPORTED, "conv_tc_run: the fused residual needs a plain fp32 NHWC output with N %% 32 == 0, N >= 64");
  if (!out.stats && out.second && out.second->stats) {     // statistics of the second destination's columns
    a.stats = out.second->stats; a.stats_oi = 1; a.stats_C = w.N - out.split_col;
    IPK_CHECK(out.second->mode == OUT_F32_NHWC, IPK_ERR_UNSUPPORTED, "conv_tc_run: fused statistics need an fp32 NHWC destination");
  }
  if (a.stats)
    IPK_CHECK(nsplit == 1 && w.Npad == w.N && w.N % 32 == 0 && w.N >= 64 && (long long)in.H * in.W >= 32 &&
                  (a.stats_oi == 1 || out.mode == OUT_F32_NHWC),
              IPK_ERR_UNSUPPORTED, "conv_tc_run: fused statistics need fp32 NHWC output, N %% 32 == 0, N >= 64, >= 32 pixels per frame");
  a.stages = 2;

  // activation maps: dims (C, W, H, F); the C extent is the true channel count so the K tail is zero-filled
  long long ad[4] = {w.K, in.W, in.H, in.F};
  long long as[3] = {(long long)in.cstride * 2, (long long)in.W * in.cstride * 2, (long long)in.H * in.W * in.cstride * 2};
  int ab[4] = {TC_BK, a.bw, a.bh, a.bf};
  const __nv_bfloat16* ahi = (const __nv_bfloat16*)in.p + in.coff;
  CUtensorMap mA_hi = make_map(ahi, 4, ad, as, ab);
  CUtensorMap mA_lo = split ? make_map((const __nv_bfloat16*)in.p_lo + in.coff, 4, ad, as, ab) : mA_hi;
  // N tile: weigh padded columns (MMA work) against A-tile re-reads (one per N tile)
  int BN = 32;
  {
    double best = 1e30;
    for (int bn : {256, 128, 64, 32}) {
      double cost = (double)cdiv(w.Npad, bn) * bn * (1.0 + 32.0 / bn);
      if (cost < best) { best = cost; BN = bn; }
    }
  }
  a.tiles_m = a.tiles_x * a.tiles_y * tiles_f;
  // CTA pairs (cta_group::2): two consecutive M tiles share one N tile and each CTA stages half of its weights (IPK_TC_CTA2=0 disables,
  // =2 also pairs short main loops).
  static const int cta2_env = []() { const char* e = getenv("IPK_TC_CTA2"); return e ? atoi(e) : 1; }();
  // Measured on B200 (profiles/r02_cta2_ab.md): a win where the W tile dominates the stage bytes and the main loop is long (NICE conv2
  // 19.2 -> 16.9 ms per step, N = K = 2048); a loss for narrow N tiles (A traffic dominates, and the pair's lock-step epilogue hand-off
  // costs more than the halved W fetch saves) and for short main loops (NICE conv1, K <= 192: epilogue-bound).
  const int CG = (cta2_env > 0 && a.tiles_m >= 2 && BN == 256 && (cta2_env > 1 || max_taps * a.nkb >= 8)) ? 2 : 1;
  long long wd[2] = {w.Kpad, (long long)w.ntaps * w.Npad};
  long long wsb[1] = {(long long)w.Kpad * 2};
  int wb[2] = {TC_BK, BN / CG};
  CUtensorMap mW_hi = make_map(w.w_hi, 2, wd, wsb, wb);
  CUtensorMap mW_lo = split ? make_map(w.w_lo, 2, wd, wsb, wb) : mW_hi;

  a.tiles_n = cdiv(w.Npad, BN);
  a.nsplit = nsplit;
  const bool fused = a.res != nullptr || a.stats != nullptr;
  // halo mode: full 3x3 tap set on a 128-wide image, one image row per tile, 64-column N tiles (the decoder's last conv2)
  static const int halo_env = []() { const char* e = getenv("IPK_TC_HALO"); return e ? atoi(e) : 2; }();     // 0 = off
  bool halo = halo_env > 0 && nsub == 1 && nsplit == 1 && in.W == TC_BM && a.bw == TC_BM && subs[0].taps.n == 9 && (BN == 64 || BN == 32);
  if (halo)
    for (int i = 0; i < 9; ++i)
      halo = halo && subs[0].taps.dy[i] == i / 3 - 1 && subs[0].taps.dx[i] == i % 3 - 1 && subs[0].taps.widx[i] == i;
  a.halo_variant = halo_env;
  if (halo) {
    int hb[4] = {TC_BK, 130, 1, 1};
    const CUtensorMap hA_hi = make_map(ahi, 4, ad, as, hb);
    const CUtensorMap hA_lo = split ? make_map((const __nv_bfloat16*)in.p_lo + in.coff, 4, ad, as, hb) : hA_hi;
#define IPK_TC_LAUNCH(bn, halo_, cg, mAh, mAl)                                              \
    do {                                                                                  \
      if (fused) {                                                                        \
        if (split) launch_tc<bn, 3, true, halo_, cg>(mAh, mAl, mW_hi, mW_lo, a, st);      \
        else launch_tc<bn, 1, true, halo_, cg>(mAh, mAl, mW_hi, mW_lo, a, st);            \
      } else {                                                                            \
        if (split) launch_tc<bn, 3, false, halo_, cg>(mAh, mAl, mW_hi, mW_lo, a, st);     \
        else launch_tc<bn, 1, false, halo_, cg>(mAh, mAl, mW_hi, mW_lo, a, st);           \
      }                                                                                   \
    } while (0)
    if (BN == 32) IPK_TC_LAUNCH(32, true, 1, hA_hi, hA_lo);        // the decoder's out_conv (N = 3)
    else if (CG == 2) IPK_TC_LAUNCH(64, true, 2, hA_hi, hA_lo);
    else IPK_TC_LAUNCH(64, true, 1, hA_hi, hA_lo);
    return nsplit;
  }
  if (CG == 2) {
    switch (BN) {
      case 64: IPK_TC_LAUNCH(64, false, 2, mA_hi, mA_lo); break;
      case 128: IPK_TC_LAUNCH(128, false, 2, mA_hi, mA_lo); break;
      default: IPK_TC_LAUNCH(256, false, 2, mA_hi, mA_lo); break;
    }
  } else {
    switch (BN) {
      case 32: IPK_TC_LAUNCH(32, false, 1, mA_hi, mA_lo); break;
      case 64: IPK_TC_LAUNCH(64, false, 1, mA_hi, mA_lo); break;
      case 128: IPK_TC_LAUNCH(128, false, 1, mA_hi, mA_lo); break;
      default: IPK_TC_LAUNCH(256, false, 1, mA_hi, mA_lo); break;
    }
  }
#undef IPK_TC_LAUNCH
  return nsplit;
}

int conv_tc_run(const ConvW& w, const ConvIn& in, const ConvOut& out, const TapList& taps, int nsplit, cudaStream_t st) {
  ConvSub sub;
  sub.taps = taps; sub.yadd = out.yadd; sub.xadd = out.xadd;
  return conv_tc_impl(w, in, out, &sub, 1, nsplit, st);
}

void conv_tc_run_multi(const ConvW& w, const ConvIn& in, const ConvOut& out, const ConvSub* subs, int nsub, cudaStream_t st) {
  conv_tc_impl(w, in, out, subs, nsub, 1, st);
}

}  // namespace ipk
