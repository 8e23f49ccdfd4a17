// tcgen05 implicit-GEMM convolution engine: kernel template + launcher (included by conv_tc.cu and by the per-tile-width translation
// units conv_tc_bn*.cu, which instantiate the kernel variants of one N tile each so that they compile in parallel).
#pragma once
#include <cuda.h>
#include <cstring>
#include <type_traits>
#include "conv.cuh"
#include "tc_ptx.cuh"

namespace ipk {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr size_t TC_SMEM_BUDGET = 192 * 1024;      // pipeline stages
constexpr size_t TC_EPI_STAGE_BYTES = 4096;          // per epilogue warp: 32 rows x 128 B transpose buffer (xor-swizzled)

// division by a run-time constant without the ~20-instruction reciprocal sequence (the per-tile index arithmetic was 14 % of the fused
// epilogue's samples): q = (umulhi(x, mul) + x) >> shr, exact for 0 <= x < 2^31
struct FastDiv {
  uint32_t mul = 0, shr = 0, d = 1;
  __host__ void set(int div) {
    d = (uint32_t)div;
    shr = 0;
    while ((1u << shr) < d) ++shr;
    mul = (uint32_t)((((unsigned long long)1 << 32) * (((unsigned long long)1 << shr) - d)) / d + 1);
  }
  __device__ __forceinline__ int div(int x) const { return (int)((__umulhi((uint32_t)x, mul) + (uint32_t)x) >> shr); }
};

constexpr int TC_MAX_SUB = 4;
constexpr int TC_MAX_NV = 16;
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
  // bf16 operand planes leave through TMA stores (tma_out[i] != 0 for destination i): the warp's 32 rows x 32 columns are one box of the
  // output viewed as (columns, x', y, f); a transposed conv's parity class (a, b) is folded into the view: x' = a * W + x, column + b * cstride
  int tma_out[2], tma_xfold, tma_cfold[2];
  // effective N tile: bn columns (a multiple of 16, <= the kernel's BN) are computed per tile -- UMMA N, the W box and the epilogue follow it,
  // so a layer with 144 or 288 output columns is one or two tiles of 144 instead of padded 256-column tiles; half_cols = columns of the
  // first epilogue column half (a multiple of 32)
  int bn, half_cols;
  // uneven N tiles (nv_tiles > 0; wide layers on CTA pairs whose uniform tiling ends in a partly filled wave): tile slot i covers columns
  // [nv_n0[i], nv_n0[i] + nv_w[i]), widths are multiples of 32 <= bn (the W box is always bn / CG rows), slots are sorted wide-first and a
  // unit index is slot * (M-tile groups) + group, so that the round-robin persistent schedule gives no pair two wide tiles
  int nv_tiles;
  short nv_n0[TC_MAX_NV], nv_w[TC_MAX_NV];
  FastDiv fd_tiles_mg;
  FastDiv fd_tiles_n, fd_per, fd_tiles_mn, fd_txy, fd_tiles_x, fd_nsub;     // divisors of the per-tile index arithmetic (per = nsub * tiles_n)
  long long* trace;               // optional per-CTA timeline (ipk_tc_trace_enable): 32 clock stamps per CTA, null = off
  int halo_variant;               // HALO kernels: 1 = row-shifted descriptors carry the swizzle base offset, 2 = they do not
};

// PTX wrappers (mbarrier, TMA, UMMA descriptors, TMEM loads): tc_ptx.cuh

// ------------------------------------------------------------------------------------------------ kernel
// activation applied to a register tile; the branch is warp-uniform.  ELU uses ex2.approx (abs error ~1e-7, far inside
// the engine's own operand rounding); the fp32 SIMT validation engine keeps expm1f.
// one MUFU.EX2 (results below 2^-126 flush to zero; exp(x) - 1 is -1 there either way)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int NV>
__device__ __forceinline__ void act_tile(float (&v)[NV], int act) {
  if (act == ACT_ELU) {
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = v[j] > 0.f ? v[j] : (ex2_approx(v[j] * 1.4426950408889634f) - 1.0f);
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

// timeline probe (ipk_tc_trace_enable): slot s of this CTA <- SM clock; slot 0 additionally carries the global timer in slot 31
__device__ __forceinline__ void tc_trace(const TcArgs& a, int slot) {
  if (a.trace != nullptr) a.trace[(size_t)blockIdx.x * 32 + slot] = clock64();
}

// linear work-unit id -> (z, M-tile group, N tile).  Split-K launches: z (the K slice) is the slowest index, n the fastest, so CTAs running
// side by side share the A tile in L2.  Multi-sub launches (the four parity classes of a transposed conv, which all read the SAME input
// tile): z is iterated inside the M-tile group, so the input tensor streams through L2 once instead of once per class (ncu, r02: 2.15 GB
// of DRAM reads per 256-frame chunk of the last up-block against 0.54 GB of input).
__device__ __forceinline__ void tc_decode_tile(const TcArgs& a, int tiles_mn, int tile, int& z, int& mg, int& nt) {
  if (a.nsub > 1) {
    const int per = a.nsub * a.tiles_n;
    mg = a.fd_per.div(tile);
    const int rem = tile - mg * per;
    const int zi = a.fd_tiles_n.div(rem);
    nt = rem - zi * a.tiles_n;
    // the classes have 4 / 2 / 2 / 1 taps: rotate them with the M-tile group, otherwise a persistent CTA (stride = a multiple of nsub)
    // would draw the same class every time (r02 timeline: 2x spread of the CTAs' finishing times)
    const int zs = zi + mg;
    z = zs - a.fd_nsub.div(zs) * a.nsub;
  } else if (a.nv_tiles > 0) {
    z = 0;
    nt = a.fd_tiles_mg.div(tile);                 // tile slot (wide-first), M-tile group fastest
    mg = tile - nt * a.fd_tiles_mg.d;
  } else {
    z = a.fd_tiles_mn.div(tile);
    const int rem = tile - z * tiles_mn;
    mg = a.fd_tiles_n.div(rem);
    nt = rem - mg * a.tiles_n;
  }
}
// first column and width of N tile `nt`
__device__ __forceinline__ void tc_tile_cols(const TcArgs& a, int nt, int& n0, int& w) {
  if (a.nv_tiles > 0) { n0 = a.nv_n0[nt]; w = a.nv_w[nt]; }
  else { n0 = nt * a.bn; w = a.bn; }
}

// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue.
// Two TMEM accumulator stages (2 x BN columns): the epilogue of tile i overlaps the main loop of tile i+1.
// Tile order: linear id -> (split z, m tile, n tile) with n fastest, so CTAs running side by side share the A tile in L2.
// FUSED: 1 = output statistics in the epilogue, 2 = residual branch (+ statistics); see ConvOut.  FUSED = 2 launches have one fp32 NHWC
// destination, so only the transposed write-out is compiled for them.
// HALO (3x3 convs on 128-wide images, one image row per tile): the producer loads each input row ONCE as a 130-pixel box
// (x = -1 .. 128, zero-filled outside) and the three dx taps are row-shifted views of that box (UMMA descriptor start advanced by
// dx * 128 bytes) instead of three separate TMA loads: A traffic / 3.
constexpr int TC_HALO_A_BYTES = 17 * 1024;      // 130 rows x 128 B = 16 640 B, padded to keep the regions 1024-byte aligned
// CG = 2 (CTA pair, launched as clusters of 2): the two CTAs own two consecutive M tiles of the same N tile; each loads its own A
// tile and HALF of the W tile (BN / 2 rows), and the pair's leader issues cta_group::2 MMAs (M = 256, N = BN) that read both CTAs'
// shared memory -- every W byte is fetched from L2 once per pair instead of once per CTA, which is what bounds these kernels
// (measured: NICE conv2 sits at the chip's TMA throughput, not at the tensor pipe).  Each CTA's TMEM holds its own 128 rows x BN
// columns, so the epilogue is the single-CTA one.
template <int BN, int NSPLIT, int FUSED, bool HALO, int CG>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
               const __grid_constant__ CUtensorMap tmO0_hi, const __grid_constant__ CUtensorMap tmO0_lo,
               const __grid_constant__ CUtensorMap tmO1_hi, const __grid_constant__ CUtensorMap tmO1_lo,
               const __grid_constant__ TcArgs a) {
  constexpr int A_BYTES = TC_BM * TC_BK * 2;         // 16 KB
  constexpr int WROWS = BN / CG;                     // W rows this CTA stages
  constexpr int W_BYTES = WROWS * TC_BK * 2;
  constexpr int NPLANES = NSPLIT == 3 ? 2 : 1;
  constexpr int STAGE_BYTES = HALO ? NPLANES * (TC_HALO_A_BYTES + 3 * W_BYTES) : NPLANES * (A_BYTES + W_BYTES);
  const uint32_t w_tx_bytes = (uint32_t)(a.bn / CG) * TC_BK * 2;       // bytes of one W plane actually loaded per stage
  constexpr int MAX_STAGES = 8;
  constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  constexpr int EPI_CHUNK = BN >= 64 ? 32 : 16;      // columns per TMEM load
  constexpr int HALF_COLS = BN / 2;                  // columns owned by one of the two epilogue warps of a lane quarter

  // Shared memory (all dynamic, so that its base is the CTA's 1024-byte aligned window and no alignment slack is needed):
  //   [stages x STAGE_BYTES] operand ring | [8 x 4 KB] epilogue transpose buffers | mbarriers + the TMEM base address
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;                                                        // SWIZZLE_128B needs 1024-byte alignment
  uint8_t* epi_stage = smem + (size_t)a.stages * STAGE_BYTES;                      // 8 x 4 KB epilogue transpose buffers
  uint64_t* full_bar = (uint64_t*)(epi_stage + TC_EPI_WARPS * TC_EPI_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + MAX_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t& tmem_base_smem = *(uint32_t*)(tmem_empty_bar + 2);

  uint32_t tid_;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid_));       // read once (the compiler otherwise re-reads SR_TID inside the epilogue loops)
  const int warp = (int)(tid_ >> 5), lane = (int)(tid_ & 31);
  const int stages = a.stages;
  if (tid_ == 0 && (smem_u32(smem_raw) & 1023u) != 0) {
    printf("ipoke_b200 conv_tc: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  if (tid_ == 0 && a.trace != nullptr) {
    tc_trace(a, 0);
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    a.trace[(size_t)blockIdx.x * 32 + 30] = (long long)gt;
  }
  const int txy = a.tiles_x * a.tiles_y;
  // work units: (split / sub-convolution z, M-tile group, N tile); a group is CG consecutive M tiles, one per CTA of the pair
  const int tiles_mg = (a.tiles_m + CG - 1) / CG;
  const int tiles_mn = tiles_mg * a.tiles_n;
  const int total_tiles = tiles_mn * (a.nsub > 1 ? a.nsub : a.nsplit);
  const uint32_t crank = CG == 2 ? cluster_ctarank() : 0u;
  const int unit0 = (int)blockIdx.x / CG, unit_step = (int)gridDim.x / CG;

  if (warp == 0 && lane == 0) {       // descriptor fetch (kernel parameters) off the first TMA's critical path: ~1.8 us per launch measured
    prefetch_tensormap(&tmA_hi); prefetch_tensormap(&tmW_hi);
    if (NSPLIT == 3) { prefetch_tensormap(&tmA_lo); prefetch_tensormap(&tmW_lo); }
  }
  if (warp == 2 && lane == 0) {
    if (a.tma_out[0]) { prefetch_tensormap(&tmO0_hi); if (a.o[0].mode == OUT_BF16_SPLIT) prefetch_tensormap(&tmO0_lo); }
    if (a.tma_out[1]) { prefetch_tensormap(&tmO1_hi); if (a.o[1].mode == OUT_BF16_SPLIT) prefetch_tensormap(&tmO1_lo); }
  }
  if (tid_ == 0) {
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
  if (tid_ == 0) tc_trace(a, 1);
  // Programmatic dependent launch: the prologue above overlapped the previous kernel's tail.  Each role waits for the previous grid
  // (griddepcontrol.wait) only where it first touches global memory -- the producer right before its first TMA load, after the first
  // tile's index arithmetic, so that the cold instruction-cache misses of that code are taken while the previous kernel still runs
  // (r02 timeline: 3 000 cycles between the wait and the first TMA load of every launch); the MMA warp touches no global memory.
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
        int z, mg, nt;                                                  // z: split-K slice, or sub-convolution when nsub > 1
        tc_decode_tile(a, tiles_mn, tile, z, mg, nt);
        const int mt = mg * CG + (int)crank;                            // beyond tiles_m: every box is out of bounds -> zero fill
        const int tf = a.fd_txy.div(mt), r2 = mt - tf * txy;
        const int ty = a.fd_tiles_x.div(r2), tx = r2 - ty * a.tiles_x;
        int n0, bnw;
        tc_tile_cols(a, nt, n0, bnw);
        n0 += (int)crank * (bnw / CG);
        const int f0 = tf * a.bf, y0 = ty * a.bh, x0 = tx * a.bw;
        const TcSub& sb = a.sub[a.nsub > 1 ? z : 0];
        if constexpr (HALO) {
          // one stage per (input row dy, k-block): the 130-pixel row box of both planes + the weights of its three dx taps
          for (int dyi = 0; dyi < 3; ++dyi)
            for (int kb = 0; kb < a.nkb; ++kb) {
              mbar_wait(&empty_bar[s], ph ^ 1);
              uint8_t* st = smem + (size_t)s * STAGE_BYTES;
              if (crank == 0) mbar_expect_tx(&full_bar[s], CG * NPLANES * (130 * 128 + 3 * w_tx_bytes));
              if (tile == unit0 && dyi == 0 && kb == 0) pdl_wait();
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
            if (crank == 0) mbar_expect_tx(&full_bar[s], CG * NPLANES * (A_BYTES + w_tx_bytes));
            if (tile == unit0 && it == it_begin) { pdl_wait(); tc_trace(a, 2); }
            lda(st, &tmA_hi, s, kb * TC_BK, x0 + dx, y0 + dy, f0);
            if (tile == unit0 && it == it_begin) tc_trace(a, 19);
            ldw(st + NPLANES * A_BYTES, &tmW_hi, s, kb * TC_BK, wrow);
            if (NSPLIT == 3) {
              lda(st + A_BYTES, &tmA_lo, s, kb * TC_BK, x0 + dx, y0 + dy, f0);
              ldw(st + NPLANES * A_BYTES + W_BYTES, &tmW_lo, s, kb * TC_BK, wrow);
            }
            if (++s == stages) { s = 0; ph ^= 1; }
            if (tile == unit0 && it == it_begin) tc_trace(a, 3);
          }
          if (++kb == a.nkb) { kb = 0; ++t; }
        }
        if (tile == unit0) tc_trace(a, 4);
      }
      tc_trace(a, 5);
    }
    __syncwarp();      // the warp reaches the final block barrier as a whole
  } else if (warp == 1) {
    // ===================== MMA issuer (the pair's leader only when CG = 2) =====================
    if (lane == 0 && crank == 0) {
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      uint32_t IDESC = umma_idesc_bf16(TC_BM * CG, a.bn);
      auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t acc) {
        if constexpr (CG == 2) umma_bf16_2cta(d, da, db, IDESC, acc);
        else umma_bf16(d, da, db, IDESC, acc);
      };
      auto commit = [&](uint64_t* bar) {
        if constexpr (CG == 2) umma_commit_2cta(bar);
        else umma_commit(bar);
      };
      for (int tile = unit0; tile < total_tiles; tile += unit_step) {
        int z, mg_, nt_;
        tc_decode_tile(a, tiles_mn, tile, z, mg_, nt_);
        if (a.nv_tiles > 0) IDESC = umma_idesc_bf16(TC_BM * CG, a.nv_w[nt_]);
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
        const int tsl = tile == unit0 ? 6 : (tile == unit0 + unit_step ? 8 : 28);
        for (int it = 0; it < iters; ++it) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (it == 0) tc_trace(a, tsl);
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
        tc_trace(a, tsl + 1);
        if (++as == 2) { as = 0; aph ^= 1; }
      }
      // drain: the epilogue warps have released both accumulator stages before this warp frees the tensor memory (waiting on a stage
      // that was never filled returns at once)
      for (int i = 0; i < 2; ++i) {
        mbar_wait(&tmem_empty_bar[as], aph ^ 1);
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue: warps 2..9; TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 ============
    // * Everything an epilogue warp needs from global memory is requested BEFORE it waits for the accumulator (the tile's bias, 4 columns
    //   per lane, and the first chunk's residual rows -> registers), the TMEM load of chunk c+1 is in flight while chunk c is processed, and
    //   the accumulator stage goes back to the MMA warp as soon as the last TMEM load has landed (ncu, r02: the first version spent
    //   > 60 % of its time on the L2 latency of per-chunk bias / residual loads).
    // * The write-out mode and the activation are dispatched ONCE per column range of a destination, outside the chunk loop, into bodies
    //   specialised at compile time: the hot loop of a launch is one compact straight-line block (ncu, r02: 591 instructions per
    //   32 x 32 chunk, 16 % of the epilogue's stall samples on instruction fetch, IPC 0.25 with two epilogue warps per scheduler).
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;                    // accumulator row == pixel index inside the box
    const int xl = r % a.bw, yl = (r / a.bw) % a.bh, fl = r / (a.bw * a.bh);
    const int seg = lane & 7, trow0 = lane >> 3;    // transposed write-out: 8 lanes = one 128-byte row segment, rows trow0 + 4 i
    const uint32_t stg_u32 = smem_u32(epi_stage + (size_t)(warp - 2) * TC_EPI_STAGE_BYTES);
    constexpr bool has_res = FUSED == 2 && EPI_CHUNK == 32;
    const bool any_tma = EPI_CHUNK == 32 && (a.tma_out[0] | a.tma_out[1]) != 0;
    const int r0w = q * 32;                         // first row of this warp: origin of its TMA store box inside the tile
    const int xl0 = r0w % a.bw, yl0 = (r0w / a.bw) % a.bh, fl0 = r0w / (a.bw * a.bh);
    int as = 0;
    uint32_t aph = 0;
    pdl_wait();                                     // residual rows / statistics are read, and outputs written, after the previous grid
    int bias_nt = -1;                               // N tile whose bias this lane holds in b4 (one N tile: loaded once per kernel)
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    // the accumulator stage goes back to the MMA issuer: the leader's barrier collects the epilogue warps of both CTAs of a pair
    const uint32_t tmem_empty_addr0 = CG == 2 ? mapa_u32(smem_u32(&tmem_empty_bar[0]), 0) : 0u;
    for (int tile = unit0; tile < total_tiles; tile += unit_step) {
      int z, mg, nt;
      tc_decode_tile(a, tiles_mn, tile, z, mg, nt);
      const int mt = mg * CG + (int)crank;
      const int tf = a.fd_txy.div(mt), r2 = mt - tf * txy;
      const int ty = a.fd_tiles_x.div(r2), tx = r2 - ty * a.tiles_x;
      const int f = tf * a.bf + fl, y = ty * a.bh + yl, x = tx * a.bw + xl;
      int n0, bnw;
      tc_tile_cols(a, nt, n0, bnw);
      const int hcols = a.nv_tiles > 0 ? (((bnw >> 1) + 31) & ~31) : a.half_cols;
      const bool valid = (mt < a.tiles_m) && (f < a.F) && (y < a.H) && (x < a.W);
      const TcSub& sb = a.sub[a.nsub > 1 ? z : 0];
      const int oy = y * a.ymul + sb.yadd, ox = x * a.xmul + sb.xadd;
      const size_t opix = ((size_t)f * a.Ho + (size_t)oy) * a.Wo + (size_t)ox;
      const size_t zoff = a.nsub > 1 ? 0 : (size_t)z * a.split_stride;
      const int c_begin = half * hcols, c_end = min(min(c_begin + hcols, bnw), a.Npad - n0);     // this warp's columns of the tile

      // rows this lane stores in the transposed (coalesced) write-out: row_i = lane/8 + 4*i; their output pixels come from
      // the lanes that own them (all-ones = row outside the image)
      uint32_t trow[8];                               // output pixel index (< 2^32: host check), all-ones = outside
      int fw = 0;
      if constexpr (EPI_CHUNK == 32) {
        const uint32_t mine = valid ? (uint32_t)opix : ~0u;
#pragma unroll
        for (int i = 0; i < 8; ++i) trow[i] = __shfl_sync(0xffffffffu, mine, trow0 + 4 * i);
        fw = __shfl_sync(0xffffffffu, f, 0);          // frame of this warp's rows (fused residual / statistics: >= 32 pixels per frame)
      }
      // ---- requests that do not depend on the accumulator
      if (nt != bias_nt) {
        bias_nt = nt;
        b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.bias != nullptr && lane * 4 < hcols && c_begin + lane * 4 < c_end) b4 = __ldg((const float4*)(a.bias + n0 + c_begin) + lane);
      }
      float4 resv[8], mr0 = make_float4(0.f, 1.f, 0.f, 1.f), mr1 = make_float4(0.f, 1.f, 0.f, 1.f);     // residual rows, their (mean, rstd) pairs
      auto load_res = [&](int c) __attribute__((always_inline)) {
#pragma unroll
        const bool cok = c + seg * 4 < c_end;        // a 16-column tail chunk: the upper segments lie outside the tile
        for (int i = 0; i < 8; ++i)
          resv[i] = (trow[i] != ~0u && cok) ? __ldg((const float4*)(a.res + (size_t)trow[i] * a.res_cstride + n0 + c) + seg) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.res_mr != nullptr && cok) {
          const float4* mp = (const float4*)(a.res_mr + ((size_t)min(fw, a.F - 1) * a.N + n0 + c + seg * 4) * 2);
          mr0 = __ldg(mp); mr1 = __ldg(mp + 1);
        }
      };
      if (has_res && c_begin < c_end) load_res(c_begin);

      mbar_wait(&tmem_full_bar[as], aph);
      tc_fence_after();
      const int esl = tile == unit0 ? 10 : (tile == unit0 + unit_step ? 13 : 26);
      if (warp == 2 && lane == 0) tc_trace(a, esl);
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);

      auto release_stage = [&]() __attribute__((always_inline)) {
        // all TMEM reads of this warp have landed: hand the accumulator stage back to the MMA warp before the stores
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_cluster(tmem_empty_addr0 + (uint32_t)(as * sizeof(uint64_t)));
          else mbar_arrive(&tmem_empty_bar[as]);
        }
      };

      // ---- one column range [c_lo, c_hi) of destination oi, write-out MODE and activation ACT fixed at compile time
      //      MODE 0: fp32 rows, transposed through shared memory (bias, activation, residual, statistics on the way out)
      //      MODE 1: bf16 planes through TMA stores        MODE 2: lane = row direct stores (NCHW frames, bf16 without TMA, narrow tiles)
      //      ACT: an Act value, or -1 = the destination's activation is applied through the run-time switch
      auto run_range = [&](auto MODE_C, auto ACT_C, auto RACT_C, const int oi, const int c_lo, const int c_hi, const bool last_range) __attribute__((always_inline)) {
        constexpr int MODE = decltype(MODE_C)::value;
        constexpr int ACT = decltype(ACT_C)::value;
        constexpr int RACT = decltype(RACT_C)::value;       // activation of the residual branch (FUSED = 2), -1 = run-time switch
        const TcOut& od = a.o[oi];
        const int act = ACT >= 0 ? ACT : od.act;
        // the TMEM load of chunk c+1 is issued while chunk c is processed where a warp has several chunks per tile and the registers
        // allow it (FUSED kernels: the statistics / residual state takes the registers; measured slower with the spills)
        constexpr bool PREFETCH = HALF_COLS > EPI_CHUNK && FUSED == 0;
        uint32_t rr[EPI_CHUNK];
        if constexpr (PREFETCH) {
          if constexpr (EPI_CHUNK == 32) tmem_ld32(tacc + (uint32_t)c_lo, rr);
          else tmem_ld16(tacc + (uint32_t)c_lo, rr);
        }
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c += EPI_CHUNK) {
          const int ncols = min(EPI_CHUNK, c_hi - c);   // Npad is a multiple of 16: a 32-column chunk may be half valid
          float v[EPI_CHUNK];
          if constexpr (!PREFETCH) {
            if constexpr (EPI_CHUNK == 32) tmem_ld32(tacc + (uint32_t)c, rr);
            else tmem_ld16(tacc + (uint32_t)c, rr);
          }
          tmem_ld_wait_regs(rr);
#pragma unroll
          for (int j = 0; j < EPI_CHUNK; ++j) v[j] = __uint_as_float(rr[j]);
          const size_t colbase = (size_t)od.coff + (size_t)(n0 + c - (oi ? a.split_col : 0));
          // the tile's bias lives in registers, 4 columns per lane (b4 = columns c_begin + 4 lane ..): a lane fetches what it needs by
          // shuffles (no shared copy: the four warps of a column half would otherwise race on it, compute-sanitizer racecheck r02)
          const int bl = (c - c_begin) >> 2;             // lane holding the chunk's first 4 columns
          float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
          if constexpr (MODE == 0) {
            // the 4 columns this lane owns on the way out
            bb.x = __shfl_sync(0xffffffffu, b4.x, bl + seg); bb.y = __shfl_sync(0xffffffffu, b4.y, bl + seg);
            bb.z = __shfl_sync(0xffffffffu, b4.z, bl + seg); bb.w = __shfl_sync(0xffffffffu, b4.w, bl + seg);
          } else {
            if (a.bias != nullptr) {                     // lane = row: all columns of the chunk
#pragma unroll
              for (int j = 0; j < EPI_CHUNK / 4; ++j) {
                v[4 * j] += __shfl_sync(0xffffffffu, b4.x, bl + j); v[4 * j + 1] += __shfl_sync(0xffffffffu, b4.y, bl + j);
                v[4 * j + 2] += __shfl_sync(0xffffffffu, b4.z, bl + j); v[4 * j + 3] += __shfl_sync(0xffffffffu, b4.w, bl + j);
              }
            }
          }
          if (c + EPI_CHUNK < c_hi) {                  // next chunk's accumulators travel while this one is processed
            if constexpr (PREFETCH) {
              if constexpr (EPI_CHUNK == 32) tmem_ld32(tacc + (uint32_t)(c + EPI_CHUNK), rr);
              else tmem_ld16(tacc + (uint32_t)(c + EPI_CHUNK), rr);
            }
          } else if (last_range) {
            release_stage();
          }
          if (any_tma) {     // the previous chunk's TMA store has finished reading this warp's staging buffer
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
          }

          if constexpr (MODE == 0) {
            // ---- fp32 rows.  The raw accumulators are transposed through shared memory (lane = row on the way in, 8 lanes = one
            //      128-byte row segment on the way out); bias, activation, the residual branch and the statistics are applied on the way
            //      out, where a lane owns 4 fixed columns (12 per-column constants instead of 96) and the residual rows are read with
            //      the same coalesced pattern as the output is written (4 rows x 128 B per instruction).
            const int sw = lane & 7;
#pragma unroll
            for (int j = 0; j < 8; ++j) st_shared_v4(stg_u32 + lane * 128 + ((j ^ sw) << 4), make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
            const float4 sc = make_float4(mr0.y, mr0.w, mr1.y, mr1.w);
            const float4 sh = make_float4(-mr0.x * mr0.y, -mr0.z * mr0.w, -mr1.x * mr1.y, -mr1.z * mr1.w);
            __syncwarp();
            const bool do_stats = FUSED != 0 && a.stats != nullptr && oi == a.stats_oi;
            const bool col_ok = seg * 4 < ncols;
            float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
            float* const obase = (float*)od.out + zoff + colbase + seg * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = trow0 + 4 * i;
              const float4 t4 = ld_shared_v4(stg_u32 + row * 128 + ((seg ^ (row & 7)) << 4));
              float t[4] = {t4.x + bb.x, t4.y + bb.y, t4.z + bb.z, t4.w + bb.w};
              act_tile<4>(t, act);
              if (trow[i] == ~0u || !col_ok) continue;       // rows outside the image: nothing stored, nothing counted
              if (has_res) {
                float rv[4] = {fmaf(resv[i].x, sc.x, sh.x), fmaf(resv[i].y, sc.y, sh.y), fmaf(resv[i].z, sc.z, sh.z), fmaf(resv[i].w, sc.w, sh.w)};
                act_tile<4>(rv, RACT >= 0 ? RACT : a.res_act);
#pragma unroll
                for (int k = 0; k < 4; ++k) t[k] += rv[k];
              }
              *(float4*)(obase + (size_t)trow[i] * od.cstride) = make_float4(t[0], t[1], t[2], t[3]);
              if (do_stats) {
#pragma unroll
                for (int k = 0; k < 4; ++k) { s1[k] += t[k]; s2[k] = fmaf(t[k], t[k], s2[k]); }
              }
            }
            if (has_res && c + EPI_CHUNK < c_hi) load_res(c + EPI_CHUNK);
            if (do_stats) {
              // per-(frame, channel) sum / sum of squares: the four lanes that share a column segment combine, one of them adds
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], 8);  s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], 8);
                s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], 16); s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], 16);
              }
              if (lane < 8 && col_ok && fw < a.F) {
                double* sp = a.stats + ((size_t)fw * a.stats_C + (n0 + c - (oi ? a.split_col : 0)) + seg * 4) * 2;
#pragma unroll
                for (int k = 0; k < 4; ++k) { atomicAdd(sp + 2 * k, (double)s1[k]); atomicAdd(sp + 2 * k + 1, (double)s2[k]); }
              }
            }
            __syncwarp();       // staging rows are rewritten by the next chunk
            if (warp == 2 && lane == 0 && esl == 10 && c == c_begin) tc_trace(a, 11);
          } else if constexpr (MODE == 1) {
            // ---- bf16 operand planes through TMA: the warp's 32 rows x 32 columns (64 B per row and plane) are staged in shared memory in
            //      the 64-byte swizzle (conflict-free 16-byte writes: chunk j of row l at l * 64 + ((j ^ (l >> 1 & 3)) << 4)) and leave as
            //      one box per plane; rows / tiles outside the image are clipped by the TMA unit
            act_tile<EPI_CHUNK>(v, act);
            const uint32_t rowb = stg_u32 + lane * 64;
            const int sx = (lane >> 1) & 3;
            const bool splitp = od.mode == OUT_BF16_SPLIT;
#pragma unroll
            for (int j = 0; j < EPI_CHUNK / 8; ++j) {
              uint32_t h[4], l[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float v0 = v[8 * j + 2 * k], v1 = v[8 * j + 2 * k + 1];
                const __nv_bfloat162 hh = __floats2bfloat162_rn(v0, v1);
                const float2 hf = __bfloat1622float2(hh);
                const __nv_bfloat162 ll = __floats2bfloat162_rn(v0 - hf.x, v1 - hf.y);
                h[k] = *(const uint32_t*)&hh;
                l[k] = *(const uint32_t*)&ll;
              }
              st_shared_v4_b32(rowb + ((j ^ sx) << 4), h[0], h[1], h[2], h[3]);
              if (splitp) st_shared_v4_b32(rowb + 2048 + ((j ^ sx) << 4), l[0], l[1], l[2], l[3]);
            }
            fence_proxy_async_shared();
            __syncwarp();
            if (lane == 0) {
              const int cc = (int)(n0 + c - (oi ? a.split_col : 0)) + sb.xadd * a.tma_cfold[oi];
              const int cx = tx * a.bw + xl0 + sb.yadd * a.tma_xfold, cy = ty * a.bh + yl0, cf = tf * a.bf + fl0;
              tma_store_4d(oi ? &tmO1_hi : &tmO0_hi, stg_u32, cc, cx, cy, cf);
              if (splitp) tma_store_4d(oi ? &tmO1_lo : &tmO0_lo, stg_u32 + 2048, cc, cx, cy, cf);
              bulk_commit_group();
            }
          } else {
            // ---- lane = row direct stores
            act_tile<EPI_CHUNK>(v, act);
            if (od.mode == OUT_F32_NCHW) {
              // frames at the ABI edge: [f][N][Ho][Wo]; consecutive lanes are consecutive x -> coalesced per channel
              if (valid) {
#pragma unroll
                for (int j = 0; j < EPI_CHUNK; ++j) {
                  const int n = n0 + c + j;
                  if (n < a.N) ((float*)od.out)[(((size_t)f * a.N + n) * a.Ho + oy) * a.Wo + ox] = v[j];
                }
              }
            } else if (valid) {
              // bf16 operand planes without a TMA view, and narrow tiles (BN = 32)
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
                __nv_bfloat16* ph = (__nv_bfloat16*)od.out + ocol;
                __nv_bfloat16* pl = od.mode == OUT_BF16_SPLIT ? (__nv_bfloat16*)od.out_lo + ocol : nullptr;
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
      };
      using std::integral_constant;
      auto dispatch = [&](const int oi, const int c_lo, const int c_hi, const bool last_range) __attribute__((always_inline)) {
        const TcOut& od = a.o[oi];
        using IC = integral_constant<int, 0>;
        if constexpr (has_res) {
          // one fp32 destination with the residual branch: the ResBlocks use (no activation, ReLU on the residual)
          if (od.act == ACT_NONE && a.res_act == ACT_RELU) run_range(IC{}, integral_constant<int, ACT_NONE>{}, integral_constant<int, ACT_RELU>{}, oi, c_lo, c_hi, last_range);
          else run_range(IC{}, integral_constant<int, -1>{}, integral_constant<int, -1>{}, oi, c_lo, c_hi, last_range);
        } else if (EPI_CHUNK == 32 && od.mode == OUT_F32_NHWC) {
          if (od.act == ACT_NONE) run_range(IC{}, integral_constant<int, ACT_NONE>{}, integral_constant<int, -1>{}, oi, c_lo, c_hi, last_range);
          else run_range(IC{}, integral_constant<int, -1>{}, integral_constant<int, -1>{}, oi, c_lo, c_hi, last_range);
        } else if (EPI_CHUNK == 32 && a.tma_out[oi]) {
          // whole 32-column chunks leave as TMA boxes; a 16-column tail (N tiles of 240 columns) takes the direct stores
          const int c_full = c_lo + ((c_hi - c_lo) & ~31);
          const bool tail = c_full < c_hi;
          if (c_full > c_lo) {
            if (od.act == ACT_ELU) run_range(integral_constant<int, 1>{}, integral_constant<int, ACT_ELU>{}, integral_constant<int, -1>{}, oi, c_lo, c_full, last_range && !tail);
            else if (od.act == ACT_RELU) run_range(integral_constant<int, 1>{}, integral_constant<int, ACT_RELU>{}, integral_constant<int, -1>{}, oi, c_lo, c_full, last_range && !tail);
            else run_range(integral_constant<int, 1>{}, integral_constant<int, -1>{}, integral_constant<int, -1>{}, oi, c_lo, c_full, last_range && !tail);
          }
          if (tail) run_range(integral_constant<int, 2>{}, integral_constant<int, -1>{}, integral_constant<int, -1>{}, oi, c_full, c_hi, last_range);
        } else {
          run_range(integral_constant<int, 2>{}, integral_constant<int, -1>{}, integral_constant<int, -1>{}, oi, c_lo, c_hi, last_range);
        }
      };
      if (c_begin >= c_end) {
        release_stage();       // this warp's column half lies entirely in the padding of the last N tile
      } else {
        // the warp's columns split into at most two destination ranges (split_col is a multiple of 32)
        const int cs = a.split_col > 0 ? min(max(a.split_col - n0, c_begin), c_end) : c_end;
        if (cs > c_begin) dispatch(0, c_begin, cs, cs >= c_end);
        if (cs < c_end) dispatch(1, cs, c_end, true);
      }
      if (lane == 0 && (warp == 2 || warp == 9)) tc_trace(a, warp == 2 ? (esl == 10 ? 12 : 14) : (esl == 10 ? 15 : 16));
      if (++as == 2) { as = 0; aph ^= 1; }
    }
  }
  if (warp >= 2 && lane == 0) bulk_wait_read0();     // TMA stores still reading this CTA's shared memory
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (tid_ == 0) tc_trace(a, 17);
  if constexpr (CG == 2) cluster_sync_all();      // the leader's MMAs read the peer's shared memory: nobody leaves before both are done
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ launcher
inline int tc_sm_count() {
  static int n[IPK_MAX_DEVICES] = {0};
  const int slot = current_device_slot();
  if (n[slot] == 0) {
    int dev = 0;
    IPK_CUDA(cudaGetDevice(&dev));
    IPK_CUDA(cudaDeviceGetAttribute(&n[slot], cudaDevAttrMultiProcessorCount, dev));
  }
  return n[slot];
}

template <int BN, int NSPLIT, int FUSED, bool HALO, int CG>
inline void launch_tc(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi, const CUtensorMap& w_lo, const CUtensorMap* mo,
                      TcArgs& a, cudaStream_t st) {
  constexpr int STAGE_BYTES = HALO ? (NSPLIT == 3 ? 2 : 1) * (TC_HALO_A_BYTES + 3 * (BN / CG) * TC_BK * 2)
                                   : (NSPLIT == 3 ? 2 : 1) * (TC_BM * TC_BK * 2 + (BN / CG) * TC_BK * 2);
  int stages = (int)std::min<size_t>(8, TC_SMEM_BUDGET / STAGE_BYTES);
  IPK_CHECK(stages >= 2, IPK_ERR_UNSUPPORTED, "conv_tc: pipeline needs at least two stages (stage %d bytes)", STAGE_BYTES);
  a.stages = stages;
  // ring + epilogue transpose buffers + barriers (see the kernel's layout comment); no alignment slack: the dynamic
  // array is declared __align__(1024) and the kernel traps if its base is not
  constexpr size_t TAIL_BYTES = TC_EPI_WARPS * TC_EPI_STAGE_BYTES + 256;
  size_t smem = (size_t)stages * STAGE_BYTES + TAIL_BYTES;
  static bool attr_set[IPK_MAX_DEVICES] = {false};      // function attributes are per device
  const int slot = current_device_slot();
  if (!attr_set[slot]) {
    IPK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, NSPLIT, FUSED, HALO, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TC_SMEM_BUDGET + TAIL_BYTES)));
    attr_set[slot] = true;
  }
  const long long units = (long long)cdiv(a.tiles_m, CG) * a.tiles_n * (a.nsub > 1 ? a.nsub : a.nsplit);
  // persistent: one CTA per SM (CG = 2: one CTA pair per SM pair, grid a multiple of the cluster size)
  const unsigned grid = (unsigned)std::min<long long>(units * CG, (tc_sm_count() / CG) * CG);
  launch_kc(conv_tc_kernel<BN, NSPLIT, FUSED, HALO, CG>, dim3(grid), dim3(TC_THREADS), smem, st, CG, a_hi, a_lo, w_hi, w_lo, mo[0], mo[1], mo[2], mo[3], a);
}

// one launcher per N tile (conv_tc_bn*.cu): run-time (split precision, fused level, halo, CTA-pair) -> kernel instantiation
struct TcMaps { CUtensorMap a_hi, a_lo, w_hi, w_lo, o[4]; };
void tc_launch_bn32(bool split, int fused, bool halo, int cg, const TcMaps& m, TcArgs& a, cudaStream_t st);
void tc_launch_bn64(bool split, int fused, bool halo, int cg, const TcMaps& m, TcArgs& a, cudaStream_t st);
void tc_launch_bn128(bool split, int fused, bool halo, int cg, const TcMaps& m, TcArgs& a, cudaStream_t st);
void tc_launch_bn256(bool split, int fused, bool halo, int cg, const TcMaps& m, TcArgs& a, cudaStream_t st);

// expands to the (split, fused) dispatch of one (BN, HALO, CG) family
#define IPK_TC_FAMILY(bn, halo_, cg)                                                                  \
  do {                                                                                                \
    if (fused == 2) {                                                                                 \
      if (split) launch_tc<bn, 3, (bn >= 64 ? 2 : 0), halo_, cg>(m.a_hi, m.a_lo, m.w_hi, m.w_lo, m.o, a, st);   \
      else launch_tc<bn, 1, (bn >= 64 ? 2 : 0), halo_, cg>(m.a_hi, m.a_lo, m.w_hi, m.w_lo, m.o, a, st);         \
    } else if (fused == 1) {                                                                          \
      if (split) launch_tc<bn, 3, 1, halo_, cg>(m.a_hi, m.a_lo, m.w_hi, m.w_lo, m.o, a, st);          \
      else launch_tc<bn, 1, 1, halo_, cg>(m.a_hi, m.a_lo, m.w_hi, m.w_lo, m.o, a, st);                \
    } else {                                                                                          \
      if (split) launch_tc<bn, 3, 0, halo_, cg>(m.a_hi, m.a_lo, m.w_hi, m.w_lo, m.o, a, st);          \
      else launch_tc<bn, 1, 0, halo_, cg>(m.a_hi, m.a_lo, m.w_hi, m.w_lo, m.o, a, st);                \
    }                                                                                                 \
  } while (0)

}  // namespace ipk
