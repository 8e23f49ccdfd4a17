// Flow segment kernel: one CTA per sample, flow state resident in shared memory.
// Reference semantics (paths relative to the CompVis/ipoke checkout):
//   ActNorm2dFlow.forward            models/modules/INN/macow2.py:490-520
//   Shuffle.forward                  models/modules/INN/flow_blocks.py:314-326
//   MaskedConvFlow.forward/backward  models/modules/INN/macow2.py:97-151,174-288
//   MCFBlock / ShiftedConv2d         models/modules/INN/macow_utils.py:407-434,446-499
//   Affine.calc_params/fwd/bwd       models/modules/INN/macow_utils.py:49-66
//   NICE2d split/unsplit             models/modules/INN/macow2.py:364-388 (done as in-place channel index lists)
//
// Masked-conv flows (the 6 400-step sequential chain of the inverse) run in one of two modes:
//   * C <= 32 ("register-resident"): every thread keeps its slice of the shifted-conv weights (<= 96 floats) and of the
//     1x1 weights (<= 32 floats) in registers for the 8 lines of one MCF; the slices of the NEXT MCF are fetched from L2
//     while the last line of the current one finishes, and the x-independent conditioning term W1h*ELU(cond)+b
//     (precomputed for all MCFs by one GEMM, flow.cu) is prefetched one MCF ahead with cp.async into a double buffer.
//     A line then costs four phases: conv partials (2 tap rows on the two thread halves) -> combine + ELU ->
//     1x1 partials (K split four ways) -> reduce + affine transform.
//   * C <= 32 on the tensor-core precisions ("mma"): the two contractions of a line run on mma.sync.m16n8k16 bf16 with the
//     bf16x3 error-compensated split (hi*hi + lo*hi + hi*lo, fp32 accumulate) -- hidden units / outputs on the M axis, the 8
//     positions of the line on the N axis, so nothing is padded.  Weight fragments live in registers for the 8 lines of an
//     MCF (prefetched like above); the last two finished lines are kept as bf16 hi/lo rows with a zero halo in a 3-row
//     ring, so the shifted-conv operand is read straight from shared memory.  A line is two phases: conv + ELU, then
//     1x1 + conditioning term + affine transform entirely in the accumulator registers (mu and log-scale of a channel sit
//     in rows g and g+8 of the same m-tile).
//   * C > 32 (generic): weights streamed from L2 every line.
#include <type_traits>
#include "flow_segment.cuh"

namespace ipk {

constexpr int SEG_THREADS = 256;
constexpr int FAST_MAXC = 32;

struct SegSmem {
  float* s;       // [64][Cs] flow state
  float* tmp;     // [64][Cs] scratch (shuffle / forward snapshot)
  float* hterm;   // [2][64][C2s] conditioning term of the current / next MCF
  float* pc;      // fast: [2][8][128] conv partials      generic: unused
  float* act;     // [8][hidS] ELU(hidden) of the current line
  float* p1;      // fast: [4][8][64] 1x1 partials        generic: [8][C2s] params of the line
  float* red;     // [32]
  unsigned char* ring;   // mma: [3 lines][10 positions (zero halo at 0 and 9)][XSB bytes], hi/lo interleaved per k-tile
  unsigned char* actb;   // mma: [8 positions][HSB bytes] ELU(hidden) of the current line
  uint4* wst;            // mma: [slot][thread] conv fragments of the NEXT MCF, landed by cp.async during the current one
};

__host__ __device__ inline int seg_hidmax(int C) { return C <= 96 ? 4 * C : (2 * C < 512 ? 2 * C : 512); }

struct SegOffsets { size_t tmp, hterm, pc, act, p1, red, ring, actb, wst, total; };

// mma-path geometry: channels padded to 16 per k-tile, rows padded by 8 bf16 so fragment loads are bank-conflict free
// Operand rows of the mma path: per position, 64 bytes per 16-element k-tile = [t = 0..3][hi: k 2t, 2t+1, 2t+8, 2t+9 | lo: same],
// so lane (g, t) fetches b0/b1 of both planes with ONE 16-byte load.  Position strides are 64 (mod 128) bytes: the two
// positions served by one 8-lane LDS.128 phase hit disjoint bank halves.
__host__ __device__ inline int mma_xsb(int C) { const int n = (C + 15) / 16; return n * 64 + ((n & 1) ? 0 : 64); }          // bytes per ring position
__host__ __device__ inline int mma_hsb(int C) { const int n = (4 * C + 15) / 16; return n * 64 + ((n & 1) ? 0 : 64); }      // bytes per act position
// byte offset of element k (0..15) of a k-tile inside its 64-byte block, hi plane (lo plane: + 8)
__host__ __device__ inline int mma_koff(int k) { return ((k & 7) >> 1) * 16 + (((k >> 3) << 1) | (k & 1)) * 2; }
__host__ __device__ inline int mma_wst_slots(int C) { return 2 * 6 * ((C + 15) / 16); }     // staged conv fragments (uint4) per thread

// mode: 0 = FFMA paths, 1 = register-resident mma path (C <= 32), 2 = streamed mma path (32 < C <= 64)
__host__ __device__ inline SegOffsets seg_layout(int C, bool has_mcf, int mode) {
  const int Cs = (C + 3) / 4 * 4;
  const int C2s = (2 * C + 3) / 4 * 4;
  const bool fast = C <= FAST_MAXC && mode != 2;
  const bool mma = mode == 1;
  const bool big = has_mcf && mode == 2;
  SegOffsets o;
  size_t off = (size_t)64 * Cs;                       // s
  o.tmp = off; off += (size_t)64 * Cs;
  o.hterm = off; off += has_mcf ? (size_t)(fast ? 2 : 1) * 64 * C2s : 0;
  const bool use_mma = has_mcf && fast && mma;
  o.pc = off; off += (has_mcf && fast && !use_mma) ? (size_t)2 * 8 * 128 : 0;
  o.act = off; off += (has_mcf && !use_mma && !big) ? (size_t)8 * ((seg_hidmax(C) + 3) / 4 * 4) : 0;
  o.p1 = off; off += (has_mcf && !use_mma && !big) ? (fast ? (size_t)4 * 8 * 64 : (size_t)8 * C2s) : 0;
  o.red = off; off += 32;
  o.ring = off; off += (use_mma || big) ? (size_t)(3 * 10 * mma_xsb(C)) / 4 : 0;
  o.actb = off; off += (use_mma || big) ? (size_t)(8 * mma_hsb(C)) / 4 : 0;
  off = (off + 3) / 4 * 4;
  o.wst = off; off += use_mma ? (size_t)mma_wst_slots(C) * SEG_THREADS * 4 : 0;     // [slot][thread] uint4
  o.total = off;
  return o;
}

size_t flow_segment_smem_bytes(int C, bool has_mcf, int mode) {
  return seg_layout(C, has_mcf, mode).total * sizeof(float);
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  acc = fmaf(a.w, b.w, acc);
  return acc;
}

// pixel index of canonical coordinates (u = sequential axis, v = along-line axis) for MCF order o
__device__ __forceinline__ int mcf_pix(int order, int u, int v) {
  switch (order) {
    case 0: return u * 8 + v;          // A: rows top -> bottom
    case 1: return (7 - u) * 8 + v;    // B: rows bottom -> top
    case 2: return v * 8 + u;          // C: columns left -> right
    default: return v * 8 + (7 - u);   // D: columns right -> left
  }
}

__device__ __forceinline__ uint32_t seg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(seg_smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- fast MCF (C <= 32)
struct McfRegs {
  float4 wc[3][8];   // shifted-conv weights of (tap row = thread half, dv, channel quad) for hidden unit n = tid & 127
  float4 w1[8];      // 1x1 weights of (hidden quad kq*per + j) for output o = tid & 63, K quarter kq = tid >> 6
};

__device__ __forceinline__ void mcf_load_wc(const MicroOp& op, McfRegs& r) {
  const int tid = threadIdx.x, n = tid & 127, half = tid >> 7;
  const int Cp4 = op.i2 >> 2, hid = op.i3;
  const float4* Wc = (const float4*)op.p0;   // [6][Cp4][hid]
#pragma unroll
  for (int dv = 0; dv < 3; ++dv)
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4)
      r.wc[dv][c4] = (n < hid && c4 < Cp4) ? __ldg(Wc + ((size_t)(half * 3 + dv) * Cp4 + c4) * hid + n) : make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ void mcf_load_w1(const MicroOp& op, McfRegs& r) {
  const int tid = threadIdx.x, o = tid & 63, kq = tid >> 6;
  const int hid4 = op.i3 >> 2, C2 = 2 * op.i1;
  const int per = (hid4 + 3) >> 2;
  const float4* W1x = (const float4*)op.p1;  // [hid4][C2]
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k4 = kq * per + j;
    r.w1[j] = (o < C2 && j < per && k4 < hid4) ? __ldg(W1x + (size_t)k4 * C2 + o) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
// conditioning term block of sample b: rows (b*64 + p) of the [M][hstride] matrix, C2s columns from op.p2
__device__ __forceinline__ void mcf_prefetch_hterm(const MicroOp& op, int b, float* dst) {
  const int C2s = (2 * op.i1 + 3) / 4 * 4, q = C2s >> 2;
  const float* src = op.p2 + (size_t)b * 64 * op.l0;
  for (int i = threadIdx.x; i < 64 * q; i += SEG_THREADS) {
    const int p = i / q, j = i - p * q;
    cp_async16(dst + p * C2s + 4 * j, src + (size_t)p * op.l0 + 4 * j);
  }
}

template <bool FWD>
__device__ __forceinline__ void mcf_fast(const MicroOp& op, const MicroOp* next, McfRegs& r, const SegSmem& sm, const float* hterm,
                                         int Cs, float& ld) {
  const int tid = threadIdx.x;
  const int order = op.i0, C = op.i1, Cp4 = op.i2 >> 2, hid = op.i3;
  const int Cs4 = Cs >> 2, hid4 = hid >> 2;
  const int C2s = (2 * C + 3) / 4 * 4;
  const float4* s4 = (const float4*)(FWD ? sm.tmp : sm.s);
  const int n = tid & 127, half = tid >> 7;
  const int o = tid & 63, kq = tid >> 6;
  const int per = (hid4 + 3) >> 2;

  if (FWD) {  // snapshot x: all lines are computed from the un-transformed input (macow2.py:113-116)
    for (int i = tid; i < 64 * Cs; i += SEG_THREADS) sm.tmp[i] = sm.s[i];
    __syncthreads();
  }

  for (int u = 0; u < 8; ++u) {
    if (u > 0) {   // line 0 sees only zero padding: hidden = 0, ELU(0) = 0, params = conditioning term
      // ---- phase A: masked (shifted) conv partials; this thread's tap row is uu = u - 2 + half, taps dv in {-1,0,1}
      {
        float acc[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) acc[v] = 0.f;
        const int uu = u - 2 + half;
        if (n < hid && uu >= 0) {
          const int pb = mcf_pix(order, uu, 0) * Cs4, pv = (mcf_pix(order, uu, 1) - mcf_pix(order, uu, 0)) * Cs4;
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            if (c4 < Cp4) {
              float4 x[8];
#pragma unroll
              for (int v = 0; v < 8; ++v) x[v] = s4[pb + v * pv + c4];
#pragma unroll
              for (int dv = 0; dv < 3; ++dv)
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                  const int vv = v + dv - 1;
                  if (vv >= 0 && vv < 8) acc[v] = dot4(r.wc[dv][c4], x[vv], acc[v]);
                }
            }
          }
        }
        if (n < hid) {
#pragma unroll
          for (int v = 0; v < 8; ++v) sm.pc[(half * 8 + v) * 128 + n] = acc[v];
        }
      }
      __syncthreads();
      if (u == 7 && next) mcf_load_wc(*next, r);     // conv weights are dead: fetch the next MCF's while this line finishes
      // ---- phase B: combine the two tap rows, ELU
      for (int i = tid; i < 8 * hid; i += SEG_THREADS) {
        const int v = i / hid, nn = i - v * hid;
        const float h = sm.pc[v * 128 + nn] + sm.pc[(8 + v) * 128 + nn];
        sm.act[v * hid + nn] = h > 0.f ? h : expm1f(h);
      }
      __syncthreads();
      // ---- phase C: weight-normed 1x1 on ELU(hidden), K split over the four thread quarters
      {
        float acc[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) acc[v] = 0.f;
        const float4* act4 = (const float4*)sm.act;
        if (o < 2 * C) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k4 = kq * per + j;
            if (j < per && k4 < hid4) {
#pragma unroll
              for (int v = 0; v < 8; ++v) acc[v] = dot4(r.w1[j], act4[v * hid4 + k4], acc[v]);
            }
          }
#pragma unroll
          for (int v = 0; v < 8; ++v) sm.p1[(kq * 8 + v) * 64 + o] = acc[v];
        }
      }
      __syncthreads();
      if (u == 7 && next) mcf_load_w1(*next, r);
    }
    // ---- phase D: (mu, log_scale) = 1x1 partials + conditioning term; affine transform of line u (Affine.fwd / Affine.bwd)
    for (int i = tid; i < 8 * C; i += SEG_THREADS) {
      const int q = i / C, c = i - q * C;
      const int pix = mcf_pix(order, u, q);
      float mu = hterm[pix * C2s + c], ls = hterm[pix * C2s + C + c];
      if (u > 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          mu += sm.p1[(k * 8 + q) * 64 + c];
          ls += sm.p1[(k * 8 + q) * 64 + C + c];
        }
      }
      const float sc = 1.0f + tanhf(0.5f * ls);
      if (FWD) {
        sm.s[pix * Cs + c] = sc * sm.tmp[pix * Cs + c] + mu;
        ld += logf(sc);
      } else {
        sm.s[pix * Cs + c] = (sm.s[pix * Cs + c] - mu) / (sc + 1e-12f);
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------- mma MCF (C <= 32)
struct McfMmaRegs {
  uint4 wa_hi[12], wa_lo[12];   // conv A fragments of m-tile = warp, k-tile kt = (tap row r, dv, channel tile)
  uint2 w1_hi[8], w1_lo[8];     // 1x1 A fragments of m-tile = warp: rows 0-3 = mu of channels 4*mt.., rows 4-7 = their
                                // log-scales, rows 8-15 = zero (only a0 / a2 are stored)
};

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

// packed fragment arrays: [m-tile][k-tile][lane] uint4, hi plane then lo plane
// register slot (tap, ct) = tap*2 + ct holds packed k-tile tap*nct + ct  (tap = tap row * 3 + dv)
__device__ __forceinline__ void mma_load_wa(const MicroOp& op, McfMmaRegs& r) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = op.i1, hid = op.i3;
  const int nmt = (hid + 15) >> 4, nct = (C + 15) >> 4, nkt = 6 * nct;
  const uint4* W = (const uint4*)op.p0;
  const size_t plane = (size_t)nmt * nkt * 32;
#pragma unroll
  for (int tap = 0; tap < 6; ++tap)
#pragma unroll
    for (int ct = 0; ct < 2; ++ct) {
      if (warp < nmt && ct < nct) {
        const size_t i = ((size_t)warp * nkt + tap * nct + ct) * 32 + lane;
        r.wa_hi[tap * 2 + ct] = __ldg(W + i);
        r.wa_lo[tap * 2 + ct] = __ldg(W + plane + i);
      }
    }
}
// the same fragments, global -> this thread's private staging slots (cp.async; joins the caller's commit group)
__device__ __forceinline__ void mma_stage_wa(const MicroOp& op, uint4* wst) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = op.i1, hid = op.i3;
  const int nmt = (hid + 15) >> 4, nct = (C + 15) >> 4, nkt = 6 * nct;
  const uint4* W = (const uint4*)op.p0;
  const size_t plane = (size_t)nmt * nkt * 32;
  if (warp < nmt) {
    for (int kt = 0; kt < nkt; ++kt) {
      const size_t i = ((size_t)warp * nkt + kt) * 32 + lane;
      cp_async16(wst + (size_t)(2 * kt) * SEG_THREADS + tid, W + i);
      cp_async16(wst + (size_t)(2 * kt + 1) * SEG_THREADS + tid, W + plane + i);
    }
  }
}
__device__ __forceinline__ void mma_unstage_wa(const MicroOp& op, const uint4* wst, McfMmaRegs& r) {
  const int tid = threadIdx.x, warp = tid >> 5;
  const int C = op.i1, hid = op.i3;
  const int nmt = (hid + 15) >> 4, nct = (C + 15) >> 4;
#pragma unroll
  for (int tap = 0; tap < 6; ++tap)
#pragma unroll
    for (int ct = 0; ct < 2; ++ct) {
      if (warp < nmt && ct < nct) {
        const int kt = tap * nct + ct;
        r.wa_hi[tap * 2 + ct] = wst[(size_t)(2 * kt) * SEG_THREADS + tid];
        r.wa_lo[tap * 2 + ct] = wst[(size_t)(2 * kt + 1) * SEG_THREADS + tid];
      }
    }
}
__device__ __forceinline__ void mma_load_w1(const MicroOp& op, McfMmaRegs& r) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = op.i1, hid = op.i3;
  const int nmt1 = (C + 3) >> 2, nkt1 = (hid + 15) >> 4;
  const uint2* W = (const uint2*)op.p1;
  const size_t plane = (size_t)nmt1 * nkt1 * 32;
#pragma unroll
  for (int kt = 0; kt < 8; ++kt) {
    if (warp < nmt1 && kt < nkt1) {
      r.w1_hi[kt] = __ldg(W + ((size_t)warp * nkt1 + kt) * 32 + lane);
      r.w1_lo[kt] = __ldg(W + plane + ((size_t)warp * nkt1 + kt) * 32 + lane);
    }
  }
}

template <bool FWD>
__device__ __forceinline__ void mcf_mma(const MicroOp& op, const MicroOp* next, McfMmaRegs& r, const SegSmem& sm, const float* hterm,
                                        int Cs, float& ld) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int order = op.i0, C = op.i1, hid = op.i3;
  const int nct = (C + 15) >> 4;
  const int nmt = (hid + 15) >> 4, nkt1 = nmt, nmt1 = (C + 3) >> 2;
  const int XSB = mma_xsb(C), HSB = mma_hsb(C);
  const int C2s = (2 * C + 3) / 4 * 4;
  // pixel of (line u, position v) = pb + u*pu + v*pv
  const int pb = order == 1 ? 56 : (order == 3 ? 7 : 0);
  const int pu = order == 0 ? 8 : (order == 1 ? -8 : (order == 2 ? 1 : -1));
  const int pv = order < 2 ? 1 : 8;

  if (FWD) {  // snapshot x: all lines are computed from the un-transformed input (macow2.py:113-116)
    for (int i = tid; i < 64 * Cs; i += SEG_THREADS) sm.tmp[i] = sm.s[i];
    __syncthreads();
  }

  // phase D: all eight warps; lane (g, t) finishes channel c = 4*warp + (g & 3) at position 2t + (g >> 2)
  const int c = warp * 4 + (g & 3);
  const int dpos = t * 2 + (g >> 2);
  const int c_ring = (c >> 4) * 64 + mma_koff(c & 15);
  // phase A epilogue: hidden units n0 = 16*warp + g and n0 + 8 -> act offsets
  const int a_off0 = warp * 64 + mma_koff(g), a_off1 = warp * 64 + mma_koff(g + 8);

  int slot2 = 1, slot1 = 2, slot0 = 0;   // ring slots of lines u-2, u-1, u
  for (int u = 0; u < 8; ++u) {
    if (u > 0) {   // line 0 sees only zero padding: hidden = 0, ELU(0) = 0, params = conditioning term
      // ---- phase A: hidden^T[n][pos] = sum_k Wc^T[n][k] * X^T[k][pos], k = (tap row, dv, channel); then ELU -> act (bf16 hi/lo)
      if (warp < nmt) {
        float d[6][4];
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int uu = u - 2 + rr;
          if (uu >= 0) {          // warp-uniform; earlier lines are zero padding
            const unsigned char* rowp = sm.ring + ((rr == 0 ? slot2 : slot1) * 10 + g) * XSB + t * 16;
#pragma unroll
            for (int dvi = 0; dvi < 3; ++dvi)
#pragma unroll
              for (int ct = 0; ct < 2; ++ct) {
                if (ct < nct) {
                  const int kt = (rr * 3 + dvi) * 2 + ct;
                  const uint4 bq = *(const uint4*)(rowp + dvi * XSB + ct * 64);     // {b0 hi, b1 hi, b0 lo, b1 lo}
                  const int ch = (dvi & 1) * 3;
                  mma_bf16(d[ch + 0], r.wa_hi[kt], bq.x, bq.y);
                  mma_bf16(d[ch + 1], r.wa_lo[kt], bq.x, bq.y);
                  mma_bf16(d[ch + 2], r.wa_hi[kt], bq.z, bq.w);
                }
              }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float h = ((d[0][j] + d[3][j]) + (d[1][j] + d[4][j])) + (d[2][j] + d[5][j]);
          h = h > 0.f ? h : (exp2f(h * 1.4426950408889634f) - 1.0f);     // ELU; abs error ~1e-7, below the bf16x3 operand split
          const int pos = t * 2 + (j & 1);
          if (warp * 16 + g + (j >> 1) * 8 < hid) {
            unsigned char* ap = sm.actb + pos * HSB + ((j >> 1) ? a_off1 : a_off0);
            const __nv_bfloat16 hi = __float2bfloat16_rn(h);
            *(__nv_bfloat16*)ap = hi;
            *(__nv_bfloat16*)(ap + 8) = __float2bfloat16_rn(h - __bfloat162float(hi));
          }
        }
      }
      __syncthreads();
      if (u == 7 && next) {     // conv fragments are dead: take the next MCF's from the staging slots (landed during this MCF)
        cp_async_wait0();
        mma_unstage_wa(*next, sm.wst, r);
      }
    }
    // ---- phase C + D: params^T[o][pos] = sum_k W1x^T[o][k] * act^T[k][pos] + conditioning term; affine transform of line u.
    // m-tile of a warp: rows 0-3 = mu of its 4 channels, rows 4-7 = their log-scales -> accumulator row g of lane (g, t) holds
    // columns (positions) 2t, 2t+1; one xor-16 shuffle pairs mu and log-scale, each lane then finishes ONE (channel, position).
    if (warp < nmt1) {
      float d[6][4];
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
      if (u > 0) {
        const unsigned char* ap = sm.actb + g * HSB + t * 16;
#pragma unroll
        for (int kt = 0; kt < 8; ++kt) {
          if (kt < nkt1) {
            const uint4 bq = *(const uint4*)(ap + kt * 64);
            const int ch = (kt & 1) * 3;
            const uint4 ahi = make_uint4(r.w1_hi[kt].x, 0u, r.w1_hi[kt].y, 0u), alo = make_uint4(r.w1_lo[kt].x, 0u, r.w1_lo[kt].y, 0u);
            mma_bf16(d[ch + 0], ahi, bq.x, bq.y);
            mma_bf16(d[ch + 1], alo, bq.x, bq.y);
            mma_bf16(d[ch + 2], ahi, bq.z, bq.w);
          }
        }
      }
      const float v0 = ((d[0][0] + d[3][0]) + (d[1][0] + d[4][0])) + (d[2][0] + d[5][0]);     // row g, position 2t
      const float v1 = ((d[0][1] + d[3][1]) + (d[1][1] + d[4][1])) + (d[2][1] + d[5][1]);     // row g, position 2t+1
      const bool is_mu = g < 4;
      const float got = __shfl_xor_sync(0xffffffffu, is_mu ? v1 : v0, 16);
      if (c < C) {
        const int pix = pb + u * pu + dpos * pv;
        const float mu = (is_mu ? v0 : got) + hterm[pix * C2s + c];
        const float ls = (is_mu ? got : v1) + hterm[pix * C2s + C + c];
        // density direction: the reference's own formulation, so the log-det rounding stays correlated with it.
        // sampling direction (latency-critical): 1 + tanh(ls/2) == 2 / (1 + exp(-ls)) on the SFU.
        const float sc = FWD ? 1.0f + tanhf(0.5f * ls) : 2.0f * __frcp_rn(1.0f + exp2f(-1.4426950408889634f * ls));
        float xin;       // value of this element in the un-transformed domain: conv input of the following lines
        if (FWD) {
          xin = sm.tmp[pix * Cs + c];
          sm.s[pix * Cs + c] = sc * xin + mu;
          ld += logf(sc);
        } else {
          xin = (sm.s[pix * Cs + c] - mu) * __frcp_rn(sc + 1e-12f);
          sm.s[pix * Cs + c] = xin;
        }
        unsigned char* rp = sm.ring + (slot0 * 10 + dpos + 1) * XSB + c_ring;
        const __nv_bfloat16 hi = __float2bfloat16_rn(xin);
        *(__nv_bfloat16*)rp = hi;
        *(__nv_bfloat16*)(rp + 8) = __float2bfloat16_rn(xin - __bfloat162float(hi));
      }
    }
    if (u == 7 && next) mma_load_w1(*next, r);
    __syncthreads();
    { const int tmp = slot2; slot2 = slot1; slot1 = slot0; slot0 = tmp; }
  }
}

// ---------------------------------------------------------------------------------------------- streamed mma MCF (32 < C <= 64)
// Same contractions, operand ring and fragment packing as mcf_mma, but hid = 4C <= 256 hidden units are 16 m-tiles and K = 6C is
// up to 24 k-tiles: the A fragments no longer fit in registers, so every warp streams the fragments of its m-tiles
// {warp, warp + 8} from L2 each line (coalesced 512-byte warp loads, one (tap row, dv) group = up to 8 fragments prefetched
// while the previous group's MMAs issue).  The conditioning term is loaded at the start of the MCF.
struct BigFrag { uint4 hi[4], lo[4]; };
__device__ __forceinline__ void big_load_group(BigFrag& f, const uint4* __restrict__ wp, size_t plane, int kt0, int nct) {
#pragma unroll
  for (int ct = 0; ct < 4; ++ct)
    if (ct < nct) {
      f.hi[ct] = __ldg(wp + (size_t)(kt0 + ct) * 32);
      f.lo[ct] = __ldg(wp + plane + (size_t)(kt0 + ct) * 32);
    }
}

template <bool FWD>
__device__ __forceinline__ void mcf_mma_big(const MicroOp& op, const SegSmem& sm, int Cs, int b, float& ld) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int order = op.i0, C = op.i1, hid = op.i3;
  const int nct = (C + 15) >> 4, nkt = 6 * nct;
  const int nmt = (hid + 15) >> 4, nkt1 = nmt, nmt1 = (C + 3) >> 2;
  const int XSB = mma_xsb(C), HSB = mma_hsb(C);
  const int C2s = (2 * C + 3) / 4 * 4;
  const int pb = order == 1 ? 56 : (order == 3 ? 7 : 0);
  const int pu = order == 0 ? 8 : (order == 1 ? -8 : (order == 2 ? 1 : -1));
  const int pv = order < 2 ? 1 : 8;
  const uint4* WA = (const uint4*)op.p0;
  const size_t planeA = (size_t)nmt * nkt * 32;
  const uint2* W1 = (const uint2*)op.p1;
  const size_t plane1 = (size_t)nmt1 * nkt1 * 32;
  float* hterm = sm.hterm;
  {  // conditioning term of this MCF (precomputed GEMM, flow.cu): rows b*64 + p, C2s columns
    const int q = C2s >> 2;
    const float* src = op.p2 + (size_t)b * 64 * op.l0;
    for (int i = tid; i < 64 * q; i += SEG_THREADS) {
      const int p = i / q, j = i - p * q;
      *(float4*)(hterm + p * C2s + 4 * j) = __ldg((const float4*)(src + (size_t)p * op.l0 + 4 * j));
    }
  }
  if (FWD) {
    for (int i = tid; i < 64 * Cs; i += SEG_THREADS) sm.tmp[i] = sm.s[i];
  }
  __syncthreads();

  const int dpos = t * 2 + (g >> 2);
  int slot2 = 1, slot1 = 2, slot0 = 0;
  for (int u = 0; u < 8; ++u) {
    if (u > 0) {
      // ---- phase A: hidden = shifted conv of the last two finished lines, ELU -> act rows (bf16 hi/lo)
      for (int mt = warp; mt < nmt; mt += SEG_THREADS / 32) {
        float d[6][4];
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
        const uint4* wp = WA + (size_t)mt * nkt * 32 + lane;
        const int g0 = u >= 2 ? 0 : 3;                    // first (tap row, dv) group with a real line behind it
        BigFrag fa, fb;
        if (g0 & 1) big_load_group(fb, wp, planeA, g0 * nct, nct);      // group gi lives in fa (even gi) / fb (odd gi)
        else big_load_group(fa, wp, planeA, g0 * nct, nct);
#pragma unroll
        for (int gi = 0; gi < 6; ++gi) {
          if (gi < g0) continue;                           // warp-uniform
          BigFrag& cur = (gi & 1) ? fb : fa;
          BigFrag& nxt = (gi & 1) ? fa : fb;
          if (gi + 1 < 6) big_load_group(nxt, wp, planeA, (gi + 1) * nct, nct);
          const int rr = gi / 3, dvi = gi - rr * 3;
          const unsigned char* rowp = sm.ring + ((rr == 0 ? slot2 : slot1) * 10 + g) * XSB + t * 16 + dvi * XSB;
          const int ch = (dvi & 1) * 3;
#pragma unroll
          for (int ct = 0; ct < 4; ++ct)
            if (ct < nct) {
              const uint4 bq = *(const uint4*)(rowp + ct * 64);
              mma_bf16(d[ch + 0], cur.hi[ct], bq.x, bq.y);
              mma_bf16(d[ch + 1], cur.lo[ct], bq.x, bq.y);
              mma_bf16(d[ch + 2], cur.hi[ct], bq.z, bq.w);
            }
        }
        const int a_off0 = mt * 64 + mma_koff(g), a_off1 = mt * 64 + mma_koff(g + 8);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float h = ((d[0][j] + d[3][j]) + (d[1][j] + d[4][j])) + (d[2][j] + d[5][j]);
          h = h > 0.f ? h : (exp2f(h * 1.4426950408889634f) - 1.0f);
          const int pos = t * 2 + (j & 1);
          if (mt * 16 + g + (j >> 1) * 8 < hid) {
            unsigned char* ap = sm.actb + pos * HSB + ((j >> 1) ? a_off1 : a_off0);
            const __nv_bfloat16 hi = __float2bfloat16_rn(h);
            *(__nv_bfloat16*)ap = hi;
            *(__nv_bfloat16*)(ap + 8) = __float2bfloat16_rn(h - __bfloat162float(hi));
          }
        }
      }
      __syncthreads();
    }
    // ---- phase C + D: 1x1 on the act rows + conditioning term, affine transform of line u (as in mcf_mma)
    for (int mt1 = warp; mt1 < nmt1; mt1 += SEG_THREADS / 32) {
      float d[6][4];
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
      if (u > 0) {
        const unsigned char* ap = sm.actb + g * HSB + t * 16;
        const uint2* wp1 = W1 + (size_t)mt1 * nkt1 * 32 + lane;
#pragma unroll 1
        for (int k0 = 0; k0 < nkt1; k0 += 8) {
          uint2 ah[8], al[8];
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k0 + k < nkt1) { ah[k] = __ldg(wp1 + (size_t)(k0 + k) * 32); al[k] = __ldg(wp1 + plane1 + (size_t)(k0 + k) * 32); }
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k0 + k < nkt1) {
              const uint4 bq = *(const uint4*)(ap + (k0 + k) * 64);
              const int ch = (k & 1) * 3;
              const uint4 ahi = make_uint4(ah[k].x, 0u, ah[k].y, 0u), alo = make_uint4(al[k].x, 0u, al[k].y, 0u);
              mma_bf16(d[ch + 0], ahi, bq.x, bq.y);
              mma_bf16(d[ch + 1], alo, bq.x, bq.y);
              mma_bf16(d[ch + 2], ahi, bq.z, bq.w);
            }
        }
      }
      const float v0 = ((d[0][0] + d[3][0]) + (d[1][0] + d[4][0])) + (d[2][0] + d[5][0]);
      const float v1 = ((d[0][1] + d[3][1]) + (d[1][1] + d[4][1])) + (d[2][1] + d[5][1]);
      const bool is_mu = g < 4;
      const float got = __shfl_xor_sync(0xffffffffu, is_mu ? v1 : v0, 16);
      const int c = mt1 * 4 + (g & 3);
      if (c < C) {
        const int pix = pb + u * pu + dpos * pv;
        const float mu = (is_mu ? v0 : got) + hterm[pix * C2s + c];
        const float ls = (is_mu ? got : v1) + hterm[pix * C2s + C + c];
        const float sc = FWD ? 1.0f + tanhf(0.5f * ls) : 2.0f * __frcp_rn(1.0f + exp2f(-1.4426950408889634f * ls));
        float xin;
        if (FWD) {
          xin = sm.tmp[pix * Cs + c];
          sm.s[pix * Cs + c] = sc * xin + mu;
          ld += logf(sc);
        } else {
          xin = (sm.s[pix * Cs + c] - mu) * __frcp_rn(sc + 1e-12f);
          sm.s[pix * Cs + c] = xin;
        }
        unsigned char* rp = sm.ring + (slot0 * 10 + dpos + 1) * XSB + (c >> 4) * 64 + mma_koff(c & 15);
        const __nv_bfloat16 hi = __float2bfloat16_rn(xin);
        *(__nv_bfloat16*)rp = hi;
        *(__nv_bfloat16*)(rp + 8) = __float2bfloat16_rn(xin - __bfloat162float(hi));
      }
    }
    __syncthreads();
    { const int tmp = slot2; slot2 = slot1; slot1 = slot0; slot0 = tmp; }
  }
}

// ---------------------------------------------------------------------------------------------- generic MCF (C > 32)
template <bool FWD>
__device__ void mcf_generic(const MicroOp& op, const SegSmem& sm, int Cs, int b, float& ld) {
  const int tid = threadIdx.x;
  const int order = op.i0, C = op.i1, Cp = op.i2, hid = op.i3;
  const int Cp4 = Cp / 4, Cs4 = Cs / 4;
  const int C2 = 2 * C;
  const int C2s = (C2 + 3) / 4 * 4;            // row stride of P / hterm
  const float4* Wc = (const float4*)op.p0;     // [6][Cp4][hid]
  const float4* W1x = (const float4*)op.p1;    // [hid/4][2C]
  const float4* s4 = (const float4*)(FWD ? sm.tmp : sm.s);
  float* P = sm.p1;

  if (FWD) {
    for (int i = tid; i < 64 * Cs; i += SEG_THREADS) sm.tmp[i] = sm.s[i];
  }
  {  // conditioning term (precomputed): hterm[p][o] = b[o] + sum_k W1h[k][o] * ELU(cond[p][k])
    const float* src = op.p2 + (size_t)b * 64 * op.l0;
    for (int i = tid; i < 64 * C2s; i += SEG_THREADS) {
      const int p = i / C2s, j = i - p * C2s;
      sm.hterm[i] = src[(size_t)p * op.l0 + j];
    }
  }
  __syncthreads();

  int nslots = min(SEG_THREADS, (hid + 31) / 32 * 32);
  int Gn = SEG_THREADS / nslots;
  Gn = Gn >= 8 ? 8 : (Gn >= 4 ? 4 : (Gn >= 2 ? 2 : 1));
  const int qper_n = 8 / Gn, gn = tid / nslots, nl = tid % nslots;
  int oslots = min(SEG_THREADS, (C2 + 31) / 32 * 32);
  int Go = SEG_THREADS / oslots;
  Go = Go >= 8 ? 8 : (Go >= 4 ? 4 : (Go >= 2 ? 2 : 1));
  const int qper_o = 8 / Go, go = tid / oslots, ol = tid % oslots;
  const int hid4 = hid / 4;

  for (int u = 0; u < 8; ++u) {
    if (gn < Gn) {
      for (int n = nl; n < hid; n += nslots) {
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        for (int t = 0; t < 6; ++t) {
          int uu = u + (t / 3) - 2;
          if (uu < 0) continue;
          int dv = (t % 3) - 1;
          const float4* wrow = Wc + (size_t)t * Cp4 * hid + n;
          for (int c4 = 0; c4 < Cp4; ++c4) {
            float4 w = __ldg(wrow + (size_t)c4 * hid);
#pragma unroll
            for (int qi = 0; qi < 8; ++qi) {
              if (qi < qper_n) {
                int v = gn * qper_n + qi + dv;
                if (v >= 0 && v < 8) acc[qi] = dot4(w, s4[mcf_pix(order, uu, v) * Cs4 + c4], acc[qi]);
              }
            }
          }
        }
#pragma unroll
        for (int qi = 0; qi < 8; ++qi)
          if (qi < qper_n) {
            float h = acc[qi];
            sm.act[(gn * qper_n + qi) * hid + n] = h > 0.f ? h : expm1f(h);
          }
      }
    }
    __syncthreads();
    if (go < Go) {
      const float4* act4 = (const float4*)sm.act;
      for (int o = ol; o < C2; o += oslots) {
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        for (int n4 = 0; n4 < hid4; ++n4) {
          float4 w = __ldg(W1x + (size_t)n4 * C2 + o);
#pragma unroll
          for (int qi = 0; qi < 8; ++qi)
            if (qi < qper_o) acc[qi] = dot4(w, act4[(go * qper_o + qi) * hid4 + n4], acc[qi]);
        }
#pragma unroll
        for (int qi = 0; qi < 8; ++qi)
          if (qi < qper_o) {
            int q = go * qper_o + qi;
            P[q * C2s + o] = acc[qi] + sm.hterm[mcf_pix(order, u, q) * C2s + o];
          }
      }
    }
    __syncthreads();
    for (int i = tid; i < 8 * C; i += SEG_THREADS) {
      int q = i / C, c = i % C;
      float mu = P[q * C2s + c];
      float sc = 1.0f + tanhf(0.5f * P[q * C2s + C + c]);
      int pix = mcf_pix(order, u, q);
      if (FWD) {
        sm.s[pix * Cs + c] = sc * sm.tmp[pix * Cs + c] + mu;
        ld += logf(sc);
      } else {
        sm.s[pix * Cs + c] = (sm.s[pix * Cs + c] - mu) / (sc + 1e-12f);
      }
    }
    __syncthreads();
  }
}

// Finishes a NICE coupling: params[p] = bias + sum over taps t and split-K slices of T[p + delta_t][t]  (the 3x3 gather of
// conv3 with zero padding), then Affine.fwd / bwd on the transformed channels.
// Step 1 (lane = parameter index e: mu of element e, or log-scale of element e - n_p): a warp reads 2*n_p contiguous floats
// per (pixel, tap, slice); PB pixels x nine taps of loads are in flight per lane.  Step 2 (lane = element).
template <bool FWD, int PB>
__device__ __forceinline__ void affine_op(const MicroOp& op, const SegSmem& sm, int Cs, int b, float& ld) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nsplit = op.i0, Npad = op.i1, n_p = op.i2, N3p = op.i3;
  const float* Tb = op.p0 + (size_t)b * 64 * Npad;
  constexpr int NW = SEG_THREADS / 32;
  for (int e = lane; e < 2 * n_p; e += 32) {
    const float bias = __ldg(op.p1 + e);
#pragma unroll 1
    for (int pg = 0; pg < 64 / NW; pg += PB) {
      float v[PB];
#pragma unroll
      for (int q = 0; q < PB; ++q) v[q] = bias;
      for (int sidx = 0; sidx < nsplit; ++sidx) {
        const float* T = Tb + (size_t)sidx * op.l0 + e;
        float pv[PB][9];
#pragma unroll
        for (int q = 0; q < PB; ++q) {
          const int pp = warp + (pg + q) * NW;
          const int y = pp >> 3, x = pp & 7;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
            const bool ok = yy >= 0 && yy < 8 && xx >= 0 && xx < 8;
            pv[q][t] = ok ? T[(size_t)(yy * 8 + xx) * Npad + t * N3p] : 0.f;
          }
        }
#pragma unroll
        for (int q = 0; q < PB; ++q)
#pragma unroll
          for (int t = 0; t < 9; ++t) v[q] += pv[q][t];
      }
#pragma unroll
      for (int q = 0; q < PB; ++q) sm.tmp[(warp + (pg + q) * NW) * Cs + e] = v[q];
    }
  }
  __syncwarp();
  for (int j = lane; j < n_p; j += 32) {
    const int c = op.idx[j];
    for (int pp = warp; pp < 64; pp += NW) {
      const float mu = sm.tmp[pp * Cs + j], ls = sm.tmp[pp * Cs + n_p + j];
      const float sc = 1.0f + tanhf(0.5f * ls);
      const float xv = sm.s[pp * Cs + c];
      if (FWD) {
        sm.s[pp * Cs + c] = sc * xv + mu;
        ld += logf(sc);
      } else {
        sm.s[pp * Cs + c] = (xv - mu) / (sc + 1e-12f);
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ int next_mcf(const MicroOp* __restrict__ ops, int from, int nops) {
  for (int i = from; i < nops; ++i)
    if (ops[i].kind == MK_MCF) return i;
  return -1;
}

// MODE: 0 = FFMA MCF paths, 1 = register-resident mma MCFs (C <= 32), 2 = streamed mma MCFs (32 < C <= 64)
template <bool FWD, int MODE>
__global__ void __launch_bounds__(SEG_THREADS, 1) flow_segment_kernel(const MicroOp* __restrict__ ops, int nops, int C, int has_mcf,
                                                                       float* __restrict__ state, int C0,
                                                                       float* __restrict__ logdet, int b0) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int b = blockIdx.x + b0;
  const int Cs = (C + 3) / 4 * 4;
  const int C2s = (2 * C + 3) / 4 * 4;
  constexpr bool MMA = MODE == 1;
  constexpr bool BIG = MODE == 2;
  const bool fast = !BIG && C <= FAST_MAXC;
  SegSmem sm;
  const SegOffsets lo = seg_layout(C, has_mcf != 0, MODE);
  sm.s = smem; sm.tmp = smem + lo.tmp; sm.hterm = smem + lo.hterm; sm.pc = smem + lo.pc; sm.act = smem + lo.act;
  sm.p1 = smem + lo.p1; sm.red = smem + lo.red;
  sm.ring = (unsigned char*)(smem + lo.ring); sm.actb = (unsigned char*)(smem + lo.actb); sm.wst = (uint4*)(smem + lo.wst);

  const int warp = tid >> 5, lane = tid & 31;
  float* gs = state + (size_t)b * 64 * C0;
  pdl_wait();
  pdl_trigger();
  // (pixel by warp, channel by lane) loops everywhere below: no integer divisions on the latency chain
  for (int p = warp; p < 64; p += SEG_THREADS / 32)
    for (int c = lane; c < Cs; c += 32) sm.s[p * Cs + c] = c < C ? gs[p * C0 + c] : 0.f;
  __syncthreads();

  float ld = 0.f;
  // A segment that follows a coupling network starts with that coupling's affine update.  Run it before the MCF weight
  // registers become live: the whole 8-pixel x 9-tap gather of a warp is then in flight at once.
  int op_begin = 0;
  if (nops > 0 && ops[0].kind == MK_AFFINE) {
    const MicroOp op0 = ops[0];
    affine_op<FWD, 8>(op0, sm, Cs, b, ld);
    op_begin = 1;
  }

  // first MCF of the segment: start fetching its weights / conditioning term
  typename std::conditional<MMA, McfMmaRegs, McfRegs>::type regs;
  int hbuf = 0;
  const MicroOp* nxt = nullptr;      // next MCF micro-op (read from global on demand: keeps ~20 registers free)
  int nxt_i = (has_mcf && fast) ? next_mcf(ops, op_begin, nops) : -1;
  if (nxt_i >= 0) {
    nxt = ops + nxt_i;
    mcf_prefetch_hterm(*nxt, b, sm.hterm);
    cp_async_commit();
    if constexpr (MMA) {
      mma_load_wa(*nxt, regs);
      mma_load_w1(*nxt, regs);
      // zero the operand ring (halo positions and channel padding stay zero for the whole segment) and the act rows
      uint32_t* z = (uint32_t*)sm.ring;
      const int nz = (int)(lo.wst - lo.ring);
      for (int i = tid; i < nz; i += SEG_THREADS) z[i] = 0u;
      __syncthreads();
    } else {
      mcf_load_wc(*nxt, regs);
      mcf_load_w1(*nxt, regs);
    }
  }

  if constexpr (BIG) {
    if (has_mcf) {   // operand ring (zero halo, channel padding) and act rows start from zero
      uint32_t* z = (uint32_t*)sm.ring;
      const int nz = (int)(lo.wst - lo.ring);
      for (int i = tid; i < nz; i += SEG_THREADS) z[i] = 0u;
      __syncthreads();
    }
  }

  for (int oi = op_begin; oi < nops; ++oi) {
    const MicroOp op = ops[oi];
    switch (op.kind) {
      case MK_ACTNORM: {
        const int coff = op.i0, cnt = op.i1;
        for (int c = lane; c < cnt; c += 32) {
          const float ls = __ldg(op.p0 + c), bb = __ldg(op.p1 + c);
          const float e = expf(ls);
          for (int p = warp; p < 64; p += SEG_THREADS / 32) {
            const float x = sm.s[p * Cs + coff + c];
            if (FWD) {
              sm.s[p * Cs + coff + c] = x * e + bb;
              ld += ls;   // H*W*sum(log_scale): one contribution per (pixel, channel)
            } else {
              sm.s[p * Cs + coff + c] = (x - bb) / (e + 1e-8f);
            }
          }
        }
        __syncthreads();
        break;
      }
      case MK_SHUFFLE: {
        const int Cn = op.i0;
        for (int c = lane; c < Cn; c += 32) {
          const int src = op.idx[c];
          for (int p = warp; p < 64; p += SEG_THREADS / 32) sm.tmp[p * Cs + c] = sm.s[p * Cs + src];
        }
        __syncthreads();
        for (int c = lane; c < Cn; c += 32)
          for (int p = warp; p < 64; p += SEG_THREADS / 32) sm.s[p * Cs + c] = sm.tmp[p * Cs + c];
        __syncthreads();
        break;
      }
      case MK_MCF: {
        if constexpr (BIG) {
          mcf_mma_big<FWD>(op, sm, Cs, b, ld);
        } else if (fast) {
          // registers hold this MCF's weights; its conditioning term is the (possibly still pending) cp.async group of
          // buffer hbuf.  Queue the NEXT MCF's conditioning term into the other buffer, then wait for ours.
          const int nn = next_mcf(ops, oi + 1, nops);
          if (nn >= 0) {
            nxt = ops + nn;
            mcf_prefetch_hterm(*nxt, b, sm.hterm + (hbuf ^ 1) * 64 * C2s);
            if constexpr (MMA) mma_stage_wa(*nxt, sm.wst);     // same commit group: lands while this MCF runs
            cp_async_commit();
            cp_async_wait1();
          } else {
            cp_async_wait0();
          }
          __syncthreads();
          if constexpr (MMA) mcf_mma<FWD>(op, nn >= 0 ? nxt : nullptr, regs, sm, sm.hterm + hbuf * 64 * C2s, Cs, ld);
          else mcf_fast<FWD>(op, nn >= 0 ? nxt : nullptr, regs, sm, sm.hterm + hbuf * 64 * C2s, Cs, ld);
          hbuf ^= 1;
        } else {
          mcf_generic<FWD>(op, sm, Cs, b, ld);
        }
        break;
      }
      case MK_AFFINE:
        affine_op<FWD, 2>(op, sm, Cs, b, ld);
        break;
      case MK_IM2COL: {
        // operand rows of the next coupling's conv1: A1[b*64 + p][k = tap*n_z + j] = z-part of the state at pixel p + delta_tap
        const int n_z = op.i0, K1 = op.i1, mode = op.i2;
        const int kv = 9 * n_z;
        constexpr int NW = SEG_THREADS / 32;
        for (int k = lane; k < K1; k += 32) {
          const bool kvalid = k < kv;
          const int t = kvalid ? k / n_z : 0;
          const int ch = kvalid ? op.idx[k - t * n_z] : 0;
          const int dy = t / 3 - 1, dx = t % 3 - 1;
          for (int pp = warp; pp < 64; pp += NW) {
            const int yy = (pp >> 3) + dy, xx = (pp & 7) + dx;
            float v = 0.f;
            if (kvalid && yy >= 0 && yy < 8 && xx >= 0 && xx < 8) v = sm.s[(yy * 8 + xx) * Cs + ch];
            const size_t di = ((size_t)b * 64 + pp) * K1 + k;
            if (mode == OUT_F32_NHWC) {
              ((float*)op.out0)[di] = v;
            } else if (mode == OUT_BF16_SPLIT) {
              __nv_bfloat16 hi, lo;
              split_bf16(v, hi, lo);
              ((__nv_bfloat16*)op.out0)[di] = hi;
              ((__nv_bfloat16*)op.out1)[di] = lo;
            } else {
              ((__nv_bfloat16*)op.out0)[di] = __float2bfloat16_rn(v);
            }
          }
        }
        __syncthreads();
        break;
      }
      default: break;
    }
  }
  __syncthreads();
  for (int p = warp; p < 64; p += SEG_THREADS / 32)
    for (int c = lane; c < C; c += 32) gs[p * C0 + c] = sm.s[p * Cs + c];
  if (FWD) {
    // block reduction of the log-det contributions (warp shuffles + one smem pass)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ld += __shfl_xor_sync(0xffffffffu, ld, o);
    if ((tid & 31) == 0) sm.red[tid >> 5] = ld;
    __syncthreads();
    if (tid < 32) {
      float v = tid < SEG_THREADS / 32 ? sm.red[tid] : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (tid == 0) logdet[b] += v;
    }
  }
}

void flow_segment_init() {
  IPK_CUDA(cudaFuncSetAttribute(flow_segment_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  IPK_CUDA(cudaFuncSetAttribute(flow_segment_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  IPK_CUDA(cudaFuncSetAttribute(flow_segment_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  IPK_CUDA(cudaFuncSetAttribute(flow_segment_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  IPK_CUDA(cudaFuncSetAttribute(flow_segment_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  IPK_CUDA(cudaFuncSetAttribute(flow_segment_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
}

void flow_segment_run(const SegmentLaunch& s, bool forward, float* state, int C0, float* logdet, int b0, int nb, cudaStream_t st) {
  const int B = nb;
  if (s.nops == 0 || B == 0) return;
  // tensor-core precisions: register-resident mma MCFs up to 32 channels, streamed mma MCFs up to MCF_MMA_MAXC, FFMA beyond
  const int mode = (s.mma && s.has_mcf) ? (s.C <= FAST_MAXC ? 1 : (s.C <= MCF_MMA_MAXC ? 2 : 0)) : 0;
  size_t smem = flow_segment_smem_bytes(s.C, s.has_mcf, mode);
  IPK_CHECK(smem <= 200 * 1024, IPK_ERR_UNSUPPORTED, "flow segment needs %zu bytes of shared memory (C=%d)", smem, s.C);
  const int hm = s.has_mcf ? 1 : 0;
  const MicroOp* ops = s.ops;
  const int nops = s.nops, C = s.C;
  auto go = [&](auto kernel) { launch_k(kernel, dim3(B), dim3(SEG_THREADS), smem, st, ops, nops, C, hm, state, C0, logdet, b0); };
  if (forward) {
    if (mode == 1) go(flow_segment_kernel<true, 1>);
    else if (mode == 2) go(flow_segment_kernel<true, 2>);
    else go(flow_segment_kernel<true, 0>);
  } else {
    if (mode == 1) go(flow_segment_kernel<false, 1>);
    else if (mode == 2) go(flow_segment_kernel<false, 2>);
    else go(flow_segment_kernel<false, 0>);
  }
}

// ---------------------------------------------------------------------------------------------- mma fragment packing
// A fragment of mma.m16n8k16 (row-major 16x16 bf16): lane (g = lane/4, t = lane%4) holds a0 = (row g, k 2t..2t+1),
// a1 = (row g+8, same k), a2 = (row g, k 2t+8..2t+9), a3 = (row g+8, k 2t+8..2t+9); low half = lower k.
__device__ __forceinline__ void pack_pair(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
  const __nv_bfloat16 l0 = __float2bfloat16_rn(v0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
  hi = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
  lo = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
}

// shifted-conv weights w[hid][C][kh][kw] -> fragments [m-tile][k-tile][lane][4] (hi plane, then lo plane);
// k-tile kt = (tap row r in {u-2, u-1}, dv in {-1,0,1}, channel tile ct), element kk -> channel ct*16 + kk
__global__ void pack_mcf_mma_conv_kernel(const float* __restrict__ w, uint32_t* __restrict__ dst, int hid, int C, int kh, int kw, int order) {
  const int nct = (C + 15) / 16, nkt = 6 * nct, nmt = (hid + 15) / 16;
  const int total = nmt * nkt * 32 * 4;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int j = e & 3, lane = (e >> 2) & 31, kt = (e >> 7) % nkt, mt = (e >> 7) / nkt;
    const int g = lane >> 2, t = lane & 3;
    const int n = mt * 16 + g + (j & 1) * 8;
    const int kk = t * 2 + (j >> 1) * 8;
    const int rr = kt / (3 * nct), rem = kt - rr * 3 * nct, dvi = rem / nct, ct = rem - dvi * nct;
    int ky, kx;
    switch (order) {
      case 0: ky = rr; kx = dvi; break;          // A (2x3)
      case 1: ky = 1 - rr; kx = dvi; break;      // B (2x3)
      case 2: kx = rr; ky = dvi; break;          // C (3x2)
      default: kx = 1 - rr; ky = dvi; break;     // D (3x2)
    }
    float v[2];
    for (int q = 0; q < 2; ++q) {
      const int c = ct * 16 + kk + q;
      v[q] = (n < hid && c < C) ? w[(((size_t)n * C + c) * kh + ky) * kw + kx] : 0.f;
    }
    uint32_t hi, lo;
    pack_pair(v[0], v[1], hi, lo);
    dst[e] = hi;
    dst[(size_t)total + e] = lo;
  }
}
// 1x1 weights v[2C][row] (weight-norm scale os[2C]) columns [0, hid) -> fragments [m-tile][k-tile][lane] uint2 = (a0, a2):
// m-tile mt rows 0-3 -> outputs 4*mt + r (mu), rows 4-7 -> outputs C + 4*mt + r - 4 (log-scale); rows 8-15 are zero and not stored
__global__ void pack_mcf_mma_1x1_kernel(const float* __restrict__ v, const float* __restrict__ os, uint32_t* __restrict__ dst, int hid, int C, int row) {
  const int nmt1 = (C + 3) / 4, nkt1 = (hid + 15) / 16;
  const int total = nmt1 * nkt1 * 32 * 2;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int j = e & 1, lane = (e >> 1) & 31, kt = (e >> 6) % nkt1, mt = (e >> 6) / nkt1;
    const int g = lane >> 2, t = lane & 3;
    const int c = mt * 4 + (g & 3);
    const int o = (g >= 4) ? C + c : c;
    const int kk = kt * 16 + t * 2 + j * 8;
    float x[2];
    for (int q = 0; q < 2; ++q) x[q] = (c < C && kk + q < hid) ? v[(size_t)o * row + kk + q] * os[o] : 0.f;
    uint32_t hi, lo;
    pack_pair(x[0], x[1], hi, lo);
    dst[e] = hi;
    dst[(size_t)total + e] = lo;
  }
}
size_t mcf_mma_conv_words(int C) { return (size_t)((4 * C + 15) / 16) * (6 * ((C + 15) / 16)) * 128 * 2; }
size_t mcf_mma_1x1_words(int C) { return (size_t)((C + 3) / 4) * ((4 * C + 15) / 16) * 64 * 2; }
void pack_mcf_mma(const float* shift_w, const float* v, const float* os, uint32_t* conv_dst, uint32_t* x1_dst, int hid, int C,
                  int kh, int kw, int order, int row, cudaStream_t st) {
  pack_mcf_mma_conv_kernel<<<std::max(1, std::min(64, (int)(mcf_mma_conv_words(C) / 2 + 255) / 256)), 256, 0, st>>>(shift_w, conv_dst, hid, C, kh, kw, order);
  IPK_LAUNCH_CHECK();
  pack_mcf_mma_1x1_kernel<<<std::max(1, std::min(64, (int)(mcf_mma_1x1_words(C) / 2 + 255) / 256)), 256, 0, st>>>(v, os, x1_dst, hid, C, row);
  IPK_LAUNCH_CHECK();
}

}  // namespace ipk
