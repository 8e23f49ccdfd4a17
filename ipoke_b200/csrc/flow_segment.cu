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
//   * C > 32 (generic): weights streamed from L2 every line.
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
};

__host__ __device__ inline int seg_hidmax(int C) { return C <= 96 ? 4 * C : (2 * C < 512 ? 2 * C : 512); }

struct SegOffsets { size_t tmp, hterm, pc, act, p1, red, total; };

__host__ __device__ inline SegOffsets seg_layout(int C, bool has_mcf) {
  const int Cs = (C + 3) / 4 * 4;
  const int C2s = (2 * C + 3) / 4 * 4;
  const bool fast = C <= FAST_MAXC;
  SegOffsets o;
  size_t off = (size_t)64 * Cs;                       // s
  o.tmp = off; off += (size_t)64 * Cs;
  o.hterm = off; off += has_mcf ? (size_t)(fast ? 2 : 1) * 64 * C2s : 0;
  o.pc = off; off += (has_mcf && fast) ? (size_t)2 * 8 * 128 : 0;
  o.act = off; off += has_mcf ? (size_t)8 * ((seg_hidmax(C) + 3) / 4 * 4) : 0;
  o.p1 = off; off += has_mcf ? (fast ? (size_t)4 * 8 * 64 : (size_t)8 * C2s) : 0;
  o.red = off; off += 32;
  o.total = off;
  return o;
}

size_t flow_segment_smem_bytes(int C, int h_ch, bool has_mcf) {
  (void)h_ch;
  return seg_layout(C, has_mcf).total * sizeof(float);
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

__device__ __forceinline__ int next_mcf(const MicroOp* __restrict__ ops, int from, int nops) {
  for (int i = from; i < nops; ++i)
    if (ops[i].kind == MK_MCF) return i;
  return -1;
}

template <bool FWD>
__global__ void __launch_bounds__(SEG_THREADS, 1) flow_segment_kernel(const MicroOp* __restrict__ ops, int nops, int C, int has_mcf,
                                                                       float* __restrict__ state, int C0,
                                                                       float* __restrict__ logdet) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const int Cs = (C + 3) / 4 * 4;
  const int C2s = (2 * C + 3) / 4 * 4;
  const bool fast = C <= FAST_MAXC;
  SegSmem sm;
  const SegOffsets lo = seg_layout(C, has_mcf != 0);
  sm.s = smem; sm.tmp = smem + lo.tmp; sm.hterm = smem + lo.hterm; sm.pc = smem + lo.pc; sm.act = smem + lo.act;
  sm.p1 = smem + lo.p1; sm.red = smem + lo.red;

  // first MCF of the segment: start fetching its weights / conditioning term before anything else
  McfRegs regs;
  int hbuf = 0;
  MicroOp nxt;
  int nxt_i = (has_mcf && fast) ? next_mcf(ops, 0, nops) : -1;
  if (nxt_i >= 0) {
    nxt = ops[nxt_i];
    mcf_prefetch_hterm(nxt, b, sm.hterm);
    cp_async_commit();
    mcf_load_wc(nxt, regs);
    mcf_load_w1(nxt, regs);
  }

  float* gs = state + (size_t)b * 64 * C0;
  for (int i = tid; i < 64 * Cs; i += SEG_THREADS) {
    int p = i / Cs, c = i % Cs;
    sm.s[i] = c < C ? gs[p * C0 + c] : 0.f;
  }
  __syncthreads();

  float ld = 0.f;
  for (int oi = 0; oi < nops; ++oi) {
    const MicroOp op = ops[oi];
    switch (op.kind) {
      case MK_ACTNORM: {
        const int coff = op.i0, cnt = op.i1;
        for (int i = tid; i < 64 * cnt; i += SEG_THREADS) {
          int p = i / cnt, c = i % cnt;
          float ls = __ldg(op.p0 + c), bb = __ldg(op.p1 + c);
          float x = sm.s[p * Cs + coff + c];
          if (FWD) {
            sm.s[p * Cs + coff + c] = x * expf(ls) + bb;
            ld += ls;   // H*W*sum(log_scale): one contribution per (pixel, channel)
          } else {
            sm.s[p * Cs + coff + c] = (x - bb) / (expf(ls) + 1e-8f);
          }
        }
        __syncthreads();
        break;
      }
      case MK_SHUFFLE: {
        const int Cn = op.i0;
        for (int i = tid; i < 64 * Cn; i += SEG_THREADS) {
          int p = i / Cn, c = i % Cn;
          sm.tmp[p * Cs + c] = sm.s[p * Cs + op.idx[c]];
        }
        __syncthreads();
        for (int i = tid; i < 64 * Cn; i += SEG_THREADS) {
          int p = i / Cn, c = i % Cn;
          sm.s[p * Cs + c] = sm.tmp[p * Cs + c];
        }
        __syncthreads();
        break;
      }
      case MK_MCF: {
        if (fast) {
          // registers hold this MCF's weights; its conditioning term is the (possibly still pending) cp.async group of
          // buffer hbuf.  Queue the NEXT MCF's conditioning term into the other buffer, then wait for ours.
          const int nn = next_mcf(ops, oi + 1, nops);
          if (nn >= 0) {
            nxt = ops[nn];
            mcf_prefetch_hterm(nxt, b, sm.hterm + (hbuf ^ 1) * 64 * C2s);
            cp_async_commit();
            cp_async_wait1();
          } else {
            cp_async_wait0();
          }
          __syncthreads();
          mcf_fast<FWD>(op, nn >= 0 ? &nxt : nullptr, regs, sm, sm.hterm + hbuf * 64 * C2s, Cs, ld);
          hbuf ^= 1;
        } else {
          mcf_generic<FWD>(op, sm, Cs, b, ld);
        }
        break;
      }
      case MK_AFFINE: {
        // finishes a NICE coupling: params[p] = bias + sum over taps t and split-K slices of T[p + delta_t][t]  (the 3x3
        // gather of conv3 with zero padding), then Affine.fwd / bwd on the transformed channels
        const int nsplit = op.i0, Npad = op.i1, n_p = op.i2, N3p = op.i3;
        for (int i = tid; i < 64 * n_p; i += SEG_THREADS) {
          const int p = i / n_p, j = i - p * n_p;
          const int y = p >> 3, x = p & 7;
          float mu = __ldg(op.p1 + j), ls = __ldg(op.p1 + n_p + j);
          for (int s = 0; s < nsplit; ++s) {
            const float* T = op.p0 + (size_t)s * op.l0 + (size_t)b * 64 * Npad;
            float pm[9], pl[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) {           // nine independent loads in flight
              const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
              const bool ok = yy >= 0 && yy < 8 && xx >= 0 && xx < 8;
              const float* r = T + (size_t)(ok ? yy * 8 + xx : 0) * Npad + t * N3p;
              pm[t] = ok ? r[j] : 0.f;
              pl[t] = ok ? r[n_p + j] : 0.f;
            }
#pragma unroll
            for (int t = 0; t < 9; ++t) { mu += pm[t]; ls += pl[t]; }
          }
          float sc = 1.0f + tanhf(0.5f * ls);
          int c = op.idx[j];
          float xv = sm.s[p * Cs + c];
          if (FWD) {
            sm.s[p * Cs + c] = sc * xv + mu;
            ld += logf(sc);
          } else {
            sm.s[p * Cs + c] = (xv - mu) / (sc + 1e-12f);
          }
        }
        __syncthreads();
        break;
      }
      case MK_IM2COL: {
        const int n_z = op.i0, K1 = op.i1, mode = op.i2;
        const int kv = 9 * n_z;
        for (int i = tid; i < 64 * K1; i += SEG_THREADS) {
          int p = i / K1, k = i % K1;
          float v = 0.f;
          if (k < kv) {
            int t = k / n_z, j = k % n_z;
            int yy = (p >> 3) + t / 3 - 1, xx = (p & 7) + t % 3 - 1;
            if (yy >= 0 && yy < 8 && xx >= 0 && xx < 8) v = sm.s[(yy * 8 + xx) * Cs + op.idx[j]];
          }
          size_t di = ((size_t)b * 64 + p) * K1 + k;
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
        __syncthreads();
        break;
      }
      default: break;
    }
  }
  __syncthreads();
  for (int i = tid; i < 64 * C; i += SEG_THREADS) {
    int p = i / C, c = i % C;
    gs[p * C0 + c] = sm.s[p * Cs + c];
  }
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
  IPK_CUDA(cudaFuncSetAttribute(flow_segment_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  IPK_CUDA(cudaFuncSetAttribute(flow_segment_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
}

void flow_segment_run(const SegmentLaunch& s, bool forward, float* state, int C0, float* logdet, int B, cudaStream_t st) {
  if (s.nops == 0 || B == 0) return;
  size_t smem = flow_segment_smem_bytes(s.C, 0, s.has_mcf);
  IPK_CHECK(smem <= 200 * 1024, IPK_ERR_UNSUPPORTED, "flow segment needs %zu bytes of shared memory (C=%d)", smem, s.C);
  if (forward)
    flow_segment_kernel<true><<<B, SEG_THREADS, smem, st>>>(s.ops, s.nops, s.C, s.has_mcf ? 1 : 0, state, C0, logdet);
  else
    flow_segment_kernel<false><<<B, SEG_THREADS, smem, st>>>(s.ops, s.nops, s.C, s.has_mcf ? 1 : 0, state, C0, logdet);
  IPK_LAUNCH_CHECK();
}

}  // namespace ipk
