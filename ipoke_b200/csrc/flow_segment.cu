// Flow segment kernel: one CTA per sample, flow state resident in shared memory.
// Reference semantics (paths relative to the CompVis/ipoke checkout):
//   ActNorm2dFlow.forward            models/modules/INN/macow2.py:490-520
//   Shuffle.forward                  models/modules/INN/flow_blocks.py:314-326
//   MaskedConvFlow.forward/backward  models/modules/INN/macow2.py:97-151,174-288
//   MCFBlock / ShiftedConv2d         models/modules/INN/macow_utils.py:407-434,446-499
//   Affine.calc_params/fwd/bwd       models/modules/INN/macow_utils.py:49-66
//   NICE2d split/unsplit             models/modules/INN/macow2.py:364-388 (done as in-place channel index lists)
#include "flow_segment.cuh"

namespace ipk {

constexpr int SEG_THREADS = 256;

struct SegSmem {
  float* s;      // [64][Cs] flow state
  float* tmp;    // [64][Cs] scratch (shuffle / forward snapshot)
  float* P;      // [8][2C]  affine params of the current line
  float* act;    // [8][hid] ELU(hidden) of the current line
  float* hterm;  // [64][2C] conditioning contribution + bias of the current MCF
  float* e;      // [64][h_ch] ELU(cond)
  float* red;    // [32]
};

__host__ __device__ inline int seg_hidmax(int C) { return C <= 96 ? 4 * C : (2 * C < 512 ? 2 * C : 512); }

__host__ __device__ inline size_t seg_layout(int C, int h_ch, bool has_mcf, size_t* o_tmp, size_t* o_P, size_t* o_act,
                                             size_t* o_hterm, size_t* o_e, size_t* o_red) {
  int Cs = (C + 3) / 4 * 4;
  size_t off = 0;
  off += (size_t)64 * Cs;            // s
  *o_tmp = off; off += (size_t)64 * Cs;
  *o_P = off; off += (size_t)8 * ((2 * C + 3) / 4 * 4);
  *o_act = off; off += has_mcf ? (size_t)8 * ((seg_hidmax(C) + 3) / 4 * 4) : 0;
  *o_hterm = off; off += has_mcf ? (size_t)64 * ((2 * C + 3) / 4 * 4) : 0;
  *o_e = off; off += has_mcf ? (size_t)64 * h_ch : 0;
  *o_red = off; off += 32;
  return off;
}

size_t flow_segment_smem_bytes(int C, int h_ch, bool has_mcf) {
  size_t a, b, c, d, e, f;
  return seg_layout(C, h_ch, has_mcf, &a, &b, &c, &d, &e, &f) * sizeof(float);
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

template <bool FWD>
__device__ void mcf_op(const MicroOp& op, const SegSmem& sm, int Cs, int h_ch, float& ld) {
  const int tid = threadIdx.x;
  const int order = op.i0, C = op.i1, Cp = op.i2, hid = op.i3;
  const int Cp4 = Cp / 4, Cs4 = Cs / 4;
  const int C2 = 2 * C;
  const int C2s = (C2 + 3) / 4 * 4;            // row stride of P / hterm
  const float4* Wc = (const float4*)op.p0;     // [6][Cp4][hid]
  const float4* W1x = (const float4*)op.p1;    // [hid/4][2C]
  const float4* W1h = (const float4*)op.p2;    // [h_ch/4][2C]
  const float* bias = op.p3;
  const float4* s4 = (const float4*)(FWD ? sm.tmp : sm.s);

  if (FWD) {  // snapshot x: all lines are computed from the un-transformed input (macow2.py:113-116)
    for (int i = tid; i < 64 * Cs; i += SEG_THREADS) sm.tmp[i] = sm.s[i];
  }

  // ---- conditioning term: hterm[p][o] = b[o] + sum_k W1h[k][o] * ELU(cond[p][k])  (concat + ELU + 1x1, macow_utils.py:429-432)
  {
    int oslots = min(SEG_THREADS, (C2 + 31) / 32 * 32);
    int G = SEG_THREADS / oslots;              // pixel groups
    G = G >= 8 ? 8 : (G >= 4 ? 4 : (G >= 2 ? 2 : 1));
    int ppg = 64 / G;
    int g = tid / oslots, ol = tid % oslots;
    const float4* e4 = (const float4*)sm.e;
    const int h4 = h_ch / 4;
    if (g < G) {
      for (int o = ol; o < C2; o += oslots) {
        float b = bias[o];
        for (int p0 = 0; p0 < ppg; p0 += 8) {
          float acc[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = 0.f;
          for (int k4 = 0; k4 < h4; ++k4) {
            float4 w = __ldg(W1h + (size_t)k4 * C2 + o);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = dot4(w, e4[(g * ppg + p0 + i) * h4 + k4], acc[i]);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) sm.hterm[(g * ppg + p0 + i) * C2s + o] = acc[i] + b;
        }
      }
    }
  }
  __syncthreads();

  // thread mapping for the masked conv (hid outputs x 8 positions) and the 1x1 (2C outputs x 8 positions)
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
    // ---- masked (shifted) conv on the two previous lines: taps du in {-2,-1}, dv in {-1,0,1}
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
    // ---- weight-normed 1x1 on ELU(cat[hidden, cond]) -> (mu, log_scale)
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
            sm.P[q * C2s + o] = acc[qi] + sm.hterm[mcf_pix(order, u, q) * C2s + o];
          }
      }
    }
    __syncthreads();
    // ---- affine transform of line u (Affine.fwd / Affine.bwd)
    for (int i = tid; i < 8 * C; i += SEG_THREADS) {
      int q = i / C, c = i % C;
      float mu = sm.P[q * C2s + c];
      float sc = 1.0f + tanhf(0.5f * sm.P[q * C2s + C + c]);
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

template <bool FWD>
__global__ void __launch_bounds__(SEG_THREADS) flow_segment_kernel(const MicroOp* __restrict__ ops, int nops, int C, int has_mcf,
                                                                    float* __restrict__ state, int C0,
                                                                    const float* __restrict__ cond, int h_ch,
                                                                    float* __restrict__ logdet) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const int Cs = (C + 3) / 4 * 4;
  SegSmem sm;
  size_t o_tmp, o_P, o_act, o_hterm, o_e, o_red;
  seg_layout(C, h_ch, has_mcf != 0, &o_tmp, &o_P, &o_act, &o_hterm, &o_e, &o_red);
  sm.s = smem; sm.tmp = smem + o_tmp; sm.P = smem + o_P; sm.act = smem + o_act; sm.hterm = smem + o_hterm;
  sm.e = smem + o_e; sm.red = smem + o_red;

  float* gs = state + (size_t)b * 64 * C0;
  for (int i = tid; i < 64 * Cs; i += SEG_THREADS) {
    int p = i / Cs, c = i % Cs;
    sm.s[i] = c < C ? gs[p * C0 + c] : 0.f;
  }
  if (has_mcf) {
    const float* gc = cond + (size_t)b * 64 * h_ch;
    for (int i = tid; i < 64 * h_ch; i += SEG_THREADS) {
      float v = gc[i];
      sm.e[i] = v > 0.f ? v : expm1f(v);
    }
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
      case MK_MCF:
        mcf_op<FWD>(op, sm, Cs, h_ch, ld);
        break;
      case MK_AFFINE: {
        const int nsplit = op.i0, Npad = op.i1, n_p = op.i2;
        for (int i = tid; i < 64 * n_p; i += SEG_THREADS) {
          int p = i / n_p, j = i % n_p;
          size_t row = ((size_t)b * 64 + p) * Npad;
          float mu = __ldg(op.p1 + j), ls = __ldg(op.p1 + n_p + j);
          for (int s = 0; s < nsplit; ++s) {
            mu += op.p0[(size_t)s * op.l0 + row + j];
            ls += op.p0[(size_t)s * op.l0 + row + n_p + j];
          }
          float sc = 1.0f + tanhf(0.5f * ls);
          int c = op.idx[j];
          float x = sm.s[p * Cs + c];
          if (FWD) {
            sm.s[p * Cs + c] = sc * x + mu;
            ld += logf(sc);
          } else {
            sm.s[p * Cs + c] = (x - mu) / (sc + 1e-12f);
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

void flow_segment_run(const SegmentLaunch& s, bool forward, float* state, int C0, const float* cond, int h_ch,
                      float* logdet, int B, cudaStream_t st) {
  if (s.nops == 0 || B == 0) return;
  size_t smem = flow_segment_smem_bytes(s.C, h_ch, s.has_mcf);
  IPK_CHECK(smem <= 200 * 1024, IPK_ERR_UNSUPPORTED, "flow segment needs %zu bytes of shared memory (C=%d)", smem, s.C);
  IPK_CHECK(h_ch % 4 == 0, IPK_ERR_UNSUPPORTED, "h_channels must be a multiple of 4");
  if (forward)
    flow_segment_kernel<true><<<B, SEG_THREADS, smem, st>>>(s.ops, s.nops, s.C, s.has_mcf ? 1 : 0, state, C0, cond, h_ch, logdet);
  else
    flow_segment_kernel<false><<<B, SEG_THREADS, smem, st>>>(s.ops, s.nops, s.C, s.has_mcf ? 1 : 0, state, C0, cond, h_ch, logdet);
  IPK_LAUNCH_CHECK();
}

}  // namespace ipk
