// Final decoder convolution: 3x3, 64 -> 3 channels, + tanh, fp32 FFMA, frames written NCHW at the ABI edge
// (SpadeCondConvDecoder.out_conv, models/modules/autoencoders/fully_conv_models.py:163-164,176).
//
// N = 3 makes this layer bandwidth-shaped: a GEMM formulation re-reads the 64-channel input once per tap for almost no
// math.  Here a CTA stages an (8+2) x (32+2) pixel halo tile of the fp32 NHWC input in shared memory ONCE (padded rows:
// conflict-free LDS.128) and the 1 728 folded weights live in registers, 108 per lane (one channel quad each).
//
// Fused SPADE (Spade.forward, util.py:494-499): when `mr` / `spade` are given the input is the last up-block's raw output and the staged
// tile is normalised in place before the taps read it,  y = (x - mean[f,c]) * rstd[f,c] * (1 + gamma)[v,p,c] + beta[v,p,c]  (zero outside
// the image: the conv pads the NORMALISED tensor) -- the separate norm pass wrote and this kernel re-read 4.3 GB per 1 024 frames.
#include "elementwise.cuh"

namespace ipk {

constexpr int OC_CIN = 64, OC_TH = 8, OC_TW = 32, OC_PS = OC_CIN + 4;     // pixel stride in floats (pad 4: bank spread)

struct OutConvParams {
  float w[9 * OC_CIN * 3];     // [tap = ky*3+kx][c][o]
  float bias[3];
};

// 256 threads = 8 warps; warp w owns output row w of the 8 x 32 tile, each half-warp a 16-column strip; lane = (strip,
// channel quad c4).  A lane keeps the 108 weights of its channel quad in registers for the whole (persistent) kernel and a
// sliding 3x3 window of float4 activations; the 16 channel-quad partial sums of a pixel are combined with four xor-shuffles.
__global__ void __launch_bounds__(256, 1)
out_conv_kernel(const float* __restrict__ in, float* __restrict__ out, int F, int S, int tiles_x, int tiles_y,
                const float* __restrict__ wpk, const float* __restrict__ mr, const float* __restrict__ spade, int T) {
  extern __shared__ __align__(16) float halo[];      // 2 x [(OC_TH+2)*(OC_TW+2)][OC_PS], then out tile [3][OC_TH][OC_TW]
  constexpr int HP = (OC_TH + 2) * (OC_TW + 2), C4 = OC_CIN / 4;
  float* otile = halo + 2 * HP * OC_PS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c4 = lane & 15, strip = lane >> 4;
  pdl_wait();
  pdl_trigger();
  float wr[9][4][3];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap)
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int o = 0; o < 3; ++o) wr[tap][k][o] = __ldg(wpk + (tap * OC_CIN + c4 * 4 + k) * 3 + o);
  const float b0 = __ldg(wpk + 9 * OC_CIN * 3), b1 = __ldg(wpk + 9 * OC_CIN * 3 + 1), b2 = __ldg(wpk + 9 * OC_CIN * 3 + 2);

  const int tpf = tiles_x * tiles_y;
  const int ntiles = F * tpf;
  // halo tiles are double-buffered: tile i+1 is fetched with cp.async (zero-fill outside the image = conv padding) while
  // tile i is being computed
  auto stage = [&](int tile, float* dst) {
    const int f = tile / tpf, r = tile - f * tpf;
    const int y0 = (r / tiles_x) * OC_TH, x0 = (r % tiles_x) * OC_TW;
    const float* inf = in + (size_t)f * S * S * OC_CIN;
    for (int i = tid; i < HP * C4; i += 256) {
      const int px = i / C4, q = i - px * C4;
      const int hy = px / (OC_TW + 2), hx = px - hy * (OC_TW + 2);
      const int y = y0 + hy - 1, x = x0 + hx - 1;
      const bool ok = y >= 0 && y < S && x >= 0 && x < S;
      const float* src = ok ? inf + ((size_t)y * S + x) * OC_CIN + q * 4 : inf;
      const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + px * OC_PS + q * 4);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(ok ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int buf = 0;
  if (blockIdx.x < ntiles) stage(blockIdx.x, halo);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int f = tile / tpf, r = tile - f * tpf;
    const int y0 = (r / tiles_x) * OC_TH, x0 = (r % tiles_x) * OC_TW;
    const float* hb = halo + (size_t)buf * HP * OC_PS;
    const int nxt = tile + gridDim.x;
    if (nxt < ntiles) {
      stage(nxt, halo + (size_t)(buf ^ 1) * HP * OC_PS);     // that buffer was released by the barrier ending the previous tile
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (spade != nullptr) {
      // in-place SPADE normalisation of the staged tile: thread = (channel quad tid % 16, pixels tid / 16 + 16 k)
      float* hw = halo + (size_t)buf * HP * OC_PS;
      const int q = tid & 15;
      const float4 m0 = __ldg((const float4*)(mr + ((size_t)f * OC_CIN + q * 4) * 2)), m1 = __ldg((const float4*)(mr + ((size_t)f * OC_CIN + q * 4) * 2) + 1);
      const float* spv = spade + (size_t)(f / T) * S * S * (2 * OC_CIN) + q * 4;
      constexpr int NB = 6;                       // independent (gamma, beta) load pairs in flight per thread
      for (int px0 = tid >> 4; px0 < HP; px0 += 16 * NB) {
        float4 g[NB], bt[NB];
        bool ok[NB];
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          const int px = px0 + 16 * k;
          const int hy = px / (OC_TW + 2), hx = px - hy * (OC_TW + 2);
          const int y = y0 + hy - 1, x = x0 + hx - 1;
          ok[k] = px < HP && y >= 0 && y < S && x >= 0 && x < S;
          if (ok[k]) {
            const float* sp = spv + ((size_t)y * S + x) * (2 * OC_CIN);
            g[k] = __ldg((const float4*)sp);
            bt[k] = __ldg((const float4*)(sp + OC_CIN));
          }
        }
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          if (ok[k]) {
            float4* hp = (float4*)(hw + (px0 + 16 * k) * OC_PS + q * 4);
            float4 v = *hp;
            v.x = (v.x - m0.x) * m0.y * g[k].x + bt[k].x;
            v.y = (v.y - m0.z) * m0.w * g[k].y + bt[k].y;
            v.z = (v.z - m1.x) * m1.y * g[k].z + bt[k].z;
            v.w = (v.w - m1.z) * m1.w * g[k].w + bt[k].w;
            *hp = v;
          }
        }
      }
      __syncthreads();
    }
    // ---- this half-warp: output row `warp`, columns 16*strip .. 16*strip+15 (halo coords: rows warp..warp+2, cols +0..+2)
    const float* hrow = hb + (warp * (OC_TW + 2) + strip * 16) * OC_PS + c4 * 4;
    float4 win[3][3];
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) {
      win[rr][0] = *(const float4*)(hrow + (rr * (OC_TW + 2) + 0) * OC_PS);
      win[rr][1] = *(const float4*)(hrow + (rr * (OC_TW + 2) + 1) * OC_PS);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
#pragma unroll
      for (int rr = 0; rr < 3; ++rr) win[rr][2] = *(const float4*)(hrow + (rr * (OC_TW + 2) + j + 2) * OC_PS);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int rr = 0; rr < 3; ++rr)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          const float av[4] = {win[rr][cc].x, win[rr][cc].y, win[rr][cc].z, win[rr][cc].w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            a0 = fmaf(av[k], wr[rr * 3 + cc][k][0], a0);
            a1 = fmaf(av[k], wr[rr * 3 + cc][k][1], a1);
            a2 = fmaf(av[k], wr[rr * 3 + cc][k][2], a2);
          }
        }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) {        // combine the 16 channel quads of this pixel (stays inside the half-warp)
        a0 += __shfl_xor_sync(0xffffffffu, a0, off);
        a1 += __shfl_xor_sync(0xffffffffu, a1, off);
        a2 += __shfl_xor_sync(0xffffffffu, a2, off);
      }
      if (c4 == 0) {
        const int col = strip * 16 + j;
        otile[(0 * OC_TH + warp) * OC_TW + col] = a0 + b0;      // tanh is applied by the coalesced store pass
        otile[(1 * OC_TH + warp) * OC_TW + col] = a1 + b1;
        otile[(2 * OC_TH + warp) * OC_TW + col] = a2 + b2;
      }
#pragma unroll
      for (int rr = 0; rr < 3; ++rr) { win[rr][0] = win[rr][1]; win[rr][1] = win[rr][2]; }
    }
    __syncthreads();
    // ---- coalesced NCHW store of the 3 x 8 x 32 tile
    for (int i = tid; i < 3 * OC_TH * OC_TW; i += 256) {
      const int col = i % OC_TW, row = (i / OC_TW) % OC_TH, o = i / (OC_TW * OC_TH);
      const int y = y0 + row, x = x0 + col;
      if (y < S && x < S) out[((size_t)f * 3 + o) * S * S + (size_t)y * S + x] = tanhf(otile[i]);
    }
    __syncthreads();     // tile buffer and output tile free for reuse
    buf ^= 1;
  }
}

// the plan only remembers the packed device weights ([9][64][3] followed by bias[3])
struct OutConvPlan { const float* wpk; };

OutConvPlan* out_conv_plan_create(const float* w_dev_packed, cudaStream_t st) {
  (void)st;
  OutConvPlan* p = new OutConvPlan();
  p->wpk = w_dev_packed;
  static bool attr_set[IPK_MAX_DEVICES] = {false};
  const int slot = current_device_slot();
  if (!attr_set[slot]) {
    IPK_CUDA(cudaFuncSetAttribute(out_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set[slot] = true;
  }
  return p;
}
void out_conv_plan_destroy(OutConvPlan* p) { delete p; }

void out_conv_run(const OutConvPlan* p, const float* in_nhwc, float* out_nchw, int F, int S, cudaStream_t st, const float* mr, const float* spade, int T) {
  const int tiles_x = cdiv(S, OC_TW), tiles_y = cdiv(S, OC_TH);
  const long long tiles = (long long)F * tiles_x * tiles_y;
  if (tiles == 0) return;
  const size_t smem = ((size_t)2 * (OC_TH + 2) * (OC_TW + 2) * OC_PS + 3 * OC_TH * OC_TW) * sizeof(float);
  const int grid = (int)std::min<long long>(tiles, 148LL);
  launch_k(out_conv_kernel, dim3(grid), dim3(256), smem, st, in_nhwc, out_nchw, F, S, tiles_x, tiles_y, p->wpk, mr, spade, std::max(1, T));
}

// packs OIHW [3][64][3][3] (optionally divided by the spectral-norm sigma) + bias into the layout above
__global__ void out_conv_pack_kernel(const float* __restrict__ w, const float* __restrict__ sigma, const float* __restrict__ bias, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 9 * OC_CIN * 3) {
    const int o = i % 3, c = (i / 3) % OC_CIN, tap = i / (3 * OC_CIN);
    float v = w[((size_t)o * OC_CIN + c) * 9 + tap];
    if (sigma) v = v / sigma[0];
    dst[i] = v;
  } else if (i < 9 * OC_CIN * 3 + 3) {
    dst[i] = bias[i - 9 * OC_CIN * 3];
  }
}
void out_conv_pack(const float* w_oihw, const float* sigma, const float* bias, float* dst, cudaStream_t st) {
  out_conv_pack_kernel<<<cdiv(9 * OC_CIN * 3 + 3, 256), 256, 0, st>>>(w_oihw, sigma, bias, dst);
  IPK_LAUNCH_CHECK();
}

}  // namespace ipk
