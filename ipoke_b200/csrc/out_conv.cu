// Final decoder convolution: 3x3, 64 -> 3 channels, + tanh, fp32 FFMA, frames written NCHW at the ABI edge
// (SpadeCondConvDecoder.out_conv, models/modules/autoencoders/fully_conv_models.py:163-164,176), with the last SPADE block
// (Spade.forward, util.py:494-499) fused in front of it.
//
// N = 3 makes this layer bandwidth-shaped: a GEMM formulation re-reads the 64-channel input once per tap for almost no math.  Here a
// CTA works on 8 x 16 output pixels: the (8+2) x (16+2) halo tile of the fp32 NHWC input arrives as ONE TMA box (out-of-bounds pixels
// zero-filled = the conv's padding; no per-thread address arithmetic), double-buffered so that tile i+1 travels while tile i is
// computed, and the 1 728 folded weights live in registers, 108 per lane (one channel quad each).
//
// Fused SPADE: when `mr` / `spade` are given the input is the last up-block's RAW output; the matching box of the (1 + gamma | beta)
// maps lands beside the tile and the tile is normalised in place before the taps read it,
//     y = (x - mean[f,c]) * rstd[f,c] * (1 + gamma)[v,p,c] + beta[v,p,c]
// (zero outside the image, because the maps' box is zero-filled there too: the conv pads the NORMALISED tensor).  The separate pass
// wrote, and this kernel re-read, a 4.3 GB normalised copy per 1 024 frames.
#include <cuda.h>
#include "elementwise.cuh"
#include "tc_ptx.cuh"

namespace ipk {

CUtensorMap tc_make_map_f32(const float* base, int rank, const long long* dims, const long long* strides_bytes, const int* box);   // conv_tc.cu

constexpr int OC_CIN = 64, OC_TH = 8, OC_TW = 16, OC_HW = OC_TW + 2, OC_HP = (OC_TH + 2) * OC_HW;     // 180 halo pixels
constexpr int OC_X_BYTES = OC_HP * OC_CIN * 4, OC_SP_BYTES = OC_HP * 2 * OC_CIN * 4;
constexpr size_t OC_SMEM = 2 * OC_X_BYTES + OC_SP_BYTES + 3 * OC_TH * OC_TW * 4 + 64;

// 256 threads = 8 warps; warp w owns output row w of the tile, each half-warp an 8-column strip; lane = (strip, channel quad c4).  A
// lane keeps the 108 weights of its channel quad in registers for the whole (persistent) kernel and a sliding 3x3 window of float4
// activations.  The 16 channel-quad partial sums of the strip's 8 pixels x 3 outputs are combined by a halving exchange (24 shuffles per
// lane instead of 96): after the xor-8 / 4 / 2 steps a lane holds the three outputs of pixel c4 >> 1.
__global__ void __launch_bounds__(256, 1)
out_conv_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmSP, float* __restrict__ out, int F, int S,
                int tiles_x, int tiles_y, const float* __restrict__ wpk, const float* __restrict__ mr, int has_spade, int T) {
  extern __shared__ __align__(1024) uint8_t oc_smem[];
  float* xb = (float*)oc_smem;                                   // [2][OC_HP][64]
  float* spb = (float*)(oc_smem + 2 * OC_X_BYTES);               // [OC_HP][128]
  float* otile = (float*)(oc_smem + 2 * OC_X_BYTES + OC_SP_BYTES);
  uint64_t* full_x = (uint64_t*)(otile + 3 * OC_TH * OC_TW);     // [2]
  uint64_t* full_sp = full_x + 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c4 = lane & 15, strip = lane >> 4;
  if (tid == 0) {
    prefetch_tensormap(&tmX);
    if (has_spade) prefetch_tensormap(&tmSP);
    mbar_init(&full_x[0], 1); mbar_init(&full_x[1], 1); mbar_init(full_sp, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
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
  auto issue = [&](int tile, int b) {      // thread 0: the boxes of `tile` -> x buffer b (+ the SPADE maps' buffer)
    const int f = tile / tpf, r = tile - f * tpf;
    const int y0 = (r / tiles_x) * OC_TH, x0 = (r % tiles_x) * OC_TW;
    mbar_expect_tx(&full_x[b], OC_X_BYTES);
    tma_load_4d(xb + (size_t)b * OC_HP * OC_CIN, &tmX, &full_x[b], 0, x0 - 1, y0 - 1, f);
    if (has_spade) {
      mbar_expect_tx(full_sp, OC_SP_BYTES);
      tma_load_4d(spb, &tmSP, full_sp, 0, x0 - 1, y0 - 1, f / T);
    }
  };
  if (tid == 0 && (int)blockIdx.x < ntiles) issue(blockIdx.x, 0);
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    const int f = tile / tpf, r = tile - f * tpf;
    const int y0 = (r / tiles_x) * OC_TH, x0 = (r % tiles_x) * OC_TW;
    float* hb = xb + (size_t)buf * OC_HP * OC_CIN;
    float4 m0 = make_float4(0.f, 1.f, 0.f, 1.f), m1 = m0;
    if (has_spade) {
      const float4* mp = (const float4*)(mr + ((size_t)f * OC_CIN + (tid & 15) * 4) * 2);
      m0 = __ldg(mp); m1 = __ldg(mp + 1);
    }
    mbar_wait(&full_x[buf], (uint32_t)((it >> 1) & 1));
    if (has_spade) {
      mbar_wait(full_sp, (uint32_t)(it & 1));
      // in-place normalisation: thread = (channel quad tid % 16, pixels tid / 16 + 16 k)
      const int q = tid & 15;
#pragma unroll 4
      for (int px = tid >> 4; px < OC_HP; px += 16) {
        float4* hp = (float4*)(hb + px * OC_CIN + q * 4);
        const float4 g = *(const float4*)(spb + px * (2 * OC_CIN) + q * 4), bt = *(const float4*)(spb + px * (2 * OC_CIN) + OC_CIN + q * 4);
        float4 v = *hp;
        v.x = (v.x - m0.x) * m0.y * g.x + bt.x;
        v.y = (v.y - m0.z) * m0.w * g.y + bt.y;
        v.z = (v.z - m1.x) * m1.y * g.z + bt.z;
        v.w = (v.w - m1.z) * m1.w * g.w + bt.w;
        *hp = v;
      }
      fence_proxy_async_shared();      // the maps' buffer is about to be overwritten by the next tile's box
    }
    __syncthreads();
    // next tile's boxes: the other x buffer was released by the barrier ending the previous tile, the maps' buffer just now
    if (tid == 0 && tile + (int)gridDim.x < ntiles) issue(tile + gridDim.x, buf ^ 1);

    // ---- this half-warp: output row `warp`, columns 8*strip .. 8*strip+7 (halo coords: rows warp..warp+2, cols +0..+2)
    const float* hrow = hb + (warp * OC_HW + strip * 8) * OC_CIN + c4 * 4;
    float4 win[3][3];
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) {
      win[rr][0] = *(const float4*)(hrow + (rr * OC_HW + 0) * OC_CIN);
      win[rr][1] = *(const float4*)(hrow + (rr * OC_HW + 1) * OC_CIN);
    }
    float acc[24];       // [pixel j][output o]
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int rr = 0; rr < 3; ++rr) win[rr][2] = *(const float4*)(hrow + (rr * OC_HW + j + 2) * OC_CIN);
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
      acc[3 * j] = a0; acc[3 * j + 1] = a1; acc[3 * j + 2] = a2;
#pragma unroll
      for (int rr = 0; rr < 3; ++rr) { win[rr][0] = win[rr][1]; win[rr][1] = win[rr][2]; }
    }
    // halving exchange over the 16 channel quads: keep the half of the values selected by this lane's bit, add the partner's
    {
      const bool h8 = (c4 & 8) != 0, h4 = (c4 & 4) != 0, h2 = (c4 & 2) != 0;
      float s12[12], s6[6], s3[3];
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        const float keep = h8 ? acc[12 + i] : acc[i], send = h8 ? acc[i] : acc[12 + i];
        s12[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const float keep = h4 ? s12[6 + i] : s12[i], send = h4 ? s12[i] : s12[6 + i];
        s6[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float keep = h2 ? s6[3 + i] : s6[i], send = h2 ? s6[i] : s6[3 + i];
        s3[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) s3[i] += __shfl_xor_sync(0xffffffffu, s3[i], 1);
      if ((c4 & 1) == 0) {
        const int col = strip * 8 + (c4 >> 1);
        otile[(0 * OC_TH + warp) * OC_TW + col] = s3[0] + b0;      // tanh is applied by the coalesced store pass
        otile[(1 * OC_TH + warp) * OC_TW + col] = s3[1] + b1;
        otile[(2 * OC_TH + warp) * OC_TW + col] = s3[2] + b2;
      }
    }
    __syncthreads();
    // ---- coalesced NCHW store of the 3 x 8 x 16 tile
    for (int i = tid; i < 3 * OC_TH * OC_TW; i += 256) {
      const int col = i % OC_TW, row = (i / OC_TW) % OC_TH, o = i / (OC_TW * OC_TH);
      const int y = y0 + row, x = x0 + col;
      if (y < S && x < S) out[((size_t)f * 3 + o) * S * S + (size_t)y * S + x] = tanhf(otile[i]);
    }
    __syncthreads();     // tile buffer and output tile free for reuse
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
    IPK_CUDA(cudaFuncSetAttribute(out_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OC_SMEM));
    attr_set[slot] = true;
  }
  return p;
}
void out_conv_plan_destroy(OutConvPlan* p) { delete p; }

void out_conv_run(const OutConvPlan* p, const float* in_nhwc, float* out_nchw, int F, int S, cudaStream_t st, const float* mr, const float* spade, int T) {
  const int tiles_x = cdiv(S, OC_TW), tiles_y = cdiv(S, OC_TH);
  const long long tiles = (long long)F * tiles_x * tiles_y;
  if (tiles == 0) return;
  IPK_CHECK((mr == nullptr) == (spade == nullptr), IPK_ERR_INVALID, "out_conv_run: the fused SPADE needs both the statistics and the maps");
  T = std::max(1, T);
  long long xd[4] = {OC_CIN, S, S, F};
  long long xs[3] = {OC_CIN * 4LL, (long long)S * OC_CIN * 4, (long long)S * S * OC_CIN * 4};
  int xbox[4] = {OC_CIN, OC_HW, OC_TH + 2, 1};
  const CUtensorMap mX = tc_make_map_f32(in_nhwc, 4, xd, xs, xbox);
  CUtensorMap mSP = mX;
  if (spade) {
    long long sd[4] = {2 * OC_CIN, S, S, (F + T - 1) / T};
    long long ss[3] = {2 * OC_CIN * 4LL, (long long)S * 2 * OC_CIN * 4, (long long)S * S * 2 * OC_CIN * 4};
    int sbox[4] = {2 * OC_CIN, OC_HW, OC_TH + 2, 1};
    mSP = tc_make_map_f32(spade, 4, sd, ss, sbox);
  }
  const int grid = (int)std::min<long long>(tiles, 148LL);
  launch_k(out_conv_kernel, dim3(grid), dim3(256), OC_SMEM, st, mX, mSP, out_nchw, F, S, tiles_x, tiles_y, p->wpk, mr, spade ? 1 : 0, T);
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
