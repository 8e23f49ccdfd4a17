// "Segment" executor of the MaCow flow: everything between two NICE coupling networks runs as ONE kernel with the
// per-sample flow state (8x8xC fp32) resident in shared memory -- ActNorm, channel shuffle, the four masked-conv
// flows of each MaCowUnit (row/column-sequential inverse), the affine update that finishes the previous coupling
// and the im2col/operand split that feeds the next one.
#pragma once
#include "common.cuh"

namespace ipk {

enum MicroKind : int { MK_ACTNORM = 0, MK_SHUFFLE = 1, MK_MCF = 2, MK_AFFINE = 3, MK_IM2COL = 4 };

struct MicroOp {
  int kind;
  int i0, i1, i2, i3;
  const float* p0;
  const float* p1;
  const float* p2;
  const float* p3;
  const int* idx;
  void* out0;
  void* out1;
  long long l0;
};
/* field use
   MK_ACTNORM : i0 = channel offset, i1 = channel count, p0 = log_scale, p1 = bias
   MK_SHUFFLE : i0 = C, idx = gather index (new[c] = old[idx[c]])
   MK_MCF     : i0 = order (0 A,1 B,2 C,3 D), i1 = C, i2 = Cp (C rounded up to 4), i3 = hid,
                p0 = Wc [6][Cp/4][hid][4], p1 = W1x [hid/4][2C][4]   (fp32 paths)
                     or the mma.sync A-fragment arrays of pack_mcf_mma (tensor-core precisions, C <= 32),
                p2 = conditioning term of this MCF: column block of the precomputed matrix
                     Hterm[M = B*64][l0] = b + W1h * ELU(cond)  (round_up(2C, 4) columns, 16-byte aligned), l0 = row stride
   MK_AFFINE  : i0 = nsplit, i1 = Npad (row length of the tap-response matrix), i2 = n_p, i3 = N3p (columns per tap),
                p0 = split-K partial slices T[nsplit][M][Npad] with T[q][t*N3p + j] = response of tap t at pixel q,
                p1 = bias [2*n_p], idx = state channel of each transformed element, l0 = slice stride (elements)
   MK_IM2COL  : i0 = n_z, i1 = K1pad, i2 = out mode (OUT_F32_NHWC / OUT_BF16_SPLIT / OUT_BF16), idx = state channels of z,
                out0 / out1 = A1 [M][K1pad] (hi / lo)
*/

struct SegmentLaunch {
  const MicroOp* ops;   // device array
  int nops;
  int C;                // active channels of this level
  bool has_mcf;
  bool mma;             // masked-conv flows on mma.sync bf16x3 (tensor-core precisions, C <= 32)
};

constexpr int MAX_NSPLIT = 9;   // upper bound on the split-K slices of a NICE conv3

// state: [*][64][C0] fp32 NHWC (in place); logdet: [*] (forward only, accumulated); processes samples [b0, b0 + nb)
void flow_segment_run(const SegmentLaunch& s, bool forward, float* state, int C0, float* logdet, int b0, int nb, cudaStream_t st);
size_t flow_segment_smem_bytes(int C, bool has_mcf, int mode);   // mode 0 FFMA, 1 register-resident mma, 2 streamed mma
// mma.sync A-fragment packing of one MCF (see flow_segment.cu); sizes in 32-bit words (hi + lo planes)
size_t mcf_mma_conv_words(int C);
size_t mcf_mma_1x1_words(int C);
void pack_mcf_mma(const float* shift_w, const float* v, const float* os, uint32_t* conv_dst, uint32_t* x1_dst, int hid, int C,
                  int kh, int kw, int order, int row, cudaStream_t st);
constexpr int MCF_MMA_MAXC = 64;   // mma.sync MCF paths: register-resident fragments up to 32 channels, streamed from L2 up to 64
void flow_segment_init();

}  // namespace ipk
