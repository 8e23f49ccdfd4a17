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
                p0 = Wc [6][Cp/4][hid][4], p1 = W1x [hid/4][2C][4],
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
};

constexpr int MAX_NSPLIT = 9;   // upper bound on the split-K slices of a NICE conv3

// state: [B][64][C0] fp32 NHWC (in place); logdet: [B] (forward only, accumulated)
void flow_segment_run(const SegmentLaunch& s, bool forward, float* state, int C0, float* logdet, int B, cudaStream_t st);
size_t flow_segment_smem_bytes(int C, int h_ch, bool has_mcf);
void flow_segment_init();

}  // namespace ipk
