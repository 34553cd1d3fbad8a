// cagebrick.cuh -- argument block and launch interface of the BRICK kernels (cagebrick.cu): the cage warp of ALIGNet's
// training case (contiguous single-channel tensors) in the separable arithmetic of VOLSTN_FAST_ARITH.
#pragma once

#include "sampler.cuh"

namespace volstn {

struct BrickArgs {
  const float *vol;   // B x iD x iH x iW (x 1), contiguous inside a sample; vB = elements per sample step
  const float *cage;  // B x 3 x cD x cH x cW, contiguous
  float *gcage;       // backward / step: same shape, accumulated into
  const float *tgt;   // backward: gradOut; step: the target; B x oD x oH x oW (x 1), contiguous
  float *out;         // forward: the output; step: optional output (may be null)
  long long vB;
  long long gv_delta;  // gradVol - vol in elements (gradVol has the volume's layout); used by the <GV> kernels only
  int B, iD, iH, iW, oD, oH, oW, cD, cH, cW;
  // plan (brick_plan)
  int nzu, nyu, nxs, zsplit, ysplit, ntask;
  float hz2, hy2, hx2;        // (size - 1) / 2 per source axis
  unsigned fbz, fby, fbx;     // IEEE bits of (float)(size - 1): the strict-tier test
  float cmz, cmy, cmx;        // size + 1: upper clamp of the general tier
  float lz, ly, lx;           // size + 2: low-corner bound of the shifted coordinate
  float hz, hy, hx;           // size + 1: high-corner bound of the shifted coordinate
  int kstrict, kgen;          // constants that cancel the 2^23 exponent bits (and the shift) in the offset multiply-adds
  // training step
  float loss_norm;
  double *loss;
  unsigned long long *iou;
};

bool brick_supported(const BrickArgs &a);
void brick_plan(BrickArgs &a);
// mode: 0 forward, 1 backward from gradOut, 2 training step; grad_vol: also scatter the gradient of the source volume
int brick_launch(const BrickArgs &a, int mode, bool grad_vol, cudaStream_t st);

}  // namespace volstn
