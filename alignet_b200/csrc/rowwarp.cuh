// rowwarp.cuh -- argument block and launch interface of the ROW kernels (rowwarp.cu).
//
// The packed kernels of sampler.cu walk a sample as a flat list of voxels and therefore need every
// tensor contiguous and channels-last.  The row kernels give a warp whole rows of the OUTPUT volume:
// (b, z, y) are warp-uniform, lanes run along x.  That buys, for one integer divide per row,
//   * arbitrary element strides on the B, D and H dimensions of every tensor and two W / channel
//     layouts (channels-last `... x W x C` or planar `C x ... x W`), i.e. the B x C x D x H x W tensors
//     the reference only reaches through materialised nn.Transpose copies
//     (stn3dtorch/test.lua:17,21; models/alignet.lua:11,14,21)                      -- SURVEY 8 f2
//   * sampling coordinates that are a function of (z, y, x) and never touch memory:
//       - T[b] . (zn, yn, xn, 1), the grid generator folded into the sampler
//         (stn3dtorch/test.lua:15-25, VolumetricGridGenerator.lua:58-60)             -- SURVEY 8 f3
//       - the cage up-sample of ALIGNet: the field is the trilinear interpolation of a small
//         control cage at the identity coordinates (models/alignet_2d.lua:136-147)   -- SURVEY 8 f1
//     with the arithmetic of the un-fused modules reproduced operation for operation, so that floor
//     indices, masks and outputs keep their bits.
#pragma once

#include "sampler.cuh"

namespace volstn {

enum RowSource { ROW_SRC_GRID = 0, ROW_SRC_AFFINE = 1, ROW_SRC_CAGE = 2 };

struct RowArgs {
  SamplerArgs<float> s;      // pointers, extents, fast-path constants (fneg for THIS volume's strides)
  long long vB, oB, gB, ggB; // per-sample element strides: volume (= gradVol), out / gradOut, grid, gradGrid
  int vD, vH, vW, vC;        // volume strides inside a sample (vW = C channels-last, 1 planar)
  int osD, osH, osW, osC;    // out (forward) / gradOut (backward)
  int gD, gH, gW, gC;        // grid; gC = component stride
  int ggD, ggH, ggW, ggC;    // gradGrid
  int rows_per_warp;
  int nblk;                  // CTAs per sample (gridDim.x)
  const float *T;            // ROW_SRC_AFFINE: B x 4 x 4
  double *partial;           // ROW_SRC_AFFINE backward: [B][nblk][16] partial gradT
  const float *cage;         // ROW_SRC_CAGE: B x 3 x cD x cH x cW
  float *gcage;              // ROW_SRC_CAGE backward: same shape, accumulated into
  int cD, cH, cW;
  float *gradT;              // ROW_SRC_AFFINE backward: B x 4 x 4 (row 3 is zero)
  // fused training step (k_row_bwd<..., LOSS>): s.gout is the TARGET, s.out (optional) receives the forward value
  float loss_norm;           // SmoothL1 gradient scale: 1 / nElement (sizeAverage) or 1
  double *loss;              // B sums of the SmoothL1 terms (accumulated into)
  unsigned long long *iou;   // B x 2 (intersection, union) counts at threshold 0.5 (accumulated into), may be null
  int g3;                    // ROW_SRC_GRID: grid width 3 (forward corner order of ...c:149-156, no 4th gradGrid component)
};

// what a launch needs beyond RowArgs
struct RowPlan {
  int src;        // RowSource
  int grid_aos4;  // ROW_SRC_GRID: grid (and gradGrid) are `... x 4` with unit component stride, 16-byte aligned
  int c1;         // single channel (compile-time body); otherwise the runtime channel loop
  int two_rows;   // oW <= 32: a warp trip covers two rows
  int fast;       // ROW_SRC_CAGE, single channel: VOLSTN_FAST_ARITH (separable cage up-sample, FMA sums, factored gradGrid)
};

// CTAs per sample and rows per warp for `rows` = oD * oH output rows (deterministic: the workspace query and the
// launch must agree)
void row_geometry(int64_t B, int64_t rows, int src, int two_rows, int *nblk, int *rows_per_warp);
int row_forward(const RowArgs &a, const RowPlan &p, int B, cudaStream_t st);
int row_backward(const RowArgs &a, const RowPlan &p, int B, bool only_grid, bool only_vol, cudaStream_t st);
int row_cage_step(const RowArgs &a, const RowPlan &p, int B, bool only_grid, cudaStream_t st);

}  // namespace volstn
