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
  int sD;                     // iH * iW: elements per source plane
  float hz2, hy2, hx2;        // (size - 1) / 2 per source axis
  unsigned fbz, fby, fbx;     // IEEE bits of (float)(size - 1): the strict-tier test
  float cmz, cmy, cmx;        // size + 1: upper clamp of the general tier
  float lz, ly, lx;           // size + 2: low-corner bound of the shifted coordinate
  float hz, hy, hx;           // size + 1: high-corner bound of the shifted coordinate
  float ez, ey, ex;           // strict tier: -eps <= c < size - 1, eps = size * 2^-23 ...
  unsigned ubz, uby, ubx;     // ... as the exclusive bound on the IEEE bits of c + eps
  unsigned uiz, uiy, uix;     // border tier: c + eps < size, same test
  float tz1, ty1, tx1;        // size - 1: in the border tier a high corner exists iff c < size - 1
  int kstrict, kgen;          // constants that cancel the 2^23 exponent bits (and the shift) in the offset multiply-adds
  // training step
  float loss_norm;
  double *loss;
  unsigned long long *iou;
};

#ifdef __CUDACC__
constexpr int kBrickThreads = 256;
constexpr int kBrickWarps = kBrickThreads / 32;

// (float)(-1 + k/(n-1)*2) in double: VolumetricGridGenerator.lua:17-19 (as rowwarp.cu::row_base_coord)
__device__ __forceinline__ float brick_base_coord(int k, int n) {
  return (float)(-1.0 + (double)k / (double)(n - 1) * 2.0);
}

struct BrickSmem {
  float2 *wz, *wy, *wx;  // (weight of the low cage corner, weight of the high one) per output position
  int *iz, *iy, *ix;     // low cage corner per output position, remapped so that the high corner exists (see brick_fill)
  int *zs, *ys;          // first position of z / y unit k (k = 0 .. units)
  float *lane;           // per-thread slots: 12 folded cage values, 12 gradient accumulators ([slot][thread])
};

__host__ __device__ inline int brick_units(int cn) { return cn > 1 ? cn - 1 : 1; }

__host__ __device__ inline size_t brick_smem_bytes(const BrickArgs &a) {
  const size_t axes = (size_t)a.oD + a.oH + a.oW;
  return 24 * 256 * 4 + axes * 12 + (size_t)(brick_units(a.cD) + brick_units(a.cH) + 2) * 4 + 16;
}

__device__ __forceinline__ void brick_fill(BrickSmem &m, unsigned char *raw, const BrickArgs &a) {
  const int oD = a.oD, oH = a.oH, oW = a.oW, axes = oD + oH + oW;
  m.lane = reinterpret_cast<float *>(raw);
  m.wz = reinterpret_cast<float2 *>(raw + 24 * 256 * 4);
  m.wy = m.wz + oD;
  m.wx = m.wy + oH;
  m.iz = reinterpret_cast<int *>(m.wz + axes);
  m.iy = m.iz + oD;
  m.ix = m.iy + oH;
  m.zs = m.iz + axes;
  m.ys = m.zs + brick_units(a.cD) + 1;
  // the sampler's own set-up (...c:507-517) of the identity coordinate of every axis position.  The identity coordinate
  // never leaves [0, cn - 1]; its last position sits ON the last cage node (i0 = cn - 1, low weight 1, high corner in
  // the zero padding): there the cell below with weights (0, 1) is the same value, and every position of an axis then
  // has its high corner inside the cage.  A one-node axis keeps (i0 = 0, high weight 0).
  for (int i = threadIdx.x; i < axes; i += blockDim.x) {
    int k, n, cn;
    if (i < oD)
      k = i, n = oD, cn = a.cD;
    else if (i < oD + oH)
      k = i - oD, n = oH, cn = a.cH;
    else
      k = i - oD - oH, n = oW, cn = a.cW;
    int i0;
    float w;
    axis_setup<float>(brick_base_coord(k, n), (float)(cn - 1), i0, w);
    float2 ww = make_float2(w, __fsub_rn(1.0f, w));
    if (cn == 1)
      i0 = 0, ww = make_float2(w, 0.0f);
    else if (i0 >= cn - 1)
      i0 = cn - 2, ww = make_float2(1.0f - w, w);  // w == 1 there: weights (0, 1) on the cell below
    else if (i0 < 0)
      i0 = 0, ww = make_float2(1.0f, 0.0f);        // cannot happen (the coordinate is >= 0); keeps the indices safe
    m.wz[i] = ww;  // wz / wy / wx and iz / iy / ix are consecutive
    m.iz[i] = i0;
  }
  __syncthreads();
  // unit k of an axis = the positions whose (remapped) low corner is k: a contiguous, possibly empty range
  const int nzu = brick_units(a.cD), nyu = brick_units(a.cH);
  for (int i = threadIdx.x; i <= nzu + nyu + 1; i += blockDim.x) {
    const bool isz = i <= nzu;
    const int k = isz ? i : i - nzu - 1, n = isz ? oD : oH, units = isz ? nzu : nyu;
    const int *tab = isz ? m.iz : m.iy;
    int first = n;
    if (k < units) {
      first = 0;
      while (first < n && min(tab[first], units - 1) < k) first++;
    }
    (isz ? m.zs : m.ys)[k] = first;
  }
  __syncthreads();
}

#endif  // __CUDACC__

bool brick_supported(const BrickArgs &a);
void brick_plan(BrickArgs &a);
// mode: 0 forward, 1 backward from gradOut, 2 training step; grad_vol: also scatter the gradient of the source volume
int brick_launch(const BrickArgs &a, int mode, bool grad_vol, cudaStream_t st);

}  // namespace volstn
