// sampler.cuh -- per-voxel device code of the trilinear sampler, shared by all sampler kernels.
//
// Arithmetic follows stn3dtorch/generic/BilinearSamplerVolumetric.c operation for operation
// (forward ...c:442-583 / 9-167, backward ...c:585-829 / 169-439); see volstn_common.cuh.
#pragma once

#include "volstn_common.cuh"

namespace volstn {

// Kernel arguments.  Strides are element strides of dims 0..3 (B, D, H, W); dim 4 is unit stride.
template <typename real>
struct SamplerArgs {
  const real *vol;
  const real *grid;
  real *out;         // forward only
  real *gvol;        // backward
  real *ggrid;       // backward
  const real *gout;  // backward
  int B, iD, iH, iW, C;
  int oD, oH, oW;
  long long n;               // oD * oH * oW, voxels per sample
  long long vs[4], gs[4];    // volume, grid
  long long os[4];           // out (fwd) / gradOut (bwd)
  long long gvs[4], ggs[4];  // gradVol, gradGrid
  real dm1, hm1, wm1;        // (real)(iD-1), (real)(iH-1), (real)(iW-1)
};

// Everything that depends only on the grid element and the volume extents.
template <typename real>
struct VoxelSetup {
  int x0, y0, z0;
  real wx, wy, wz;  // weight of the LOW corner on each axis
  unsigned mask;    // bit k = corner k in bounds

  __device__ __forceinline__ void init(real zf, real yf, real xf, const SamplerArgs<real> &a) {
    axis_setup<real>(xf, a.wm1, x0, wx);
    axis_setup<real>(yf, a.hm1, y0, wy);
    axis_setup<real>(zf, a.dm1, z0, wz);
    const bool ix0 = in_range(x0, a.iW), ix1 = in_range(wrap_add(x0, 1), a.iW);
    const bool iy0 = in_range(y0, a.iH), iy1 = in_range(wrap_add(y0, 1), a.iH);
    const bool iz0 = in_range(z0, a.iD), iz1 = in_range(wrap_add(z0, 1), a.iD);
    mask = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const bool in = ((k & 1) ? ix1 : ix0) && ((k & 2) ? iy1 : iy0) && ((k & 4) ? iz1 : iz0);
      mask |= (in ? 1u : 0u) << k;
    }
  }
  __device__ __forceinline__ bool in(int k) const { return (mask >> k) & 1u; }

  // ((wx' * wy') * wz') for the eight corners, products exactly as ...c:566-573
  __device__ __forceinline__ void corner_weights(real (&cw)[8]) const {
    using E = Ex<real>;
    const real wxs[2] = {wx, E::sub(real(1), wx)};
    const real wys[2] = {wy, E::sub(real(1), wy)};
    const real wzs[2] = {wz, E::sub(real(1), wz)};
    real xy[4];
#pragma unroll
    for (int k = 0; k < 4; k++) xy[k] = E::mul(wxs[k & 1], wys[k >> 1]);
#pragma unroll
    for (int k = 0; k < 8; k++) cw[k] = E::mul(xy[k & 3], wzs[k >> 2]);
  }

  // element offset of corner 0 inside a sample whose (D, H, W) strides are (sD, sH, sW).
  // 32-bit wrapping arithmetic: out-of-range indices give garbage that is never dereferenced.
  __device__ __forceinline__ int base32(int sD, int sH, int sW) const {
    return wrap_add(wrap_add(wrap_mul(z0, sD), wrap_mul(y0, sH)), wrap_mul(x0, sW));
  }
  __device__ __forceinline__ long long base64(long long sD, long long sH, long long sW) const {
    return (long long)z0 * sD + (long long)y0 * sH + (long long)x0 * sW;
  }
};

// gradGrid from the eight corner dot products (...c:792-822): every sum runs over the corners
// in the order FTL, BTL, FTR, BTR, FBL, BBL, FBR, BBR; each term is ((dot * a) * b) * (+-1).
template <typename real>
__device__ __forceinline__ void grad_grid_from_dots(const real (&dot)[8], const VoxelSetup<real> &s,
                                                    const SamplerArgs<real> &a, real &gz, real &gy, real &gx) {
  using E = Ex<real>;
  const real wxs[2] = {s.wx, E::sub(real(1), s.wx)};
  const real wys[2] = {s.wy, E::sub(real(1), s.wy)};
  const real wzs[2] = {s.wz, E::sub(real(1), s.wz)};
  real dz = 0, dy = 0, dx = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int k = order_zxy(j);
    const real wa = wxs[k & 1], wb = wys[(k >> 1) & 1], wc = wzs[(k >> 2) & 1];
    real ty = E::mul(E::mul(dot[k], wa), wc);
    const real p = E::mul(dot[k], wb);
    real tx = E::mul(p, wc);
    real tz = E::mul(p, wa);
    if (!(k & 2)) ty = -ty;
    if (!(k & 1)) tx = -tx;
    if (!(k & 4)) tz = -tz;
    dy = (j == 0) ? ty : E::add(dy, ty);
    dx = (j == 0) ? tx : E::add(dx, tx);
    dz = (j == 0) ? tz : E::add(dz, tz);
  }
  gz = E::mul(E::mul(dz, a.dm1), real(0.5));
  gy = E::mul(E::mul(dy, a.hm1), real(0.5));
  gx = E::mul(E::mul(dx, a.wm1), real(0.5));
}

// ---------------------------------------------------------------------------------------------
// channel-vector loads / stores.  CT = compile-time channel count (1..4).  Vector instructions
// are used for float with CT = 2 or 4 (the dispatcher guarantees 8 / 16-byte alignment).
// ---------------------------------------------------------------------------------------------
template <typename real, int CT>
__device__ __forceinline__ void ld_channels(const real *p, real (&v)[CT]) {
  if constexpr (sizeof(real) == 4 && CT == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (sizeof(real) == 4 && CT == 2) {
    const float2 t = __ldg(reinterpret_cast<const float2 *>(p));
    v[0] = t.x; v[1] = t.y;
  } else {
#pragma unroll
    for (int t = 0; t < CT; t++) v[t] = __ldg(p + t);
  }
}

template <typename real, int CT>
__device__ __forceinline__ void ld_channels_stream(const real *p, real (&v)[CT]) {
  if constexpr (sizeof(real) == 4 && CT == 4) {
    const float4 t = ld_stream_f4(reinterpret_cast<const float4 *>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (sizeof(real) == 4 && CT == 2) {
    const float2 t = ld_stream_f2(reinterpret_cast<const float2 *>(p));
    v[0] = t.x; v[1] = t.y;
  } else if constexpr (sizeof(real) == 4) {
#pragma unroll
    for (int t = 0; t < CT; t++) v[t] = ld_stream_f1(p + t);
  } else {
#pragma unroll
    for (int t = 0; t < CT; t++) v[t] = __ldg(p + t);
  }
}

template <typename real, int CT>
__device__ __forceinline__ void st_channels_stream(real *p, const real (&v)[CT]) {
  if constexpr (sizeof(real) == 4 && CT == 4) {
    st_stream_f4(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3]));
  } else if constexpr (sizeof(real) == 4 && CT == 2) {
    st_stream_f2(reinterpret_cast<float2 *>(p), make_float2(v[0], v[1]));
  } else if constexpr (sizeof(real) == 4) {
#pragma unroll
    for (int t = 0; t < CT; t++) st_stream_f1(p + t, v[t]);
  } else {
#pragma unroll
    for (int t = 0; t < CT; t++) p[t] = v[t];
  }
}

template <typename real, int CT>
__device__ __forceinline__ void red_channels(real *p, const real (&v)[CT]) {
  if constexpr (sizeof(real) == 4 && CT == 4) {
    red_add4(p, v[0], v[1], v[2], v[3]);
  } else if constexpr (sizeof(real) == 4 && CT == 2) {
    red_add2(p, v[0], v[1]);
  } else {
#pragma unroll
    for (int t = 0; t < CT; t++) red_add(p + t, v[t]);
  }
}

// grid element (z, y, x): one 128-bit streaming load for float G = 4
template <typename real, int G>
__device__ __forceinline__ void ld_grid(const real *g, real &zf, real &yf, real &xf) {
  if constexpr (sizeof(real) == 4 && G == 4) {
    const float4 t = ld_stream_f4(reinterpret_cast<const float4 *>(g));
    zf = t.x; yf = t.y; xf = t.z;
  } else if constexpr (sizeof(real) == 4) {
    zf = ld_stream_f1(g); yf = ld_stream_f1(g + 1); xf = ld_stream_f1(g + 2);
  } else {
    zf = __ldg(g); yf = __ldg(g + 1); xf = __ldg(g + 2);
  }
}

template <typename real, int G>
__device__ __forceinline__ void st_grad_grid(real *g, real gz, real gy, real gx) {
  if constexpr (sizeof(real) == 4 && G == 4) {
    st_stream_f4(reinterpret_cast<float4 *>(g), make_float4(gz, gy, gx, 0.0f));
  } else {
    g[0] = gz; g[1] = gy; g[2] = gx;
    if constexpr (G == 4) g[3] = real(0);  // ...c:822
  }
}

}  // namespace volstn
