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
  // fast path (float, packed): IEEE bits of wm1/hm1/dm1 (0 disables it, e.g. an extent > 2^22) and the
  // constant that cancels the 2^23 exponent bits in the offset multiply-adds (see finish_fast)
  unsigned fbx, fby, fbz;
  int fneg;
  // packed kernels: voxels per sample and elements per volume sample as 32-bit numbers (both < 2^31), so
  // that the per-CTA base pointers are one widening multiply each
  unsigned n32, vs32;
  int border_ok;  // 0 disables the border tier (VOLSTN_DEBUG_NO_BORDER, A/B timing only)
  int quad;       // backward fast tier, C = 1: 16-byte vector reductions on aligned quads: 1 = always (VOLSTN_ALGO_QUAD),
                  // 2 = per warp, when its cells are scattered (VOLSTN_ALGO_AUTO)
  int tile_agg;   // privatised backward: warp-aggregate the shared atomics (volume much smaller than the output)
  // generic kernels only: element stride of the LAST dimension (channels / grid components) of every tensor.
  // The reference assumes 1 (`+ t`, `+ 1`, `+ 2`: ...c:492-495, 556); honouring it lets a B x C x D x H x W
  // tensor be passed as a permuted view instead of through a materialised nn.Transpose copy
  long long vs4, gs4, os4, gvs4, ggs4;
};

// Everything that depends only on the grid element and the volume extents.
template <typename real>
struct VoxelSetup {
  int x0, y0, z0;
  real wx, wy, wz;  // weight of the LOW corner on each axis
  // per-axis bounds predicates of the low (i0) and high (i0+1) corner: `0 <= i && i <= n-1`
  // (...c:542-550).  Kept as six predicates (never packed) so that ptxas keeps them in predicate
  // registers; corner k is in bounds iff its three axis predicates hold.
  bool ix0, ix1, iy0, iy1, iz0, iz1;

  __device__ __forceinline__ void init(real zf, real yf, real xf, const SamplerArgs<real> &a) {
    axis_setup<real>(xf, a.wm1, x0, wx);
    axis_setup<real>(yf, a.hm1, y0, wy);
    axis_setup<real>(zf, a.dm1, z0, wz);
    ix0 = in_range(x0, a.iW), ix1 = in_range(wrap_add(x0, 1), a.iW);
    iy0 = in_range(y0, a.iH), iy1 = in_range(wrap_add(y0, 1), a.iH);
    iz0 = in_range(z0, a.iD), iz1 = in_range(wrap_add(z0, 1), a.iD);
  }

  // ---- two-phase set-up used by the packed kernels --------------------------------------------
  // begin()  : the three coordinates (identical operations to axis_setup) and a SUFFICIENT test for
  //            "all eight corners in bounds": 0 <= c < size-1 on every axis (false for NaN).
  // finish_fast()  : only if the whole warp passed the test.  floor() by the 2^23 trick -- for
  //            0 <= c < 2^23, fl(c + 2^23) rounded DOWN is exactly 2^23 + floor(c), so neither F2I
  //            nor I2F (quarter-rate XU pipe) nor the INT_MIN patch nor integer bounds tests are
  //            needed.  Same i0, same weight bits as the exact path.
  // finish_exact() : everything else.
  real cz, cy, cx;
  int fbase;  // finish_fast only: element offset of corner 0 inside the sample
  // tier of this voxel: 2 = every corner in bounds (0 <= c < size-1 on every axis), 1 = inside the volume
  // including its high faces (0 <= c <= size-1: only high corners with weight exactly 0 can be out of
  // bounds), 0 = anything else (outside, NaN, huge)
  __device__ __forceinline__ int begin(real zf, real yf, real xf, const SamplerArgs<real> &a) {
    cx = axis_coord<real>(xf, a.wm1);
    cy = axis_coord<real>(yf, a.hm1);
    cz = axis_coord<real>(zf, a.dm1);
    if constexpr (sizeof(real) == 4) {
      // for c >= +0 the IEEE bit pattern orders like the value, negative numbers and NaNs have larger
      // unsigned patterns than any finite positive bound: ONE unsigned compare per axis tests
      // 0 <= c < size-1 (or <=)
      const unsigned ux = (unsigned)__float_as_int(cx), uy = (unsigned)__float_as_int(cy), uz = (unsigned)__float_as_int(cz);
      const bool strict = ux < a.fbx && uy < a.fby && uz < a.fbz;
      const bool inside = ux <= a.fbx && uy <= a.fby && uz <= a.fbz && a.border_ok;
      return strict ? 2 : (inside ? 1 : 0);
    } else {
      return 0;
    }
  }
  // sD, sH, sW: element strides of the packed sample (sW = channels)
  template <bool BORDER = false>
  __device__ __forceinline__ void finish_fast(const SamplerArgs<real> &a, int sD, int sH, int sW) {
    if constexpr (sizeof(real) == 4) {
      const float rx = __fadd_rd(cx, 8388608.0f), ry = __fadd_rd(cy, 8388608.0f), rz = __fadd_rd(cz, 8388608.0f);
      wx = __fsub_rn(1.0f, __fsub_rn(cx, __fsub_rn(rx, 8388608.0f)));
      wy = __fsub_rn(1.0f, __fsub_rn(cy, __fsub_rn(ry, 8388608.0f)));
      wz = __fsub_rn(1.0f, __fsub_rn(cz, __fsub_rn(rz, 8388608.0f)));
      // bits(r) = 0x4B000000 + i0.  The offset z0*sD + y0*sH + x0*sW is formed from the raw bits in
      // wrapping 32-bit arithmetic; a.fneg = -0x4B000000*(sD+sH+sW) cancels the exponent bits, so the
      // three subtractions cost nothing.
      fbase = wrap_add(wrap_add(wrap_add(wrap_mul(__float_as_int(rz), sD), a.fneg), wrap_mul(__float_as_int(ry), sH)),
                       wrap_mul(__float_as_int(rx), sW));
      // the integers themselves (dead code in the direct kernels; the privatised backward needs them)
      x0 = __float_as_int(rx) - 0x4B000000, y0 = __float_as_int(ry) - 0x4B000000, z0 = __float_as_int(rz) - 0x4B000000;
      ix0 = iy0 = iz0 = true;
      if constexpr (BORDER) {  // c == size-1: i0 = size-1, the high corner is out of bounds (its weight is 0)
        ix1 = (unsigned)__float_as_int(cx) < a.fbx;
        iy1 = (unsigned)__float_as_int(cy) < a.fby;
        iz1 = (unsigned)__float_as_int(cz) < a.fbz;
      } else {
        ix1 = iy1 = iz1 = true;
      }
    }
  }
  __device__ __forceinline__ void finish_exact(const SamplerArgs<real> &a) {
    axis_finish<real>(cx, x0, wx);
    axis_finish<real>(cy, y0, wy);
    axis_finish<real>(cz, z0, wz);
    ix0 = in_range(x0, a.iW), ix1 = in_range(wrap_add(x0, 1), a.iW);
    iy0 = in_range(y0, a.iH), iy1 = in_range(wrap_add(y0, 1), a.iH);
    iz0 = in_range(z0, a.iD), iz1 = in_range(wrap_add(z0, 1), a.iD);
  }
  // a voxel that takes part in nothing (tail lanes): every predicate false, finite weights
  __device__ __forceinline__ void init_idle() {
    x0 = y0 = z0 = 0;
    wx = wy = wz = real(0);
    ix0 = ix1 = iy0 = iy1 = iz0 = iz1 = false;
  }
  __device__ __forceinline__ bool in(int k) const {
    return ((k & 1) ? ix1 : ix0) && ((k & 2) ? iy1 : iy0) && ((k & 4) ? iz1 : iz0);
  }
  __device__ __forceinline__ bool all_in() const { return ix0 && ix1 && iy0 && iy1 && iz0 && iz1; }
  __device__ __forceinline__ bool any_in() const { return (ix0 || ix1) && (iy0 || iy1) && (iz0 || iz1); }
  __device__ __forceinline__ unsigned mask() const {  // bit k = corner k in bounds (index dump only)
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) m |= (in(k) ? 1u : 0u) << k;
    return m;
  }

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

// grid element (z, y, x): one 128-bit streaming load for float G = 4.
// G >= 10 (packed float kernels only) = PLANAR grid with G % 10 components: the `B x G x D x H x W` tensor a
// convolutional network emits, handed over as its permuted view; `plane` = voxels per sample, `g`
// points at component 0 of the voxel and the three loads of a warp are three coalesced 128-byte rows.
template <typename real, int G>
__device__ __forceinline__ void ld_grid(const real *g, real &zf, real &yf, real &xf, unsigned plane = 0) {
  if constexpr (G >= 10) {
    zf = ld_stream_f1(g), yf = ld_stream_f1(g + plane), xf = ld_stream_f1(g + 2 * (size_t)plane);
  } else if constexpr (sizeof(real) == 4 && G == 4) {
    const float4 t = ld_stream_f4(reinterpret_cast<const float4 *>(g));
    zf = t.x; yf = t.y; xf = t.z;
  } else if constexpr (sizeof(real) == 4) {
    zf = ld_stream_f1(g); yf = ld_stream_f1(g + 1); xf = ld_stream_f1(g + 2);
  } else {
    zf = __ldg(g); yf = __ldg(g + 1); xf = __ldg(g + 2);
  }
}

template <typename real, int G>
__device__ __forceinline__ void st_grad_grid(real *g, real gz, real gy, real gx, unsigned plane = 0) {
  if constexpr (G >= 10) {
    st_stream_f1(g, gz), st_stream_f1(g + plane, gy), st_stream_f1(g + 2 * (size_t)plane, gx);
    if constexpr (G % 10 == 4) st_stream_f1(g + 3 * (size_t)plane, real(0));  // ...c:822
  } else if constexpr (sizeof(real) == 4 && G == 4) {
    st_stream_f4(reinterpret_cast<float4 *>(g), make_float4(gz, gy, gx, 0.0f));
  } else {
    g[0] = gz; g[1] = gy; g[2] = gx;
    if constexpr (G == 4) g[3] = real(0);  // ...c:822
  }
}

// =============================================================================================
// per-voxel bodies (packed layout: everything contiguous, 32-bit in-sample offsets)
//
// Instruction economy matters here: on a B200 the first version of these bodies (one int offset
// and one 64-bit address per corner, a packed bounds mask) issued 205 (forward) / 410 (backward)
// instructions per voxel and was ISSUE-bound (75 % issue-active at 47 % of HBM, profiles/r01).
// Now: four row pointers (the x+1 corner is an immediate offset), bounds as predicates, and a
// FAST variant for warps whose 8 x 32 corners are all in bounds (no predicates, no selects).
// =============================================================================================
// Makes a pointer opaque to the optimiser: ptxas otherwise re-derives every corner address from the
// kernel parameters (batch index * stride + ...) with 64-bit multiplies per corner.
template <typename T>
__device__ __forceinline__ T *opaque(T *p) {
  asm volatile("" : "+l"(p));
  return p;
}

template <typename real>
struct CornerRows {
  const real *p[4];  // rows (y0,z0) (y0+1,z0) (y0,z0+1) (y0+1,z0+1), each pointing at x0
  // `base`: element offset of corner 0 inside the sample (garbage for out-of-range indices, never
  // dereferenced then).  Four 32-bit offsets, four widening multiply-adds.
  __device__ __forceinline__ CornerRows(const real *sample, int base, int sD, int sH) {
    const int t1 = wrap_add(base, sH), t2 = wrap_add(base, sD), t3 = wrap_add(t2, sH);
    p[0] = sample + base;
    p[1] = sample + t1;
    p[2] = sample + t2;
    p[3] = sample + t3;
  }
  template <int CT>
  __device__ __forceinline__ const real *corner(int k) const { return p[k >> 1] + ((k & 1) ? CT : 0); }
};

template <typename real, int CT, int G, int MODE>
__device__ __forceinline__ void voxel_forward_ct(const VoxelSetup<real> &s, const real *__restrict__ vb, int sD,
                                                 int sH, real *__restrict__ o, bool valid = true) {
  using E = Ex<real>;
  static_assert(CT > 0, "compile-time channel count");
  real cw[8];
  s.corner_weights(cw);
  constexpr bool FAST = MODE == 1;
  const CornerRows<real> rows(vb, MODE ? s.fbase : s.base32(sD, sH, CT), sD, sH);
  real v[8][CT];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    if (FAST || s.in(k)) {
      ld_channels<real, CT>(rows.template corner<CT>(k), v[k]);
    } else {
#pragma unroll
      for (int t = 0; t < CT; t++) v[k][t] = real(0);  // out-of-bounds corners read as 0 (...c:531-564)
    }
  }
  real acc[CT];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int k = (G % 10 == 4) ? order_xyz(j) : order_zxy(j);  // ...c:566-573 / ...c:149-156
#pragma unroll
    for (int t = 0; t < CT; t++) {
      const real term = E::mul(cw[k], v[k][t]);
      acc[t] = (j == 0) ? term : E::add(acc[t], term);
    }
  }
  if (valid) st_channels_stream<real, CT>(o, acc);  // tail lanes compute a clamped voxel and drop it
}

template <typename real, int CT, int G>
__device__ __forceinline__ void voxel_forward_packed(const SamplerArgs<real> &a, const VoxelSetup<real> &s,
                                                     const real *__restrict__ vb, int sD, int sH, int sW,
                                                     real *__restrict__ o) {
  using E = Ex<real>;
  if constexpr (CT > 0) {
    voxel_forward_ct<real, CT, G, 0>(s, vb, sD, sH, o);
  } else {
    real cw[8];
    s.corner_weights(cw);
    const CornerRows<real> rows(vb, s.base32(sD, sH, sW), sD, sH);
    for (int t = 0; t < a.C; t++) {
      real acc = 0;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int k = (G == 4) ? order_xyz(j) : order_zxy(j);
        const real v = s.in(k) ? __ldg(rows.p[k >> 1] + ((k & 1) ? sW : 0) + t) : real(0);
        const real term = E::mul(cw[k], v);
        acc = (j == 0) ? term : E::add(acc, term);
      }
      o[t] = acc;
    }
  }
}

// go[] = gradOut of this voxel (already loaded).  gv_delta = gradVol sample base - volume sample base
// in elements (both tensors are packed with identical strides), so corner k of gradVol is
// rows.corner(k) + gv_delta.
template <typename real, int CT, int G, bool ONLY_GRID, bool ONLY_VOL, int MODE, bool QUAD = false>
__device__ __forceinline__ void voxel_backward_ct(const SamplerArgs<real> &a, const VoxelSetup<real> &s,
                                                  const real *__restrict__ vb, long long gv_delta, int sD, int sH,
                                                  const real (&go)[CT], real *__restrict__ ggp, bool valid = true) {
  using E = Ex<real>;
  static_assert(CT > 0, "compile-time channel count");
  real cw[8];
  if constexpr (!ONLY_GRID) s.corner_weights(cw);
  constexpr bool FAST = MODE == 1;
  const int base = MODE ? s.fbase : s.base32(sD, sH, CT);
  const CornerRows<real> rows(vb, base, sD, sH);
  real dot[8];
  if constexpr (!ONLY_VOL) {
    real v[8][CT];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (FAST || s.in(k)) {
        ld_channels<real, CT>(rows.template corner<CT>(k), v[k]);
      } else {
#pragma unroll
        for (int t = 0; t < CT; t++) v[k][t] = real(0);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
      real d = real(0);  // dot += v * go, channel by channel (...c:736)
#pragma unroll
      for (int t = 0; t < CT; t++) d = E::add(d, E::mul(v[k][t], go[t]));
      dot[k] = (FAST || s.in(k)) ? d : real(0);  // an out-of-bounds corner never sees go (it may be inf/nan)
    }
  }
  if constexpr (!ONLY_GRID) {
    // Measured and rejected on a B200 (gpurun_out/r1_kbench_{nzskip,rowskip,xmerge}.txt): skipping
    // reductions whose addend is an exact zero (per lane, or per row with a warp vote: 20 M -> 5.8 M L2
    // reduction sector-ops with the identity field, same run time -- the kernel is then issue-bound) and
    // merging x-adjacent contributions across lanes with shuffles (the identity grid's coordinates are one
    // ulp below an integer 36 % of the time, so only ~30 % of the neighbours line up).
    if constexpr (QUAD && FAST && CT == 1 && sizeof(real) == 4) {
      // VOLSTN_ALGO_QUAD: the x0 / x0+1 pair of a row goes out as ONE 16-byte vector reduction on the
      // aligned quad that contains x0 (zeros in the unused slots; a second scalar reduction only when x0
      // is the last cell of its quad).  With a sheared field every lane is its own L2 sector-op, so this is
      // 5 instead of 8 per voxel.  Needs 16-byte aligned rows (W % 4 == 0): the host checks.
      const unsigned q = (unsigned)(reinterpret_cast<uintptr_t>(rows.p[0] + gv_delta) >> 2) & 3u;
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const float lo = E::mul(cw[2 * r], go[0]), hi = E::mul(cw[2 * r + 1], go[0]);  // ...c:737
        float *p = const_cast<float *>(rows.p[r]) + gv_delta - q;
        const float vx = q == 0 ? lo : 0.f;
        const float vy = q == 0 ? hi : (q == 1 ? lo : 0.f);
        const float vz = q == 1 ? hi : (q == 2 ? lo : 0.f);
        const float vw = q == 2 ? hi : (q == 3 ? lo : 0.f);
        if (valid) {
          red_add4(p, vx, vy, vz, vw);
          if (q == 3) red_add(p + 4, hi);
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if (valid && (FAST || s.in(k))) {
          real r[CT];
#pragma unroll
          for (int t = 0; t < CT; t++) r[t] = E::mul(cw[k], go[t]);  // ...c:737
          red_channels<real, CT>(const_cast<real *>(rows.template corner<CT>(k)) + gv_delta, r);
        }
      }
    }
  }
  if constexpr (!ONLY_VOL) {
    real gz, gy, gx;
    grad_grid_from_dots<real>(dot, s, a, gz, gy, gx);
    if (valid) st_grad_grid<real, G>(ggp, gz, gy, gx, a.n32);
  }
}

template <typename real, int CT, int G, bool ONLY_GRID, bool ONLY_VOL>
__device__ __forceinline__ void voxel_backward_packed(const SamplerArgs<real> &a, const VoxelSetup<real> &s,
                                                      const real *__restrict__ vb, real *__restrict__ gvb, int sD,
                                                      int sH, int sW, const real *__restrict__ gop,
                                                      real *__restrict__ ggp) {
  using E = Ex<real>;
  if constexpr (CT > 0) {
    real go[CT];
    ld_channels_stream<real, CT>(gop, go);
    voxel_backward_ct<real, CT, G, ONLY_GRID, ONLY_VOL, 0>(a, s, vb, gvb - vb, sD, sH, go, ggp);
  } else {
    real cw[8];
    if constexpr (!ONLY_GRID) s.corner_weights(cw);
    const CornerRows<real> rows(vb, s.base32(sD, sH, sW), sD, sH);
    const long long gv_delta = gvb - vb;
    real dot[8];
#pragma unroll
    for (int k = 0; k < 8; k++) dot[k] = real(0);
    for (int t = 0; t < a.C; t++) {
      const real go = __ldg(gop + t);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if (s.in(k)) {
          const real *pk = rows.p[k >> 1] + ((k & 1) ? sW : 0) + t;
          if constexpr (!ONLY_VOL) dot[k] = E::add(dot[k], E::mul(__ldg(pk), go));
          if constexpr (!ONLY_GRID) red_add(const_cast<real *>(pk) + gv_delta, E::mul(cw[k], go));
        }
      }
    }
    if constexpr (!ONLY_VOL) {
      real gz, gy, gx;
      grad_grid_from_dots<real>(dot, s, a, gz, gy, gx);
      st_grad_grid<real, G>(ggp, gz, gy, gx);
    }
  }
}

// shared-memory privatised backward (sampler_tile.cu): packed float tensors, C in {1, 2, 4}, 16-byte rows
// TMA-staged kernels (sampler_tma.cu): packed single-channel float tensors, 16-byte volume rows
bool tma_supported(const SamplerArgs<float> &a, const float *gvol);
int launch_fwd_tma(const SamplerArgs<float> &a, int G, cudaStream_t st);
int launch_bwd_tma(const SamplerArgs<float> &a, int G, bool only_grid, cudaStream_t st);
bool tile_supported(const SamplerArgs<float> &a);
int launch_bwd_tile(const SamplerArgs<float> &a, int G, bool only_vol, cudaStream_t st);

}  // namespace volstn
