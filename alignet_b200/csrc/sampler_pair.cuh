// sampler_pair.cuh -- the float packed kernels' arithmetic on PAIRS of voxels (sm_100a FADD2 / FMUL2).
//
// The direct kernels are bound by instruction issue, not by HBM (profiles/r01_*: ~70 % of the issue slots
// busy at ~60 % of HBM; a do-nothing skeleton with the same memory operations runs 15-25 % faster,
// tools/ubench_stream.cu).  Blackwell adds packed fp32x2 arithmetic -- `add.rn.f32x2`, `mul.rn.f32x2`
// (SASS FADD2 / FMUL2), reached here through the __fadd2_rn / __fmul2_rn intrinsics -- which performs two
// independent IEEE round-to-nearest operations per issued instruction.  A thread already owns two voxels
// (VPT = 2), so every floating-point step of the reference's per-voxel recipe is done once for both:
// .x is the thread's first voxel, .y its second.  Each half is exactly the scalar operation of
// sampler.cuh (no contraction, same order), so out and gradGrid stay bit-identical to the reference.
// Integer work (floor bits, offsets, bounds predicates) and memory operations stay per voxel.
#pragma once

#include "sampler.cuh"

namespace volstn {

using f2 = float2;

// Packed arithmetic.  ptxas 12.9 contracts `mul.rn.f32x2` + `add.rn.f32x2` into FFMA2 even though both
// carry an explicit rounding mode (it never does that to the scalar forms; -fmad=false does not stop it),
// which would change the rounding of every sum of products.  So an ADD is issued as FFMA2(b, one, a)
// with `one` = 1.0f read from the kernel parameters, opaque to the optimiser: b * 1 is exact, the result
// is round(a + b) like add.rn, it costs the same single issue slot, and a multiply feeding it cannot be
// fused any further.  tests/ compare bit patterns with the reference, so a regression here is caught.
struct P2 {
  float one;  // SamplerArgs::one
  __device__ __forceinline__ f2 splat(float s) const { return make_float2(s, s); }
  __device__ __forceinline__ f2 neg(f2 a) const { return make_float2(-a.x, -a.y); }
  __device__ __forceinline__ f2 add(f2 a, f2 b) const { return __ffma2_rn(b, make_float2(one, one), a); }
  __device__ __forceinline__ f2 sub(f2 a, f2 b) const { return __ffma2_rn(neg(b), make_float2(one, one), a); }
  __device__ __forceinline__ f2 mul(f2 a, f2 b) const { return __fmul2_rn(a, b); }
};

// Set-up of two voxels.  MODE as in sampler.cuh: 1 = fast (all corners in bounds), 2 = border, 0 = exact.
struct PairSetup {
  f2 cx, cy, cz;  // coordinates ((g + 1) * (size - 1)) / 2
  f2 wx, wy, wz;  // weight of the low corner
  int base[2];    // element offset of corner 0 inside the sample
  bool ix0[2], ix1[2], iy0[2], iy1[2], iz0[2], iz1[2];

  // returns the tier both voxels qualify for (2 fast, 1 border, 0 exact) -- see VoxelSetup::begin
  __device__ __forceinline__ int begin(f2 zf, f2 yf, f2 xf, const SamplerArgs<float> &a) {
    const P2 E{a.one};
    const f2 one = E.splat(1.0f), half = E.splat(0.5f);
    cx = E.mul(E.mul(E.add(xf, one), E.splat(a.wm1)), half);
    cy = E.mul(E.mul(E.add(yf, one), E.splat(a.hm1)), half);
    cz = E.mul(E.mul(E.add(zf, one), E.splat(a.dm1)), half);
    int tier = 2;
#pragma unroll
    for (int v = 0; v < 2; v++) {
      const unsigned ux = (unsigned)__float_as_int(v ? cx.y : cx.x), uy = (unsigned)__float_as_int(v ? cy.y : cy.x),
                     uz = (unsigned)__float_as_int(v ? cz.y : cz.x);
      const bool strict = ux < a.fbx && uy < a.fby && uz < a.fbz;
      const bool inside = ux <= a.fbx && uy <= a.fby && uz <= a.fbz && a.border_ok;
      tier = min(tier, strict ? 2 : (inside ? 1 : 0));
    }
    return tier;
  }

  template <int MODE>
  __device__ __forceinline__ void finish(const SamplerArgs<float> &a, int sD, int sH, int sW) {
    const P2 E{a.one};
    const f2 one = E.splat(1.0f);
    if constexpr (MODE != 0) {
      // floor by the 2^23 trick (see VoxelSetup::finish_fast): the round-down adds stay scalar, the rest packs
      const f2 magic = E.splat(8388608.0f);
      const f2 rx = make_float2(__fadd_rd(cx.x, 8388608.0f), __fadd_rd(cx.y, 8388608.0f));
      const f2 ry = make_float2(__fadd_rd(cy.x, 8388608.0f), __fadd_rd(cy.y, 8388608.0f));
      const f2 rz = make_float2(__fadd_rd(cz.x, 8388608.0f), __fadd_rd(cz.y, 8388608.0f));
      wx = E.sub(one, E.sub(cx, E.sub(rx, magic)));
      wy = E.sub(one, E.sub(cy, E.sub(ry, magic)));
      wz = E.sub(one, E.sub(cz, E.sub(rz, magic)));
#pragma unroll
      for (int v = 0; v < 2; v++) {
        const int bx = __float_as_int(v ? rx.y : rx.x), by = __float_as_int(v ? ry.y : ry.x),
                  bz = __float_as_int(v ? rz.y : rz.x);
        base[v] = wrap_add(wrap_add(wrap_add(wrap_mul(bz, sD), a.fneg), wrap_mul(by, sH)), wrap_mul(bx, sW));
        ix0[v] = iy0[v] = iz0[v] = true;
        if constexpr (MODE == 2) {
          ix1[v] = (unsigned)__float_as_int(v ? cx.y : cx.x) < a.fbx;
          iy1[v] = (unsigned)__float_as_int(v ? cy.y : cy.x) < a.fby;
          iz1[v] = (unsigned)__float_as_int(v ? cz.y : cz.x) < a.fbz;
        } else {
          ix1[v] = iy1[v] = iz1[v] = true;
        }
      }
    } else {
      using S = Ex<float>;
      int x0[2], y0[2], z0[2];
      x0[0] = S::floor_to_int(cx.x), x0[1] = S::floor_to_int(cx.y);
      y0[0] = S::floor_to_int(cy.x), y0[1] = S::floor_to_int(cy.y);
      z0[0] = S::floor_to_int(cz.x), z0[1] = S::floor_to_int(cz.y);
      wx = E.sub(one, E.sub(cx, make_float2(S::from_int(x0[0]), S::from_int(x0[1]))));
      wy = E.sub(one, E.sub(cy, make_float2(S::from_int(y0[0]), S::from_int(y0[1]))));
      wz = E.sub(one, E.sub(cz, make_float2(S::from_int(z0[0]), S::from_int(z0[1]))));
#pragma unroll
      for (int v = 0; v < 2; v++) {
        base[v] = wrap_add(wrap_add(wrap_mul(z0[v], sD), wrap_mul(y0[v], sH)), wrap_mul(x0[v], sW));
        ix0[v] = in_range(x0[v], a.iW), ix1[v] = in_range(wrap_add(x0[v], 1), a.iW);
        iy0[v] = in_range(y0[v], a.iH), iy1[v] = in_range(wrap_add(y0[v], 1), a.iH);
        iz0[v] = in_range(z0[v], a.iD), iz1[v] = in_range(wrap_add(z0[v], 1), a.iD);
      }
    }
  }

  __device__ __forceinline__ bool in(int v, int k) const {
    return ((k & 1) ? ix1[v] : ix0[v]) && ((k & 2) ? iy1[v] : iy0[v]) && ((k & 4) ? iz1[v] : iz0[v]);
  }

  // ((wx' * wy') * wz') for the eight corners, products exactly as ...c:566-573
  __device__ __forceinline__ void corner_weights(f2 (&cw)[8], const P2 E) const {
    const f2 one = E.splat(1.0f);
    const f2 wxs[2] = {wx, E.sub(one, wx)};
    const f2 wys[2] = {wy, E.sub(one, wy)};
    const f2 wzs[2] = {wz, E.sub(one, wz)};
    f2 xy[4];
#pragma unroll
    for (int k = 0; k < 4; k++) xy[k] = E.mul(wxs[k & 1], wys[k >> 1]);
#pragma unroll
    for (int k = 0; k < 8; k++) cw[k] = E.mul(xy[k & 3], wzs[k >> 2]);
  }
};

// gradGrid of both voxels from the corner dot products (...c:792-822), see grad_grid_from_dots
__device__ __forceinline__ void grad_grid_from_dots2(const f2 (&dot)[8], const PairSetup &s, const SamplerArgs<float> &a,
                                                     f2 &gz, f2 &gy, f2 &gx) {
  const P2 E{a.one};
  const f2 one = E.splat(1.0f);
  const f2 wxs[2] = {s.wx, E.sub(one, s.wx)};
  const f2 wys[2] = {s.wy, E.sub(one, s.wy)};
  const f2 wzs[2] = {s.wz, E.sub(one, s.wz)};
  f2 dz = E.splat(0.f), dy = dz, dx = dz;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int k = order_zxy(j);
    const f2 wa = wxs[k & 1], wb = wys[(k >> 1) & 1], wc = wzs[(k >> 2) & 1];
    f2 ty = E.mul(E.mul(dot[k], wa), wc);
    const f2 p = E.mul(dot[k], wb);
    f2 tx = E.mul(p, wc);
    f2 tz = E.mul(p, wa);
    if (!(k & 2)) ty = E.neg(ty);
    if (!(k & 1)) tx = E.neg(tx);
    if (!(k & 4)) tz = E.neg(tz);
    dy = (j == 0) ? ty : E.add(dy, ty);
    dx = (j == 0) ? tx : E.add(dx, tx);
    dz = (j == 0) ? tz : E.add(dz, tz);
  }
  const f2 half = E.splat(0.5f);
  gz = E.mul(E.mul(dz, E.splat(a.dm1)), half);
  gy = E.mul(E.mul(dy, E.splat(a.hm1)), half);
  gx = E.mul(E.mul(dx, E.splat(a.wm1)), half);
}

// corner values of both voxels: v[k][t].x from voxel 0, .y from voxel 1; out-of-bounds corners read as 0
template <int CT, int MODE>
__device__ __forceinline__ void load_corners2(const PairSetup &s, const float *__restrict__ vb, int sD, int sH,
                                              f2 (&v)[8][CT]) {
  const CornerRows<float> r0(vb, s.base[0], sD, sH), r1(vb, s.base[1], sD, sH);
#pragma unroll
  for (int k = 0; k < 8; k++) {
    float a[CT], b[CT];
#pragma unroll
    for (int t = 0; t < CT; t++) a[t] = b[t] = 0.f;
    if (MODE == 1 || s.in(0, k)) ld_channels<float, CT>(r0.template corner<CT>(k), a);
    if (MODE == 1 || s.in(1, k)) ld_channels<float, CT>(r1.template corner<CT>(k), b);
#pragma unroll
    for (int t = 0; t < CT; t++) v[k][t] = make_float2(a[t], b[t]);
  }
}

template <int CT, int G, int MODE>
__device__ __forceinline__ void pair_forward(const SamplerArgs<float> &a, const PairSetup &s,
                                             const float *__restrict__ vb, int sD, int sH, float *__restrict__ o0,
                                             float *__restrict__ o1, bool valid0, bool valid1) {
  const P2 E{a.one};
  f2 cw[8];
  s.corner_weights(cw, E);
  f2 v[8][CT];
  load_corners2<CT, MODE>(s, vb, sD, sH, v);
  f2 acc[CT];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int k = (G == 4) ? order_xyz(j) : order_zxy(j);  // ...c:566-573 / ...c:149-156
#pragma unroll
    for (int t = 0; t < CT; t++) {
      const f2 term = E.mul(cw[k], v[k][t]);
      acc[t] = (j == 0) ? term : E.add(acc[t], term);
    }
  }
  float a0[CT], a1[CT];
#pragma unroll
  for (int t = 0; t < CT; t++) a0[t] = acc[t].x, a1[t] = acc[t].y;
  if (valid0) st_channels_stream<float, CT>(o0, a0);
  if (valid1) st_channels_stream<float, CT>(o1, a1);
}

// go[t].x / .y = gradOut of voxel 0 / 1.  gv_delta: gradVol sample base - volume sample base (elements).
template <int CT, int G, bool ONLY_GRID, bool ONLY_VOL, int MODE>
__device__ __forceinline__ void pair_backward(const SamplerArgs<float> &a, const PairSetup &s, const float *__restrict__ vb,
                                              long long gv_delta, int sD, int sH, const f2 (&go)[CT],
                                              float *__restrict__ gg0, float *__restrict__ gg1, bool valid0, bool valid1) {
  const P2 E{a.one};
  f2 cw[8];
  if constexpr (!ONLY_GRID) s.corner_weights(cw, E);
  f2 dot[8];
  if constexpr (!ONLY_VOL) {
    f2 v[8][CT];
    load_corners2<CT, MODE>(s, vb, sD, sH, v);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      f2 d = E.splat(0.f);  // dot += v * go, channel by channel (...c:736)
#pragma unroll
      for (int t = 0; t < CT; t++) d = E.add(d, E.mul(v[k][t], go[t]));
      // an out-of-bounds corner never sees go (it may be inf / nan)
      if constexpr (MODE != 1) d = make_float2(s.in(0, k) ? d.x : 0.f, s.in(1, k) ? d.y : 0.f);
      dot[k] = d;
    }
  }
  if constexpr (!ONLY_GRID) {
    const CornerRows<float> r0(vb, s.base[0], sD, sH), r1(vb, s.base[1], sD, sH);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      f2 r[CT];
#pragma unroll
      for (int t = 0; t < CT; t++) r[t] = E.mul(cw[k], go[t]);  // ...c:737
      float q0[CT], q1[CT];
#pragma unroll
      for (int t = 0; t < CT; t++) q0[t] = r[t].x, q1[t] = r[t].y;
      if (valid0 && (MODE == 1 || s.in(0, k)))
        red_channels<float, CT>(const_cast<float *>(r0.template corner<CT>(k)) + gv_delta, q0);
      if (valid1 && (MODE == 1 || s.in(1, k)))
        red_channels<float, CT>(const_cast<float *>(r1.template corner<CT>(k)) + gv_delta, q1);
    }
  }
  if constexpr (!ONLY_VOL) {
    f2 gz, gy, gx;
    grad_grid_from_dots2(dot, s, a, gz, gy, gx);
    if (valid0) st_grad_grid<float, G>(gg0, gz.x, gy.x, gx.x);
    if (valid1) st_grad_grid<float, G>(gg1, gz.y, gy.y, gx.y);
  }
}

}  // namespace volstn
