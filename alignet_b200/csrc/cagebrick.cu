// cagebrick.cu -- BRICK kernels: ALIGNet's cage warp (models/alignet_2d.lua:136-147 lifted to 3-D) for
// its training case -- contiguous single-channel tensors -- in the separable arithmetic of VOLSTN_FAST_ARITH.
//
// The row kernels (rowwarp.cu) spend two thirds of their 596 warp instructions per 32 voxels on addresses, predicates
// and shared-memory traffic: every voxel re-derives its cage cell, reloads eight cage corners from shared
// memory and stages its gradient for a per-row segment sum.  The mapping here removes that work instead of making the
// arithmetic cheaper:
//
//   * a WARP owns a brick: one 32-wide x slot of the rows of ONE cage cell in y and (a part of) one cage cell in z.  A
//     lane's x never changes, so its cage column (ix, wx) is a constant of the task, and so is the cell (kz, ky): the
//     twelve x-interpolated cage values  P[j][k][c] = C[c][kz+k][ky+j][ix] + wx.(C[..][ix+1] - C[..][ix])  are formed ONCE
//     per task (shared memory, [slot][thread]), folded along z and into voxel units of the source once per z
//     (C0[c], D[c] = C1[c] - C0[c]: six registers) and a voxel's three sampling coordinates are ONE FMA each,
//     c = C0 + wy.D.  The `low + w.(high - low)` form is exact where two nodes coincide: the identity cage yields
//     coordinates ON the integers;
//   * the gradient takes the same road back: per voxel  T0[c] += g_c, T1[c] += wy.g_c  with g = go . d out/d coord
//     (six instructions), per z twelve FMAs into shared-memory accumulators A[j][k][c], and per TASK one segmented warp
//     reduction over the lanes that share a cage column (consecutive lanes; shuffles by doubling) and one red.global per
//     (cell corner, channel) -- no staging, no per-row sums;
//   * the trilinear value and its three coordinate derivatives come from the seven lerps of the separable form
//     (22 flops instead of the reference's 8 + 7 and 40 + 21);
//   * three tiers per warp and trip of rows, chosen by warp votes on one unsigned compare per axis: STRICT (every corner
//     of every lane inside the volume; a coordinate a few ulp below 0 -- the identity cage's low faces -- is read as 0:
//     conversion-free floor, no predicates), BORDER (0 <= c < size: in the last cell of an axis the high corner is the
//     reference's zero padding, ...c:542-564 -- the identity cage puts a whole face exactly ON size - 1, where the
//     reference takes that cell -- three predicates) and GENERAL (coordinates clamped to [-2, size + 1] and shifted by
//     2 so that the same floor trick applies; corners outside the volume are predicated off: zero value AND zero
//     derivative beyond one voxel outside).
//
// Same mathematics as the reference chain, different rounding: these kernels exist only behind VOLSTN_FAST_ARITH
// (include/volstn_b200.h), whose guarantees are measured in tests/test_gpu_rowwarp.py::test_fast_arith_*.
#include <cstdlib>
#include <cstring>

#include "cagebrick.cuh"

namespace volstn {

// ---------------------------------------------------------------------------------------------
// NR consecutive rows of the task: coordinates -> ONE tier vote -> all gathers in flight -> values, derivatives, sinks
// ---------------------------------------------------------------------------------------------
struct BrickLane {
  float T0[3], T1[3];  // sums over the rows of one z of g_c and of wy_hi . g_c (g = gradient of the voxel's coordinates)
  float lsum;          // SmoothL1 terms of this lane (float: at most a few thousand terms of O(1) per task)
  int ni, nu;          // IoU counts of this lane
};

// MODE: 0 = forward (store out), 1 = backward from gradOut, 2 = training step (SmoothL1 + IoU + backward, optional out)
// The rows of a trip share one tier vote so that all their gathers are in flight together, and the target / gradOut
// element of every row is requested before the vote.
// TIER: 0 = strict (no predicates), 1 = border (high corners may lie in the padding), 2 = general
template <int MODE, bool GV, int NR, int TIER>
__device__ __forceinline__ void brick_body(const BrickArgs &a, const float *__restrict__ vb, const float (&c)[NR][3],
                                           const float (&wyh)[NR], const float (&tg)[NR], const float *__restrict__ tp,
                                           float *__restrict__ op, bool keep, bool valid, BrickLane &L) {
  const int sH = a.iW, sD = a.sD;
  float f[NR][3];
  const float *p0[NR];
  bool zl[NR], zh[NR], yl[NR], yh[NR], xl[NR], xh[NR];
#pragma unroll
  for (int r = 0; r < NR; r++) {
    int off;
    if constexpr (TIER == 2) {
      // clamp to [-2, size + 1] (NaN -> -2), shift by 2: t in [0, size + 3], floor by the 2^23 trick; corner i0 = j - 2
      const float tz = __fadd_rn(fminf(fmaxf(c[r][0], -2.0f), a.cmz), 2.0f);
      const float ty = __fadd_rn(fminf(fmaxf(c[r][1], -2.0f), a.cmy), 2.0f);
      const float tx = __fadd_rn(fminf(fmaxf(c[r][2], -2.0f), a.cmx), 2.0f);
      const float rz = __fadd_rd(tz, 8388608.0f), ry = __fadd_rd(ty, 8388608.0f), rx = __fadd_rd(tx, 8388608.0f);
      f[r][0] = __fsub_rn(tz, __fsub_rn(rz, 8388608.0f));
      f[r][1] = __fsub_rn(ty, __fsub_rn(ry, 8388608.0f));
      f[r][2] = __fsub_rn(tx, __fsub_rn(rx, 8388608.0f));
      // low corner inside: 2 <= t < size + 2; high corner inside: 1 <= t < size + 1
      zl[r] = tz >= 2.0f && tz < a.lz, zh[r] = tz >= 1.0f && tz < a.cmz;
      yl[r] = ty >= 2.0f && ty < a.ly, yh[r] = ty >= 1.0f && ty < a.cmy;
      xl[r] = tx >= 2.0f && tx < a.lx, xh[r] = tx >= 1.0f && tx < a.cmx;
      off = wrap_add(wrap_add(wrap_add(wrap_mul(__float_as_int(rz), sD), a.kgen), wrap_mul(__float_as_int(ry), sH)),
                     __float_as_int(rx));
    } else {
      // strict: -eps <= c < size - 1 on every axis; border: -eps <= c < size.  A coordinate that rounding left a hair
      // below 0 is read as 0: the value moves by less than eps x slope.  ON the high face (the identity cage puts a whole
      // face exactly there) the reference takes the cell outside, high corners in the zero padding (...c:542-564): that
      // is the border tier's predicate, never a clamp.
      const float cz = fmaxf(c[r][0], 0.0f), cy = fmaxf(c[r][1], 0.0f), cx = fmaxf(c[r][2], 0.0f);
      const float rz = __fadd_rd(cz, 8388608.0f), ry = __fadd_rd(cy, 8388608.0f), rx = __fadd_rd(cx, 8388608.0f);
      f[r][0] = __fsub_rn(cz, __fsub_rn(rz, 8388608.0f));
      f[r][1] = __fsub_rn(cy, __fsub_rn(ry, 8388608.0f));
      f[r][2] = __fsub_rn(cx, __fsub_rn(rx, 8388608.0f));
      zl[r] = yl[r] = xl[r] = true;
      // border tier: 0 <= c < size; in the last cell [size - 1, size) the high corner is the zero padding (...c:542-564)
      zh[r] = TIER == 0 || cz < a.tz1, yh[r] = TIER == 0 || cy < a.ty1, xh[r] = TIER == 0 || cx < a.tx1;
      off = wrap_add(wrap_add(wrap_add(wrap_mul(__float_as_int(rz), sD), a.kstrict), wrap_mul(__float_as_int(ry), sH)),
                     __float_as_int(rx));
    }
    p0[r] = vb + off;  // corner 0; garbage when that corner is outside (never dereferenced then)
  }
  float v[NR][8];
#pragma unroll
  for (int r = 0; r < NR; r++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const bool in = TIER == 0 || (((k & 4) ? zh[r] : zl[r]) && ((k & 2) ? yh[r] : yl[r]) && ((k & 1) ? xh[r] : xl[r]));
      v[r][k] = in ? __ldg(p0[r] + (k & 1) + ((k & 2) ? sH : 0) + ((k & 4) ? sD : 0)) : 0.0f;
    }
  }
#pragma unroll
  for (int r = 0; r < NR; r++) {
    const float fz = f[r][0], fy = f[r][1], fx = f[r][2];
    // seven lerps: x on the four rows, y on the two planes, z
    const float dx0 = v[r][1] - v[r][0], dx1 = v[r][3] - v[r][2], dx2 = v[r][5] - v[r][4], dx3 = v[r][7] - v[r][6];
    const float e0v = fmaf(fx, dx0, v[r][0]), e1 = fmaf(fx, dx1, v[r][2]), e2 = fmaf(fx, dx2, v[r][4]), e3 = fmaf(fx, dx3, v[r][6]);
    const float dy0 = e1 - e0v, dy1 = e3 - e2;
    const float g0 = fmaf(fy, dy0, e0v), g1 = fmaf(fy, dy1, e2);
    const float dzv = g1 - g0;
    if constexpr (MODE == 0) {
      if (valid) st_stream_f1(op + r * a.oW, fmaf(fz, dzv, g0));
    } else {
      float go;
      if constexpr (MODE == 2) {
        // a lane beyond the row (ragged last x slot) carries out = target = 0: every term below vanishes without a branch
        const float o = valid ? fmaf(fz, dzv, g0) : 0.0f;
        const float x = o - tg[r];
        // THNN SmoothL1Criterion: z = |x|; loss += z < 1 ? 0.5 z^2 : z - 0.5; gradient norm * clamp(x, -1, 1)
        const float cl = fminf(fmaxf(x, -1.0f), 1.0f);
        go = a.loss_norm * cl;
        // z < 1 ? 0.5 z^2 : z - 0.5  ==  cl . (x - 0.5 cl)  with cl = clamp(x, -1, 1): two FMAs instead of a branchy seven
        L.lsum = fmaf(cl, fmaf(-0.5f, cl, x), L.lsum);
        const bool p = o > 0.5f, q = tg[r] > 0.5f;  // computeIOU (util.lua:110-120)
        L.ni += (p & q) ? 1 : 0;
        L.nu += (p | q) ? 1 : 0;
        if (keep && valid) st_stream_f1(op + r * a.oW, o);
      } else {
        go = valid ? tg[r] : 0.0f;
      }
      // d out / d coordinate (voxel units): z: dzv; y: lerp_z of the y differences; x: lerp_z lerp_y of the x differences
      const float dyv = fmaf(fz, dy1 - dy0, dy0);
      const float hx0 = fmaf(fy, dx1 - dx0, dx0), hx1 = fmaf(fy, dx3 - dx2, dx2);
      const float dxv = fmaf(fz, hx1 - hx0, hx0);
      const float gz = go * dzv, gy = go * dyv, gx = go * dxv;
      L.T0[0] += gz, L.T1[0] = fmaf(wyh[r], gz, L.T1[0]);
      L.T0[1] += gy, L.T1[1] = fmaf(wyh[r], gy, L.T1[1]);
      L.T0[2] += gx, L.T1[2] = fmaf(wyh[r], gx, L.T1[2]);
      if constexpr (GV) {
        // gradVol += corner weight * go (...c:737); gradVol has the volume's layout
        const float ux = 1.0f - fx, uy = 1.0f - fy, uz = 1.0f - fz;
        const float q[4] = {uz * uy * go, uz * fy * go, fz * uy * go, fz * fy * go};
        float *g0p = const_cast<float *>(p0[r]) + a.gv_delta;
        if (valid) {
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const bool in = TIER == 0 || (((k & 4) ? zh[r] : zl[r]) && ((k & 2) ? yh[r] : yl[r]) && ((k & 1) ? xh[r] : xl[r]));
            if (in) red_add(g0p + (k & 1) + ((k & 2) ? sH : 0) + ((k & 4) ? sD : 0), q[k >> 1] * ((k & 1) ? fx : ux));
          }
        }
      }
    }
  }
}

// C0[c]: the voxel coordinate (source voxel units) at the low y corner of the row's cage cell, D[c] = high - low
template <int MODE, bool GV, int NR>
__device__ __forceinline__ void brick_rows(const BrickArgs &a, const float *__restrict__ vb, const float (&C0)[3],
                                           const float (&D)[3], const float (&wyh)[NR], const float *__restrict__ tp,
                                           float *__restrict__ op, bool keep, bool valid, BrickLane &L) {
  float c[NR][3], tg[NR];
  unsigned u[NR][3];
  bool strict = true;
#pragma unroll
  for (int r = 0; r < NR; r++) {
#pragma unroll
    for (int i = 0; i < 3; i++) c[r][i] = fmaf(wyh[r], D[i], C0[i]);
    // strict: -eps <= c < size - 1 on every axis, inside: -eps <= c < size - eps: unsigned compares on the IEEE bits of
    // c + eps (negative and NaN patterns are larger than any positive bound)
    u[r][0] = (unsigned)__float_as_int(c[r][0] + a.ez), u[r][1] = (unsigned)__float_as_int(c[r][1] + a.ey);
    u[r][2] = (unsigned)__float_as_int(c[r][2] + a.ex);
    strict = strict & (u[r][0] < a.ubz) & (u[r][1] < a.uby) & (u[r][2] < a.ubx);
    tg[r] = 0.0f;
    if constexpr (MODE != 0) {
      if (valid) tg[r] = ld_stream_f1(tp + r * a.oW);
    }
  }
  if (__all_sync(0xffffffffu, strict)) {
    brick_body<MODE, GV, NR, 0>(a, vb, c, wyh, tg, tp, op, keep, valid, L);
    return;
  }
  bool inside = true;  // only the trips that are not strict pay for the second test
#pragma unroll
  for (int r = 0; r < NR; r++) inside = inside & (u[r][0] < a.uiz) & (u[r][1] < a.uiy) & (u[r][2] < a.uix);
  if (__all_sync(0xffffffffu, inside))
    brick_body<MODE, GV, NR, 1>(a, vb, c, wyh, tg, tp, op, keep, valid, L);
  else
    brick_body<MODE, GV, NR, 2>(a, vb, c, wyh, tg, tp, op, keep, valid, L);
}

#ifndef BRICK_MIN_CTAS
#define BRICK_MIN_CTAS 3
#endif
// rows per trip: two for the kernels with a backward (three cost registers: step 163 -> 192 us), three for the forward
// (identity cage 84.6 -> 77.8 us, jittered 140 -> 138 us); profiles/r02_fbench_brick_variants.txt
#ifndef BRICK_NR
#define BRICK_NR 2
#endif
#ifndef BRICK_NR_FWD
#define BRICK_NR_FWD 3
#endif

template <int MODE, bool GV>
__global__ void __launch_bounds__(kBrickThreads, BRICK_MIN_CTAS) k_cage_brick(const BrickArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BrickSmem m;
  brick_fill(m, smem_raw, a);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // PERSISTENT warps: the axis tables above are built once per CTA (double divisions and two barriers: 10-15 % of a
  // short-lived CTA), then every warp walks the task list of the whole batch with the stride of the launch; nothing
  // below synchronises across warps.
  const long long total = (long long)a.B * a.ntask;
  for (long long task = (long long)blockIdx.x * kBrickWarps + wid; task < total; task += (long long)gridDim.x * kBrickWarps) {
  const int b = (int)(task / a.ntask);
  // task -> (z unit, z part, y unit, y part, x slot); consecutive warps take consecutive x slots, then y
  int t = (int)(task - (long long)b * a.ntask);
  const int xs = t % a.nxs;
  t /= a.nxs;
  const int yp = t % a.ysplit;
  t /= a.ysplit;
  const int yu = t % a.nyu;
  t /= a.nyu;
  const int zp = t % a.zsplit, zu = t / a.zsplit;
  int za = m.zs[zu], zb = m.zs[zu + 1], ya = m.ys[yu], yb = m.ys[yu + 1];
  {
    const int lz = zb - za, ly = yb - ya;
    zb = za + (int)((long long)lz * (zp + 1) / a.zsplit), za += (int)((long long)lz * zp / a.zsplit);
    yb = ya + (int)((long long)ly * (yp + 1) / a.ysplit), ya += (int)((long long)ly * yp / a.ysplit);
  }
  if (za >= zb || ya >= yb) continue;

  const int x = xs * 32 + lane;
  const bool valid = x < a.oW;
  const int xc = valid ? x : a.oW - 1;
  const int kz0 = zu, kz1 = min(zu + 1, a.cD - 1), ky0 = yu, ky1 = min(yu + 1, a.cH - 1);
  const long long plane = (long long)a.cD * a.cH * a.cW;
  // per-lane state that is touched once per z lives in shared memory (registers are what limits the resident warps):
  // sP[(j, k, c)]: the cage folded along x for this lane's column, corner (y: j, z: k) of the task's cell;
  // sA[(j, k, c)]: its gradient
  float *sP = m.lane + threadIdx.x, *sA = sP + 12 * kBrickThreads;
  {
    const int ix0 = m.ix[xc], ix1 = min(ix0 + 1, a.cW - 1);
    const float wxh = m.wx[xc].y;
    const float *cb = a.cage + (long long)b * 3 * plane;
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const long long o = ((long long)(k ? kz1 : kz0) * a.cH + (j ? ky1 : ky0)) * a.cW;
#pragma unroll
        for (int c = 0; c < 3; c++) {
          // low + w_hi . (high - low): exactly `low` where the two nodes coincide (e.g. the identity cage's own axis)
          const float lo = __ldg(cb + c * plane + o + ix0), hi = __ldg(cb + c * plane + o + ix1);
          sP[((j * 2 + k) * 3 + c) * kBrickThreads] = fmaf(wxh, hi - lo, lo);
          if constexpr (MODE != 0) sA[((j * 2 + k) * 3 + c) * kBrickThreads] = 0.0f;
        }
      }
  }
  const float hs[3] = {a.hz2, a.hy2, a.hx2};
  BrickLane L;
  L.lsum = 0.0f, L.ni = 0, L.nu = 0;

  const float *vb = opaque(a.vol + (long long)b * a.vB);
  const long long osample = (long long)a.oD * a.oH * a.oW;  // < 2^31: in-sample element offsets are 32-bit
  const float *tb = (MODE != 0) ? opaque(a.tgt + (long long)b * osample + xc) : nullptr;
  const bool keep = a.out != nullptr;  // uniform: the forward value is stored (always in the forward kernel, on request in the step)
  float *ob = opaque(a.out + (keep ? (long long)b * osample + xc : 0));

  for (int z = za; z < zb; z++) {
    const float2 wz = m.wz[z];
    float C0[3], D[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      // the normalised field at the two y corners, then ((f + 1) * (size - 1)) / 2 = f * h + h: source voxel units
      const float p00 = sP[((0 * 2 + 0) * 3 + c) * kBrickThreads], p01 = sP[((0 * 2 + 1) * 3 + c) * kBrickThreads];
      const float p10 = sP[((1 * 2 + 0) * 3 + c) * kBrickThreads], p11 = sP[((1 * 2 + 1) * 3 + c) * kBrickThreads];
      const float c0 = fmaf(fmaf(wz.y, p01 - p00, p00), hs[c], hs[c]), c1 = fmaf(fmaf(wz.y, p11 - p10, p10), hs[c], hs[c]);
      C0[c] = c0, D[c] = c1 - c0;
      L.T0[c] = L.T1[c] = 0.0f;
    }
    unsigned eo = (unsigned)(z * a.oH + ya) * (unsigned)a.oW;
    int y = ya;
    constexpr int NRT = MODE == 0 ? BRICK_NR_FWD : BRICK_NR;
    for (; y + NRT <= yb; y += NRT, eo += NRT * a.oW) {
      float wyh[NRT];
#pragma unroll
      for (int r = 0; r < NRT; r++) wyh[r] = m.wy[y + r].y;
      brick_rows<MODE, GV, NRT>(a, vb, C0, D, wyh, MODE != 0 ? tb + eo : nullptr, ob + eo, keep, valid, L);
    }
    for (; y < yb; y++, eo += a.oW) {
      const float wyh[1] = {m.wy[y].y};
      brick_rows<MODE, GV, 1>(a, vb, C0, D, wyh, MODE != 0 ? tb + eo : nullptr, ob + eo, keep, valid, L);
    }
    if constexpr (MODE != 0) {
      // c = C0 + wy_hi (C1 - C0): d/dC1 = T1, d/dC0 = T0 - T1; C_j = F_j h + h and F_j = P_j0 + wz_hi (P_j1 - P_j0)
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const float d1 = L.T1[c], d0 = L.T0[c] - d1;
        float *q = sA + c * kBrickThreads;
        q[0] = fmaf(wz.x, d0, q[0]);
        q[3 * kBrickThreads] = fmaf(wz.y, d0, q[3 * kBrickThreads]);
        q[6 * kBrickThreads] = fmaf(wz.x, d1, q[6 * kBrickThreads]);
        q[9 * kBrickThreads] = fmaf(wz.y, d1, q[9 * kBrickThreads]);
      }
    }
  }

  if constexpr (MODE == 2) {
    double ls = (double)L.lsum;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) ls += __shfl_xor_sync(0xffffffffu, ls, d);
    const int ni = __reduce_add_sync(0xffffffffu, L.ni), nu = __reduce_add_sync(0xffffffffu, L.nu);
    if (lane == 0) {
      atomicAdd(a.loss + b, ls);
      if (a.iou) {
        atomicAdd(a.iou + 2 * b, (unsigned long long)ni);
        atomicAdd(a.iou + 2 * b + 1, (unsigned long long)nu);
      }
    }
  }
  if constexpr (MODE != 0) {
    // the lanes of one cage column are consecutive: segmented reduction by doubling, heads add to the cage gradient
    const int ix0 = m.ix[xc], ix1 = min(ix0 + 1, a.cW - 1);
    const float2 wx = m.wx[xc];
    bool same[5];
#pragma unroll
    for (int s = 0; s < 5; s++) {
      const int ok = __shfl_down_sync(0xffffffffu, ix0, 1 << s);
      same[s] = lane + (1 << s) < 32 && ok == ix0;
    }
    const int ix_prev = __shfl_up_sync(0xffffffffu, ix0, 1);  // every lane takes part (no short-circuit around a shuffle)
    const bool head = lane == 0 || ix_prev != ix0;
    float *gb = a.gcage + (long long)b * 3 * plane;
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int k = 0; k < 2; k++)
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const float g = sA[((j * 2 + k) * 3 + c) * kBrickThreads] * hs[c];
          float lo = g * wx.x, hi = g * wx.y;
#pragma unroll
          for (int s = 0; s < 5; s++) {
            const float l2 = __shfl_down_sync(0xffffffffu, lo, 1 << s), h2 = __shfl_down_sync(0xffffffffu, hi, 1 << s);
            if (same[s]) lo += l2, hi += h2;
          }
          if (head) {
            float *q = gb + c * plane + ((long long)(k ? kz1 : kz0) * a.cH + (j ? ky1 : ky0)) * a.cW;
            if (lo != 0.0f) red_add(q + ix0, lo);
            if (hi != 0.0f) red_add(q + ix1, hi);
          }
        }
  }
  }  // task loop
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool brick_supported(const BrickArgs &a) {
  if (a.iD < 1 || a.iH < 1 || a.iW < 1 || a.iD > (1 << 22) || a.iH > (1 << 22) || a.iW > (1 << 22)) return false;
  if ((long long)a.iD * a.iH * a.iW >= 0x7fffff00LL) return false;  // 32-bit in-sample offsets
  if (a.oD < 2 || a.oH < 2 || a.oW < 2) return false;
  if ((long long)a.oD * a.oH * a.oW >= 0x7fffff00LL) return false;
  if (a.cD < 1 || a.cH < 1 || a.cW < 1) return false;
  BrickArgs t = a;
  return brick_smem_bytes(t) <= 160 * 1024;
}

// units, splits and the fast-tier constants; `B` decides how far the cells are split (about two waves of warps)
void brick_plan(BrickArgs &a) {
  a.nzu = brick_units(a.cD), a.nyu = brick_units(a.cH), a.nxs = (a.oW + 31) / 32;
  a.zsplit = a.ysplit = 1;
  const long long want = (long long)kSMs * BRICK_MIN_CTAS * kBrickWarps * 2;
  double lz = (double)a.oD / a.nzu, ly = (double)a.oH / a.nyu;
  long long tasks = (long long)a.B * a.nzu * a.nyu * a.nxs;
  while (tasks < want && (lz / a.zsplit) * (ly / a.ysplit) >= 24.0) {
    if (lz / a.zsplit >= ly / a.ysplit)
      a.zsplit++;
    else
      a.ysplit++;
    tasks = (long long)a.B * a.nzu * a.nyu * a.nxs * a.zsplit * a.ysplit;
  }
  if (const char *e = getenv("VOLSTN_BRICK_ZSPLIT")) {  // tuning only
    const int v = atoi(e);
    if (v >= 1 && v <= 64) a.zsplit = v;
  }
  if (const char *e = getenv("VOLSTN_BRICK_YSPLIT")) {
    const int v = atoi(e);
    if (v >= 1 && v <= 64) a.ysplit = v;
  }
  a.ntask = a.nzu * a.zsplit * a.nyu * a.ysplit * a.nxs;
  a.sD = a.iH * a.iW;
  const float dm1 = (float)(a.iD - 1), hm1 = (float)(a.iH - 1), wm1 = (float)(a.iW - 1);
  a.hz2 = dm1 * 0.5f, a.hy2 = hm1 * 0.5f, a.hx2 = wm1 * 0.5f;
  memcpy(&a.fbz, &dm1, 4), memcpy(&a.fby, &hm1, 4), memcpy(&a.fbx, &wm1, 4);
  a.cmz = (float)a.iD + 1.0f, a.cmy = (float)a.iH + 1.0f, a.cmx = (float)a.iW + 1.0f;
  a.lz = (float)a.iD + 2.0f, a.ly = (float)a.iH + 2.0f, a.lx = (float)a.iW + 2.0f;
  a.hz = a.cmz, a.hy = a.cmy, a.hx = a.cmx;
  const unsigned sH = (unsigned)a.iW, sD = (unsigned)a.iH * sH;
  a.kstrict = (int)(0u - 0x4B000000u * (sD + sH + 1u));
  a.kgen = (int)(0u - 0x4B000002u * (sD + sH + 1u));
  // strict tier: -eps <= c < size - 1 with eps = size * 2^-23 (a few ulp of the coordinate), tested as
  // bits(c + eps) < bits(size - 1 + eps): rounding is monotone, so c = size - 1 itself never passes
  {
    const int sz[3] = {a.iD, a.iH, a.iW};
    float *eps[3] = {&a.ez, &a.ey, &a.ex};
    unsigned *ub[3] = {&a.ubz, &a.uby, &a.ubx}, *ui[3] = {&a.uiz, &a.uiy, &a.uix};
    float *top[3] = {&a.tz1, &a.ty1, &a.tx1};
    for (int i = 0; i < 3; i++) {
      const float e = (float)sz[i] * 1.1920929e-7f, hi = (float)(sz[i] - 1) + e;
      *eps[i] = e;
      *ub[i] = 0;  // a one-voxel axis has no cell: never strict
      if (sz[i] > 1) memcpy(ub[i], &hi, 4);
      // border tier: c + eps < size, i.e. the low corner exists (the high one may be the padding); top = size - 1
      const float sizef = (float)sz[i];
      memcpy(ui[i], &sizef, 4);
      *top[i] = (float)(sz[i] - 1);
    }
  }
}

template <int MODE, bool GV>
static int brick_launch_t(const BrickArgs &a, cudaStream_t st) {
  const size_t smem = brick_smem_bytes(a);
  if (smem > 48 * 1024)
    VOLSTN_CUDA_TRY(cudaFuncSetAttribute(k_cage_brick<MODE, GV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long need = ((long long)a.B * a.ntask + kBrickWarps - 1) / kBrickWarps;
  long long nblk = need;
  if (!getenv("VOLSTN_BRICK_NO_PERSIST")) {  // A/B timing only
    const long long resident = (long long)kSMs * BRICK_MIN_CTAS;
    if (nblk > resident) nblk = resident;
  }
  if (nblk > 0x7fffffffLL) nblk = 0x7fffffffLL;
  if (nblk < 1) return VOLSTN_OK;
  k_cage_brick<MODE, GV><<<(unsigned)nblk, kBrickThreads, smem, st>>>(a);
  return after_launch();
}

int brick_launch(const BrickArgs &a, int mode, bool grad_vol, cudaStream_t st) {
  switch (mode) {
    case 0: return brick_launch_t<0, false>(a, st);
    case 1: return grad_vol ? brick_launch_t<1, true>(a, st) : brick_launch_t<1, false>(a, st);
    default: return grad_vol ? brick_launch_t<2, true>(a, st) : brick_launch_t<2, false>(a, st);
  }
}

}  // namespace volstn
