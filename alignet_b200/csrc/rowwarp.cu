// rowwarp.cu -- ROW kernels: a warp owns whole rows of the output volume (see rowwarp.cuh).
//
//   k_row_fwd<SRC, AOS4, C1, TWO_ROWS>            out = sample(vol, coords(b, z, y, x))
//   k_row_bwd<SRC, AOS4, C1, TWO_ROWS, OG, OV>    gradVol scatter-add + the gradient of the coordinate source
//
// SRC selects where the sampling coordinates come from and where their gradient goes:
//   ROW_SRC_GRID    a grid tensor with arbitrary B/D/H strides, `... x G` or planar   -> gradGrid tensor
//   ROW_SRC_AFFINE  T[b] . (zn, yn, xn, 1) computed in registers                       -> gradT (two-stage, deterministic)
//   ROW_SRC_CAGE    trilinear up-sample of a cD x cH x cW x 3 control cage held in     -> gradCage (privatised in shared
//                   shared memory, evaluated at the identity coordinates                  memory, flushed with red.global)
// The voxel arithmetic is the one of sampler.cuh (VoxelSetup tiers, corner weights, grad_grid_from_dots), so
// `out` and gradGrid have the bits of the reference CPU code; the AFFINE / CAGE coordinates are computed with the
// operation order of k_gridgen_fwd and of the sampler itself, so a fused call has the bits of the un-fused
// module chain (tests/test_gpu_rowwarp.py).
#include <cstdlib>
#include <cstring>

#include "rowwarp.cuh"
#include "volstn_internal.h"

namespace volstn {

constexpr int kRowThreads = 256;
constexpr int kRowWarps = kRowThreads / 32;

// (float)(-1 + k/(n-1)*2) in double: VolumetricGridGenerator.lua:17-19 (same as gridgen.cu::base_coord)
__device__ __forceinline__ float row_base_coord(int k, int n) {
  return (float)(-1.0 + (double)k / (double)(n - 1) * 2.0);
}

// ---------------------------------------------------------------------------------------------
// shared memory
// ---------------------------------------------------------------------------------------------
// Cage cells in shared memory as (z, y) pairs plus x values instead of 16-byte (z, y, x, 1) cells: a 16-byte shared load of a
// warp is four wavefronts whatever the lanes' addresses, an 8-byte one two, a 4-byte one one -- three per corner instead
// of four.  The L1/LSU pipe is what bounds the cage kernels (87 % busy): forward 198 -> 186 us, full backward 556 -> 537 us
// at 64^3 x 64 (profiles/r01_fbench_history.txt).
#ifndef ROW_CAGE_SPLIT
#define ROW_CAGE_SPLIT 1
#endif

struct RowSmem {
  // AFFINE: normalised coordinates of the three axes.  CAGE: low-corner index and weight per axis position.
  float *tz, *ty, *tx;
  int *iz, *iy, *ix;
  float4 *cg;    // CAGE: (cD+1) x (cH+1) x (cW+1) cells (z, y, x, 1), one layer of zeros on every high face
  float2 *cgzy;  // ROW_CAGE_SPLIT: the same cells as (z, y) pairs ...
  float *cgx;    // ... and x values (8 + 4 byte loads: 3 shared-memory wavefronts per corner instead of 4)
  float *gc;     // CAGE backward: gradient of the cage, same cells x 3
  int *xs;       // CAGE backward: xs[t] = first x whose low cage cell is >= t  (t = 0 .. cW)
  float *stage;  // CAGE backward: per warp, per slot: 6 x oWp products (wx * gg[c], (1 - wx) * gg[c])
  float *nodes;  // CAGE, fast arithmetic: per warp, per row of the trip: nstride (z, y) pairs, then nstride x values (cage_nodes)
  int4 *nitem;   // CAGE, fast arithmetic: the (node, component) item lane `i` (and i + 32) computes for every row
};

__host__ __device__ inline int row_cells(const RowArgs &a) { return (a.cD + 1) * (a.cH + 1) * (a.cW + 1); }
// staging row length: a multiple of 32 plus 8, so that the three channels of a cage cell (rows c and 3 + c of the
// stage) fall into different shared-memory banks when the item lanes sum their segments
__host__ __device__ inline int row_owp(const RowArgs &a) { return ((a.s.oW + 31) & ~31) + 8; }
// fast arithmetic: floats per component of a row's node vector, EVEN so that the (z, y) pairs of the second row of a trip
// (3 * nstride floats further) stay 8-byte aligned (an odd cage width faulted with `misaligned address`)
__host__ __device__ inline int row_nstride(const RowArgs &a) { return (a.cW + 3) & ~1; }

template <int SRC, bool BWD>
__host__ __device__ inline size_t row_smem_bytes(const RowArgs &a, bool two_rows) {
  const size_t axes = (size_t)a.s.oD + a.s.oH + a.s.oW;
  if (SRC == ROW_SRC_AFFINE) return axes * 4;
  if (SRC == ROW_SRC_CAGE) {
    size_t n = (size_t)row_cells(a) * 16 + axes * 8 + 8 + 64 * 16 + (size_t)kRowWarps * 2 * row_nstride(a) * 12;
    if (BWD) n += (size_t)row_cells(a) * 12 + (size_t)(a.cW + 2) * 4 + (size_t)kRowWarps * (two_rows ? 2 : 1) * 6 * row_owp(a) * 4;
    return n;
  }
  return 0;
}

template <int SRC, bool BWD>
__device__ __forceinline__ void row_carve(RowSmem &m, unsigned char *raw, const RowArgs &a) {
  const int oD = a.s.oD, oH = a.s.oH, oW = a.s.oW;
  if constexpr (SRC == ROW_SRC_AFFINE) {
    m.tz = reinterpret_cast<float *>(raw);
    m.ty = m.tz + oD;
    m.tx = m.ty + oH;
  } else if constexpr (SRC == ROW_SRC_CAGE) {
    m.cg = reinterpret_cast<float4 *>(raw);
    m.cgzy = reinterpret_cast<float2 *>(raw);
    m.cgx = reinterpret_cast<float *>(m.cgzy + row_cells(a));
    m.tz = reinterpret_cast<float *>(m.cg + row_cells(a));
    m.ty = m.tz + oD;
    m.tx = m.ty + oH;
    m.iz = reinterpret_cast<int *>(m.tx + oW);
    m.iy = m.iz + oD;
    m.ix = m.iy + oH;
    m.nitem = reinterpret_cast<int4 *>(reinterpret_cast<unsigned char *>(m.ix + oW) + ((oD + oH + oW) & 1) * 8);  // 16-byte aligned
    m.nodes = reinterpret_cast<float *>(m.nitem + 64);
    if constexpr (BWD) {
      m.gc = m.nodes + kRowWarps * 2 * 3 * row_nstride(a);
      m.xs = reinterpret_cast<int *>(m.gc + 3 * row_cells(a));
      m.stage = reinterpret_cast<float *>(m.xs + a.cW + 2);
    }
  }
}

template <int SRC, bool BWD>
__device__ __forceinline__ void row_fill(const RowSmem &m, const RowArgs &a, int b) {
  const int oD = a.s.oD, oH = a.s.oH, oW = a.s.oW;
  if constexpr (SRC == ROW_SRC_AFFINE) {
    for (int i = threadIdx.x; i < oD + oH + oW; i += blockDim.x) {
      if (i < oD)
        m.tz[i] = row_base_coord(i, oD);
      else if (i < oD + oH)
        m.ty[i - oD] = row_base_coord(i - oD, oH);
      else
        m.tx[i - oD - oH] = row_base_coord(i - oD - oH, oW);
    }
  } else if constexpr (SRC == ROW_SRC_CAGE) {
    // the cage of sample b, channels-last with the constant 4th component of the un-fused cage volume
    const int cHp = a.cH + 1, cWp = a.cW + 1;
    const long long plane = (long long)a.cD * a.cH * a.cW;
    const float *cb = a.cage + (long long)b * 3 * plane;
    for (int i = threadIdx.x; i < row_cells(a); i += blockDim.x) {
      const int x = i % cWp, q = i / cWp, y = q % cHp, z = q / cHp;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (x < a.cW && y < a.cH && z < a.cD) {
        const long long o = ((long long)z * a.cH + y) * a.cW + x;
        v = make_float4(__ldg(cb + o), __ldg(cb + plane + o), __ldg(cb + 2 * plane + o), 1.0f);
      }
#if ROW_CAGE_SPLIT
      m.cgzy[i] = make_float2(v.x, v.y);
      m.cgx[i] = v.z;
#else
      m.cg[i] = v;
#endif
      if constexpr (BWD) m.gc[3 * i] = m.gc[3 * i + 1] = m.gc[3 * i + 2] = 0.f;
    }
    // the sampler's own set-up (...c:507-517) of the identity coordinate of every axis position: with the identity
    // field the low corner and its weight depend on one axis index only
    for (int i = threadIdx.x; i < oD + oH + oW; i += blockDim.x) {
      int k, n, cn;
      if (i < oD)
        k = i, n = oD, cn = a.cD;
      else if (i < oD + oH)
        k = i - oD, n = oH, cn = a.cH;
      else
        k = i - oD - oH, n = oW, cn = a.cW;
      int i0;
      float w;
      axis_setup<float>(row_base_coord(k, n), (float)(cn - 1), i0, w);
      m.tz[i] = w;  // tz / ty / tx and iz / iy / ix are consecutive
      m.iz[i] = i0;
    }
    // fast arithmetic: item i = (component c, x-node t) of a row's node vector; where its four cage values live in the
    // cage's shared-memory copy viewed as floats ((z, y) pairs first, then the x plane) and where the node goes
    if (threadIdx.x < 64) {
      const int cWp = a.cW + 1, it = threadIdx.x;
      const int c = it / cWp, t = it - c * cWp;
      int4 e = make_int4(-1, 0, 0, 0);
      if (c < 3) e = c < 2 ? make_int4(t, c, 2, 2 * t + c) : make_int4(t, 2 * row_cells(a), 1, 2 * row_nstride(a) + t);
      m.nitem[it] = e;
    }
    if constexpr (BWD) {
      __syncthreads();
      if (threadIdx.x == 0) {
        int t = 0;
        for (int x = 0; x < oW; x++)
          while (t <= m.ix[x] && t <= a.cW) m.xs[t++] = x;
        while (t <= a.cW + 1) m.xs[t++] = oW;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// coordinate sources
// ---------------------------------------------------------------------------------------------
// ((T0*zn + T1*yn) + T2*xn) + T3 -- the operation order of k_gridgen_fwd
__device__ __forceinline__ float affine_row(const float (&t)[12], int i, float zn, float yn, float xn) {
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t[4 * i], zn), __fmul_rn(t[4 * i + 1], yn)), __fmul_rn(t[4 * i + 2], xn)),
                   t[4 * i + 3]);
}

// the cage volume sampled at the identity coordinate (z, y, x): BilinearSamplerVolumetric forward with C = 4
// (...c:554-576), corner weights ((wx'.wy').wz'), corners summed in the order 0..7; cells beyond the high face
// are the zeros the sampler substitutes for out-of-bounds corners
__device__ __forceinline__ void cage_field(const RowSmem &m, const RowArgs &a, int z, int y, int x, float &zf, float &yf,
                                           float &xf) {
  const int cWp = a.cW + 1, cHWp = (a.cH + 1) * cWp;
  const float wx = m.tx[x], wy = m.ty[y], wz = m.tz[z];
  const float wxs[2] = {wx, __fsub_rn(1.0f, wx)};
  const float wys[2] = {wy, __fsub_rn(1.0f, wy)};
  const float wzs[2] = {wz, __fsub_rn(1.0f, wz)};
  const int cbase = (m.iz[z] * (a.cH + 1) + m.iy[y]) * cWp + m.ix[x];
  float az = 0.f, ay = 0.f, ax = 0.f;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const float w = __fmul_rn(__fmul_rn(wxs[k & 1], wys[(k >> 1) & 1]), wzs[k >> 2]);
    const int ci = cbase + (k & 1) + ((k >> 1) & 1) * cWp + (k >> 2) * cHWp;
#if ROW_CAGE_SPLIT
    const float2 vzy = m.cgzy[ci];
    const float vx = m.cgx[ci];
    const float pz = __fmul_rn(w, vzy.x), py = __fmul_rn(w, vzy.y), px = __fmul_rn(w, vx);
#else
    const float4 v = m.cg[ci];
    const float pz = __fmul_rn(w, v.x), py = __fmul_rn(w, v.y), px = __fmul_rn(w, v.z);
#endif
    az = k ? __fadd_rn(az, pz) : pz;
    ay = k ? __fadd_rn(ay, py) : py;
    ax = k ? __fadd_rn(ax, px) : px;
  }
  zf = az, yf = ay, xf = ax;
}

// ---- fast arithmetic (VOLSTN_FAST_ARITH) ------------------------------------------------------------------------
// Along an output row (z, y fixed) the up-sampled field is piecewise LINEAR in x: inside cage x-cell t it is
//   wx * N[t] + (1 - wx) * N[t + 1],     N[t] = sum over the row's four (y, z) cage corners of (wy' * wz') * cage[.., .., t]
// so the eight 12-byte corner loads and 24 multiplies per voxel of cage_field() become, per ROW, one node vector N
// (cW + 1 nodes x 3 components, one (node, component) item per lane) and, per VOXEL, two node loads and three lerps.
// Same mathematics, different rounding: the field agrees with cage_field() to a few ulp (tests: <= 1e-6 absolute on
// coordinates in [-1, 1]), NOT bit for bit -- hence a flag, off by default.
// nodes of output row (z, y) into `nbuf` (one row's 3 * nstride floats of this warp's buffer)
__device__ __forceinline__ void cage_nodes(const RowSmem &m, const RowArgs &a, float *nbuf, int z, int y, int lane) {
  const int cWp = a.cW + 1, cHWp = (a.cH + 1) * cWp;
  const float wy = m.ty[y], wz = m.tz[z];
  const float wy1 = 1.0f - wy, wz1 = 1.0f - wz;
  const float w00 = wy * wz, w10 = wy1 * wz, w01 = wy * wz1, w11 = wy1 * wz1;  // (y', z')
  const int cbase = (m.iz[z] * (a.cH + 1) + m.iy[y]) * cWp;
  const float *cf = reinterpret_cast<const float *>(m.cgzy);
  const int items = 3 * cWp;
#pragma unroll
  for (int i = 0; i < 2; i++) {
    if (lane + 32 * i < items) {
      const int4 e = m.nitem[lane + 32 * i];  // (t, source base, source stride, destination)
      const float *p = cf + e.y + (cbase + e.x) * e.z;
      const int sy = cWp * e.z, sz = cHWp * e.z;
      float v = w00 * p[0];
      v = fmaf(w10, p[sy], v);
      v = fmaf(w01, p[sz], v);
      v = fmaf(w11, p[sz + sy], v);
      nbuf[e.w] = v;
    }
  }
}
__device__ __forceinline__ void cage_field_sep(const RowSmem &m, const float *nbuf, int nstride, int x, float &zf, float &yf,
                                               float &xf) {
  const float wx = m.tx[x], wx1 = 1.0f - wx;
  const int i = m.ix[x];
  const float2 *nzy = reinterpret_cast<const float2 *>(nbuf);
  const float *nx = nbuf + 2 * nstride;
  const float2 a0 = nzy[i], a1 = nzy[i + 1];
  const float b0 = nx[i], b1 = nx[i + 1];
  zf = fmaf(wx, a0.x, wx1 * a1.x);
  yf = fmaf(wx, a0.y, wx1 * a1.y);
  xf = fmaf(wx, b0, wx1 * b1);
}

// (Keeping the lane's eight cage corners in registers across rows was measured and rejected: 48 more registers spill or
// halve the occupancy -- forward 192-270 us against 199 us, profiles/r01_fbench_history.txt.  The brick kernels of
// cagebrick.cu get the same effect from their mapping.)
template <int SRC, bool AOS4, bool FASTM = false>
struct RowCoords {
  // per-CTA state of the source
  float t[12];
  const float *gb;
  // fast arithmetic: this warp's node buffer (two rows of 3 * nstride floats)
  float *nbuf;
  int nstride, nlane;
  __device__ __forceinline__ void init_nodes(const RowSmem &m, const RowArgs &a, int wid, int lane) {
    if constexpr (SRC == ROW_SRC_CAGE && FASTM) {
      nstride = row_nstride(a);
      nbuf = m.nodes + wid * 2 * 3 * nstride;
      nlane = lane;
    }
  }
  // the nodes of the trip's rows: (z, y) and, for a two-row trip, the next row.  Bracketed by __syncwarp(): the lanes
  // that wrote a node are not the ones that read it.
  template <bool TWO_ROWS>
  __device__ __forceinline__ void row_begin(const RowSmem &m, const RowArgs &a, int z, int y, bool second_row) {
    if constexpr (SRC == ROW_SRC_CAGE && FASTM) {
      __syncwarp();
      cage_nodes(m, a, nbuf, z, y, nlane);
      if constexpr (TWO_ROWS) {
        if (second_row) {
          const bool wrap = y + 1 == a.s.oH;
          cage_nodes(m, a, nbuf + 3 * nstride, wrap ? z + 1 : z, wrap ? 0 : y + 1, nlane);
        }
      }
      __syncwarp();
    }
  }
  __device__ __forceinline__ void init(const RowArgs &a, int b) {
    if constexpr (SRC == ROW_SRC_AFFINE) {
#pragma unroll
      for (int i = 0; i < 12; i++) t[i] = __ldg(a.T + (long long)b * 16 + i);
    }
    if constexpr (SRC == ROW_SRC_GRID) gb = opaque(a.s.grid + (long long)b * a.gB);
  }
  template <int V>  // V: slot of the trip (compile time: the corner registers of the two slots are distinct)
  __device__ __forceinline__ void get(const RowSmem &m, const RowArgs &a, int z, int y, int x, float &zf, float &yf,
                                      float &xf) {
    if constexpr (SRC == ROW_SRC_GRID) {
      const float *g = gb + z * a.gD + y * a.gH + x * a.gW;
      if constexpr (AOS4) {
        const float4 v = ld_stream_f4(reinterpret_cast<const float4 *>(g));
        zf = v.x, yf = v.y, xf = v.z;
      } else {
        zf = ld_stream_f1(g), yf = ld_stream_f1(g + a.gC), xf = ld_stream_f1(g + 2 * a.gC);
      }
    } else if constexpr (SRC == ROW_SRC_AFFINE) {
      const float zn = m.tz[z], yn = m.ty[y], xn = m.tx[x];
      zf = affine_row(t, 0, zn, yn, xn);
      yf = affine_row(t, 1, zn, yn, xn);
      xf = affine_row(t, 2, zn, yn, xn);
    } else if constexpr (FASTM) {
      // V = slot of the trip: with two rows per trip slot 1 is the second row (clamped to the first at the end)
      cage_field_sep(m, nbuf + ((V && second_slot_row) ? 3 * nstride : 0), nstride, x, zf, yf, xf);
    } else {
      cage_field(m, a, z, y, x, zf, yf, xf);
    }
  }
  bool second_slot_row = false;  // set per trip by the kernels (TWO_ROWS and the second row exists)
};

// ---------------------------------------------------------------------------------------------
// voxel bodies with runtime strides (x stride sW, channel stride sC)
// ---------------------------------------------------------------------------------------------
template <int ORDER>  // 0: corners 0..7 (...c:566-573), 1: 0,4,1,5,2,6,3,7 (...c:149-156)
__device__ __forceinline__ float sum8(const float (&cw)[8], const float (&v)[8]) {
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int k = ORDER ? order_zxy(j) : order_xyz(j);
    const float term = __fmul_rn(cw[k], v[k]);
    acc = j ? __fadd_rn(acc, term) : term;
  }
  return acc;
}

// fast arithmetic: the same eight terms as one FMA chain
__device__ __forceinline__ float sum8_fma(const float (&cw)[8], const float (&v)[8]) {
  float acc = cw[0] * v[0];
#pragma unroll
  for (int k = 1; k < 8; k++) acc = fmaf(cw[k], v[k], acc);
  return acc;
}

// fast arithmetic: gradGrid factored per axis -- d/dx = sum over the four (y', z') of (wy' wz') (dot[x+1] - dot[x]) etc. --
// 12 weight products, 12 differences, 12 FMAs instead of the reference's 40 products and 21 sums (...c:792-817)
__device__ __forceinline__ void grad_grid_fast(const float (&dot)[8], const VoxelSetup<float> &s, const SamplerArgs<float> &a,
                                               float &gz, float &gy, float &gx) {
  const float wx1 = 1.0f - s.wx, wy1 = 1.0f - s.wy, wz1 = 1.0f - s.wz;
  // corner k: x = k & 1, y = (k >> 1) & 1, z = k >> 2
  const float yz[4] = {s.wy * s.wz, wy1 * s.wz, s.wy * wz1, wy1 * wz1};  // index y + 2 z
  const float xz[4] = {s.wx * s.wz, wx1 * s.wz, s.wx * wz1, wx1 * wz1};  // index x + 2 z
  const float xy[4] = {s.wx * s.wy, wx1 * s.wy, s.wx * wy1, wx1 * wy1};  // index x + 2 y
  float dx = yz[0] * (dot[1] - dot[0]);
  dx = fmaf(yz[1], dot[3] - dot[2], dx);
  dx = fmaf(yz[2], dot[5] - dot[4], dx);
  dx = fmaf(yz[3], dot[7] - dot[6], dx);
  float dy = xz[0] * (dot[2] - dot[0]);
  dy = fmaf(xz[1], dot[3] - dot[1], dy);
  dy = fmaf(xz[2], dot[6] - dot[4], dy);
  dy = fmaf(xz[3], dot[7] - dot[5], dy);
  float dz = xy[0] * (dot[4] - dot[0]);
  dz = fmaf(xy[1], dot[5] - dot[1], dz);
  dz = fmaf(xy[2], dot[6] - dot[2], dz);
  dz = fmaf(xy[3], dot[7] - dot[3], dz);
  gz = dz * (a.dm1 * 0.5f), gy = dy * (a.hm1 * 0.5f), gx = dx * (a.wm1 * 0.5f);
}

template <bool C1, int MODE, bool FASTM = false>
__device__ __forceinline__ void row_voxel_forward(const RowArgs &a, const VoxelSetup<float> &s, const float *__restrict__ vb,
                                                  float *__restrict__ o, bool valid) {
  constexpr bool FAST = MODE == 1;
  const int sW = C1 ? 1 : a.vW;
  float cw[8];
  s.corner_weights(cw);
  const CornerRows<float> rows(vb, MODE ? s.fbase : s.base32(a.vD, a.vH, sW), a.vD, a.vH);
  const int C = C1 ? 1 : a.s.C;
  for (int t = 0; t < C; t++) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; k++)
      v[k] = (FAST || s.in(k)) ? __ldg(rows.p[k >> 1] + ((k & 1) ? sW : 0) + (C1 ? 0 : t * a.vC)) : 0.f;
    const float r = FASTM ? sum8_fma(cw, v) : (a.g3 ? sum8<1>(cw, v) : sum8<0>(cw, v));
    if (valid) st_stream_f1(o + (C1 ? 0 : t * a.osC), r);
  }
}

// returns the gradient of the three coordinates (before any sink)
// LOSS (single channel, fused training step): `gop` points at the TARGET voxel; the forward value is formed from the
// gathers that the backward needs anyway, gradOut is the SmoothL1 gradient norm * clamp(out - target, -1, 1)
// (THNN SmoothL1Criterion_updateGradInput) and (out, target) are handed back for the loss / IoU accumulators.
template <bool C1, bool OG, bool OV, int MODE, bool LOSS = false, bool FASTM = false>
__device__ __forceinline__ void row_voxel_backward(const RowArgs &a, const VoxelSetup<float> &s, const float *__restrict__ vb,
                                                   long long gv_delta, const float *__restrict__ gop, bool valid,
                                                   float &gz, float &gy, float &gx, float *fwd = nullptr,
                                                   float *tgt = nullptr) {
  static_assert(!FASTM || C1, "fast arithmetic: single-channel cage kernels only");
  static_assert(!LOSS || (C1 && !OV), "the fused step is single-channel and needs the gathers");
  constexpr bool FAST = MODE == 1;
  const int sW = C1 ? 1 : a.vW;
  float cw[8];
  if constexpr (!OG || LOSS) s.corner_weights(cw);
  const CornerRows<float> rows(vb, MODE ? s.fbase : s.base32(a.vD, a.vH, sW), a.vD, a.vH);
  float dot[8];
#pragma unroll
  for (int k = 0; k < 8; k++) dot[k] = 0.f;
  const int C = C1 ? 1 : a.s.C;
  for (int t = 0; t < C; t++) {
    float go = ld_stream_f1(gop + (C1 ? 0 : t * a.osC));
    const int co = C1 ? 0 : t * a.vC;
    if constexpr (!OV) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; k++) v[k] = (FAST || s.in(k)) ? __ldg(rows.p[k >> 1] + ((k & 1) ? sW : 0) + co) : 0.f;
      if constexpr (LOSS) {
        // the forward value: bit for bit (...c:566-573 / 149-156), or one FMA chain with fast arithmetic
        const float o = FASTM ? sum8_fma(cw, v) : (a.g3 ? sum8<1>(cw, v) : sum8<0>(cw, v));
        const float x = __fsub_rn(o, go);
        *fwd = o, *tgt = go;
        go = x < -1.0f ? -a.loss_norm : (x > 1.0f ? a.loss_norm : __fmul_rn(a.loss_norm, x));
      }
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const float d = FASTM ? v[k] * go : __fadd_rn(dot[k], __fmul_rn(v[k], go));  // dot += v * go (...c:736)
        dot[k] = (FAST || s.in(k)) ? d : dot[k];                 // an out-of-bounds corner never sees go
      }
    }
    if constexpr (!OG) {
#pragma unroll
      for (int k = 0; k < 8; k++)
        if (valid && (FAST || s.in(k)))
          red_add(const_cast<float *>(rows.p[k >> 1]) + ((k & 1) ? sW : 0) + co + gv_delta, __fmul_rn(cw[k], go));  // ...c:737
    }
  }
  if constexpr (!OV) {
    if constexpr (FASTM)
      grad_grid_fast(dot, s, a.s, gz, gy, gx);
    else
      grad_grid_from_dots<float>(dot, s, a.s, gz, gy, gx);
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
struct RowSlot {
  int z, y, x;  // x clamped to the row
  bool valid;
};

// (z, y) of the warp's current row is carried along (one divide per warp, not per row)
template <bool TWO_ROWS>
__device__ __forceinline__ void row_slots(RowSlot (&sl)[2], int r, int r_end, int z, int y, int xb, int lane, int oH,
                                          int oW) {
#pragma unroll
  for (int v = 0; v < 2; v++) {
    const bool second = TWO_ROWS && v == 1 && r + 1 < r_end;  // the second row of the trip (clamped to the first at the end)
    const bool wrap = second && y + 1 == oH;
    sl[v].z = wrap ? z + 1 : z;
    sl[v].y = second ? (wrap ? 0 : y + 1) : y;
    const int x = xb + lane + (TWO_ROWS ? 0 : 32 * v);
    sl[v].valid = (!TWO_ROWS || v == 0 || r + 1 < r_end) && x < oW;
    sl[v].x = min(x, oW - 1);
  }
}
__device__ __forceinline__ void row_advance(int &z, int &y, int step, int oH) {
  y += step;
  if (y >= oH) y -= oH, z++;  // step <= 2 <= oH
}

// Resident CTAs per SM the kernels are compiled for (a register cap of 48 / 64).  Measured on a B200 with every
// value from 1 to 6 (profiles/r02_fbench_history.txt): without a cap ptxas takes 56-128 registers and occupancy
// falls to 24-45 %; caps of 5-6 spill in the backward kernels and lose more than the occupancy gains.
#ifndef ROW_MIN_CTAS
#define ROW_MIN_CTAS 0
#endif
// fast-arithmetic cage backward kernels: 1 = the 80-register build (3 resident CTAs), 0 = 64 registers (4 CTAs)
#ifndef ROW_FAST_HEAVY
#define ROW_FAST_HEAVY 1
#endif
#ifndef ROW_MIN_CTAS_HEAVY
#define ROW_MIN_CTAS_HEAVY 3
#endif
constexpr int row_min_ctas(int src, bool bwd, bool c1, bool heavy = false) {
  if (ROW_MIN_CTAS) return ROW_MIN_CTAS;
  if (!c1) return 3;
  // cage backward WITH the gradVol scatter (552 us at 3, 618 us at 4, 577 us at 2) and the affine backward with its
  // twelve gradT accumulators (277 us at 3, 287 us at 4, 306 us at 2)
  if (heavy) return ROW_MIN_CTAS_HEAVY;
  return (src == ROW_SRC_GRID && !bwd) ? 5 : 4;
}

template <int SRC, bool AOS4, bool C1, bool TWO_ROWS, bool FASTM = false>
__global__ void __launch_bounds__(kRowThreads, row_min_ctas(SRC, false, C1)) k_row_fwd(const RowArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SamplerArgs<float> &s = a.s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int b = blockIdx.y;
  RowSmem m;
  row_carve<SRC, false>(m, smem_raw, a);
  row_fill<SRC, false>(m, a, b);
  if constexpr (SRC != ROW_SRC_GRID) __syncthreads();
  RowCoords<SRC, AOS4, FASTM> src;
  src.init(a, b);
  src.init_nodes(m, a, wid, lane);
  const float *vb = opaque(s.vol + (long long)b * a.vB);
  float *ob = opaque(s.out + (long long)b * a.oB);
  const int sW = C1 ? 1 : a.vW;
  int r = (blockIdx.x * kRowWarps + wid) * a.rows_per_warp;
  const int r_end = min(r + a.rows_per_warp, s.oD * s.oH);
  int z = r / s.oH, y = r - z * s.oH;
  for (; r < r_end; r += TWO_ROWS ? 2 : 1, row_advance(z, y, TWO_ROWS ? 2 : 1, s.oH)) {
    src.second_slot_row = TWO_ROWS && r + 1 < r_end;
    src.template row_begin<TWO_ROWS>(m, a, z, y, src.second_slot_row);
    for (int xb = 0; xb < s.oW; xb += TWO_ROWS ? 32 : 64) {
      RowSlot sl[2];
      row_slots<TWO_ROWS>(sl, r, r_end, z, y, xb, lane, s.oH, s.oW);
      VoxelSetup<float> su[2];
      int tier = 2;
      {
        float zf, yf, xf;
        src.template get<0>(m, a, sl[0].z, sl[0].y, sl[0].x, zf, yf, xf);
        tier = min(tier, su[0].begin(zf, yf, xf, s));
        src.template get<1>(m, a, sl[1].z, sl[1].y, sl[1].x, zf, yf, xf);
        tier = min(tier, su[1].begin(zf, yf, xf, s));
      }
      float *o[2];
#pragma unroll
      for (int v = 0; v < 2; v++) o[v] = ob + sl[v].z * a.osD + sl[v].y * a.osH + sl[v].x * a.osW;
      if (__all_sync(0xffffffffu, tier == 2)) {
#pragma unroll
        for (int v = 0; v < 2; v++) su[v].template finish_fast<false>(s, a.vD, a.vH, sW);
#pragma unroll
        for (int v = 0; v < 2; v++) row_voxel_forward<C1, 1, FASTM>(a, su[v], vb, o[v], sl[v].valid);
      } else if (__all_sync(0xffffffffu, tier >= 1)) {
#pragma unroll
        for (int v = 0; v < 2; v++) su[v].template finish_fast<true>(s, a.vD, a.vH, sW);
#pragma unroll
        for (int v = 0; v < 2; v++) row_voxel_forward<C1, 2, FASTM>(a, su[v], vb, o[v], sl[v].valid);
      } else {
#pragma unroll
        for (int v = 0; v < 2; v++) {
          su[v].finish_exact(s);
          row_voxel_forward<C1, 0, FASTM>(a, su[v], vb, o[v], sl[v].valid);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// CAGE sink, one output row: the lanes have left wx*gg[c] and (1-wx)*gg[c] of their voxels in `st`; lane `it` owns
// (cage x cell t, channel c) = (it / 3, it % 3), sums the row's contributions to that cell and scales them by
// the four (y, z) corner weights of the row into its register accumulators, which are flushed to the
// shared-memory cage gradient only when the row's cage cell (z0, y0) changes.
struct CageAcc {
  float acc[2][4];
  int z0, y0;  // cell of the rows accumulated so far (-1: nothing yet)
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int q = 0; q < 4; q++) acc[i][q] = 0.f;
  }
  __device__ __forceinline__ void flush(const RowSmem &m, const RowArgs &a, int lane) {
    if (z0 < 0) return;
    const int cWp = a.cW + 1, cHp = a.cH + 1, items = 3 * cWp;
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const int it = lane + 32 * i;
      if (it < items) {
        const int t = it / 3, c = it - 3 * t;
#pragma unroll
        for (int q = 0; q < 4; q++)  // q: bit0 = +y, bit1 = +z
          atomicAdd(m.gc + 3 * (((z0 + (q >> 1)) * cHp + y0 + (q & 1)) * cWp + t) + c, acc[i][q]);
      }
    }
    reset();
  }
  __device__ __forceinline__ void row(const RowSmem &m, const RowArgs &a, const float *st, int owp, int z, int y, int lane) {
    const int rz0 = m.iz[z], ry0 = m.iy[y];
    if (rz0 != z0 || ry0 != y0) {
      flush(m, a, lane);
      z0 = rz0, y0 = ry0;
    }
    const float wy = m.ty[y], wz = m.tz[z];
    const float wys[2] = {wy, 1.0f - wy}, wzs[2] = {wz, 1.0f - wz};
    const int items = 3 * (a.cW + 1);
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const int it = lane + 32 * i;
      if (it < items) {
        const int t = it / 3, c = it - 3 * t;
        float A = 0.f;
        if (t < a.cW)
          for (int x = m.xs[t]; x < m.xs[t + 1]; x++) A += st[c * owp + x];
        if (t > 0)
          for (int x = m.xs[t - 1]; x < m.xs[t]; x++) A += st[(3 + c) * owp + x];
#pragma unroll
        for (int q = 0; q < 4; q++) acc[i][q] = fmaf(A, wys[q & 1] * wzs[q >> 1], acc[i][q]);
      }
    }
  }
};

// HEAVY: the 80-register build of the cage kernels that skip the gradVol scatter (OG).  At 64^3 with an 8^3 cage the
// 64-register build is 2 % faster (365 vs 373 us); with rows of more than 64 voxels (two trips per row) or more than
// 32 (cell, channel) items per row it spills: 615 vs 355 us at 128^3 x 8, 606 vs 547 us with a 12^3 cage.  The launcher picks.
template <int SRC, bool AOS4, bool C1, bool TWO_ROWS, bool OG, bool OV, bool LOSS = false, bool HEAVY = false, bool FASTM = false>
__global__ void __launch_bounds__(kRowThreads, row_min_ctas(SRC, true, C1, (SRC == ROW_SRC_CAGE && (!OG || HEAVY || (FASTM && ROW_FAST_HEAVY))) || SRC == ROW_SRC_AFFINE))
    k_row_bwd(const RowArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[kRowWarps][12];
  const SamplerArgs<float> &s = a.s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int b = blockIdx.y;
  RowSmem m;
  row_carve<SRC, true>(m, smem_raw, a);
  row_fill<SRC, true>(m, a, b);
  if constexpr (SRC != ROW_SRC_GRID) __syncthreads();
  RowCoords<SRC, AOS4, FASTM> src;
  src.init(a, b);
  src.init_nodes(m, a, wid, lane);
  const float *vb = opaque(s.vol + (long long)b * a.vB);
  const long long gv_delta = s.gvol - s.vol;  // gradVol has the volume's strides
  const float *gob = opaque(s.gout + (long long)b * a.oB);
  float *ggb = (SRC == ROW_SRC_GRID && !OV) ? opaque(s.ggrid + (long long)b * a.ggB) : nullptr;
  const int sW = C1 ? 1 : a.vW;
  constexpr bool SINK = !OV;
  float tacc[12];  // AFFINE: gradT rows 0..2 of this lane
#pragma unroll
  for (int i = 0; i < 12; i++) tacc[i] = 0.f;
  CageAcc cacc;
  cacc.z0 = cacc.y0 = -1;
  cacc.reset();
  const int owp = row_owp(a);
  float *st = (SRC == ROW_SRC_CAGE && SINK) ? m.stage + wid * (TWO_ROWS ? 2 : 1) * 6 * owp : nullptr;
  // fused step: SmoothL1 sum of this lane (THNN accumulates in double) and the IoU counts of this warp
  double lsum = 0.0;
  unsigned n_inter = 0, n_union = 0;
  float *outb = (LOSS && s.out) ? opaque(s.out + (long long)b * a.oB) : nullptr;

  int r = (blockIdx.x * kRowWarps + wid) * a.rows_per_warp;
  const int r_end = min(r + a.rows_per_warp, s.oD * s.oH);
  int z = r / s.oH, y = r - z * s.oH;
  for (; r < r_end; r += TWO_ROWS ? 2 : 1, row_advance(z, y, TWO_ROWS ? 2 : 1, s.oH)) {
    src.second_slot_row = TWO_ROWS && r + 1 < r_end;
    src.template row_begin<TWO_ROWS>(m, a, z, y, src.second_slot_row);
    for (int xb = 0; xb < s.oW; xb += TWO_ROWS ? 32 : 64) {
      RowSlot sl[2];
      row_slots<TWO_ROWS>(sl, r, r_end, z, y, xb, lane, s.oH, s.oW);
      VoxelSetup<float> su[2];
      int tier = 2;
      {
        float zf, yf, xf;
        src.template get<0>(m, a, sl[0].z, sl[0].y, sl[0].x, zf, yf, xf);
        tier = min(tier, su[0].begin(zf, yf, xf, s));
        src.template get<1>(m, a, sl[1].z, sl[1].y, sl[1].x, zf, yf, xf);
        tier = min(tier, su[1].begin(zf, yf, xf, s));
      }
      const float *gop[2];
#pragma unroll
      for (int v = 0; v < 2; v++) gop[v] = gob + sl[v].z * a.osD + sl[v].y * a.osH + sl[v].x * a.osW;
      float gz[2], gy[2], gx[2], fo[2], ft[2];
      if (__all_sync(0xffffffffu, tier == 2)) {
#pragma unroll
        for (int v = 0; v < 2; v++) su[v].template finish_fast<false>(s, a.vD, a.vH, sW);
#pragma unroll
        for (int v = 0; v < 2; v++)
          row_voxel_backward<C1, OG, OV, 1, LOSS, FASTM>(a, su[v], vb, gv_delta, gop[v], sl[v].valid, gz[v], gy[v], gx[v], &fo[v], &ft[v]);
      } else if (__all_sync(0xffffffffu, tier >= 1)) {
#pragma unroll
        for (int v = 0; v < 2; v++) su[v].template finish_fast<true>(s, a.vD, a.vH, sW);
#pragma unroll
        for (int v = 0; v < 2; v++)
          row_voxel_backward<C1, OG, OV, 2, LOSS, FASTM>(a, su[v], vb, gv_delta, gop[v], sl[v].valid, gz[v], gy[v], gx[v], &fo[v], &ft[v]);
      } else {
#pragma unroll
        for (int v = 0; v < 2; v++) {
          su[v].finish_exact(s);
          row_voxel_backward<C1, OG, OV, 0, LOSS, FASTM>(a, su[v], vb, gv_delta, gop[v], sl[v].valid, gz[v], gy[v], gx[v], &fo[v], &ft[v]);
        }
      }
      if constexpr (LOSS) {
#pragma unroll
        for (int v = 0; v < 2; v++) {
          const bool ok = sl[v].valid;
          if (ok) {
            // SmoothL1Criterion_updateOutput: z = |input - target|; sum += z < 1 ? 0.5 z^2 : z - 0.5  (double)
            const double zz = (double)fabsf(__fsub_rn(fo[v], ft[v]));
            lsum += zz < 1.0 ? 0.5 * zz * zz : zz - 0.5;
            if (outb) st_stream_f1(outb + sl[v].z * a.osD + sl[v].y * a.osH + sl[v].x * a.osW, fo[v]);
          }
          // computeIOU (util.lua:110-120): thresholds at 0.5, intersection and union counts per sample
          const bool p = ok && fo[v] > 0.5f, q = ok && ft[v] > 0.5f;
          n_inter += __popc(__ballot_sync(0xffffffffu, p && q));
          n_union += __popc(__ballot_sync(0xffffffffu, p || q));
        }
      }
      // ---- sinks of the coordinate gradient ----
      if constexpr (SINK) {
#pragma unroll
        for (int v = 0; v < 2; v++) {
          if constexpr (SRC == ROW_SRC_GRID) {
            float *g = ggb + sl[v].z * a.ggD + sl[v].y * a.ggH + sl[v].x * a.ggW;
            if (sl[v].valid) {
              if constexpr (AOS4) {
                st_stream_f4(reinterpret_cast<float4 *>(g), make_float4(gz[v], gy[v], gx[v], 0.0f));  // ...c:819-822
              } else {
                g[0] = gz[v], g[a.ggC] = gy[v], g[2 * a.ggC] = gx[v];
                if (!a.g3) g[3 * a.ggC] = 0.0f;
              }
            }
          } else if constexpr (SRC == ROW_SRC_AFFINE) {
            // gradT[i][j] += gradGrid[i] * base[j]  (VolumetricGridGenerator.lua:79-82), as k_gridgen_bwd_partial
            if (sl[v].valid) {
              const float zn = m.tz[sl[v].z], yn = m.ty[sl[v].y], xn = m.tx[sl[v].x];
              const float g3[3] = {gz[v], gy[v], gx[v]};
#pragma unroll
              for (int i = 0; i < 3; i++) {
                tacc[4 * i + 0] = fmaf(g3[i], zn, tacc[4 * i + 0]);
                tacc[4 * i + 1] = fmaf(g3[i], yn, tacc[4 * i + 1]);
                tacc[4 * i + 2] = fmaf(g3[i], xn, tacc[4 * i + 2]);
                tacc[4 * i + 3] += g3[i];
              }
            }
          } else {
            if (sl[v].valid) {
              float *sv = st + (TWO_ROWS ? v * 6 * owp : 0);
              const float wx = m.tx[sl[v].x], wx1 = 1.0f - wx;
              const int x = sl[v].x;
              sv[0 * owp + x] = wx * gz[v], sv[1 * owp + x] = wx * gy[v], sv[2 * owp + x] = wx * gx[v];
              sv[3 * owp + x] = wx1 * gz[v], sv[4 * owp + x] = wx1 * gy[v], sv[5 * owp + x] = wx1 * gx[v];
            }
          }
        }
      }
    }
    if constexpr (SRC == ROW_SRC_CAGE && SINK) {
      __syncwarp();
#pragma unroll
      for (int v = 0; v < (TWO_ROWS ? 2 : 1); v++) {
        if (r + v < r_end) {
          const bool wrap = v == 1 && y + 1 == s.oH;
          cacc.row(m, a, st + v * 6 * owp, owp, wrap ? z + 1 : z, v == 1 ? (wrap ? 0 : y + 1) : y, lane);
        }
      }
      __syncwarp();
    }
  }

  if constexpr (LOSS) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, d);
    if (lane == 0) {
      atomicAdd(a.loss + b, lsum);
      if (a.iou) {
        atomicAdd(a.iou + 2 * b, (unsigned long long)n_inter);
        atomicAdd(a.iou + 2 * b + 1, (unsigned long long)n_union);
      }
    }
  }
  if constexpr (SRC == ROW_SRC_AFFINE && SINK) {
    // warp tree in double, then the warps in order: one partial 4x4 per CTA (row 3 is zero: gradGrid[...,3] = 0)
#pragma unroll
    for (int i = 0; i < 12; i++) {
      double v = (double)tacc[i];
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      if (lane == 0) red[wid][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
      double v = 0.0;
      if (threadIdx.x < 12)
#pragma unroll
        for (int w = 0; w < kRowWarps; w++) v += red[w][threadIdx.x];
      a.partial[((long long)b * gridDim.x + blockIdx.x) * 16 + threadIdx.x] = v;
    }
  }
  if constexpr (SRC == ROW_SRC_CAGE && SINK) {
    cacc.flush(m, a, lane);
    __syncthreads();
    const int cHp = a.cH + 1, cWp = a.cW + 1;
    const long long plane = (long long)a.cD * a.cH * a.cW;
    float *gcb = a.gcage + (long long)b * 3 * plane;
    for (int i = threadIdx.x; i < row_cells(a); i += blockDim.x) {
      const int x = i % cWp, q = i / cWp, y = q % cHp, z = q / cHp;
      if (x < a.cW && y < a.cH && z < a.cD) {
        const long long o = ((long long)z * a.cH + y) * a.cW + x;
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const float v = m.gc[3 * i + c];
          if (v != 0.0f) red_add(gcb + c * plane + o, v);
        }
      }
    }
  }
}

// gradT[b] = sum of the CTA partials in block order, rounded once
__global__ void k_row_affine_final(const double *__restrict__ partial, float *__restrict__ gradT, int nblk) {
  const int b = blockIdx.x, e = threadIdx.x;  // 16 threads
  double v = 0.0;
  for (int k = 0; k < nblk; k++) v += partial[((long long)b * nblk + k) * 16 + e];
  gradT[(long long)b * 16 + e] = (float)v;
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
namespace {

template <typename K>
int set_smem(K kernel, size_t bytes) {
  if (bytes > 200 * 1024) return VOLSTN_ERR_UNSUPPORTED;
  if (bytes > 48 * 1024)
    VOLSTN_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return VOLSTN_OK;
}

template <int SRC, bool AOS4, bool C1, bool TWO, bool FASTM = false>
int launch_fwd(const RowArgs &a, int B, cudaStream_t st) {
  const size_t smem = row_smem_bytes<SRC, false>(a, TWO);
  int rc = set_smem(k_row_fwd<SRC, AOS4, C1, TWO, FASTM>, smem);
  if (rc) return rc;
  for (int b0 = 0; b0 < B; b0 += 65535) {
    RowArgs c = a;
    const int nb = B - b0 < 65535 ? B - b0 : 65535;
    c.s.vol += (long long)b0 * a.vB;
    c.s.out += (long long)b0 * a.oB;
    if (c.s.grid) c.s.grid += (long long)b0 * a.gB;
    if (c.T) c.T += (long long)b0 * 16;
    if (c.cage) c.cage += (long long)b0 * 3 * a.cD * a.cH * a.cW;
    k_row_fwd<SRC, AOS4, C1, TWO, FASTM><<<dim3(a.nblk, nb), kRowThreads, smem, st>>>(c);
    if ((rc = after_launch())) return rc;
  }
  return VOLSTN_OK;
}

template <int SRC, bool AOS4, bool C1, bool TWO, bool OG, bool OV, bool LOSS = false, bool HEAVY = false, bool FASTM = false>
int launch_bwd(const RowArgs &a, int B, cudaStream_t st) {
  if constexpr (SRC == ROW_SRC_CAGE && OG && C1 && !HEAVY) {
    if (a.s.oW > 64 || 3 * (a.cW + 1) > 32) return launch_bwd<SRC, AOS4, C1, TWO, OG, OV, LOSS, true, FASTM>(a, B, st);
  }
  const size_t smem = row_smem_bytes<SRC, true>(a, TWO);
  int rc = set_smem(k_row_bwd<SRC, AOS4, C1, TWO, OG, OV, LOSS, HEAVY, FASTM>, smem);
  if (rc) return rc;
  for (int b0 = 0; b0 < B; b0 += 65535) {
    RowArgs c = a;
    const int nb = B - b0 < 65535 ? B - b0 : 65535;
    c.s.vol += (long long)b0 * a.vB;
    if (c.s.gvol) c.s.gvol += (long long)b0 * a.vB;
    c.s.gout += (long long)b0 * a.oB;
    if (c.s.grid) c.s.grid += (long long)b0 * a.gB;
    if (c.s.ggrid) c.s.ggrid += (long long)b0 * a.ggB;
    if (c.T) c.T += (long long)b0 * 16;
    if (c.partial) c.partial += (long long)b0 * a.nblk * 16;
    if (c.cage) c.cage += (long long)b0 * 3 * a.cD * a.cH * a.cW;
    if (c.gcage) c.gcage += (long long)b0 * 3 * a.cD * a.cH * a.cW;
    if (c.s.out) c.s.out += (long long)b0 * a.oB;
    if (c.loss) c.loss += b0;
    if (c.iou) c.iou += 2 * (long long)b0;
    k_row_bwd<SRC, AOS4, C1, TWO, OG, OV, LOSS, HEAVY, FASTM><<<dim3(a.nblk, nb), kRowThreads, smem, st>>>(c);
    if ((rc = after_launch())) return rc;
  }
  return VOLSTN_OK;
}

template <int SRC, bool AOS4>
int fwd_pick(const RowArgs &a, const RowPlan &p, int B, cudaStream_t st) {
  if (p.c1) return p.two_rows ? launch_fwd<SRC, AOS4, true, true>(a, B, st) : launch_fwd<SRC, AOS4, true, false>(a, B, st);
  return p.two_rows ? launch_fwd<SRC, AOS4, false, true>(a, B, st) : launch_fwd<SRC, AOS4, false, false>(a, B, st);
}

template <int SRC, bool AOS4, bool OG, bool OV>
int bwd_pick(const RowArgs &a, const RowPlan &p, int B, cudaStream_t st) {
  if (p.c1)
    return p.two_rows ? launch_bwd<SRC, AOS4, true, true, OG, OV>(a, B, st)
                      : launch_bwd<SRC, AOS4, true, false, OG, OV>(a, B, st);
  return p.two_rows ? launch_bwd<SRC, AOS4, false, true, OG, OV>(a, B, st)
                    : launch_bwd<SRC, AOS4, false, false, OG, OV>(a, B, st);
}

template <int SRC, bool AOS4>
int bwd_mode(const RowArgs &a, const RowPlan &p, int B, bool og, bool ov, cudaStream_t st) {
  if (og) return bwd_pick<SRC, AOS4, true, false>(a, p, B, st);
  if (ov) return bwd_pick<SRC, AOS4, false, true>(a, p, B, st);
  return bwd_pick<SRC, AOS4, false, false>(a, p, B, st);
}

}  // namespace

void row_geometry(int64_t B, int64_t rows, int src, int two_rows, int *nblk, int *rows_per_warp) {
  // ~4 waves of 148 SMs x 8 resident CTAs over the whole batch, but at least 4 rows per warp; the cage source has
  // a heavier prologue / epilogue (cage, axis tables and the cage gradient in shared memory), so its CTAs are twice
  // as long and never shorter than 8 rows per warp.  Swept on a B200 at 32^3 x 32 and 64^3 x 64
  // (profiles/r02_fbench_history.txt): flat between 4 and 14 rows, -15 % at 28, -30 % at 56.
  int64_t want = (int64_t)kSMs * 8 * 4 / (B > 0 ? B : 1);
  if (src == ROW_SRC_CAGE) want /= 2;
  if (want < 1) want = 1;
  int64_t rpw = (rows + want * kRowWarps - 1) / (want * kRowWarps);
  const int64_t floor_rows = src == ROW_SRC_CAGE ? 8 : 4;
  if (rpw < floor_rows) rpw = floor_rows;
  if (const char *e = getenv("VOLSTN_ROW_RPW")) {  // tuning only
    const long long v = atoll(e);
    if (v > 0) rpw = v;
  }
  if (two_rows) rpw = (rpw + 1) & ~(int64_t)1;
  int64_t blocks = (rows + rpw * kRowWarps - 1) / (rpw * kRowWarps);
  if (blocks < 1) blocks = 1;
  *nblk = (int)blocks;
  *rows_per_warp = (int)rpw;
}

int row_forward(const RowArgs &a, const RowPlan &p, int B, cudaStream_t st) {
  switch (p.src) {
    case ROW_SRC_GRID: return p.grid_aos4 ? fwd_pick<ROW_SRC_GRID, true>(a, p, B, st) : fwd_pick<ROW_SRC_GRID, false>(a, p, B, st);
    case ROW_SRC_AFFINE: return fwd_pick<ROW_SRC_AFFINE, false>(a, p, B, st);
    default:
      if (p.fast && p.c1)
        return p.two_rows ? launch_fwd<ROW_SRC_CAGE, false, true, true, true>(a, B, st)
                          : launch_fwd<ROW_SRC_CAGE, false, true, false, true>(a, B, st);
      return fwd_pick<ROW_SRC_CAGE, false>(a, p, B, st);
  }
}

int row_backward(const RowArgs &a, const RowPlan &p, int B, bool og, bool ov, cudaStream_t st) {
  int rc;
  switch (p.src) {
    case ROW_SRC_GRID:
      return p.grid_aos4 ? bwd_mode<ROW_SRC_GRID, true>(a, p, B, og, ov, st) : bwd_mode<ROW_SRC_GRID, false>(a, p, B, og, ov, st);
    case ROW_SRC_AFFINE:
      if ((rc = bwd_mode<ROW_SRC_AFFINE, false>(a, p, B, og, ov, st))) return rc;
      if (!ov) {
        k_row_affine_final<<<B, 16, 0, st>>>(a.partial, a.gradT, a.nblk);
        return after_launch();
      }
      return VOLSTN_OK;
    default:
      if (p.fast && p.c1 && !ov) {
        if (og)
          return p.two_rows ? launch_bwd<ROW_SRC_CAGE, false, true, true, true, false, false, false, true>(a, B, st)
                            : launch_bwd<ROW_SRC_CAGE, false, true, false, true, false, false, false, true>(a, B, st);
        return p.two_rows ? launch_bwd<ROW_SRC_CAGE, false, true, true, false, false, false, false, true>(a, B, st)
                          : launch_bwd<ROW_SRC_CAGE, false, true, false, false, false, false, false, true>(a, B, st);
      }
      return bwd_mode<ROW_SRC_CAGE, false>(a, p, B, og, ov, st);
  }
}

// fused training step: forward + SmoothL1 + backward in one pass (single channel); cage source or grid tensor
int row_cage_step(const RowArgs &a, const RowPlan &p, int B, bool og, cudaStream_t st) {
  if (!p.c1) return VOLSTN_ERR_UNSUPPORTED;
  if (p.src == ROW_SRC_GRID) {
    if (p.grid_aos4) {
      if (og)
        return p.two_rows ? launch_bwd<ROW_SRC_GRID, true, true, true, true, false, true>(a, B, st)
                          : launch_bwd<ROW_SRC_GRID, true, true, false, true, false, true>(a, B, st);
      return p.two_rows ? launch_bwd<ROW_SRC_GRID, true, true, true, false, false, true>(a, B, st)
                        : launch_bwd<ROW_SRC_GRID, true, true, false, false, false, true>(a, B, st);
    }
    if (og)
      return p.two_rows ? launch_bwd<ROW_SRC_GRID, false, true, true, true, false, true>(a, B, st)
                        : launch_bwd<ROW_SRC_GRID, false, true, false, true, false, true>(a, B, st);
    return p.two_rows ? launch_bwd<ROW_SRC_GRID, false, true, true, false, false, true>(a, B, st)
                      : launch_bwd<ROW_SRC_GRID, false, true, false, false, false, true>(a, B, st);
  }
  if (p.src != ROW_SRC_CAGE) return VOLSTN_ERR_UNSUPPORTED;
  if (p.fast) {
    if (og)
      return p.two_rows ? launch_bwd<ROW_SRC_CAGE, false, true, true, true, false, true, false, true>(a, B, st)
                        : launch_bwd<ROW_SRC_CAGE, false, true, false, true, false, true, false, true>(a, B, st);
    return p.two_rows ? launch_bwd<ROW_SRC_CAGE, false, true, true, false, false, true, false, true>(a, B, st)
                      : launch_bwd<ROW_SRC_CAGE, false, true, false, false, false, true, false, true>(a, B, st);
  }
  if (og)
    return p.two_rows ? launch_bwd<ROW_SRC_CAGE, false, true, true, true, false, true>(a, B, st)
                      : launch_bwd<ROW_SRC_CAGE, false, true, false, true, false, true>(a, B, st);
  return p.two_rows ? launch_bwd<ROW_SRC_CAGE, false, true, true, false, false, true>(a, B, st)
                    : launch_bwd<ROW_SRC_CAGE, false, true, false, false, false, true>(a, B, st);
}

}  // namespace volstn
