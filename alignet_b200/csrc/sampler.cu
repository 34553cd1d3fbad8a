// sampler.cu -- nn.BilinearSamplerVolumetric on sm_100a: forward gather, backward
// (gradVol scatter-add + gradGrid), index dump, and the C-ABI dispatchers for DEVICE pointers.
//
// Replaces stn3dtorch/generic/BilinearSamplerVolumetric.c (CPU) and
// stn3dtorch/BilinearSamplerVolumetric.cu (the sm_20-era CUDA kernels) -- see include/volstn_b200.h.
//
// Thread mapping (the opposite of the reference's warp-per-voxel / lanes-over-channels,
// ...cu:590,669): one thread per output voxel, consecutive lanes along W, so that
//   - the grid element is ONE aligned 128-bit streaming load per voxel (G = 4),
//   - the C output / gradOut channels are one 4C-byte vector access,
//   - gradGrid is one 128-bit streaming store,
//   - the eight corner gathers of a warp fall on four short runs of volume rows (L1-resident for
//     coherent fields), and gradVol reductions are vector `red.global.add` (C = 2, 4).
#include <cstdlib>
#include <cstring>

#include "sampler.cuh"
#include "volstn_internal.h"

namespace volstn {

constexpr int kThreads = 256;

// =============================================================================================
// packed kernels: grid.x covers the voxels of one sample (VPT per thread, stride kThreads so every
// access instruction of a warp is contiguous), grid.y walks the batch.
//
// The warp vote selects between two bodies that compute the same values: FAST (all eight corners of
// all 32 lanes in bounds: no predicates, conversion-free floor) and the exact predicated one.
// Tail lanes re-read the last voxel (no divergent set-up code) and only skip their stores.
//
// Measured and rejected on a B200 (gpurun_out/r1_kbench_pf.txt, r1_kbench_persist.txt): an L2 prefetch
// of the grid one wave of CTAs ahead (no change), a persistent, register-pipelined loop (64 registers,
// -30 %) and a plain strided loop without prefetch (r1_kbench_loop.txt: 40 registers, -20 % even when the
// launch covers the sample in one trip).
// =============================================================================================
template <typename real, int CT, int G, int VPT>
__global__ void __launch_bounds__(kThreads) k_sampler_fwd_packed(const SamplerArgs<real> a) {
  const int C = CT ? CT : a.C;
  const int sW = C, sH = a.iW * C, sD = a.iH * a.iW * C;
  const unsigned n = a.n32;
  const unsigned j0 = blockIdx.x * (kThreads * VPT) + threadIdx.x;
  for (unsigned b = blockIdx.y; b < (unsigned)a.B; b += gridDim.y) {
    const size_t sb = (size_t)b * n;  // first voxel of the sample in the packed batch
    const real *vb = opaque(a.vol + (size_t)b * a.vs32);
    constexpr int GC = G % 10, GS = G >= 10 ? 1 : G;  // components; element stride of a voxel (planar grid: 1)
    const real *gb = opaque(a.grid + sb * GC);
    real *ob = opaque(a.out + sb * C);
    real zf[VPT], yf[VPT], xf[VPT];
#pragma unroll
    for (int v = 0; v < VPT; v++) {
      const unsigned j = min(j0 + v * kThreads, n - 1);
      ld_grid<real, G>(gb + (size_t)j * GS, zf[v], yf[v], xf[v], n);
    }
    if constexpr (CT > 0) {
      // one vote for all VPT voxels of the thread: inside the fast block the bodies are straight-line
      // code, so the gathers of the second voxel are in flight while the first one is summed
      VoxelSetup<real> s[VPT];
      int tier = 2;
#pragma unroll
      for (int v = 0; v < VPT; v++) tier = min(tier, s[v].begin(zf[v], yf[v], xf[v], a));
      if (__all_sync(0xffffffffu, tier == 2)) {
#pragma unroll
        for (int v = 0; v < VPT; v++) s[v].template finish_fast<false>(a, sD, sH, sW);
#pragma unroll
        for (int v = 0; v < VPT; v++) {
          const unsigned j = j0 + v * kThreads;
          voxel_forward_ct<real, CT, G, 1>(s[v], vb, sD, sH, ob + (size_t)j * C, j < n);
        }
      } else if (__all_sync(0xffffffffu, tier >= 1)) {
#pragma unroll
        for (int v = 0; v < VPT; v++) s[v].template finish_fast<true>(a, sD, sH, sW);
#pragma unroll
        for (int v = 0; v < VPT; v++) {
          const unsigned j = j0 + v * kThreads;
          voxel_forward_ct<real, CT, G, 2>(s[v], vb, sD, sH, ob + (size_t)j * C, j < n);
        }
      } else {
#pragma unroll
        for (int v = 0; v < VPT; v++) {
          const unsigned j = j0 + v * kThreads;
          s[v].finish_exact(a);
          voxel_forward_ct<real, CT, G, 0>(s[v], vb, sD, sH, ob + (size_t)j * C, j < n);
        }
      }
    } else {
#pragma unroll
      for (int v = 0; v < VPT; v++) {
        const unsigned j = j0 + v * kThreads;
        VoxelSetup<real> s;
        s.init(zf[v], yf[v], xf[v], a);
        if (j < n) voxel_forward_packed<real, CT, G>(a, s, vb, sD, sH, sW, ob + (size_t)j * C);
      }
    }
  }
}

// resident CTAs per SM the single-channel float backward is compiled for (register cap 65536 / (256 * n)): 6 -> 40
// registers and 24-byte spills (6 STL + 6 LDL), 5 -> 48 registers, no spills, 4 -> 56.  Swept again in round 2 with the
// per-warp quad path in the body (profiles/r02_bwd_launch_bounds.md): identity field 187.0 / 187.5 / 202.4 us, jittered
// cage 468 / 467 / 474 us, uniformly random 682 / 662 / 661 us at 6 / 5 / 4 -- 5 is as fast as 6 and spill-free.
#ifndef VOLSTN_BWD_CTAS
#define VOLSTN_BWD_CTAS 5
#endif
template <typename real, int CT, int G, int VPT, bool ONLY_GRID, bool ONLY_VOL>
__global__ void __launch_bounds__(kThreads, (CT == 1 && sizeof(real) == 4) ? VOLSTN_BWD_CTAS : 1) k_sampler_bwd_packed(const SamplerArgs<real> a) {
  const int C = CT ? CT : a.C;
  constexpr int CR = CT ? CT : 1;
  const int sW = C, sH = a.iW * C, sD = a.iH * a.iW * C;
  const unsigned n = a.n32;
  const unsigned j0 = blockIdx.x * (kThreads * VPT) + threadIdx.x;
  // gradVol is packed like the volume: one pointer difference serves every sample
  const long long gv_delta = a.gvol - a.vol;
  for (unsigned b = blockIdx.y; b < (unsigned)a.B; b += gridDim.y) {
    const size_t sb = (size_t)b * n;
    const real *vb = opaque(a.vol + (size_t)b * a.vs32);
    real *gvb = const_cast<real *>(vb) + gv_delta;
    constexpr int GC = G % 10, GS = G >= 10 ? 1 : G;
    const real *gb = opaque(a.grid + sb * GC);
    real *ggb = opaque(a.ggrid + sb * GC);
    const real *gob = opaque(a.gout + sb * C);
    real zf[VPT], yf[VPT], xf[VPT], go[VPT][CR];
#pragma unroll
    for (int v = 0; v < VPT; v++) {
      const size_t j = min(j0 + v * kThreads, n - 1);
      ld_grid<real, G>(gb + j * GS, zf[v], yf[v], xf[v], n);
      if constexpr (CT > 0) ld_channels_stream<real, CR>(gob + j * C, go[v]);
    }
    if constexpr (CT > 0) {
      VoxelSetup<real> s[VPT];
      int tier = 2;
#pragma unroll
      for (int v = 0; v < VPT; v++) tier = min(tier, s[v].begin(zf[v], yf[v], xf[v], a));
      if (__all_sync(0xffffffffu, tier == 2)) {
#pragma unroll
        for (int v = 0; v < VPT; v++) s[v].template finish_fast<false>(a, sD, sH, sW);
#pragma unroll
        for (int v = 0; v < VPT; v++) {
          const unsigned j = j0 + v * kThreads;
          // a.quad: 1 = forced (VOLSTN_ALGO_QUAD), 2 = chosen per warp (VOLSTN_ALGO_AUTO): the aligned 16-byte reductions
          // when the warp's 32 corner-0 cells span 48 elements or more, i.e. when its scalar reductions would not coalesce
          // anyway.  Identity-like fields never qualify (span <= 33: 187.4 vs 187.3 us); jittered cage 488 -> 467 us,
          // rotations 508 -> 471 us, uniformly random 867 -> 682 us.  (Testing for exactly consecutive cells instead sent
          // the identity field to the quad path -- its coordinates sit one ulp below an integer 36 % of the time: 245 us.)
          bool quad = a.quad == 1;
          if (a.quad == 2)
            quad = __reduce_max_sync(0xffffffffu, s[v].fbase) - __reduce_min_sync(0xffffffffu, s[v].fbase) >= 48;
          if (quad)
            voxel_backward_ct<real, CR, G, ONLY_GRID, ONLY_VOL, 1, true>(a, s[v], vb, gv_delta, sD, sH, go[v],
                                                                         ggb + (size_t)j * GS, j < n);
          else
            voxel_backward_ct<real, CR, G, ONLY_GRID, ONLY_VOL, 1, false>(a, s[v], vb, gv_delta, sD, sH, go[v],
                                                                          ggb + (size_t)j * GS, j < n);
        }
      } else if (__all_sync(0xffffffffu, tier >= 1)) {
#pragma unroll
        for (int v = 0; v < VPT; v++) s[v].template finish_fast<true>(a, sD, sH, sW);
#pragma unroll
        for (int v = 0; v < VPT; v++) {
          const unsigned j = j0 + v * kThreads;
          voxel_backward_ct<real, CR, G, ONLY_GRID, ONLY_VOL, 2>(a, s[v], vb, gv_delta, sD, sH, go[v],
                                                                 ggb + (size_t)j * GS, j < n);
        }
      } else {
#pragma unroll
        for (int v = 0; v < VPT; v++) {
          const unsigned j = j0 + v * kThreads;
          s[v].finish_exact(a);
          voxel_backward_ct<real, CR, G, ONLY_GRID, ONLY_VOL, 0>(a, s[v], vb, gv_delta, sD, sH, go[v],
                                                                 ggb + (size_t)j * GS, j < n);
        }
      }
    } else {
#pragma unroll
      for (int v = 0; v < VPT; v++) {
        const unsigned j = j0 + v * kThreads;
        VoxelSetup<real> s;
        s.init(zf[v], yf[v], xf[v], a);
        if (j < n)
          voxel_backward_packed<real, CT, G, ONLY_GRID, ONLY_VOL>(a, s, vb, gvb, sD, sH, sW, gob + (size_t)j * C,
                                                                  ggb + (size_t)j * GS);
      }
    }
  }
}

// =============================================================================================
// grid generator folded into the packed forward (forward only: the loader-side augmentation and the
// projector of stn3dtorch/test.lua:15-25 have no backward).  Same flat voxel mapping and tiers as k_sampler_fwd_packed;
// the sampling coordinate is T[b] . (zn, yn, xn, 1) in the operation order of k_gridgen_fwd.  The three axis tables
// (float)(-1 + k/(n-1)*2) are evaluated in double ON THE HOST (the same correctly rounded IEEE operations as on the
// device: identical bits) and arrive as kernel parameters; a thread's (z, y, x) -- hence its three table values --
// are fixed for the whole walk over the samples, so the per-voxel cost over the plain forward is 18 flops against a
// 16-byte grid load.  128 instructions per voxel-warp against 169 of the row kernel.
// =============================================================================================
constexpr int kAffineTab = 768;  // D + H + W table entries that fit the parameter block

struct AffineArgs {
  SamplerArgs<float> s;
  const float *T;
  float tab[kAffineTab];  // zn[oD], yn[oH], xn[oW]
};

template <int CT>
__global__ void __launch_bounds__(kThreads) k_affine_fwd_packed(const __grid_constant__ AffineArgs p) {
  const SamplerArgs<float> &a = p.s;
  __shared__ float tab[kAffineTab];
  for (int i = threadIdx.x; i < a.oD + a.oH + a.oW; i += kThreads) tab[i] = p.tab[i];
  __syncthreads();
  constexpr int VPT = 2;
  const int sW = CT, sH = a.iW * CT, sD = a.iH * a.iW * CT;
  const unsigned n = a.n32;
  const unsigned j0 = blockIdx.x * (kThreads * VPT) + threadIdx.x;
  float zn[VPT], yn[VPT], xn[VPT];
#pragma unroll
  for (int v = 0; v < VPT; v++) {
    const unsigned j = min(j0 + v * kThreads, n - 1);
    const unsigned x = j % (unsigned)a.oW, r = j / (unsigned)a.oW, y = r % (unsigned)a.oH, z = r / (unsigned)a.oH;
    zn[v] = tab[z], yn[v] = tab[a.oD + y], xn[v] = tab[a.oD + a.oH + x];
  }
  for (unsigned b = blockIdx.y; b < (unsigned)a.B; b += gridDim.y) {
    const float *vb = opaque(a.vol + (size_t)b * a.vs32);
    float *ob = opaque(a.out + (size_t)b * n * CT);
    float t[12];
#pragma unroll
    for (int i = 0; i < 12; i++) t[i] = __ldg(p.T + (size_t)b * 16 + i);
    VoxelSetup<float> s[VPT];
    int tier = 2;
#pragma unroll
    for (int v = 0; v < VPT; v++) {
      float c3[3];
#pragma unroll
      for (int i = 0; i < 3; i++)  // ((T0*zn + T1*yn) + T2*xn) + T3
        c3[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t[4 * i], zn[v]), __fmul_rn(t[4 * i + 1], yn[v])),
                                    __fmul_rn(t[4 * i + 2], xn[v])), t[4 * i + 3]);
      tier = min(tier, s[v].begin(c3[0], c3[1], c3[2], a));
    }
    if (__all_sync(0xffffffffu, tier == 2)) {
#pragma unroll
      for (int v = 0; v < VPT; v++) s[v].template finish_fast<false>(a, sD, sH, sW);
#pragma unroll
      for (int v = 0; v < VPT; v++) {
        const unsigned j = j0 + v * kThreads;
        voxel_forward_ct<float, CT, 4, 1>(s[v], vb, sD, sH, ob + (size_t)j * CT, j < n);
      }
    } else if (__all_sync(0xffffffffu, tier >= 1)) {
#pragma unroll
      for (int v = 0; v < VPT; v++) s[v].template finish_fast<true>(a, sD, sH, sW);
#pragma unroll
      for (int v = 0; v < VPT; v++) {
        const unsigned j = j0 + v * kThreads;
        voxel_forward_ct<float, CT, 4, 2>(s[v], vb, sD, sH, ob + (size_t)j * CT, j < n);
      }
    } else {
#pragma unroll
      for (int v = 0; v < VPT; v++) {
        const unsigned j = j0 + v * kThreads;
        s[v].finish_exact(a);
        voxel_forward_ct<float, CT, 4, 0>(s[v], vb, sD, sH, ob + (size_t)j * CT, j < n);
      }
    }
  }
}

// =============================================================================================
// generic kernels: arbitrary element strides on dims 0..3 of every tensor (the CPU reference
// honours them, ...c:460-473), any C, float or double, 64-bit offsets.  One thread per voxel.
// =============================================================================================
// Launch geometry: grid.x covers the voxels of ONE output plane (y, x) -- one 32-bit divide per thread --, grid.y walks the
// planes z and grid.z the batch (in rounds of at most 65535).  The first version decomposed a flat 64-bit voxel index per
// tensor: the double-precision forward spent 218 IMADs and six 64-bit division calls per voxel on that against 45 flops.
struct GenericVoxel {
  int b, z, y, x;
  bool ok;
};
template <typename real>
__device__ __forceinline__ GenericVoxel generic_voxel(const SamplerArgs<real> &a, int b) {
  GenericVoxel v;
  const unsigned p = blockIdx.x * kThreads + threadIdx.x, plane = (unsigned)a.oH * (unsigned)a.oW;
  v.ok = p < plane;
  const unsigned q = v.ok ? p : 0;
  v.y = (int)(q / (unsigned)a.oW);
  v.x = (int)(q - (unsigned)v.y * (unsigned)a.oW);
  v.z = blockIdx.y;
  v.b = b;
  return v;
}
__device__ __forceinline__ long long generic_offset(const GenericVoxel &v, const long long (&st)[4]) {
  return (long long)v.b * st[0] + (long long)v.z * st[1] + (long long)v.y * st[2] + (long long)v.x * st[3];
}

template <typename real, int G>
__global__ void __launch_bounds__(kThreads) k_sampler_fwd_generic(const SamplerArgs<real> a) {
  using E = Ex<real>;
  for (int b = blockIdx.z; b < a.B; b += gridDim.z) {
    const GenericVoxel vx = generic_voxel(a, b);
    if (!vx.ok) continue;
    const real *g = a.grid + generic_offset(vx, a.gs);
    VoxelSetup<real> s;
    s.init(g[0], g[a.gs4], g[2 * a.gs4], a);
    real cw[8];
    s.corner_weights(cw);
    const real *vb = a.vol + (long long)b * a.vs[0];
    const long long base = s.base64(a.vs[1], a.vs[2], a.vs[3]);
    long long off[8];
#pragma unroll
    for (int k = 0; k < 8; k++)
      off[k] = base + ((k & 1) ? a.vs[3] : 0) + ((k & 2) ? a.vs[2] : 0) + ((k & 4) ? a.vs[1] : 0);
    real *o = a.out + generic_offset(vx, a.os);
    for (int t = 0; t < a.C; t++) {
      real acc = 0;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int k = (G == 4) ? order_xyz(j) : order_zxy(j);
        const real v = s.in(k) ? vb[off[k] + t * a.vs4] : real(0);
        const real term = E::mul(cw[k], v);
        acc = (j == 0) ? term : E::add(acc, term);
      }
      o[t * a.os4] = acc;
    }
  }
}

template <typename real, int G, bool ONLY_GRID, bool ONLY_VOL>
__global__ void __launch_bounds__(kThreads) k_sampler_bwd_generic(const SamplerArgs<real> a) {
  using E = Ex<real>;
  for (int b = blockIdx.z; b < a.B; b += gridDim.z) {
    const GenericVoxel vx = generic_voxel(a, b);
    if (!vx.ok) continue;
    const real *g = a.grid + generic_offset(vx, a.gs);
    VoxelSetup<real> s;
    s.init(g[0], g[a.gs4], g[2 * a.gs4], a);
    real cw[8];
    s.corner_weights(cw);
    const real *vb = a.vol + (long long)b * a.vs[0];
    real *gvb = a.gvol + (long long)b * a.gvs[0];
    const long long base = s.base64(a.vs[1], a.vs[2], a.vs[3]);
    const long long gbase = s.base64(a.gvs[1], a.gvs[2], a.gvs[3]);  // gradVol has its own strides (...c:686)
    real dot[8];
    long long off[8], goffs[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      off[k] = base + ((k & 1) ? a.vs[3] : 0) + ((k & 2) ? a.vs[2] : 0) + ((k & 4) ? a.vs[1] : 0);
      goffs[k] = gbase + ((k & 1) ? a.gvs[3] : 0) + ((k & 2) ? a.gvs[2] : 0) + ((k & 4) ? a.gvs[1] : 0);
      dot[k] = real(0);
    }
    const real *gop = a.gout + generic_offset(vx, a.os);
    for (int t = 0; t < a.C; t++) {
      const real go = gop[t * a.os4];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if (s.in(k)) {
          if constexpr (!ONLY_VOL) dot[k] = E::add(dot[k], E::mul(vb[off[k] + t * a.vs4], go));
          if constexpr (!ONLY_GRID) red_add(gvb + goffs[k] + t * a.gvs4, E::mul(cw[k], go));
        }
      }
    }
    if constexpr (!ONLY_VOL) {
      real gz, gy, gx;
      grad_grid_from_dots<real>(dot, s, a, gz, gy, gx);
      real *gg = a.ggrid + generic_offset(vx, a.ggs);
      gg[0] = gz;
      gg[a.ggs4] = gy;
      gg[2 * a.ggs4] = gx;
      if constexpr (G == 4) gg[3 * a.ggs4] = real(0);
    }
  }
}

// index / mask dump (debug; bit-exact parity test)
template <typename real>
__global__ void __launch_bounds__(kThreads) k_sampler_indices(const SamplerArgs<real> a, int32_t *idx_out,
                                                              uint8_t *mask_out) {
  for (int b = blockIdx.z; b < a.B; b += gridDim.z) {
    const GenericVoxel vx = generic_voxel(a, b);
    if (!vx.ok) continue;
    const real *g = a.grid + generic_offset(vx, a.gs);
    VoxelSetup<real> s;
    s.init(g[0], g[a.gs4], g[2 * a.gs4], a);
    const long long idx = (((long long)b * a.oD + vx.z) * a.oH + vx.y) * a.oW + vx.x;
    idx_out[idx * 3 + 0] = s.z0;
    idx_out[idx * 3 + 1] = s.y0;
    idx_out[idx * 3 + 2] = s.x0;
    mask_out[idx] = (uint8_t)s.mask();
  }
}

// =============================================================================================
// host-side validation and dispatch
// =============================================================================================
namespace {

bool is_contiguous(const volstn_tensor5 *t) {
  int64_t expect = 1;
  for (int i = 4; i >= 0; i--) {
    if (t->size[i] != 1 && t->stride[i] != expect) return false;
    expect *= t->size[i];
  }
  return true;
}

bool fits_int(int64_t v) { return v >= 0 && v <= 0x7fffffffLL; }

int check_desc(const volstn_tensor5 *t) {
  if (!t) return VOLSTN_ERR_ARG;
  for (int i = 0; i < 5; i++)
    if (!fits_int(t->size[i])) return VOLSTN_ERR_ARG;
  return VOLSTN_OK;  // any element strides: the packed kernels need contiguity (packed_ok), the others honour them
}

int64_t numel(const volstn_tensor5 *t) { return t->size[0] * t->size[1] * t->size[2] * t->size[3] * t->size[4]; }

bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// shape rules of BilinearSamplerVolumetric.lua:26-42,73
int check_sampler_shapes(int G, const volstn_tensor5 *vol, const volstn_tensor5 *grid, const volstn_tensor5 *outlike) {
  if (G != 3 && G != 4) return VOLSTN_ERR_ARG;
  int st;
  if ((st = check_desc(vol)) || (st = check_desc(grid)) || (st = check_desc(outlike))) return st;
  if (grid->size[0] != vol->size[0]) return VOLSTN_ERR_SHAPE;  // lua:33
  if (grid->size[4] != G) return VOLSTN_ERR_SHAPE;             // lua:34
  for (int i = 0; i < 4; i++)
    if (outlike->size[i] != grid->size[i]) return VOLSTN_ERR_SHAPE;  // lua:36-41,73
  if (outlike->size[4] != vol->size[4]) return VOLSTN_ERR_SHAPE;
  return VOLSTN_OK;
}

template <typename real>
void fill_fast(SamplerArgs<real> &a) {
  a.fbx = a.fby = a.fbz = 0;
  a.fneg = 0;
  a.border_ok = 0;
  a.quad = 0;
  a.tile_agg = 0;
  if constexpr (sizeof(real) == 4) {
    if (a.iD <= (1 << 22) && a.iH <= (1 << 22) && a.iW <= (1 << 22)) {
      memcpy(&a.fbx, &a.wm1, 4);
      memcpy(&a.fby, &a.hm1, 4);
      memcpy(&a.fbz, &a.dm1, 4);
      const unsigned sW = (unsigned)a.C, sH = (unsigned)a.iW * sW, sD = (unsigned)a.iH * sH;
      a.fneg = (int)(0u - 0x4B000000u * (sD + sH + sW));
      a.border_ok = 1;
    }
  }
}

template <typename real>
void fill_common(SamplerArgs<real> &a, const volstn_tensor5 *vol, const volstn_tensor5 *grid,
                 const volstn_tensor5 *outlike) {
  a.vol = static_cast<const real *>(vol->data);
  a.grid = static_cast<const real *>(grid->data);
  a.out = nullptr;
  a.gvol = nullptr;
  a.ggrid = nullptr;
  a.gout = nullptr;
  a.B = (int)vol->size[0];
  a.iD = (int)vol->size[1];
  a.iH = (int)vol->size[2];
  a.iW = (int)vol->size[3];
  a.C = (int)vol->size[4];
  a.oD = (int)grid->size[1];
  a.oH = (int)grid->size[2];
  a.oW = (int)grid->size[3];
  a.n = (long long)a.oD * a.oH * a.oW;
  for (int i = 0; i < 4; i++) {
    a.vs[i] = vol->stride[i];
    a.gs[i] = grid->stride[i];
    a.os[i] = outlike->stride[i];
    a.gvs[i] = 0;
    a.ggs[i] = 0;
  }
  a.vs4 = vol->stride[4];
  a.gs4 = grid->stride[4];
  a.os4 = outlike->stride[4];
  a.gvs4 = a.ggs4 = 1;
  a.dm1 = (real)(a.iD - 1);
  a.hm1 = (real)(a.iH - 1);
  a.wm1 = (real)(a.iW - 1);
  fill_fast(a);
  a.n32 = (unsigned)a.n;
  a.vs32 = (unsigned)vol->stride[0];
}

// grid of the generic kernels (generic_voxel): x = blocks per output plane, y = planes, z = batch rounds.  -1 = not launchable
// (more than 65535 planes or 2^31 voxels per plane: no caller gets there -- the sizes come from int extents -- but say so)
template <typename real>
bool generic_grid(const SamplerArgs<real> &a, dim3 &grid) {
  const long long plane = (long long)a.oH * a.oW;
  if (a.oD > 65535 || plane >= 0x7fffffffLL) return false;
  grid = dim3((unsigned)((plane + kThreads - 1) / kThreads), (unsigned)a.oD, (unsigned)(a.B < 65535 ? a.B : 65535));
  return true;
}

unsigned ctas_per_sample(long long n, int vpt) {
  return (unsigned)((n + (long long)kThreads * vpt - 1) / ((long long)kThreads * vpt));
}

// gridDim.y: how many samples are walked side by side; a CTA then loops over the samples b, b + gridDim.y, ...  About
// 16 K CTAs per launch instead of one CTA per 512 voxels of the whole batch (32 K at 64^3 x 64) is 2-10 % faster on a
// B200 -- forward 83.4 -> 78.7 us, backward 190.3 -> 186.4 us at 64^3 x 64, C = 1; 181 -> 163 us and 558 -> 516 us at C = 4;
// 184 -> 164 us and 596 -> 544 us at 128^3 x 8, C = 4; fewer than 8 K CTAs loses again (profiles/r01_kbench_history.txt,
// section batch_rows).  The number of rounds is an integer so that every CTA walks the same number of samples (+-1).
unsigned batch_rows(int B, unsigned blocks_x) {
  const long long kTargetCtas = 16384;
  long long rounds = ((long long)B * blocks_x + kTargetCtas - 1) / kTargetCtas;
  if (rounds < 1) rounds = 1;
  long long by = (B + rounds - 1) / rounds;
  if (const char *e = getenv("VOLSTN_BY")) {  // tuning only
    const long long v = atoll(e);
    if (v >= 1) by = v;
  }
  if (by > B) by = B;
  if (by > 65535) by = 65535;
  if (by < 1) by = 1;
  return (unsigned)by;
}

int pick_vpt(unsigned flags, int C, long long n) {
  int vpt = (int)((flags & VOLSTN_VPT_MASK) >> VOLSTN_VPT_SHIFT);
  if (vpt == 1 || vpt == 2 || vpt == 4) return vpt;
  (void)n;
  return (C <= 2) ? 2 : 1;
}

// ---- forward launchers ----
template <int CT, int G>
int launch_fwd_packed(const SamplerArgs<float> &a, int vpt, cudaStream_t st) {
  dim3 block(kThreads);
  const unsigned by = batch_rows(a.B, ctas_per_sample(a.n, vpt));
  auto blocks_x = [&](int v) { return ctas_per_sample(a.n, v); };
  switch (vpt) {
    case 4: k_sampler_fwd_packed<float, CT, G, 4><<<dim3(blocks_x(4), by), block, 0, st>>>(a); break;
    case 2: k_sampler_fwd_packed<float, CT, G, 2><<<dim3(blocks_x(2), by), block, 0, st>>>(a); break;
    default: k_sampler_fwd_packed<float, CT, G, 1><<<dim3(blocks_x(1), by), block, 0, st>>>(a); break;
  }
  return after_launch();
}

template <int G>
int dispatch_fwd_packed(const SamplerArgs<float> &a, int vpt, cudaStream_t st) {
  switch (a.C) {
    case 1: return launch_fwd_packed<1, G>(a, vpt, st);
    case 2: return launch_fwd_packed<2, G>(a, vpt, st);
    case 3: return launch_fwd_packed<3, G>(a, vpt, st);
    case 4: return launch_fwd_packed<4, G>(a, vpt, st);
    default: return launch_fwd_packed<0, G>(a, vpt > 2 ? 2 : vpt, st);
  }
}

template <typename real>
int launch_fwd_generic(const SamplerArgs<real> &a, int G, cudaStream_t st) {
  dim3 blocks;
  if (!generic_grid(a, blocks)) return VOLSTN_ERR_UNSUPPORTED;
  if (G == 4)
    k_sampler_fwd_generic<real, 4><<<blocks, kThreads, 0, st>>>(a);
  else
    k_sampler_fwd_generic<real, 3><<<blocks, kThreads, 0, st>>>(a);
  return after_launch();
}

// ---- backward launchers ----
template <int CT, int G, bool OG, bool OV>
int launch_bwd_packed(const SamplerArgs<float> &a, int vpt, cudaStream_t st) {
  dim3 block(kThreads);
  const unsigned by = batch_rows(a.B, ctas_per_sample(a.n, vpt));
  auto blocks_x = [&](int v) { return ctas_per_sample(a.n, v); };
  switch (vpt) {
    case 2: k_sampler_bwd_packed<float, CT, G, 2, OG, OV><<<dim3(blocks_x(2), by), block, 0, st>>>(a); break;
    default: k_sampler_bwd_packed<float, CT, G, 1, OG, OV><<<dim3(blocks_x(1), by), block, 0, st>>>(a); break;
  }
  return after_launch();
}

template <int G, bool OG, bool OV>
int dispatch_bwd_packed_c(const SamplerArgs<float> &a, int vpt, cudaStream_t st) {
  switch (a.C) {
    case 1: return launch_bwd_packed<1, G, OG, OV>(a, vpt, st);
    case 2: return launch_bwd_packed<2, G, OG, OV>(a, vpt, st);
    case 3: return launch_bwd_packed<3, G, OG, OV>(a, vpt, st);
    case 4: return launch_bwd_packed<4, G, OG, OV>(a, vpt, st);
    default: return launch_bwd_packed<0, G, OG, OV>(a, 1, st);
  }
}

template <int G>
int dispatch_bwd_packed(const SamplerArgs<float> &a, bool og, bool ov, int vpt, cudaStream_t st) {
  if (og) return dispatch_bwd_packed_c<G, true, false>(a, vpt, st);
  if (ov) return dispatch_bwd_packed_c<G, false, true>(a, vpt, st);
  return dispatch_bwd_packed_c<G, false, false>(a, vpt, st);
}

template <typename real, int G>
int launch_bwd_generic_g(const SamplerArgs<real> &a, bool og, bool ov, cudaStream_t st) {
  dim3 blocks;
  if (!generic_grid(a, blocks)) return VOLSTN_ERR_UNSUPPORTED;
  if (og)
    k_sampler_bwd_generic<real, G, true, false><<<blocks, kThreads, 0, st>>>(a);
  else if (ov)
    k_sampler_bwd_generic<real, G, false, true><<<blocks, kThreads, 0, st>>>(a);
  else
    k_sampler_bwd_generic<real, G, false, false><<<blocks, kThreads, 0, st>>>(a);
  return after_launch();
}

template <typename real>
int launch_bwd_generic(const SamplerArgs<real> &a, int G, bool og, bool ov, cudaStream_t st) {
  return G == 4 ? launch_bwd_generic_g<real, 4>(a, og, ov, st) : launch_bwd_generic_g<real, 3>(a, og, ov, st);
}

// can the float fast path be used for this set of tensors?
bool packed_ok(int G, const volstn_tensor5 *vol, std::initializer_list<const volstn_tensor5 *> others) {
  if (!is_contiguous(vol)) return false;
  const int64_t C = vol->size[4];
  const int64_t per_sample = vol->size[1] * vol->size[2] * vol->size[3] * C;
  if (per_sample >= 0x7fffffffLL) return false;
  for (const volstn_tensor5 *t : others)
    if (t && t->size[1] * t->size[2] * t->size[3] >= 0x7fffffffLL) return false;  // 32-bit voxel indices
  const size_t cal = (C == 4) ? 16 : (C == 2) ? 8 : 4;
  if (!aligned(vol->data, cal)) return false;
  for (const volstn_tensor5 *t : others) {
    if (!t) continue;
    if (!is_contiguous(t)) return false;
    const int64_t last = t->size[4];
    const size_t al = (last == 4) ? 16 : (last == 2) ? 8 : 4;
    if (!aligned(t->data, al)) return false;
  }
  (void)G;
  return true;
}

// single-channel contiguous volume (channels-last and channel-first coincide) with a PLANAR contiguous grid, i.e. the
// `B x D x H x W x G` view of a contiguous `B x G x D x H x W` tensor: the packed kernels with three coalesced
// component loads instead of one 128-bit load
bool planar_grid(const volstn_tensor5 *g) {
  const int64_t n = g->size[1] * g->size[2] * g->size[3];
  return g->stride[4] == n && g->stride[3] == 1 && g->stride[2] == g->size[3] && g->stride[1] == g->size[2] * g->size[3] &&
         g->stride[0] == g->size[4] * n && n < 0x7fffffffLL && aligned(g->data, 4);
}
bool planar_ok(const volstn_tensor5 *vol, const volstn_tensor5 *grid, const volstn_tensor5 *gradGrid,
               std::initializer_list<const volstn_tensor5 *> packed) {
  if (vol->size[4] != 1 || !is_contiguous(vol) || !planar_grid(grid)) return false;
  if (gradGrid && !planar_grid(gradGrid)) return false;
  if (vol->size[1] * vol->size[2] * vol->size[3] >= 0x7fffffffLL) return false;
  for (const volstn_tensor5 *t : packed)
    if (t && !is_contiguous(t)) return false;
  return true;
}

}  // namespace

// returns -1 when the flat-mapping kernel does not apply (the caller uses the row kernel)
int affine_forward_packed(const float *T, const volstn_tensor5 *vol, const volstn_tensor5 *out, unsigned flags,
                          cudaStream_t st) {
  const int64_t C = vol->size[4];
  if (C < 1 || C > 4 || !packed_ok(4, vol, {out})) return -1;
  if (out->size[1] + out->size[2] + out->size[3] > kAffineTab) return -1;
  AffineArgs p;  // 3 KB on the stack; the launch copies it into the parameter block
  volstn_tensor5 g = *out;  // shape of the (never materialised) grid
  g.size[4] = 4;
  fill_common(p.s, vol, &g, out);
  p.s.out = static_cast<float *>(out->data);
  p.s.grid = nullptr;
  p.T = T;
  if (flags & VOLSTN_DEBUG_NO_FAST) p.s.fbx = p.s.fby = p.s.fbz = 0, p.s.border_ok = 0;
  if (flags & VOLSTN_DEBUG_NO_BORDER) p.s.border_ok = 0;
  int o = 0;
  for (int ax = 1; ax <= 3; ax++) {
    const int n = (int)out->size[ax];
    for (int k = 0; k < n; k++) p.tab[o++] = (float)(-1.0 + (double)k / (double)(n - 1) * 2.0);  // VolumetricGridGenerator.lua:17-19
  }
  // long-lived CTAs: a thread's table look-ups and index arithmetic are amortised over the samples it walks
  const unsigned bx = ctas_per_sample(p.s.n, 2);
  unsigned by = batch_rows(p.s.B, bx);
  // walk up to eight samples per CTA, as long as about 2 K CTAs remain (32^3 x 32 has 64 CTAs per sample: one sample each)
  int rounds = p.s.B < 8 ? p.s.B : 8;
  while (rounds > 1 && (long long)bx * ((p.s.B + rounds - 1) / rounds) < 2048) rounds--;
  if ((p.s.B + (int)by - 1) / (int)by < rounds) by = (unsigned)((p.s.B + rounds - 1) / rounds);
  if (const char *e = getenv("VOLSTN_BY")) {  // tuning only
    const long long v = atoll(e);
    if (v >= 1 && v <= p.s.B) by = (unsigned)v;
  }
  dim3 grid(bx, by), block(kThreads);
  switch ((int)C) {
    case 1: k_affine_fwd_packed<1><<<grid, block, 0, st>>>(p); break;
    case 2: k_affine_fwd_packed<2><<<grid, block, 0, st>>>(p); break;
    case 3: k_affine_fwd_packed<3><<<grid, block, 0, st>>>(p); break;
    default: k_affine_fwd_packed<4><<<grid, block, 0, st>>>(p); break;
  }
  return after_launch();
}

int sampler_forward_impl(int dtype, int G, const volstn_tensor5 *vol, const volstn_tensor5 *grid,
                         const volstn_tensor5 *out, unsigned flags, cudaStream_t st) {
  if (dtype != VOLSTN_F32 && dtype != VOLSTN_F64) return VOLSTN_ERR_ARG;
  int rc = check_sampler_shapes(G, vol, grid, out);
  if (rc) return rc;
  if (numel(out) == 0) return VOLSTN_OK;  // nothing to write
  if (vol->size[1] < 1 || vol->size[2] < 1 || vol->size[3] < 1) return VOLSTN_ERR_SHAPE;
  if (!vol->data || !grid->data || !out->data) return VOLSTN_ERR_ARG;
  if (dtype == VOLSTN_F32) {
    SamplerArgs<float> a;
    fill_common(a, vol, grid, out);
    a.out = static_cast<float *>(out->data);
    if (flags & VOLSTN_DEBUG_NO_FAST) a.fbx = a.fby = a.fbz = 0, a.border_ok = 0;
    if (flags & VOLSTN_DEBUG_NO_BORDER) a.border_ok = 0;
    const bool force_row = (flags & VOLSTN_ALGO_MASK) == VOLSTN_ALGO_ROW;
    if (!force_row && packed_ok(G, vol, {grid, out})) {
      if ((flags & VOLSTN_ALGO_MASK) == VOLSTN_ALGO_TMA && tma_supported(a, nullptr)) return launch_fwd_tma(a, G, st);
      const int vpt = pick_vpt(flags, a.C, a.n);
      return G == 4 ? dispatch_fwd_packed<4>(a, vpt, st) : dispatch_fwd_packed<3>(a, vpt, st);
    }
    if (!force_row && planar_ok(vol, grid, nullptr, {out}))  // channel-first field, single-channel volume
      return G == 4 ? launch_fwd_packed<1, 14>(a, 2, st) : launch_fwd_packed<1, 13>(a, 2, st);
    // strided or planar tensors (e.g. permuted views of B x C x D x H x W tensors): the row kernels
    if ((rc = row_sampler_forward(G, vol, grid, out, flags, st)) != -1) return rc;
    return launch_fwd_generic<float>(a, G, st);
  }
  SamplerArgs<double> a;
  fill_common(a, vol, grid, out);
  a.out = static_cast<double *>(out->data);
  return launch_fwd_generic<double>(a, G, st);
}

int sampler_backward_impl(int dtype, int G, const volstn_tensor5 *vol, const volstn_tensor5 *grid,
                          const volstn_tensor5 *gradVol, const volstn_tensor5 *gradGrid,
                          const volstn_tensor5 *gradOut, unsigned flags, cudaStream_t st) {
  if (dtype != VOLSTN_F32 && dtype != VOLSTN_F64) return VOLSTN_ERR_ARG;
  const bool og = flags & VOLSTN_BWD_ONLY_GRID, ov = flags & VOLSTN_BWD_ONLY_VOLUME;
  if (og && ov) return VOLSTN_ERR_ARG;
  int rc = check_sampler_shapes(G, vol, grid, gradOut);
  if (rc) return rc;
  if (!og) {
    if ((rc = check_desc(gradVol))) return rc;
    for (int i = 0; i < 5; i++)
      if (gradVol->size[i] != vol->size[i]) return VOLSTN_ERR_SHAPE;  // lua:103 resizeAs
    if (numel(gradVol) && !gradVol->data) return VOLSTN_ERR_ARG;
  }
  if (!ov) {
    if ((rc = check_desc(gradGrid))) return rc;
    for (int i = 0; i < 5; i++)
      if (gradGrid->size[i] != grid->size[i]) return VOLSTN_ERR_SHAPE;
    if (numel(gradGrid) && !gradGrid->data) return VOLSTN_ERR_ARG;
  }
  const size_t esz = dtype == VOLSTN_F32 ? 4 : 8;
  if (!og && (flags & VOLSTN_BWD_ZERO_GRADVOL) && numel(gradVol)) {
    if (!desc_dense(gradVol)) return VOLSTN_ERR_LAYOUT;
    VOLSTN_CUDA_TRY(cudaMemsetAsync(gradVol->data, 0, (size_t)numel(gradVol) * esz, st));
  }
  if (numel(gradOut) == 0) return VOLSTN_OK;
  if (vol->size[1] < 1 || vol->size[2] < 1 || vol->size[3] < 1) return VOLSTN_ERR_SHAPE;
  if (!vol->data || !grid->data || !gradOut->data) return VOLSTN_ERR_ARG;

  if (dtype == VOLSTN_F32) {
    SamplerArgs<float> a;
    fill_common(a, vol, grid, gradOut);
    a.gout = static_cast<const float *>(gradOut->data);
    if (!og) {
      a.gvol = static_cast<float *>(gradVol->data);
      for (int i = 0; i < 4; i++) a.gvs[i] = gradVol->stride[i];
      a.gvs4 = gradVol->stride[4];
    }
    if (!ov) {
      a.ggrid = static_cast<float *>(gradGrid->data);
      for (int i = 0; i < 4; i++) a.ggs[i] = gradGrid->stride[i];
      a.ggs4 = gradGrid->stride[4];
    }
    if (flags & VOLSTN_DEBUG_NO_FAST) a.fbx = a.fby = a.fbz = 0, a.border_ok = 0;
    if (flags & VOLSTN_DEBUG_NO_BORDER) a.border_ok = 0;
    const bool force_row = (flags & VOLSTN_ALGO_MASK) == VOLSTN_ALGO_ROW;
    if (!force_row && packed_ok(G, vol, {grid, gradOut, og ? nullptr : gradVol, ov ? nullptr : gradGrid})) {
      a.quad = ((flags & VOLSTN_ALGO_MASK) == VOLSTN_ALGO_QUAD && !og && a.C == 1 && a.iW % 4 == 0 &&
                reinterpret_cast<uintptr_t>(a.gvol) % 16 == 0) ? 1 : 0;
      if ((flags & VOLSTN_ALGO_MASK) == VOLSTN_ALGO_AUTO && !og && a.C == 1 && a.iW % 4 == 0 &&
          reinterpret_cast<uintptr_t>(a.gvol) % 16 == 0)
        a.quad = 2;
      a.tile_agg = a.n >= 128LL * a.iD * a.iH * a.iW;
      if ((flags & VOLSTN_ALGO_MASK) == VOLSTN_ALGO_TMA && !ov && tma_supported(a, og ? nullptr : a.gvol))
        return launch_bwd_tma(a, G, og, st);
      if ((flags & VOLSTN_ALGO_MASK) == VOLSTN_ALGO_TILE && !og && tile_supported(a))
        return launch_bwd_tile(a, G, ov, st);
      // Automatic choice: a volume much smaller than the output (ALIGNet's cage up-sample: 16.8 M voxels scatter
      // into 4^3 .. 12^3 cells per sample, models/alignet_2d.lua:136-147) serialises the direct kernel's global
      // reductions on a handful of addresses -- 22.5 ms at 4^3, 3.0 ms at 8^3, 1.15 ms at 12^3 for 64^3 x 64 -- while the
      // privatised kernel folds a brick's 2048 contributions into a few cells first (and warp-aggregates its shared
      // atomics, sampler_tile.cu): 0.47 / 0.55 / 0.67 ms (tools/upbench.py, profiles/r01_upsample_backward.txt).
      if ((flags & VOLSTN_ALGO_MASK) == VOLSTN_ALGO_AUTO && !og && tile_supported(a) && a.tile_agg)
        return launch_bwd_tile(a, G, ov, st);
      int vpt = (int)((flags & VOLSTN_VPT_MASK) >> VOLSTN_VPT_SHIFT);
      if (vpt != 1 && vpt != 2) vpt = 2;
      return G == 4 ? dispatch_bwd_packed<4>(a, og, ov, vpt, st)
                    : dispatch_bwd_packed<3>(a, og, ov, vpt, st);
    }
    if (!force_row && planar_ok(vol, grid, ov ? nullptr : gradGrid, {gradOut, og ? nullptr : gradVol})) {
      if (og) return G == 4 ? launch_bwd_packed<1, 14, true, false>(a, 2, st) : launch_bwd_packed<1, 13, true, false>(a, 2, st);
      if (ov) return G == 4 ? launch_bwd_packed<1, 14, false, true>(a, 2, st) : launch_bwd_packed<1, 13, false, true>(a, 2, st);
      return G == 4 ? launch_bwd_packed<1, 14, false, false>(a, 2, st) : launch_bwd_packed<1, 13, false, false>(a, 2, st);
    }
    if ((rc = row_sampler_backward(G, vol, grid, gradVol, gradGrid, gradOut, og, ov, flags, st)) != -1) return rc;
    return launch_bwd_generic<float>(a, G, og, ov, st);
  }
  SamplerArgs<double> a;
  fill_common(a, vol, grid, gradOut);
  a.gout = static_cast<const double *>(gradOut->data);
  if (!og) {
    a.gvol = static_cast<double *>(gradVol->data);
    for (int i = 0; i < 4; i++) a.gvs[i] = gradVol->stride[i];
    a.gvs4 = gradVol->stride[4];
  }
  if (!ov) {
    a.ggrid = static_cast<double *>(gradGrid->data);
    for (int i = 0; i < 4; i++) a.ggs[i] = gradGrid->stride[i];
    a.ggs4 = gradGrid->stride[4];
  }
  return launch_bwd_generic<double>(a, G, og, ov, st);
}

int sampler_indices_impl(int dtype, int G, const int64_t vol_size[5], const volstn_tensor5 *grid, int32_t *idx,
                         uint8_t *mask, cudaStream_t st) {
  if (dtype != VOLSTN_F32 && dtype != VOLSTN_F64) return VOLSTN_ERR_ARG;
  if (G != 3 && G != 4) return VOLSTN_ERR_ARG;
  if (!vol_size || !idx || !mask) return VOLSTN_ERR_ARG;
  int rc = check_desc(grid);
  if (rc) return rc;
  if (grid->size[4] != G) return VOLSTN_ERR_SHAPE;
  volstn_tensor5 vol;
  vol.data = nullptr;
  for (int i = 0; i < 5; i++) {
    if (!fits_int(vol_size[i])) return VOLSTN_ERR_ARG;
    vol.size[i] = vol_size[i];
    vol.stride[i] = 0;
  }
  if (numel(grid) == 0) return VOLSTN_OK;
  if (dtype == VOLSTN_F32) {
    SamplerArgs<float> a;
    fill_common(a, &vol, grid, grid);
    a.B = (int)grid->size[0];
    dim3 blocks;
    if (!generic_grid(a, blocks)) return VOLSTN_ERR_UNSUPPORTED;
    k_sampler_indices<float><<<blocks, kThreads, 0, st>>>(a, idx, mask);
  } else {
    SamplerArgs<double> a;
    fill_common(a, &vol, grid, grid);
    a.B = (int)grid->size[0];
    dim3 blocks;
    if (!generic_grid(a, blocks)) return VOLSTN_ERR_UNSUPPORTED;
    k_sampler_indices<double><<<blocks, kThreads, 0, st>>>(a, idx, mask);
  }
  return after_launch();
}

}  // namespace volstn

// =============================================================================================
// C ABI (DEVICE pointers)
// =============================================================================================
extern "C" {

int volstn_b200_sampler_forward(int dtype, int G, const volstn_tensor5 *vol, const volstn_tensor5 *grid,
                                const volstn_tensor5 *out, unsigned flags, void *cuda_stream) {
  return volstn::sampler_forward_impl(dtype, G, vol, grid, out, flags, static_cast<cudaStream_t>(cuda_stream));
}

int volstn_b200_sampler_backward(int dtype, int G, const volstn_tensor5 *vol, const volstn_tensor5 *grid,
                                 const volstn_tensor5 *gradVol, const volstn_tensor5 *gradGrid,
                                 const volstn_tensor5 *gradOut, unsigned flags, void *cuda_stream) {
  return volstn::sampler_backward_impl(dtype, G, vol, grid, gradVol, gradGrid, gradOut, flags,
                                       static_cast<cudaStream_t>(cuda_stream));
}

int volstn_b200_sampler_indices(int dtype, int G, const int64_t vol_size[5], const volstn_tensor5 *grid,
                                int32_t *idx, uint8_t *mask, void *cuda_stream) {
  return volstn::sampler_indices_impl(dtype, G, vol_size, grid, idx, mask, static_cast<cudaStream_t>(cuda_stream));
}

}  // extern "C"
