// sampler_tile.cu -- brick-tiled sampler kernels for sm_100a (VOLSTN_ALGO_TILE, the automatic choice
// for packed float tensors with C <= 4).
//
// Why: measured on a B200 (profiles/r01_*.md), the one-thread-per-voxel kernels of sampler.cu are
// not HBM-bound on sheared warp fields.  A warp's 32 lanes land in ~20 different 32-byte sectors
// per corner, so (i) the gathers are bound by L1 sector look-ups and (ii) the gradVol scatter is
// bound by L2 reduction operations, whose cost is per touched SECTOR (~175 G sector-ops/s chip
// wide, tools/ubench_red.cu), not per byte: 5.2 sector-ops per voxel = 500 us at 64^3 x 64.
//
// Here one CTA owns an 8 x 8 x BZ brick of OUTPUT voxels (one thread per voxel):
//   1. every thread loads its grid element (one 128-bit streaming load) and does the coordinate
//      set-up of ...c:507-550 (bit-exact, volstn_common.cuh);
//   2. the CTA reduces the bounding box of the brick's in-bounds corner indices (REDUX + a few
//      shared-memory integer atomics);
//   3. if the box fits in shared memory the volume box is staged with row-coalesced loads and the
//      eight corner reads become shared-memory gathers (32 banks per cycle instead of one L1 tag
//      look-up per sector); otherwise the CTA falls back to the direct per-voxel path -- results
//      are identical either way;
//   4. backward only: gradVol contributions are accumulated in a second, privatised shared-memory
//      box (shared float atomicAdd) and flushed once per CTA with row-coalesced `red.global.add`,
//      skipping zero cells, so L2 sees ~0.3-1 reduction sector-ops per voxel instead of ~5.
// The arithmetic per voxel is the code of sampler.cuh: out and gradGrid stay bit-identical to the
// reference CPU build; gradVol differs only in summation order.
#include <atomic>

#include "sampler.cuh"
#include "volstn_internal.h"

namespace volstn {

namespace {

constexpr int kBX = 8, kBY = 8;  // brick extent in x and y; z extent is the template parameter BZ

struct BrickBox {
  int lo[3];   // z, y, x of the first cell of the box
  int ext[3];  // extents (<= 0 on any axis: no corner of the brick is in bounds)
};

// Bounding box of the in-bounds corners of all voxels of the CTA.  Corner indices along one axis are
// i0 and i0+1; the in-bounds ones lie in [max(i0,0), min(i0+1,size-1)].  s_box: lo z,y,x, hi z,y,x.
template <typename real>
__device__ __forceinline__ BrickBox brick_box(const VoxelSetup<real> &s, bool has, const SamplerArgs<real> &a,
                                              int *s_box) {
  if (threadIdx.x < 3) s_box[threadIdx.x] = INT32_MAX;
  else if (threadIdx.x < 6) s_box[threadIdx.x] = INT32_MIN;
  __syncthreads();
  int lo[3], hi[3];
  lo[0] = has ? max(s.z0, 0) : INT32_MAX;
  lo[1] = has ? max(s.y0, 0) : INT32_MAX;
  lo[2] = has ? max(s.x0, 0) : INT32_MAX;
  hi[0] = has ? min(wrap_add(s.z0, 1), a.iD - 1) : INT32_MIN;
  hi[1] = has ? min(wrap_add(s.y0, 1), a.iH - 1) : INT32_MIN;
  hi[2] = has ? min(wrap_add(s.x0, 1), a.iW - 1) : INT32_MIN;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    lo[i] = __reduce_min_sync(0xffffffffu, lo[i]);
    hi[i] = __reduce_max_sync(0xffffffffu, hi[i]);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
      if (lo[i] != INT32_MAX) atomicMin(&s_box[i], lo[i]);
      if (hi[i] != INT32_MIN) atomicMax(&s_box[3 + i], hi[i]);
    }
  }
  __syncthreads();
  BrickBox b;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    b.lo[i] = s_box[i];
    const int h = s_box[3 + i];
    b.ext[i] = (b.lo[i] == INT32_MAX || h == INT32_MIN) ? 0 : h - b.lo[i] + 1;
  }
  return b;
}

// cell-vector helpers for the shared-memory boxes (CT floats per cell, channels-last like the volume)
template <int CT>
__device__ __forceinline__ void lds_cell(const float *p, float (&v)[CT]) {
  if constexpr (CT == 4) {
    const float4 t = *reinterpret_cast<const float4 *>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (CT == 2) {
    const float2 t = *reinterpret_cast<const float2 *>(p);
    v[0] = t.x; v[1] = t.y;
  } else {
#pragma unroll
    for (int t = 0; t < CT; t++) v[t] = p[t];
  }
}
template <int CT>
__device__ __forceinline__ void sts_cell(float *p, const float (&v)[CT]) {
  if constexpr (CT == 4) {
    *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else if constexpr (CT == 2) {
    *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
  } else {
#pragma unroll
    for (int t = 0; t < CT; t++) p[t] = v[t];
  }
}

// Row-coalesced walk over the cells of a box: calls f(cell_index_in_tile, element_offset_in_sample)
// for every cell.  A row of the box is `ex` consecutive cells in global memory; 8, 16 or 32 lanes
// share a row so that short rows do not idle most of the warp.
template <int NT, typename F>
__device__ __forceinline__ void for_each_box_cell(const BrickBox &b, int pitch, int sD, int sH, int sW, F f) {
  const int ex = b.ext[2], ey = b.ext[1];
  const int rows = b.ext[0] * ey;
  const int lpr = ex <= 8 ? 8 : (ex <= 16 ? 16 : 32);  // lanes per row
  const int rpi = NT / lpr;                             // rows per CTA iteration
  const int li = threadIdx.x & (lpr - 1);
  for (int r = threadIdx.x / lpr; r < rows; r += rpi) {
    const int zz = r / ey, yy = r - zz * ey;
    const int goff = (b.lo[0] + zz) * sD + (b.lo[1] + yy) * sH + b.lo[2] * sW;
    const int toff = r * pitch;
    for (int i = li; i < ex; i += lpr) f(toff + i, goff + i * sW);
  }
}

template <int BZ>
__device__ __forceinline__ bool brick_voxel(const SamplerArgs<float> &a, int nbx, int nby, long long &j) {
  const int brick = blockIdx.x;
  const int bx = brick % nbx, t = brick / nbx;
  const int by = t % nby, bz = t / nby;
  const int x = bx * kBX + (threadIdx.x & 7);
  const int y = by * kBY + ((threadIdx.x >> 3) & 7);
  const int z = bz * BZ + (threadIdx.x >> 6);
  j = ((long long)z * a.oH + y) * a.oW + x;
  return x < a.oW && y < a.oH && z < a.oD;
}

}  // namespace

// =============================================================================================
// forward
// =============================================================================================
template <int CT, int G, int BZ>
__global__ void __launch_bounds__(BZ * 64, BZ == 8 ? 2 : 4)
    k_sampler_fwd_tile(const SamplerArgs<float> a, int nbx, int nby, int cap_cells) {
  using E = Ex<float>;
  constexpr int NT = BZ * 64;
  extern __shared__ __align__(16) float s_vol[];
  __shared__ int s_box[6];
  const int sW = CT, sH = a.iW * CT, sD = a.iH * a.iW * CT;
  long long j;
  const bool valid = brick_voxel<BZ>(a, nbx, nby, j);
  for (int b = blockIdx.y; b < a.B; b += gridDim.y) {
    const float *vb = a.vol + (long long)b * a.vs[0];
    VoxelSetup<float> s;
    s.init_idle();
    if (valid) {
      float zf, yf, xf;
      ld_grid<float, G>(a.grid + ((long long)b * a.n + j) * G, zf, yf, xf);
      s.init(zf, yf, xf, a);
    }
    const BrickBox box = brick_box<float>(s, valid && s.any_in(), a, s_box);
    const int pitch = box.ext[2] | 1;  // odd row pitch spreads rows over the banks
    const long long cells = (long long)box.ext[0] * box.ext[1] * pitch;
    float *o = a.out + ((long long)b * a.n + j) * CT;
    if (cells <= 0) {
      // nothing of the brick is in bounds.  The result is w * 0 summed, which is 0 unless a weight is
      // inf/nan (non-finite grid): the predicated per-voxel path reproduces that exactly, without loads.
      if (valid) voxel_forward_packed<float, CT, G>(a, s, vb, sD, sH, sW, o);
    } else if (cells <= cap_cells) {
      for_each_box_cell<NT>(box, pitch, sD, sH, sW, [&](int tc, int go) {
        float v[CT];
        ld_channels<float, CT>(vb + go, v);
        sts_cell<CT>(s_vol + tc * CT, v);
      });
      __syncthreads();
      if (valid) {
        float cw[8];
        s.corner_weights(cw);
        const int tz = box.ext[1] * pitch;
        const int tb = (s.z0 - box.lo[0]) * tz + (s.y0 - box.lo[1]) * pitch + (s.x0 - box.lo[2]);
        float v[8][CT];
#pragma unroll
        for (int k = 0; k < 8; k++) {
#pragma unroll
          for (int t = 0; t < CT; t++) v[k][t] = 0.f;
          if (s.in(k)) lds_cell<CT>(s_vol + (tb + ((k & 1) ? 1 : 0) + ((k & 2) ? pitch : 0) + ((k & 4) ? tz : 0)) * CT, v[k]);
        }
        float acc[CT];
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
          const int k = (G == 4) ? order_xyz(jj) : order_zxy(jj);  // ...c:566-573 / ...c:149-156
#pragma unroll
          for (int t = 0; t < CT; t++) {
            const float term = E::mul(cw[k], v[k][t]);
            acc[t] = (jj == 0) ? term : E::add(acc[t], term);
          }
        }
        st_channels_stream<float, CT>(o, acc);
      }
    } else if (valid) {
      voxel_forward_packed<float, CT, G>(a, s, vb, sD, sH, sW, o);
    }
    if (b + (int)gridDim.y < a.B) __syncthreads();  // the box is reused by the next sample
  }
}

// =============================================================================================
// backward
// =============================================================================================
template <int CT, int G, int BZ, bool ONLY_GRID, bool ONLY_VOL>
__global__ void __launch_bounds__(BZ * 64, BZ == 8 ? 2 : 4)
    k_sampler_bwd_tile(const SamplerArgs<float> a, int nbx, int nby, int cap_cells) {
  using E = Ex<float>;
  constexpr int NT = BZ * 64;
  extern __shared__ __align__(16) float s_mem[];
  __shared__ int s_box[6];
  // layout: [ volume box | gradVol box ], each cap_cells * CT floats (only the ones needed exist)
  float *s_vol = s_mem;
  float *s_gv = ONLY_VOL ? s_mem : s_mem + (size_t)cap_cells * CT;
  const int sW = CT, sH = a.iW * CT, sD = a.iH * a.iW * CT;
  long long j;
  const bool valid = brick_voxel<BZ>(a, nbx, nby, j);
  for (int b = blockIdx.y; b < a.B; b += gridDim.y) {
    const float *vb = a.vol + (long long)b * a.vs[0];
    float *gvb = a.gvol + (long long)b * a.gvs[0];
    const float *gop = a.gout + ((long long)b * a.n + j) * CT;
    float *ggp = a.ggrid + ((long long)b * a.n + j) * G;
    VoxelSetup<float> s;
    s.init_idle();
    float go[CT];
#pragma unroll
    for (int t = 0; t < CT; t++) go[t] = 0.f;
    if (valid) {
      float zf, yf, xf;
      ld_grid<float, G>(a.grid + ((long long)b * a.n + j) * G, zf, yf, xf);
      ld_channels_stream<float, CT>(gop, go);
      s.init(zf, yf, xf, a);
    }
    const BrickBox box = brick_box<float>(s, valid && s.any_in(), a, s_box);
    const int pitch = box.ext[2] | 1;
    const long long cells = (long long)box.ext[0] * box.ext[1] * pitch;
    if (cells <= 0) {
      // every corner of every voxel is out of bounds: no loads, no reductions; gradGrid = 0 * weights
      voxel_backward_ct<float, CT, G, ONLY_GRID, ONLY_VOL, 0>(a, s, vb, gvb - vb, sD, sH, go, ggp, valid);
    } else if (cells <= cap_cells) {
      if constexpr (!ONLY_GRID) {
        for (int i = threadIdx.x; i < (int)cells * CT; i += NT) s_gv[i] = 0.f;
      }
      if constexpr (!ONLY_VOL) {
        for_each_box_cell<NT>(box, pitch, sD, sH, sW, [&](int tc, int go_) {
          float v[CT];
          ld_channels<float, CT>(vb + go_, v);
          sts_cell<CT>(s_vol + tc * CT, v);
        });
      }
      __syncthreads();
      if (valid) {
        const int tz = box.ext[1] * pitch;
        const int tb = (s.z0 - box.lo[0]) * tz + (s.y0 - box.lo[1]) * pitch + (s.x0 - box.lo[2]);
        float cw[8];
        if constexpr (!ONLY_GRID) s.corner_weights(cw);
        float dot[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const int tc = (tb + ((k & 1) ? 1 : 0) + ((k & 2) ? pitch : 0) + ((k & 4) ? tz : 0)) * CT;
          dot[k] = 0.f;
          if (s.in(k)) {
            if constexpr (!ONLY_VOL) {
              float v[CT];
              lds_cell<CT>(s_vol + tc, v);
              float d = 0.f;  // dot += v * go, channel by channel (...c:736)
#pragma unroll
              for (int t = 0; t < CT; t++) d = E::add(d, E::mul(v[t], go[t]));
              dot[k] = d;
            }
            if constexpr (!ONLY_GRID) {
#pragma unroll
              for (int t = 0; t < CT; t++) atomicAdd(s_gv + tc + t, E::mul(cw[k], go[t]));  // ...c:737
            }
          }
        }
        if constexpr (!ONLY_VOL) {
          float gz, gy, gx;
          grad_grid_from_dots<float>(dot, s, a, gz, gy, gx);
          st_grad_grid<float, G>(ggp, gz, gy, gx);
        }
      }
      if constexpr (!ONLY_GRID) {
        __syncthreads();
        // flush: one reduction per non-zero cell; lanes run along x, so the L2 sees one operation per
        // touched sector.  Adding an exact zero is the identity, so skipping it changes nothing.
        for_each_box_cell<NT>(box, pitch, sD, sH, sW, [&](int tc, int go_) {
          float v[CT];
          lds_cell<CT>(s_gv + tc * CT, v);
          bool nz = false;
#pragma unroll
          for (int t = 0; t < CT; t++) nz |= (v[t] != 0.f);
          if (nz) red_channels<float, CT>(gvb + go_, v);
        });
      }
    } else {
      voxel_backward_ct<float, CT, G, ONLY_GRID, ONLY_VOL, 0>(a, s, vb, gvb - vb, sD, sH, go, ggp, valid);
    }
    if (b + (int)gridDim.y < a.B) __syncthreads();
  }
}

// =============================================================================================
// launchers
// =============================================================================================
namespace {

constexpr int kSmemBudget = 226 * 1024;  // per SM, leaving the 1 KB per-CTA reservation some room

template <typename K>
int set_smem(K kernel, size_t bytes, std::atomic<size_t> &have) {
  if (bytes > 48 * 1024 && have.load(std::memory_order_relaxed) < bytes) {
    VOLSTN_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    have.store(bytes, std::memory_order_relaxed);
  }
  return VOLSTN_OK;
}

// cells per shared-memory box such that `ctas` CTAs with `boxes` boxes of CT floats per cell fit an SM
int cap_for(int ctas, int boxes, int CT) {
  const int per_cta = kSmemBudget / ctas - 1024 - 64;
  return per_cta / (boxes * CT * 4);
}

template <int CT, int G, int BZ>
int launch_fwd_tile_t(const SamplerArgs<float> &a, cudaStream_t st) {
  static std::atomic<size_t> have{0};
  const int nbx = (a.oW + kBX - 1) / kBX, nby = (a.oH + kBY - 1) / kBY, nbz = (a.oD + BZ - 1) / BZ;
  const long long bricks = (long long)nbx * nby * nbz;
  if (bricks > 0x7fffffffLL) return VOLSTN_ERR_UNSUPPORTED;
  const int cap = cap_for(BZ == 8 ? 2 : 4, 1, CT);
  const size_t smem = (size_t)cap * CT * 4;
  int rc = set_smem(k_sampler_fwd_tile<CT, G, BZ>, smem, have);
  if (rc) return rc;
  const unsigned by = (unsigned)(a.B < 65535 ? a.B : 65535);
  k_sampler_fwd_tile<CT, G, BZ><<<dim3((unsigned)bricks, by), BZ * 64, smem, st>>>(a, nbx, nby, cap);
  return after_launch();
}

template <int CT, int G, int BZ, bool OG, bool OV>
int launch_bwd_tile_t(const SamplerArgs<float> &a, cudaStream_t st) {
  static std::atomic<size_t> have{0};
  const int nbx = (a.oW + kBX - 1) / kBX, nby = (a.oH + kBY - 1) / kBY, nbz = (a.oD + BZ - 1) / BZ;
  const long long bricks = (long long)nbx * nby * nbz;
  if (bricks > 0x7fffffffLL) return VOLSTN_ERR_UNSUPPORTED;
  const int boxes = (OG || OV) ? 1 : 2;
  const int cap = cap_for(BZ == 8 ? 2 : 4, boxes, CT);
  const size_t smem = (size_t)cap * CT * 4 * boxes;
  int rc = set_smem(k_sampler_bwd_tile<CT, G, BZ, OG, OV>, smem, have);
  if (rc) return rc;
  const unsigned by = (unsigned)(a.B < 65535 ? a.B : 65535);
  k_sampler_bwd_tile<CT, G, BZ, OG, OV><<<dim3((unsigned)bricks, by), BZ * 64, smem, st>>>(a, nbx, nby, cap);
  return after_launch();
}

template <int CT, int G>
int launch_fwd_tile_c(const SamplerArgs<float> &a, cudaStream_t st) {
  if constexpr (CT <= 2) return launch_fwd_tile_t<CT, G, 8>(a, st);
  return launch_fwd_tile_t<CT, G, 4>(a, st);
}

template <int CT, int G, bool OG, bool OV>
int launch_bwd_tile_c(const SamplerArgs<float> &a, cudaStream_t st) {
  if constexpr (CT <= 2) return launch_bwd_tile_t<CT, G, 8, OG, OV>(a, st);
  return launch_bwd_tile_t<CT, G, 4, OG, OV>(a, st);
}

template <int G, bool OG, bool OV>
int launch_bwd_tile_g(const SamplerArgs<float> &a, cudaStream_t st) {
  switch (a.C) {
    case 1: return launch_bwd_tile_c<1, G, OG, OV>(a, st);
    case 2: return launch_bwd_tile_c<2, G, OG, OV>(a, st);
    case 3: return launch_bwd_tile_c<3, G, OG, OV>(a, st);
    case 4: return launch_bwd_tile_c<4, G, OG, OV>(a, st);
    default: return VOLSTN_ERR_UNSUPPORTED;
  }
}

template <int G>
int launch_fwd_tile_g(const SamplerArgs<float> &a, cudaStream_t st) {
  switch (a.C) {
    case 1: return launch_fwd_tile_c<1, G>(a, st);
    case 2: return launch_fwd_tile_c<2, G>(a, st);
    case 3: return launch_fwd_tile_c<3, G>(a, st);
    case 4: return launch_fwd_tile_c<4, G>(a, st);
    default: return VOLSTN_ERR_UNSUPPORTED;
  }
}

}  // namespace

bool tile_supported(int C) { return C >= 1 && C <= 4; }

int launch_fwd_tile(const SamplerArgs<float> &a, int G, cudaStream_t st) {
  return G == 4 ? launch_fwd_tile_g<4>(a, st) : launch_fwd_tile_g<3>(a, st);
}

int launch_bwd_tile(const SamplerArgs<float> &a, int G, bool og, bool ov, cudaStream_t st) {
  if (G == 4) {
    if (og) return launch_bwd_tile_g<4, true, false>(a, st);
    if (ov) return launch_bwd_tile_g<4, false, true>(a, st);
    return launch_bwd_tile_g<4, false, false>(a, st);
  }
  if (og) return launch_bwd_tile_g<3, true, false>(a, st);
  if (ov) return launch_bwd_tile_g<3, false, true>(a, st);
  return launch_bwd_tile_g<3, false, false>(a, st);
}

}  // namespace volstn
