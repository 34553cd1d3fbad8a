// sampler_tile.cu -- backward sampler with the gradVol scatter PRIVATISED IN SHARED MEMORY
// (VOLSTN_ALGO_TILE): for warp fields that shear.
//
// Why (all measured on a B200, profiles/r01_*): the L2 executes `red.global.add` at ~175 G operations per
// second, where one operation = one 32-byte sector touched by one instruction -- whether 1 or 8 of its
// lanes hit that sector (tools/ubench_red.cu).  With the identity field a warp's 32 lanes fall in 4-5
// sectors per corner and the direct kernel (sampler.cu) pays 1.2 sector-ops per voxel; with a sheared
// field (jittered ALIGNet cage, rotations) every lane lands in its own sector: 5.2 sector-ops per voxel,
// 480 us at 64^3 x 64, three times the rest of the kernel.
//
// Here one CTA owns an 8 x 8 x 4 brick of OUTPUT voxels (one thread per voxel):
//   1. set-up per voxel exactly as in the direct kernel (bit-identical gradGrid, same dot products);
//   2. the CTA reduces the bounding box of its in-bounds corner indices (REDUX + six shared atomics) and
//      aligns it to 16 bytes in x;
//   3. contributions are accumulated into a shared-memory copy of that box in FIXED POINT with native
//      integer shared atomics (ATOMS.ADD).  A shared float atomicAdd is a CAS loop on sm_100a: with the
//      bank conflicts and same-address retries of a scatter it cost ~12 shared-memory wavefronts per
//      corner and made this kernel slower than the one it replaces (gpurun_out/r1_kbench_tile4.txt).
//      The scale is a power of two chosen per CTA from max|gradOut| of the brick (2^21 steps up to the
//      next power of two above it), so one contribution is rounded by at most 2^-22 of that maximum --
//      the order of an fp32 rounding of a full-size contribution -- and 256 of them cannot overflow.
//      gradVol's contract is a tolerance (1e-4 of max|ref|, summation order is free); measured error
//      stays ~1e-6.  Bricks whose gradOut is not finite take the direct float path.  Every touched
//      (z, y) row is flagged;
//   4. each flagged row is flushed with 16-byte `red.global.add.v4.f32`, lanes along x: ~2 sector-ops per
//      row instead of one per lane per corner.
// A brick whose box does not fit the shared-memory budget (or any tensor the alignment rules exclude)
// falls back to the direct reductions, CTA by CTA: the result is the same sum in a different order.
#include <atomic>

#include "sampler.cuh"
#include "volstn_internal.h"

namespace volstn {

namespace {

constexpr int kBX = 8, kBY = 8, kBZ = 4;  // output brick
constexpr int kTileThreads = kBX * kBY * kBZ;

struct BrickBox {
  int lo[3];   // z, y, x of the first cell of the box
  int ext[3];  // extents (0 on every axis: no corner of the brick is in bounds)
};

// Bounding box of the in-bounds corners of all voxels of the CTA.  Corner indices along one axis are
// i0 and i0+1; the in-bounds ones lie in [max(i0,0), min(i0+1,size-1)].  s_box: lo z,y,x, hi z,y,x.
__device__ __forceinline__ BrickBox brick_box(const VoxelSetup<float> &s, bool has, const SamplerArgs<float> &a,
                                              int *s_box, unsigned go_abs_bits, unsigned &go_max_bits) {
  if (threadIdx.x < 3) s_box[threadIdx.x] = INT32_MAX;
  else if (threadIdx.x < 6) s_box[threadIdx.x] = INT32_MIN;
  else if (threadIdx.x == 6) s_box[6] = 0;
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
  // |x| of a float orders like its bit pattern (sign cleared); inf / nan have the largest patterns
  const unsigned gm = __reduce_max_sync(0xffffffffu, go_abs_bits);
  if ((threadIdx.x & 31) == 0) {
    if (gm) atomicMax(reinterpret_cast<unsigned *>(&s_box[6]), gm);
#pragma unroll
    for (int i = 0; i < 3; i++) {
      if (lo[i] != INT32_MAX) atomicMin(&s_box[i], lo[i]);
      if (hi[i] != INT32_MIN) atomicMax(&s_box[3 + i], hi[i]);
    }
  }
  __syncthreads();
  go_max_bits = (unsigned)s_box[6];
  BrickBox b;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    b.lo[i] = s_box[i];
    const int h = s_box[3 + i];
    b.ext[i] = (b.lo[i] == INT32_MAX || h == INT32_MIN) ? 0 : h - b.lo[i] + 1;
  }
  return b;
}

}  // namespace

// CT floats per cell.  XA = cells per 16 bytes (4 / CT): the box is aligned to XA cells in x.
template <int CT, int G, bool ONLY_VOL>
__global__ void __launch_bounds__(kTileThreads, 5)
    k_sampler_bwd_tile(const SamplerArgs<float> a, int nbx, int nby, int cap_floats, int max_rows) {
  using E = Ex<float>;
  constexpr int XA = 4 / CT;
  extern __shared__ __align__(128) int s_gv[];                                    // cap_floats fixed-point cells
  unsigned char *s_row = reinterpret_cast<unsigned char *>(s_gv + cap_floats);  // max_rows flags
  __shared__ int s_box[8];
  const int sW = CT, sH = a.iW * CT, sD = a.iH * a.iW * CT;

  const int brick = blockIdx.x;
  const int bx = brick % nbx, bt = brick / nbx;
  const int by = bt % nby, bz = bt / nby;
  const int x = bx * kBX + (threadIdx.x & 7), y = by * kBY + ((threadIdx.x >> 3) & 7), z = bz * kBZ + (threadIdx.x >> 6);
  const bool valid = x < a.oW && y < a.oH && z < a.oD;
  const size_t j = ((size_t)z * a.oH + y) * a.oW + x;

  for (unsigned b = blockIdx.y; b < (unsigned)a.B; b += gridDim.y) {
    const size_t sb = (size_t)b * a.n32;
    const float *vb = a.vol + (size_t)b * a.vs32;
    float *gvb = a.gvol + (size_t)b * a.vs32;
    float *ggp = a.ggrid + (sb + j) * G;
    VoxelSetup<float> s;
    s.init_idle();
    float go[CT];
#pragma unroll
    for (int t = 0; t < CT; t++) go[t] = 0.f;
    if (valid) {
      float zf, yf, xf;
      ld_grid<float, G>(a.grid + (sb + j) * G, zf, yf, xf);
      ld_channels_stream<float, CT>(a.gout + (sb + j) * CT, go);
      s.init(zf, yf, xf, a);
    }
    unsigned go_abs = 0;
#pragma unroll
    for (int t = 0; t < CT; t++) go_abs = max(go_abs, (unsigned)__float_as_int(go[t]) & 0x7fffffffu);
    unsigned go_max;
    const BrickBox box = brick_box(s, valid && s.any_in(), a, s_box, go_abs, go_max);
    // 16-byte alignment of every row: start on a multiple of XA cells, extent a multiple of XA cells
    // (the host guarantees iW % XA == 0, so the rounded-up end stays inside the row)
    const int x_lo = box.lo[2] & ~(XA - 1);
    const int pitch = box.ext[2] > 0 ? ((box.lo[2] + box.ext[2] + XA - 1) & ~(XA - 1)) - x_lo : 0;
    const int rows = box.ext[0] * box.ext[1];
    const long long floats = (long long)rows * pitch * CT;

    if (floats > 0 && floats <= cap_floats && rows <= max_rows && go_max < 0x7f800000u) {
      // ---- privatised accumulation ----
      // scale = 2^21 / (the power of two above max|go|), as exponent arithmetic; clamped to normal floats
      const int e_s = min(253, max(1, 274 - (int)(go_max >> 23)));
      const float scale = __int_as_float(e_s << 23), inv_scale = __int_as_float((254 - e_s) << 23);
      for (int i = threadIdx.x; i < (int)floats / 4; i += kTileThreads)
        reinterpret_cast<int4 *>(s_gv)[i] = make_int4(0, 0, 0, 0);
      for (int i = threadIdx.x; i < rows; i += kTileThreads) s_row[i] = 0;
      __syncthreads();
      float cw[8];
      s.corner_weights(cw);
      const CornerRows<float> vrows(vb, s.base32(sD, sH, sW), sD, sH);
      const int tr = (s.z0 - box.lo[0]) * box.ext[1] + (s.y0 - box.lo[1]);  // tile row of corner 0
      const int tx = s.x0 - x_lo;
      float dot[8];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        dot[k] = 0.f;
        if (s.in(k)) {
          if constexpr (!ONLY_VOL) {
            float v[CT];
            ld_channels<float, CT>(vrows.template corner<CT>(k), v);
            float d = 0.f;  // dot += v * go, channel by channel (...c:736)
#pragma unroll
            for (int t = 0; t < CT; t++) d = E::add(d, E::mul(v[t], go[t]));
            dot[k] = d;
          }
          const int r = tr + ((k & 2) ? 1 : 0) + ((k & 4) ? box.ext[1] : 0);
          int *cell = s_gv + ((size_t)r * pitch + tx + (k & 1)) * CT;
          bool any = false;
#pragma unroll
          for (int t = 0; t < CT; t++) {
            const float c = E::mul(cw[k], go[t]);  // ...c:737, the reference's product
            // round(c * scale) without a conversion instruction: |c * scale| < 2^21, so adding 1.5 * 2^23
            // leaves the integer in the mantissa
            const int q = __float_as_int(__fmaf_rn(c, scale, 12582912.0f)) - 0x4B400000;
            if (q != 0) {
              atomicAdd(cell + t, q);
              any = true;
            }
          }
          if (any) s_row[r] = 1;  // benign race: every writer stores the same value
        }
      }
      if constexpr (!ONLY_VOL) {
        float gz, gy, gx;
        grad_grid_from_dots<float>(dot, s, a, gz, gy, gx);
        if (valid) st_grad_grid<float, G>(ggp, gz, gy, gx);
      }
      // ---- flush: 16-byte vector reductions, lanes along x, a power-of-two group of lanes per row ----
      // (One `cp.reduce.async.bulk ... .add.f32` (UBLKRED) per row was measured first: its operands are
      // uniform registers, so ptxas serialises the rows lane by lane -- 3x the instructions of the whole
      // kernel, gpurun_out/r1_kbench_tile3.txt.)
      __syncthreads();
      const int qpr = pitch * CT / 4;                                   // 16-byte groups per row
      const int lpr = qpr <= 4 ? 4 : (qpr <= 8 ? 8 : (qpr <= 16 ? 16 : 32));  // lanes serving one row
      const int qi = threadIdx.x & (lpr - 1);
      const float inv_ey = 1.0f / (float)box.ext[1];
      for (int r = threadIdx.x / lpr; r < rows; r += kTileThreads / lpr) {
        if (!s_row[r]) continue;
        const int zz = (int)(((float)r + 0.5f) * inv_ey), yy = r - zz * box.ext[1];  // exact for r < 2^20
        float *g = gvb + (size_t)(box.lo[0] + zz) * sD + (size_t)(box.lo[1] + yy) * sH + (size_t)x_lo * CT;
        for (int q = qi; q < qpr; q += lpr) {
          const int4 v = reinterpret_cast<const int4 *>(s_gv + (size_t)r * pitch * CT)[q];
          if (v.x | v.y | v.z | v.w)
            red_add4(g + 4 * q, (float)v.x * inv_scale, (float)v.y * inv_scale, (float)v.z * inv_scale,
                     (float)v.w * inv_scale);
        }
      }
    } else {
      // nothing in bounds, or the box does not fit: the direct path (predicated, no loads for excluded corners)
      voxel_backward_ct<float, CT, G, false, ONLY_VOL, 0>(a, s, vb, gvb - vb, sD, sH, go, ggp, valid);
    }
    if (b + gridDim.y < (unsigned)a.B) __syncthreads();  // the box is reused by the next sample
  }
}

// =============================================================================================
// launcher
// =============================================================================================
namespace {

constexpr int kSmemBudget = 226 * 1024;  // per SM
constexpr int kTileCtasPerSm = 5;

template <int CT, int G, bool OV>
int launch_bwd_tile_t(const SamplerArgs<float> &a, cudaStream_t st) {
  static std::atomic<size_t> have{0};
  const int nbx = (a.oW + kBX - 1) / kBX, nby = (a.oH + kBY - 1) / kBY, nbz = (a.oD + kBZ - 1) / kBZ;
  const long long bricks = (long long)nbx * nby * nbz;
  if (bricks > 0x7fffffffLL) return VOLSTN_ERR_UNSUPPORTED;
  const int per_cta = kSmemBudget / kTileCtasPerSm - 1024 - 128;
  // cap_floats floats + one flag byte per row; a row has at least 4 floats
  int cap_floats = (int)((per_cta / 17LL) * 16 / 4);
  cap_floats &= ~31;
  const int max_rows = cap_floats / 4;
  const size_t smem = (size_t)cap_floats * 4 + (size_t)max_rows;
  if (smem > 48 * 1024 && have.load(std::memory_order_relaxed) < smem) {
    VOLSTN_CUDA_TRY(cudaFuncSetAttribute(k_sampler_bwd_tile<CT, G, OV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
    have.store(smem, std::memory_order_relaxed);
  }
  const unsigned by = (unsigned)(a.B < 65535 ? a.B : 65535);
  k_sampler_bwd_tile<CT, G, OV><<<dim3((unsigned)bricks, by), kTileThreads, smem, st>>>(a, nbx, nby, cap_floats,
                                                                                       max_rows);
  return after_launch();
}

template <int G, bool OV>
int launch_bwd_tile_g(const SamplerArgs<float> &a, cudaStream_t st) {
  switch (a.C) {
    case 1: return launch_bwd_tile_t<1, G, OV>(a, st);
    case 2: return launch_bwd_tile_t<2, G, OV>(a, st);
    case 4: return launch_bwd_tile_t<4, G, OV>(a, st);
    default: return VOLSTN_ERR_UNSUPPORTED;
  }
}

}  // namespace

// 16-byte rows: C in {1, 2, 4}, the volume's W a multiple of 4 / C cells, gradVol 16-byte aligned
bool tile_supported(const SamplerArgs<float> &a) {
  if (a.C != 1 && a.C != 2 && a.C != 4) return false;
  if ((a.iW * a.C) % 4 != 0) return false;
  if (reinterpret_cast<uintptr_t>(a.gvol) % 16 != 0) return false;
  return true;
}

int launch_bwd_tile(const SamplerArgs<float> &a, int G, bool only_vol, cudaStream_t st) {
  if (G == 4) return only_vol ? launch_bwd_tile_g<4, true>(a, st) : launch_bwd_tile_g<4, false>(a, st);
  return only_vol ? launch_bwd_tile_g<3, true>(a, st) : launch_bwd_tile_g<3, false>(a, st);
}

}  // namespace volstn
