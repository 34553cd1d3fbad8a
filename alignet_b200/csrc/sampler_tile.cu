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
//   1. set-up per voxel with the conversion-free path of the direct kernel (bit-identical gradGrid and
//      dot products).  The privatised path is taken when every voxel of the brick lies inside the volume
//      (tier >= 1, i.e. 0 <= c <= size-1), gradOut is finite and the box fits; any other brick runs the
//      direct code.  (Privatising the bricks that straddle a face as well, with per-corner predicates, was
//      measured: slower overall, gpurun_out/r1_kbench_tile7.txt.);
//   2. the CTA reduces the bounding box of the brick's corner indices [i0, i0+1] (REDUX + six shared
//      atomics).  The box may stick out of the volume by one cell at a high face: those cells only ever
//      receive exact zeros (their weight is 0) and are not flushed, so no corner needs a predicate;
//   3. contributions are accumulated into a shared-memory copy of the box in FIXED POINT with native
//      integer shared atomics (ATOMS.ADD).  (A shared float atomicAdd is a CAS loop on sm_100a: with the
//      bank conflicts and same-address retries of a scatter it cost ~12 shared-memory wavefronts per
//      corner.)  The scale is a power of two chosen per CTA from max|gradOut| of the brick (2^21 steps up
//      to the next power of two above it): one contribution is rounded by at most 2^-22 of that maximum
//      -- the order of an fp32 rounding of a full-size contribution -- and 256 of them cannot overflow.
//      gradVol's contract is a tolerance (1e-4 of max|ref|; its summation order was never defined);
//      measured error ~1e-6;
//   4. the box is flushed with 16-byte `red.global.add.v4.f32`, lanes along x, skipping all-zero groups.
//
// Status [B200, 64^3 x 64, C = 1]: 10-17 % faster than the direct kernel on sheared fields (440 vs 490 us
// on the jittered cage, 419 vs 507 us on rotations), 1.7x slower on the identity field (334 vs 192 us) and
// on uniformly random fields: 580 instructions per voxel-warp against 270, and its eight scattered gathers
// keep the L1 at 71 %.  Opt-in only.
#include <atomic>

#include "sampler.cuh"
#include "volstn_internal.h"

namespace volstn {

namespace {

constexpr int kBX = 8, kBY = 8, kBZ = 4;  // output brick
constexpr int kTileThreads = kBX * kBY * kBZ;

}  // namespace

// CT floats per cell.  XA = cells per 16 bytes (4 / CT): the box is aligned to XA cells in x.
// s_meta: [0..2] box lo z,y,x  [3..5] box hi z,y,x  [6] max |gradOut| bits  [7] min tier
template <int CT, int G, bool ONLY_VOL>
__global__ void __launch_bounds__(kTileThreads, 5)
    k_sampler_bwd_tile(const SamplerArgs<float> a, int nbz, int cap_floats) {
  using E = Ex<float>;
  constexpr int XA = 4 / CT;
  extern __shared__ __align__(128) int s_gv[];  // cap_floats fixed-point cells
  __shared__ int s_meta[8];
  const int sH = a.iW * CT, sD = a.iH * a.iW * CT;

  const unsigned b = blockIdx.z / nbz, bz = blockIdx.z - b * nbz;
  const int x = blockIdx.x * kBX + (threadIdx.x & 7), y = blockIdx.y * kBY + ((threadIdx.x >> 3) & 7),
            z = bz * kBZ + (threadIdx.x >> 6);
  const bool valid = x < a.oW && y < a.oH && z < a.oD;
  // tail threads re-read the brick's first voxel: no divergent set-up, they only skip stores and atomics
  const unsigned j = valid ? ((unsigned)z * a.oH + y) * a.oW + x
                           : ((unsigned)(bz * kBZ) * a.oH + blockIdx.y * kBY) * a.oW + blockIdx.x * kBX;
  const size_t sb = (size_t)b * a.n32;
  const float *vb = opaque(a.vol + (size_t)b * a.vs32);
  const long long gv_delta = a.gvol - a.vol;
  float *ggp = a.ggrid + (sb + j) * G;

  if (threadIdx.x < 3) s_meta[threadIdx.x] = INT32_MAX;
  else if (threadIdx.x < 6) s_meta[threadIdx.x] = INT32_MIN;
  else if (threadIdx.x == 6) s_meta[6] = 0;
  else if (threadIdx.x == 7) s_meta[7] = 2;

  float zf, yf, xf, go[CT];
  ld_grid<float, G>(a.grid + (sb + j) * G, zf, yf, xf);
  ld_channels_stream<float, CT>(a.gout + (sb + j) * CT, go);
  if (!valid) {
#pragma unroll
    for (int t = 0; t < CT; t++) go[t] = 0.f;
  }
  VoxelSetup<float> s;
  const int tier = s.begin(zf, yf, xf, a);
  unsigned go_abs = 0;
#pragma unroll
  for (int t = 0; t < CT; t++) go_abs = max(go_abs, (unsigned)__float_as_int(go[t]) & 0x7fffffffu);
  __syncthreads();
  // ---- CTA reductions: tier, max|go|, bounding box ----
  const int wtier = __reduce_min_sync(0xffffffffu, tier);
  const unsigned wgo = __reduce_max_sync(0xffffffffu, go_abs);  // |x| orders like its bit pattern; inf/nan largest
  if (wtier >= 1) s.template finish_fast<true>(a, sD, sH, CT);  // warp-uniform; gives x0, y0, z0, weights, fbase
  int lo[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, hi[3] = {INT32_MIN, INT32_MIN, INT32_MIN};
  if (wtier >= 1) {
    lo[0] = __reduce_min_sync(0xffffffffu, s.z0), hi[0] = __reduce_max_sync(0xffffffffu, s.z0);
    lo[1] = __reduce_min_sync(0xffffffffu, s.y0), hi[1] = __reduce_max_sync(0xffffffffu, s.y0);
    lo[2] = __reduce_min_sync(0xffffffffu, s.x0), hi[2] = __reduce_max_sync(0xffffffffu, s.x0);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&s_meta[7], wtier);
    if (wgo) atomicMax(reinterpret_cast<unsigned *>(&s_meta[6]), wgo);
    if (wtier >= 1) {
#pragma unroll
      for (int i = 0; i < 3; i++) {
        atomicMin(&s_meta[i], lo[i]);
        atomicMax(&s_meta[3 + i], hi[i]);
      }
    }
  }
  __syncthreads();
  const unsigned go_max = (unsigned)s_meta[6];
  // box of corner indices: [min i0, max i0 + 1]; x aligned to 16 bytes
  const int z_lo = s_meta[0], y_lo = s_meta[1], ey = s_meta[4] - y_lo + 2, ez = s_meta[3] - z_lo + 2;
  const int x_lo = s_meta[2] & ~(XA - 1);
  const int pitch = ((s_meta[5] + 2 + XA - 1) & ~(XA - 1)) - x_lo;  // cells per tile row
  const int rows = ez * ey;
  const bool tiled = s_meta[7] >= 1 && go_max < 0x7f800000u && (long long)rows * pitch * CT <= cap_floats;

  if (!tiled) {
    // direct path (the reference's per-corner reductions): exact tier, every predicate
    s.finish_exact(a);
    voxel_backward_ct<float, CT, G, false, ONLY_VOL, 0>(a, s, vb, gv_delta, sD, sH, go, ggp, valid);
    return;
  }

  // ---- privatised accumulation ----
  const int floats = rows * pitch * CT;
  for (int i = threadIdx.x; i < floats / 4; i += kTileThreads) reinterpret_cast<int4 *>(s_gv)[i] = make_int4(0, 0, 0, 0);
  // scale = 2^21 / (the power of two above max|go|), as exponent arithmetic; clamped to normal floats
  const int e_s = min(253, max(1, 274 - (int)(go_max >> 23)));
  const float scale = __int_as_float(e_s << 23), inv_scale = __int_as_float((254 - e_s) << 23);
  float cw[8];
  s.corner_weights(cw);
  const CornerRows<float> vrows(vb, s.fbase, sD, sH);
  float dot[8];
  if constexpr (!ONLY_VOL) {
    float v[8][CT];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (s.in(k)) {  // only high corners on a high face are excluded (border tier)
        ld_channels<float, CT>(vrows.template corner<CT>(k), v[k]);
      } else {
#pragma unroll
        for (int t = 0; t < CT; t++) v[k][t] = 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
      float d = 0.f;  // dot += v * go, channel by channel (...c:736)
#pragma unroll
      for (int t = 0; t < CT; t++) d = E::add(d, E::mul(v[k][t], go[t]));
      dot[k] = d;  // go is finite here, so an excluded corner's 0 * go is 0 as in the reference
    }
  }
  __syncthreads();  // the box is zero
  {
    const int t0 = (((s.z0 - z_lo) * ey + (s.y0 - y_lo)) * pitch + (s.x0 - x_lo)) * CT;
    const int ty = pitch * CT, tz = ey * pitch * CT;
    // Warp aggregation.  When the volume is much smaller than the output (ALIGNet's cage up-sample: 64^3 voxels
    // over a 4^3 .. 12^3 cage) the 32 voxels of a warp fall into one or two cells, and 32-lane shared atomics on
    // one address serialise: 22 wavefronts per ATOMS, the L1 pipe 97 % busy, 1.7 ms per launch
    // (profiles/r01_upsample_backward.txt).  Lanes with the same corner-0 cell are therefore summed with REDUX
    // (the contributions are integers) and one lane per cell issues the atomic.  The groups of a warp execute their
    // REDUX one after the other (different member masks), so this pays only for very few, large groups: taken when
    // the launch is in the small-volume regime (a.tile_agg, set by the host) AND the warp holds at most four
    // distinct cells (warp-uniform).  With a threshold of eight cells the 12^3 cage lost 20 %; sheared fields on
    // large volumes never enter (no MATCH).
    // (REDUX with a partial member mask is emulated by a loop -- 940 instructions per voxel-warp -- so each of the
    // at most three groups is reduced with a full-mask REDUX over `lane in group ? q : 0`.)
    bool lead = false;
    int gid = 0, ngroups = 0;  // ngroups = 0: per-lane atomics
    if (a.tile_agg) {
      const unsigned grp = __match_any_sync(0xffffffffu, t0);
      const int leader = __ffs(grp) - 1;
      lead = (int)(threadIdx.x & 31) == leader;
      const unsigned leaders = __ballot_sync(0xffffffffu, lead);
      gid = __popc(leaders & ((1u << leader) - 1u));
      ngroups = __popc(leaders) <= 4 ? __popc(leaders) : 0;
    }
    // one straight-line body per group count (warp-uniform switch): a loop over the groups with its tests inside cost
    // 1 000 instructions per voxel-warp (profiles/r01_upsample_backward.txt)
    auto contribution = [&](int k, int t) {
      const float c = E::mul(cw[k], go[t]);  // ...c:737, the reference's product (0 for tail threads)
      // round(c * scale) without a conversion instruction: |c * scale| < 2^21, so adding 1.5 * 2^23
      // leaves the integer in the mantissa
      return __float_as_int(__fmaf_rn(c, scale, 12582912.0f)) - 0x4B400000;
    };
    auto cell_of = [&](int k) { return s_gv + t0 + ((k & 1) ? CT : 0) + ((k & 2) ? ty : 0) + ((k & 4) ? tz : 0); };
    // a channel whose gradOut is zero in the whole warp contributes exact zeros: skipped (the cage up-sample's 4th
    // channel always is -- the field's 4th component has no gradient, ...c:822)
    bool live[CT];
#pragma unroll
    for (int t = 0; t < CT; t++) live[t] = !ngroups || __any_sync(0xffffffffu, go[t] != 0.f);
    if (ngroups == 1) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        int *cell = cell_of(k);
#pragma unroll
        for (int t = 0; t < CT; t++) {
          if (!live[t]) continue;
          const int sum = __reduce_add_sync(0xffffffffu, contribution(k, t));
          if (lead) atomicAdd(cell + t, sum);
        }
      }
    } else if (ngroups == 2) {
      const bool g0 = gid == 0;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        int *cell = cell_of(k);
#pragma unroll
        for (int t = 0; t < CT; t++) {
          if (!live[t]) continue;
          const int q = contribution(k, t);
          const int s0 = __reduce_add_sync(0xffffffffu, g0 ? q : 0), s1 = __reduce_add_sync(0xffffffffu, g0 ? 0 : q);
          if (lead) atomicAdd(cell + t, g0 ? s0 : s1);
        }
      }
    } else if (ngroups == 3) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        int *cell = cell_of(k);
#pragma unroll
        for (int t = 0; t < CT; t++) {
          if (!live[t]) continue;
          const int q = contribution(k, t);
          const int s0 = __reduce_add_sync(0xffffffffu, gid == 0 ? q : 0), s1 = __reduce_add_sync(0xffffffffu, gid == 1 ? q : 0),
                    s2 = __reduce_add_sync(0xffffffffu, gid == 2 ? q : 0);
          if (lead) atomicAdd(cell + t, gid == 0 ? s0 : (gid == 1 ? s1 : s2));
        }
      }
    } else if (ngroups == 4) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        int *cell = cell_of(k);
#pragma unroll
        for (int t = 0; t < CT; t++) {
          if (!live[t]) continue;
          const int q = contribution(k, t);
          const int s0 = __reduce_add_sync(0xffffffffu, gid == 0 ? q : 0), s1 = __reduce_add_sync(0xffffffffu, gid == 1 ? q : 0),
                    s2 = __reduce_add_sync(0xffffffffu, gid == 2 ? q : 0), s3 = __reduce_add_sync(0xffffffffu, gid == 3 ? q : 0);
          if (lead) atomicAdd(cell + t, gid < 2 ? (gid == 0 ? s0 : s1) : (gid == 2 ? s2 : s3));
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        int *cell = cell_of(k);
#pragma unroll
        for (int t = 0; t < CT; t++) atomicAdd(cell + t, contribution(k, t));
      }
    }
  }
  if constexpr (!ONLY_VOL) {
    float gz, gy, gx;
    grad_grid_from_dots<float>(dot, s, a, gz, gy, gx);
    if (valid) st_grad_grid<float, G>(ggp, gz, gy, gx);
  }
  __syncthreads();
  // ---- flush: 16-byte vector reductions, lanes along x, a power-of-two group of lanes per row; rows and
  // cells outside the volume (the one-cell overhang at a high face) hold zeros and are skipped ----
  float *gvb = a.gvol + (size_t)b * a.vs32;
  const int qpr = pitch * CT / 4;                                         // 16-byte groups per row
  const int lpr = qpr <= 4 ? 4 : (qpr <= 8 ? 8 : (qpr <= 16 ? 16 : 32));  // lanes serving one row
  const int qi = threadIdx.x & (lpr - 1);
  const float inv_ey = 1.0f / (float)ey;
  for (int r = threadIdx.x / lpr; r < rows; r += kTileThreads / lpr) {
    const int zz = (int)(((float)r + 0.5f) * inv_ey), yy = r - zz * ey;  // exact for r < 2^20
    if (z_lo + zz >= a.iD || y_lo + yy >= a.iH) continue;
    float *g = gvb + (size_t)(z_lo + zz) * sD + (size_t)(y_lo + yy) * sH + (size_t)x_lo * CT;
    for (int q = qi; q < qpr; q += lpr) {
      const int4 v = reinterpret_cast<const int4 *>(s_gv + (size_t)r * pitch * CT)[q];
      if ((v.x | v.y | v.z | v.w) && (x_lo * CT + 4 * q) < sH)
        red_add4(g + 4 * q, (float)v.x * inv_scale, (float)v.y * inv_scale, (float)v.z * inv_scale,
                 (float)v.w * inv_scale);
    }
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
  const int per_cta = kSmemBudget / kTileCtasPerSm - 1024 - 128;
  const int cap_floats = (per_cta / 4) & ~31;
  const size_t smem = (size_t)cap_floats * 4;
  if (smem > 48 * 1024 && have.load(std::memory_order_relaxed) < smem) {
    VOLSTN_CUDA_TRY(cudaFuncSetAttribute(k_sampler_bwd_tile<CT, G, OV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
    have.store(smem, std::memory_order_relaxed);
  }
  k_sampler_bwd_tile<CT, G, OV><<<dim3(nbx, nby, (unsigned)(nbz * a.B)), kTileThreads, smem, st>>>(a, nbz, cap_floats);
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

// 16-byte rows: C in {1, 2, 4}, the volume's W a multiple of 4 / C cells, gradVol 16-byte aligned; the
// conversion-free set-up must be valid (extents <= 2^22) and the launch grid must fit
bool tile_supported(const SamplerArgs<float> &a) {
  if (a.C != 1 && a.C != 2 && a.C != 4) return false;
  if ((a.iW * a.C) % 4 != 0) return false;
  if (reinterpret_cast<uintptr_t>(a.gvol) % 16 != 0) return false;
  if (!a.border_ok) return false;
  const long long nbz = (a.oD + kBZ - 1) / kBZ, nby = (a.oH + kBY - 1) / kBY;
  if (nbz * a.B > 65535 || nby > 65535) return false;
  return true;
}

int launch_bwd_tile(const SamplerArgs<float> &a, int G, bool only_vol, cudaStream_t st) {
  if (G == 4) return only_vol ? launch_bwd_tile_g<4, true>(a, st) : launch_bwd_tile_g<4, false>(a, st);
  return only_vol ? launch_bwd_tile_g<3, true>(a, st) : launch_bwd_tile_g<3, false>(a, st);
}

}  // namespace volstn
