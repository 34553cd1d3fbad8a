// sampler_tma.cu -- the sampler with the volume gather STAGED BY TMA (VOLSTN_ALGO_TMA): for warp fields that shear.
//
// What the north star names for the gather ("shared-memory or TMA staging of volume tiles") and what round 1 never
// measured.  With a sheared field (jittered ALIGNet cage `g2`, rotations, random fields) the direct kernels' eight
// gathers of a warp land in ~20 different 32-byte sectors each: the forward is L1-sector bound, the backward bound by
// L2 reduction sector-operations (5.2 per voxel, profiles/r02_fields_ncu.md).  Here
//
//   * a CTA owns an 8 x 8 x 8 BRICK of output voxels (256 threads, two voxels each) -- a cube keeps the bounding box
//     of the brick's corner cells small: about (8 |J| + 2)^3 cells for a field with Jacobian J;
//   * the CTA reduces the box [min i0, max i0 + 1] of its corner indices (REDUX + six shared atomics) and ONE thread
//     issues ONE `cp.async.bulk.tensor.4d` (a BX x BY x BZ x 1 box of the B x D x H x W volume) into shared memory.
//     TMA's out-of-bounds ZERO FILL is the reference's padding rule (an out-of-bounds corner reads as 0,
//     ...c:542-564) beyond the high faces; below zero the box is clamped (a TMA reduction faults on a negative start
//     coordinate, measured) and the corner predicates of the exact tier do the rest -- any finite coordinate is handled;
//   * the eight corners are then shared-memory loads, and the arithmetic is the direct kernels' (sampler.cuh:
//     corner_weights, the reference's summation order): `out` and gradGrid keep the reference's bits;
//   * backward: gradVol is privatised in a second shared-memory box in fixed point (native ATOMS.ADD; a float
//     shared atomicAdd is a CAS loop on sm_100a), converted in place and flushed with ONE
//     `cp.reduce.async.bulk.tensor.4d ... .add` -- one L2 reduction sector-op per 8 cells instead of one per
//     scattered lane, and cells outside the volume are dropped by the hardware;
//   * a brick whose box does not fit (a fold, a far out-of-bounds excursion, NaN) runs the direct per-voxel code.
#include <cuda.h>

#include <atomic>
#include <mutex>

#include "sampler.cuh"
#include "volstn_internal.h"

namespace volstn {

namespace {

constexpr int kTB = 8;                       // output brick edge
constexpr int kTmaThreads = 256;             // two voxels per thread
// The box must START on a 16-byte boundary of the innermost dimension (x a multiple of 4 floats): an unaligned start
// coordinate raises an illegal-instruction fault on the B200 (tools/probe/tma_probe*.cu, profiles/r02_tma.md), so the box
// is 4 cells longer in x than in y and z to absorb the alignment.
#ifndef VOLSTN_TMA_BOX
#define VOLSTN_TMA_BOX 16
#endif
#ifndef VOLSTN_TMA_CTAS_FWD
#define VOLSTN_TMA_CTAS_FWD 4
#endif
#ifndef VOLSTN_TMA_CTAS_BWD
#define VOLSTN_TMA_CTAS_BWD 3
#endif
constexpr int kBoxX = VOLSTN_TMA_BOX + 4, kBoxY = VOLSTN_TMA_BOX, kBoxZ = VOLSTN_TMA_BOX;
// ... and a SMALL box for bricks of fields that barely shear (identity-like fields: 8 + 2 cells per axis): what a brick
// costs is, first of all, the bytes its boxes move (a 24^3 box: 4x the time, profiles/r02_tma.md), so each brick takes
// the smallest of the two boxes that holds it.  A tensor map fixes its box size: two maps per tensor.
// MEASURED AND SWITCHED OFF (profiles/r02_tma.md): the identity field's backward gains 10 % (322 -> 291 us) but every
// forward and the rotations lose 5-20 % -- with two box sizes the box pitches are run-time values and the eight corner
// addresses of a voxel cost integer multiplies.  VOLSTN_TMA_DUAL = 1 re-enables it.
#ifndef VOLSTN_TMA_DUAL
#define VOLSTN_TMA_DUAL 0
#endif
#ifndef VOLSTN_TMA_SMALL
#define VOLSTN_TMA_SMALL 12
#endif
constexpr int kSmX = VOLSTN_TMA_SMALL + 4, kSmY = VOLSTN_TMA_SMALL, kSmZ = VOLSTN_TMA_SMALL;
constexpr int kSmCells = kSmX * kSmY * kSmZ;
constexpr int kBoxCells = kBoxX * kBoxY * kBoxZ;

// ---- PTX wrappers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// box load: coordinates innermost first (x, y, z, b); out-of-bounds elements arrive as zeros
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, unsigned long long *bar, int x, int y, int z,
                                            int b) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<unsigned long long>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z), "r"(b)
      : "memory");
}
// box reduction: global[box] += shared[box]; elements outside the tensor are dropped
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap *map, const void *src, int x, int y, int z, int b) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<unsigned long long>(map)),
               "r"(smem_u32(src)), "r"(x), "r"(y), "r"(z), "r"(b)
               : "memory");
}
__device__ __forceinline__ void tma_commit_and_wait_read() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct TmaArgs {
  SamplerArgs<float> s;
  int nbx, nby, nbz;  // bricks per sample
};

// per-voxel state that survives the staging barrier
struct BrickVoxel {
  VoxelSetup<float> su;
  bool valid;     // inside the output
  unsigned j;     // voxel index inside the sample
};

// s_meta: [0..2] min z0, y0, x0   [3..5] max z0, y0, x0   [6] any lane not stageable   [7] max |gradOut| bits (backward)
template <int G, int VPT>
__device__ __forceinline__ bool brick_setup(const TmaArgs &p, BrickVoxel (&v)[VPT], int *s_meta, unsigned b, int bx, int by, int bz) {
  const SamplerArgs<float> &a = p.s;
  const size_t sb = (size_t)b * a.n32;
  int lo[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, hi[3] = {INT32_MIN, INT32_MIN, INT32_MIN};
  bool bad = false;
#pragma unroll
  for (int q = 0; q < VPT; q++) {
    const int t = threadIdx.x + q * kTmaThreads;  // 0 .. 511: x fastest
    const int x = bx * kTB + (t & 7), y = by * kTB + ((t >> 3) & 7), z = bz * kTB + (t >> 6);
    v[q].valid = x < a.oW && y < a.oH && z < a.oD;
    // tail voxels re-read the brick's first voxel (always inside): no divergent set-up
    v[q].j = v[q].valid ? ((unsigned)z * a.oH + y) * a.oW + x : ((unsigned)(bz * kTB) * a.oH + by * kTB) * a.oW + bx * kTB;
    float zf, yf, xf;
    ld_grid<float, G>(a.grid + (sb + v[q].j) * G, zf, yf, xf);
    v[q].su.begin(zf, yf, xf, a);  // the three coordinates (...c:507-517)
    v[q].su.finish_exact(a);       // the reference's floor, weights and bounds predicates: any coordinate
    // stageable: |coordinate| small enough that the box arithmetic cannot overflow (NaN / Inf / 1e30 are not)
    const bool ok = fabsf(v[q].su.cz) < 4194304.0f && fabsf(v[q].su.cy) < 4194304.0f && fabsf(v[q].su.cx) < 4194304.0f;
    bad |= !ok;
    if (ok) {
      lo[0] = min(lo[0], v[q].su.z0), hi[0] = max(hi[0], v[q].su.z0);
      lo[1] = min(lo[1], v[q].su.y0), hi[1] = max(hi[1], v[q].su.y0);
      lo[2] = min(lo[2], v[q].su.x0), hi[2] = max(hi[2], v[q].su.x0);
    }
  }
  const bool wbad = __any_sync(0xffffffffu, bad);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    lo[i] = __reduce_min_sync(0xffffffffu, lo[i]);
    hi[i] = __reduce_max_sync(0xffffffffu, hi[i]);
  }
  if ((threadIdx.x & 31) == 0) {
    if (wbad) s_meta[6] = 1;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      atomicMin(&s_meta[i], lo[i]);
      atomicMax(&s_meta[3 + i], hi[i]);
    }
  }
  return true;
}

}  // namespace

// =============================================================================================
// forward
// =============================================================================================
template <int G>
__global__ void __launch_bounds__(kTmaThreads, VOLSTN_TMA_CTAS_FWD) k_sampler_fwd_tma(const TmaArgs p, const __grid_constant__ CUtensorMap vmap,
                                                                                        const __grid_constant__ CUtensorMap vmap_s) {
  using E = Ex<float>;
  const SamplerArgs<float> &a = p.s;
  extern __shared__ __align__(128) unsigned char tma_smem[];
  float *s_box = reinterpret_cast<float *>(tma_smem);
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ int s_meta[8];
  const int bx = blockIdx.x, by = blockIdx.y;
  const unsigned b = blockIdx.z / p.nbz;
  const int bz = blockIdx.z - b * p.nbz;
  if (threadIdx.x < 3) s_meta[threadIdx.x] = INT32_MAX;
  else if (threadIdx.x < 6) s_meta[threadIdx.x] = INT32_MIN;
  else if (threadIdx.x < 8) s_meta[threadIdx.x] = 0;
  if (threadIdx.x == 32) mbar_init(&s_bar, 1);
  __syncthreads();
  BrickVoxel v[2];
  brick_setup<G, 2>(p, v, s_meta, b, bx, by, bz);
  __syncthreads();
  // box origin: the low corner cells, clamped at 0 (cells below zero are out of bounds: predicated away, never staged)
  // and, in x, aligned down to 16 bytes
  const int z_lo = max(s_meta[0], 0), y_lo = max(s_meta[1], 0), x_lo = max(s_meta[2], 0) & ~3;
  const int ez = s_meta[3] - z_lo + 2, ey = s_meta[4] - y_lo + 2, ex = s_meta[5] - x_lo + 2;  // cells the brick touches
  const bool staged = !s_meta[6] && ez <= kBoxZ && ey <= kBoxY && ex <= kBoxX;
  const bool small = VOLSTN_TMA_DUAL && ez <= kSmZ && ey <= kSmY && ex <= kSmX;
  const int pX = small ? kSmX : kBoxX, pXY = small ? kSmX * kSmY : kBoxX * kBoxY;  // pitches of the staged box
  float *ob = a.out + (size_t)b * a.n32;
  if (staged) {
    if (threadIdx.x == 0) {
      mbar_expect_tx(&s_bar, (small ? kSmCells : kBoxCells) * 4);
      tma_load_4d(s_box, small ? &vmap_s : &vmap, &s_bar, x_lo, y_lo, z_lo, (int)b);
    }
    // the corner weights do not need the data: computed while the box is in flight
    float cw[2][8];
    int base[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      v[q].su.corner_weights(cw[q]);
      base[q] = (v[q].su.z0 - z_lo) * pXY + (v[q].su.y0 - y_lo) * pX + (v[q].su.x0 - x_lo);
    }
    mbar_wait(&s_bar, 0);
#pragma unroll
    for (int q = 0; q < 2; q++) {
      float acc = 0.f;
#pragma unroll
      for (int jj = 0; jj < 8; jj++) {
        const int k = (G == 4) ? order_xyz(jj) : order_zxy(jj);  // ...c:566-573 / ...c:149-156
        // out-of-bounds corners read as 0 (...c:531-564): beyond the high faces TMA has filled zeros, below zero the
        // cell is not in the box at all
        const float val = v[q].su.in(k) ? s_box[base[q] + (k & 1) + ((k >> 1) & 1) * pX + (k >> 2) * pXY] : 0.f;
        const float term = E::mul(cw[q][k], val);
        acc = jj ? E::add(acc, term) : term;
      }
      if (v[q].valid) st_stream_f1(ob + v[q].j, acc);
    }
  } else {
    // the direct code of the exact tier (per-corner predicates, gathers from global memory)
    const float *vb = a.vol + (size_t)b * a.vs32;
    const int sH = a.iW, sD = a.iH * a.iW;
#pragma unroll
    for (int q = 0; q < 2; q++) voxel_forward_ct<float, 1, G, 0>(v[q].su, vb, sD, sH, ob + v[q].j, v[q].valid);
  }
}

// =============================================================================================
// backward: gradGrid (bit-exact) + gradVol privatised in shared memory, flushed by one TMA reduction
// =============================================================================================
template <int G, bool ONLY_GRID>
__global__ void __launch_bounds__(kTmaThreads, VOLSTN_TMA_CTAS_BWD) k_sampler_bwd_tma(const TmaArgs p, const __grid_constant__ CUtensorMap vmap,
                                                                   const __grid_constant__ CUtensorMap gmap,
                                                                   const __grid_constant__ CUtensorMap vmap_s,
                                                                   const __grid_constant__ CUtensorMap gmap_s) {
  using E = Ex<float>;
  const SamplerArgs<float> &a = p.s;
  extern __shared__ __align__(128) unsigned char tma_smem[];
  float *s_box = reinterpret_cast<float *>(tma_smem);
  int *s_acc = reinterpret_cast<int *>(tma_smem + (size_t)kBoxCells * 4);  // only with the gradVol scatter
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ int s_meta[8];
  const int bx = blockIdx.x, by = blockIdx.y;
  const unsigned b = blockIdx.z / p.nbz;
  const int bz = blockIdx.z - b * p.nbz;
  if (threadIdx.x < 3) s_meta[threadIdx.x] = INT32_MAX;
  else if (threadIdx.x < 6) s_meta[threadIdx.x] = INT32_MIN;
  else if (threadIdx.x < 8) s_meta[threadIdx.x] = 0;
  if (threadIdx.x == 32) mbar_init(&s_bar, 1);
  __syncthreads();
  BrickVoxel v[2];
  brick_setup<G, 2>(p, v, s_meta, b, bx, by, bz);
  const size_t sb = (size_t)b * a.n32;
  float go[2];
  unsigned go_abs = 0;
#pragma unroll
  for (int q = 0; q < 2; q++) {
    go[q] = ld_stream_f1(a.gout + sb + v[q].j);
    if (!v[q].valid) go[q] = 0.f;
    go_abs = max(go_abs, (unsigned)__float_as_int(go[q]) & 0x7fffffffu);
  }
  if constexpr (!ONLY_GRID) {
    go_abs = __reduce_max_sync(0xffffffffu, go_abs);  // |x| orders like its bit pattern; inf / nan are largest
    if ((threadIdx.x & 31) == 0 && go_abs) atomicMax(reinterpret_cast<unsigned *>(&s_meta[7]), go_abs);
  }
  __syncthreads();
  // box origin clamped at 0 and 16-byte aligned in x: a TMA REDUCTION with a negative start coordinate faults (a load
  // does not; tools/probe/tma_probe.cu), and cells below zero never receive anything anyway
  const int z_lo = max(s_meta[0], 0), y_lo = max(s_meta[1], 0), x_lo = max(s_meta[2], 0) & ~3;
  const unsigned go_max = (unsigned)s_meta[7];
  const int ez = s_meta[3] - z_lo + 2, ey = s_meta[4] - y_lo + 2, ex = s_meta[5] - x_lo + 2;  // cells the brick touches
  const bool staged = !s_meta[6] && go_max < 0x7f800000u && ez <= kBoxZ && ey <= kBoxY && ex <= kBoxX;
  const bool small = VOLSTN_TMA_DUAL && ez <= kSmZ && ey <= kSmY && ex <= kSmX;
  const int pX = small ? kSmX : kBoxX, pXY = small ? kSmX * kSmY : kBoxX * kBoxY;
  const int cells = small ? kSmCells : kBoxCells;
  float *ggb = a.ggrid + sb * G;
  const float *vb = a.vol + (size_t)b * a.vs32;
  const int sH = a.iW, sD = a.iH * a.iW;
  if (!staged) {
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const float g1[1] = {go[q]};
      voxel_backward_ct<float, 1, G, ONLY_GRID, false, 0>(a, v[q].su, vb, a.gvol - a.vol, sD, sH, g1, ggb + (size_t)v[q].j * G, v[q].valid);
    }
    return;
  }
  if (threadIdx.x == 0) {
    mbar_expect_tx(&s_bar, cells * 4);
    tma_load_4d(s_box, small ? &vmap_s : &vmap, &s_bar, x_lo, y_lo, z_lo, (int)b);
  }
  if constexpr (!ONLY_GRID) {  // the privatised gradVol box of this brick's size, zeroed while the volume box is in flight
    for (int i = threadIdx.x; i < cells / 4; i += kTmaThreads) reinterpret_cast<int4 *>(s_acc)[i] = make_int4(0, 0, 0, 0);
    __syncthreads();
  }
  // fixed point for the privatised gradVol: 2^20 steps up to the power of two above max |gradOut| of the brick, so that
  // one contribution (|w| <= 1) is rounded by at most 2^-21 of that maximum and the 512 voxels of a brick cannot overflow
  // (scale = 2^20 / the power of two above max |gradOut|, as exponent arithmetic, clamped to normal floats)
  const int e_s = min(253, max(1, 273 - (int)(go_max >> 23)));
  const float scale = __int_as_float(e_s << 23), inv_scale = __int_as_float((254 - e_s) << 23);
  float cw[2][8];
  int base[2];
#pragma unroll
  for (int q = 0; q < 2; q++) {
    v[q].su.corner_weights(cw[q]);
    base[q] = (v[q].su.z0 - z_lo) * pXY + (v[q].su.y0 - y_lo) * pX + (v[q].su.x0 - x_lo);
  }
  if constexpr (!ONLY_GRID) {
    // the scatter does not need the staged volume either
#pragma unroll
    for (int q = 0; q < 2; q++) {
      if (v[q].valid) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
          // ...c:737, then round(c * scale) without a conversion instruction (|c * scale| < 2^21: adding 1.5 * 2^23 leaves
          // the integer in the mantissa)
          const int qv = __float_as_int(__fmaf_rn(E::mul(cw[q][k], go[q]), scale, 12582912.0f)) - 0x4B400000;
          if (qv && v[q].su.in(k)) atomicAdd(&s_acc[base[q] + (k & 1) + ((k >> 1) & 1) * pX + (k >> 2) * pXY], qv);
        }
      }
    }
  }
  mbar_wait(&s_bar, 0);
#pragma unroll
  for (int q = 0; q < 2; q++) {
    float dot[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      // dot = 0 + v * go (...c:736); an out-of-bounds corner stays 0, never sees go and is never read
      const bool in = v[q].su.in(k);
      const float val = in ? s_box[base[q] + (k & 1) + ((k >> 1) & 1) * pX + (k >> 2) * pXY] : 0.f;
      dot[k] = in ? E::add(0.f, E::mul(val, go[q])) : 0.f;
    }
    float gz, gy, gx;
    grad_grid_from_dots<float>(dot, v[q].su, a, gz, gy, gx);
    if (v[q].valid) st_grad_grid<float, G>(ggb + (size_t)v[q].j * G, gz, gy, gx);
  }
  if constexpr (!ONLY_GRID) {
    __syncthreads();
    // fixed point -> float in place, then ONE box reduction into gradVol (cells outside the volume are dropped)
    for (int i = threadIdx.x; i < cells / 4; i += kTmaThreads) {
      const int4 qv = reinterpret_cast<const int4 *>(s_acc)[i];
      reinterpret_cast<float4 *>(s_acc)[i] =
          make_float4((float)qv.x * inv_scale, (float)qv.y * inv_scale, (float)qv.z * inv_scale, (float)qv.w * inv_scale);
    }
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
      tma_reduce_add_4d(small ? &gmap_s : &gmap, s_acc, x_lo, y_lo, z_lo, (int)b);
      tma_commit_and_wait_read();  // shared memory must stay valid until the engine has read it
    }
  }
}

// =============================================================================================
// host side
// =============================================================================================
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup: the library links no libcuda
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      (void)cudaGetLastError();
  });
  return fn;
}

// B x D x H x W float tensor (C = 1, contiguous) as a rank-4 tiled map with a kBoxX x kBoxY x kBoxZ x 1 box
bool make_map(CUtensorMap *m, const float *base, int B, int D, int H, int W, bool small_box) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * D * 4};
  const cuuint32_t box[4] = {(cuuint32_t)(small_box ? kSmX : kBoxX), (cuuint32_t)(small_box ? kSmY : kBoxY),
                             (cuuint32_t)(small_box ? kSmZ : kBoxZ), 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename K>
int big_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) VOLSTN_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return VOLSTN_OK;
}

bool fill_bricks(TmaArgs &p, const SamplerArgs<float> &a) {
  p.s = a;
  p.nbx = (a.oW + kTB - 1) / kTB, p.nby = (a.oH + kTB - 1) / kTB, p.nbz = (a.oD + kTB - 1) / kTB;
  return (long long)p.nbz * a.B <= 65535 && p.nby <= 65535;
}

}  // namespace

// single-channel packed float tensors, rows of the volume 16 bytes wide (TMA global strides), 16-byte aligned base
bool tma_supported(const SamplerArgs<float> &a, const float *gvol) {
  if (a.C != 1 || a.iW % 4 != 0) return false;
  if (reinterpret_cast<uintptr_t>(a.vol) % 16 != 0 || (gvol && reinterpret_cast<uintptr_t>(gvol) % 16 != 0)) return false;
  if ((long long)((a.oD + kTB - 1) / kTB) * a.B > 65535 || (a.oH + kTB - 1) / kTB > 65535) return false;
  return encode_tiled() != nullptr;
}

int launch_fwd_tma(const SamplerArgs<float> &a, int G, cudaStream_t st) {
  TmaArgs p;
  CUtensorMap vmap, vmap_s;
  if (!fill_bricks(p, a) || !make_map(&vmap, a.vol, a.B, a.iD, a.iH, a.iW, false) || !make_map(&vmap_s, a.vol, a.B, a.iD, a.iH, a.iW, true))
    return VOLSTN_ERR_UNSUPPORTED;
  const dim3 grid(p.nbx, p.nby, (unsigned)(p.nbz * a.B));
  const size_t smem = (size_t)kBoxCells * 4;
  int rc;
  if (G == 4) {
    if ((rc = big_smem(k_sampler_fwd_tma<4>, smem))) return rc;
    k_sampler_fwd_tma<4><<<grid, kTmaThreads, smem, st>>>(p, vmap, vmap_s);
  } else {
    if ((rc = big_smem(k_sampler_fwd_tma<3>, smem))) return rc;
    k_sampler_fwd_tma<3><<<grid, kTmaThreads, smem, st>>>(p, vmap, vmap_s);
  }
  return after_launch();
}

int launch_bwd_tma(const SamplerArgs<float> &a, int G, bool only_grid, cudaStream_t st) {
  TmaArgs p;
  CUtensorMap vmap, gmap, vmap_s, gmap_s;
  if (!fill_bricks(p, a) || !make_map(&vmap, a.vol, a.B, a.iD, a.iH, a.iW, false) || !make_map(&vmap_s, a.vol, a.B, a.iD, a.iH, a.iW, true))
    return VOLSTN_ERR_UNSUPPORTED;
  if (only_grid)
    gmap = vmap, gmap_s = vmap_s;
  else if (!make_map(&gmap, a.gvol, a.B, a.iD, a.iH, a.iW, false) || !make_map(&gmap_s, a.gvol, a.B, a.iD, a.iH, a.iW, true))
    return VOLSTN_ERR_UNSUPPORTED;
  const dim3 grid(p.nbx, p.nby, (unsigned)(p.nbz * a.B));
  const size_t smem = (size_t)kBoxCells * 4 * (only_grid ? 1 : 2);
  int rc;
  if (G == 4) {
    if (only_grid) {
      if ((rc = big_smem(k_sampler_bwd_tma<4, true>, smem))) return rc;
      k_sampler_bwd_tma<4, true><<<grid, kTmaThreads, smem, st>>>(p, vmap, gmap, vmap_s, gmap_s);
    } else {
      if ((rc = big_smem(k_sampler_bwd_tma<4, false>, smem))) return rc;
      k_sampler_bwd_tma<4, false><<<grid, kTmaThreads, smem, st>>>(p, vmap, gmap, vmap_s, gmap_s);
    }
  } else {
    if (only_grid) {
      if ((rc = big_smem(k_sampler_bwd_tma<3, true>, smem))) return rc;
      k_sampler_bwd_tma<3, true><<<grid, kTmaThreads, smem, st>>>(p, vmap, gmap, vmap_s, gmap_s);
    } else {
      if ((rc = big_smem(k_sampler_bwd_tma<3, false>, smem))) return rc;
      k_sampler_bwd_tma<3, false><<<grid, kTmaThreads, smem, st>>>(p, vmap, gmap, vmap_s, gmap_s);
    }
  }
  return after_launch();
}

}  // namespace volstn
