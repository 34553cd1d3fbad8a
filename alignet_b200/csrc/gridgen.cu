// gridgen.cu -- nn.VolumetricGridGenerator on sm_100a.
//
// Reference: stn3dtorch/VolumetricGridGenerator.lua.  The Lua module materialises a base grid
// D x H x W x 4 of (zn, yn, xn, 1) (lua:12-23), replicates it B times (lua:50-55) and calls
// torch.bmm / baddbmm with K = 4 (lua:58-60, 79-82).  Here nothing is materialised:
//   forward  : write-only, 16 B per voxel.  The three axis tables (D + H + W numbers, computed in
//              double and rounded to `real` exactly like the Lua constructor) live in shared
//              memory; a warp owns whole W-rows, so (b, z, y) -- and with them T[b] and the
//              z / y partial sums -- are warp-uniform and one lane emits one 128-bit store.
//   backward : read-only, 16 B per voxel, 16 FMAs per voxel, deterministic two-stage reduction
//              (per-block partials in double -> one block per sample folds them in fixed order).
#include "volstn_common.cuh"

namespace volstn {

constexpr int kGGThreads = 256;
constexpr int kGGWarps = kGGThreads / 32;

// (real)(-1 + k/(n-1)*2) in double, as Lua evaluates it (VolumetricGridGenerator.lua:17-19)
template <typename real>
__device__ __forceinline__ real base_coord(int k, int n) {
  return (real)(-1.0 + (double)k / (double)(n - 1) * 2.0);
}

template <typename real>
__device__ __forceinline__ void fill_axis_tables(real *tz, real *ty, real *tx, int D, int H, int W) {
  for (int i = threadIdx.x; i < D + H + W; i += blockDim.x) {
    if (i < D)
      tz[i] = base_coord<real>(i, D);
    else if (i < D + H)
      ty[i - D] = base_coord<real>(i - D, H);
    else
      tx[i - D - H] = base_coord<real>(i - D - H, W);
  }
}

template <typename real> struct Vec4;
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<double> { using type = double4; };

// out[b, z, y, x, i] = ((T[i][0]*zn + T[i][1]*yn) + T[i][2]*xn) + T[i][3]*1   (k = 0..3 in order)
template <typename real>
__global__ void __launch_bounds__(kGGThreads) k_gridgen_fwd(const real *__restrict__ T, real *__restrict__ out,
                                                            int B, int D, int H, int W, long long rows_per_warp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  real *tz = reinterpret_cast<real *>(smem_raw);
  real *ty = tz + D;
  real *tx = ty + H;
  fill_axis_tables<real>(tz, ty, tx, D, H, W);
  __syncthreads();

  using E = Ex<real>;
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * kGGWarps + (threadIdx.x >> 5);
  const long long total_rows = (long long)B * D * H;
  long long r = warp * rows_per_warp;
  const long long r_end = (r + rows_per_warp < total_rows) ? r + rows_per_warp : total_rows;
  int cur_b = -1;
  real t[16];
  for (; r < r_end; r++) {
    const int y = (int)(r % H);
    const long long q = r / H;
    const int z = (int)(q % D);
    const int b = (int)(q / D);
    if (b != cur_b) {  // warp-uniform
      cur_b = b;
#pragma unroll
      for (int i = 0; i < 16; i++) t[i] = __ldg(T + (long long)b * 16 + i);
    }
    const real zn = tz[z], yn = ty[y];
    real part[4];
#pragma unroll
    for (int i = 0; i < 4; i++) part[i] = E::add(E::mul(t[i * 4 + 0], zn), E::mul(t[i * 4 + 1], yn));
    real *orow = out + r * (long long)W * 4;
    for (int x = lane; x < W; x += 32) {
      const real xn = tx[x];
      real o[4];
#pragma unroll
      for (int i = 0; i < 4; i++) o[i] = E::add(E::add(part[i], E::mul(t[i * 4 + 2], xn)), t[i * 4 + 3]);
      if constexpr (sizeof(real) == 4) {
        *reinterpret_cast<float4 *>(orow + (long long)x * 4) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
        *reinterpret_cast<double4 *>(orow + (long long)x * 4) = make_double4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

// stage 1: partial[b][blk][i*4 + j] = sum over this block's rows of gradGrid[b,v,i] * base[v,j]
// Tunables of the backward, swept on a B200 at 64^3 x 64 (tools/ggbench.py, profiles/r01_gridgen_backward.txt): rows of
// loads in flight per warp trip, resident CTAs per SM (a 64-register cap) and waves of CTAs over the batch.
//   4 / none (80 registers, 3 CTAs) / 4 waves : 90 us      4 / 4 / 4 : 74 us      4 / 4 / 2 : 70 us      4 / 4 / 1 : 65.5 us
//   8 rows in flight: 84-141 us (registers)      caps of 5, 6, 8 CTAs: 92-174 us (spills)
// One wave of long-lived CTAs also leaves fewer partials for the ordered final stage.  Folding that stage into the
// last CTA of each sample (completion counters zeroed by a memset node) was measured too: 69.6 us against 65.5 us for
// the two launches -- the memset node costs more than the 16-thread kernel.
#ifndef GG_RPT
#define GG_RPT 4
#endif
#ifndef GG_CTAS
#define GG_CTAS 4
#endif
#ifndef GG_WAVES
#define GG_WAVES 1
#endif
template <typename real>
__global__ void __launch_bounds__(kGGThreads, GG_CTAS) k_gridgen_bwd_partial(const real *__restrict__ gg,
                                                                    double *__restrict__ partial, int D, int H,
                                                                    int W, int rows_per_warp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  real *tz = reinterpret_cast<real *>(smem_raw);
  real *ty = tz + D;
  real *tx = ty + H;
  __shared__ double red[kGGWarps][16];
  fill_axis_tables<real>(tz, ty, tx, D, H, W);
  __syncthreads();

  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const long long rows = (long long)D * H;
  long long r = ((long long)blockIdx.x * kGGWarps + wid) * rows_per_warp;
  const long long r_end = (r + rows_per_warp < rows) ? r + rows_per_warp : rows;
  real acc[16];
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = real(0);
  const real *gb = gg + (long long)b * rows * W * 4;
  // Two rows per trip with all their loads issued before the first multiply-add: the accumulators are the
  // only loop-carried state, so this doubles the bytes each warp keeps in flight.
  auto accumulate = [&](const real (&g)[4], real zn, real yn, real xn) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      acc[i * 4 + 0] = fma(g[i], zn, acc[i * 4 + 0]);
      acc[i * 4 + 1] = fma(g[i], yn, acc[i * 4 + 1]);
      acc[i * 4 + 2] = fma(g[i], xn, acc[i * 4 + 2]);
      acc[i * 4 + 3] += g[i];
    }
  };
  auto load4 = [&](const real *p, real (&g)[4]) {
    if constexpr (sizeof(real) == 4) {
      const float4 v = ld_stream_f4(reinterpret_cast<const float4 *>(p));
      g[0] = v.x; g[1] = v.y; g[2] = v.z; g[3] = v.w;
    } else {
      const double4 v = *reinterpret_cast<const double4 *>(p);
      g[0] = v.x; g[1] = v.y; g[2] = v.z; g[3] = v.w;
    }
  };
  constexpr int kRowsPerTrip = GG_RPT;
  for (; r < r_end; r += kRowsPerTrip) {
    for (int x = lane; x < W; x += 32) {
      real g[kRowsPerTrip][4];
#pragma unroll
      for (int q = 0; q < kRowsPerTrip; q++)
        if (r + q < r_end) load4(gb + ((r + q) * (long long)W + x) * 4, g[q]);
      const real xn = tx[x];
#pragma unroll
      for (int q = 0; q < kRowsPerTrip; q++) {
        if (r + q < r_end) {
          const int y = (int)((r + q) % H), z = (int)((r + q) / H);
          accumulate(g[q], tz[z], ty[y], xn);
        }
      }
    }
  }
  // warp tree in double, then the block's warps in order
#pragma unroll
  for (int i = 0; i < 16; i++) {
    double v = (double)acc[i];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (lane == 0) red[wid][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kGGWarps; w++) v += red[w][threadIdx.x];
    partial[((long long)b * gridDim.x + blockIdx.x) * 16 + threadIdx.x] = v;
  }
}

// stage 2: gradT[b][i][j] = sum over blocks (fixed order), rounded once to `real`
template <typename real>
__global__ void k_gridgen_bwd_final(const double *__restrict__ partial, real *__restrict__ gradT, int nblk) {
  const int b = blockIdx.x, e = threadIdx.x;  // 16 threads
  // loads issued eight at a time (they are independent; only the additions are ordered)
  double v = 0.0;
  int k = 0;
  for (; k + 8 <= nblk; k += 8) {
    double t[8];
#pragma unroll
    for (int q = 0; q < 8; q++) t[q] = partial[((long long)b * nblk + k + q) * 16 + e];
#pragma unroll
    for (int q = 0; q < 8; q++) v += t[q];
  }
  for (; k < nblk; k++) v += partial[((long long)b * nblk + k) * 16 + e];
  gradT[(long long)b * 16 + e] = (real)v;
}

namespace {
int bwd_blocks_per_sample(int64_t B, int64_t D, int64_t H, int *rows_per_warp) {
  const int64_t rows = D * H;
  // one wave of 148 SMs x 4 resident CTAs over the whole batch (see the sweep above), at least 4 rows per warp
  int64_t want_blocks = (int64_t)kSMs * (GG_CTAS > 1 ? GG_CTAS : 3) * GG_WAVES / (B > 0 ? B : 1);
  if (want_blocks < 1) want_blocks = 1;
  int64_t rpw = (rows + want_blocks * kGGWarps - 1) / (want_blocks * kGGWarps);
  if (rpw < 4) rpw = 4;
  int64_t blocks = (rows + rpw * kGGWarps - 1) / (rpw * kGGWarps);
  if (blocks < 1) blocks = 1;
  *rows_per_warp = (int)rpw;
  return (int)blocks;
}
}  // namespace

template <typename real>
int gridgen_forward_t(const void *T, void *out, int64_t B, int64_t D, int64_t H, int64_t W, cudaStream_t st) {
  const long long total_rows = (long long)B * D * H;
  // persistent-style: ~8 CTAs per SM, each warp owns a contiguous run of rows
  long long blocks = (long long)kSMs * 8;
  long long rows_per_warp = (total_rows + blocks * kGGWarps - 1) / (blocks * kGGWarps);
  if (rows_per_warp < 1) rows_per_warp = 1;
  blocks = (total_rows + rows_per_warp * kGGWarps - 1) / (rows_per_warp * kGGWarps);
  const size_t smem = (size_t)(D + H + W) * sizeof(real);
  if (smem > 200 * 1024) return VOLSTN_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    VOLSTN_CUDA_TRY(cudaFuncSetAttribute(k_gridgen_fwd<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_gridgen_fwd<real><<<(unsigned)blocks, kGGThreads, smem, st>>>(static_cast<const real *>(T),
                                                                   static_cast<real *>(out), (int)B, (int)D, (int)H,
                                                                   (int)W, rows_per_warp);
  return after_launch();
}

template <typename real>
int gridgen_backward_t(const void *gg, void *gradT, int64_t B, int64_t D, int64_t H, int64_t W, void *ws,
                       cudaStream_t st) {
  int rpw;
  const int nblk = bwd_blocks_per_sample(B, D, H, &rpw);
  const size_t smem = (size_t)(D + H + W) * sizeof(real);
  if (smem > 200 * 1024) return VOLSTN_ERR_UNSUPPORTED;
  if (smem > 40 * 1024)
    VOLSTN_CUDA_TRY(
        cudaFuncSetAttribute(k_gridgen_bwd_partial<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int64_t b0 = 0; b0 < B; b0 += 65535) {
    const int nb = (int)((B - b0 < 65535) ? B - b0 : 65535);
    double *part = static_cast<double *>(ws) + b0 * nblk * 16;
    k_gridgen_bwd_partial<real><<<dim3(nblk, nb), kGGThreads, smem, st>>>(
        static_cast<const real *>(gg) + b0 * D * H * W * 4, part, (int)D, (int)H, (int)W, rpw);
    int rc = after_launch();
    if (rc) return rc;
    k_gridgen_bwd_final<real><<<nb, 16, 0, st>>>(part, static_cast<real *>(gradT) + b0 * 16, nblk);
    rc = after_launch();
    if (rc) return rc;
  }
  return VOLSTN_OK;
}

}  // namespace volstn

extern "C" {

int volstn_b200_gridgen_forward(int dtype, const void *T, void *out, int64_t B, int64_t D, int64_t H, int64_t W,
                                void *cuda_stream) {
  if (B < 0 || D < 2 || H < 2 || W < 2) return VOLSTN_ERR_SHAPE;  // lua:6-8 assert(depth > 1) ...
  if (D > 0x7fffffffLL || H > 0x7fffffffLL || W > 0x7fffffffLL || B > 0x7fffffffLL) return VOLSTN_ERR_ARG;
  if (B == 0) return VOLSTN_OK;
  if (!T || !out) return VOLSTN_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (dtype == VOLSTN_F32) return volstn::gridgen_forward_t<float>(T, out, B, D, H, W, st);
  if (dtype == VOLSTN_F64) return volstn::gridgen_forward_t<double>(T, out, B, D, H, W, st);
  return VOLSTN_ERR_ARG;
}

size_t volstn_b200_gridgen_backward_workspace(int64_t B, int64_t D, int64_t H, int64_t W) {
  (void)W;
  if (B <= 0 || D <= 0 || H <= 0) return 0;
  int rpw;
  const int nblk = volstn::bwd_blocks_per_sample(B, D, H, &rpw);
  return (size_t)B * nblk * 16 * sizeof(double);
}

int volstn_b200_gridgen_backward(int dtype, const void *gradGrid, void *gradT, int64_t B, int64_t D, int64_t H,
                                 int64_t W, void *workspace, size_t workspace_bytes, void *cuda_stream) {
  if (B < 0 || D < 2 || H < 2 || W < 2) return VOLSTN_ERR_SHAPE;
  if (D > 0x7fffffffLL || H > 0x7fffffffLL || W > 0x7fffffffLL || B > 0x7fffffffLL) return VOLSTN_ERR_ARG;
  if (B == 0) return VOLSTN_OK;
  if (!gradGrid || !gradT || !workspace) return VOLSTN_ERR_ARG;
  if (workspace_bytes < volstn_b200_gridgen_backward_workspace(B, D, H, W)) return VOLSTN_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (dtype == VOLSTN_F32) return volstn::gridgen_backward_t<float>(gradGrid, gradT, B, D, H, W, workspace, st);
  if (dtype == VOLSTN_F64) return volstn::gridgen_backward_t<double>(gradGrid, gradT, B, D, H, W, workspace, st);
  return VOLSTN_ERR_ARG;
}

}  // extern "C"
