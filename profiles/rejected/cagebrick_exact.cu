// cagebrick_exact.cu -- the BRICK mapping of cagebrick.cu with the REFERENCE's arithmetic: `out` (and the gradGrid that
// feeds the cage gradient) keep the bits of the un-fused module chain (models/alignet_2d.lua:136-147 lifted to 3-D:
// cage up-sample = BilinearSamplerVolumetric with C = 4 at the identity coordinates, then the source warp).
//
// What the mapping still buys without touching a rounding (row kernels: 596 warp instructions per 32 voxels of the step,
// 213 of them the reference's flops):
//   * a lane's x is fixed and a task stays inside one cage cell (the last unit of an axis also holds the position ON the
//     last cage node, whose cell is the zero-padded one -- the only cell change inside a task), so the lane's eight cage
//     corners x three channels are REGISTERS for the whole task: no shared-memory corner loads, no per-voxel cell index;
//   * the up-sampled field is the C = 4 sampler forward of ...c:554-576, operation for operation:
//     w_k = ((wx'.wy').wz'), sum over the corners 0..7 of w_k.v_k (products and sums separately rounded);
//   * the source warp is sampler.cuh's (VoxelSetup tiers, corner weights, sum order, grad_grid_from_dots);
//   * the gradient of the cage is NOT order-sensitive in the reference either (its CUDA path scatters with atomicAdd):
//     it is accumulated in the separable form of cagebrick.cu -- S[j][c] += wy_j . g_c per voxel, A[j][k][c] += wz_k . S
//     per z, one segmented warp reduction per cell -- and checked against the oracle chain at 1e-4.
#include <cstdlib>
#include <cstring>

#include "cagebrick.cuh"

namespace volstn {

// corner k: x = k & 1, y = (k >> 1) & 1, z = k >> 2 (the sampler's numbering, volstn_common.cuh)
struct ExactLane {
  float V[8][3];  // the lane's cage corners (zeros beyond the cage: the sampler's padding, ...c:542-564)
  float S[2][3];
  double lsum;    // THNN accumulates the criterion in double
  int ni, nu;
};

template <int ORDER>
__device__ __forceinline__ float brick_sum8(const float (&cw)[8], const float (&v)[8]) {
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int k = ORDER ? order_zxy(j) : order_xyz(j);
    const float term = __fmul_rn(cw[k], v[k]);
    acc = j ? __fadd_rn(acc, term) : term;
  }
  return acc;
}

// TIER: 1 = every corner of the warp's rows in bounds, 2 = inside including the high faces, 0 = exact predicated path
template <int MODE, bool GV, int VS, int NR, int TIER>
__device__ __forceinline__ void exact_body(const BrickArgs &a, const float *__restrict__ vb, VoxelSetup<float> (&su)[NR],
                                           const float2 (&wy)[NR], const float (&tg)[NR], long long e0, bool valid,
                                           ExactLane &L) {
  const SamplerArgs<float> &s = a.s;
  const int sH = VS ? VS : a.iW, sD = VS ? VS * VS : a.iH * a.iW;
  constexpr bool FAST = TIER == 1;
#pragma unroll
  for (int r = 0; r < NR; r++) {
    if constexpr (TIER == 1) su[r].template finish_fast<false>(s, sD, sH, 1);
    if constexpr (TIER == 2) su[r].template finish_fast<true>(s, sD, sH, 1);
    if constexpr (TIER == 0) su[r].finish_exact(s);
  }
  float v[NR][8];
  const float *p0[NR];
#pragma unroll
  for (int r = 0; r < NR; r++) {
    p0[r] = vb + (TIER ? su[r].fbase : su[r].base32(sD, sH, 1));
#pragma unroll
    for (int k = 0; k < 8; k++)
      v[r][k] = (FAST || su[r].in(k)) ? __ldg(p0[r] + (k & 1) + ((k & 2) ? sH : 0) + ((k & 4) ? sD : 0)) : 0.f;
  }
#pragma unroll
  for (int r = 0; r < NR; r++) {
    const long long e = e0 + (long long)r * a.oW;
    float cw[8];
    if constexpr (MODE != 1 || GV) su[r].corner_weights(cw);
    if constexpr (MODE == 0) {
      if (valid) st_stream_f1(a.out + e, brick_sum8<0>(cw, v[r]));  // ...c:566-573
      continue;
    }
    float go;
    if constexpr (MODE == 2) {
      const float o = brick_sum8<0>(cw, v[r]);
      const float x = __fsub_rn(o, tg[r]);
      // THNN SmoothL1Criterion_updateGradInput: norm * clamp(x, -1, 1); _updateOutput: z = |x|, z < 1 ? 0.5 z^2 : z - 0.5 (double)
      go = x < -1.0f ? -a.loss_norm : (x > 1.0f ? a.loss_norm : __fmul_rn(a.loss_norm, x));
      if (valid) {
        const double zz = (double)fabsf(x);
        L.lsum += zz < 1.0 ? 0.5 * zz * zz : zz - 0.5;
        const bool p = o > 0.5f, q = tg[r] > 0.5f;  // computeIOU (util.lua:110-120)
        L.ni += (p && q) ? 1 : 0;
        L.nu += (p || q) ? 1 : 0;
        if (a.out) st_stream_f1(a.out + e, o);
      }
    } else {
      go = tg[r];
    }
    float dot[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const float d = __fadd_rn(0.f, __fmul_rn(v[r][k], go));  // dot += v * go (...c:736)
      dot[k] = (FAST || su[r].in(k)) ? d : 0.f;                // an out-of-bounds corner never sees go
    }
    if constexpr (GV) {
      float *g0p = const_cast<float *>(p0[r]) + a.gv_delta;
#pragma unroll
      for (int k = 0; k < 8; k++)
        if (valid && (FAST || su[r].in(k)))
          red_add(g0p + (k & 1) + ((k & 2) ? sH : 0) + ((k & 4) ? sD : 0), __fmul_rn(cw[k], go));  // ...c:737
    }
    float gz, gy, gx;
    grad_grid_from_dots<float>(dot, su[r], s, gz, gy, gx);  // ...c:792-822
    if (!valid) gz = gy = gx = 0.f;
    L.S[0][0] = fmaf(wy[r].x, gz, L.S[0][0]), L.S[1][0] = fmaf(wy[r].y, gz, L.S[1][0]);
    L.S[0][1] = fmaf(wy[r].x, gy, L.S[0][1]), L.S[1][1] = fmaf(wy[r].y, gy, L.S[1][1]);
    L.S[0][2] = fmaf(wy[r].x, gx, L.S[0][2]), L.S[1][2] = fmaf(wy[r].y, gx, L.S[1][2]);
  }
}

template <int MODE, bool GV, int VS, int NR>
__device__ __forceinline__ void exact_rows(const BrickArgs &a, const float *__restrict__ vb, float2 wx, float2 wz,
                                           const float2 (&wy)[NR], long long e0, bool valid, ExactLane &L) {
  VoxelSetup<float> su[NR];
  float tg[NR];
  int tier = 2;
#pragma unroll
  for (int r = 0; r < NR; r++) {
    // the cage volume sampled at the identity coordinate: corner weights ((wx'.wy').wz'), corners summed 0..7 (...c:554-576)
    float xy[4];
    xy[0] = __fmul_rn(wx.x, wy[r].x), xy[1] = __fmul_rn(wx.y, wy[r].x), xy[2] = __fmul_rn(wx.x, wy[r].y), xy[3] = __fmul_rn(wx.y, wy[r].y);
    float f[3];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const float w = __fmul_rn(xy[k & 3], (k & 4) ? wz.y : wz.x);
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const float p = __fmul_rn(w, L.V[k][c]);
        f[c] = k ? __fadd_rn(f[c], p) : p;
      }
    }
    tier = min(tier, su[r].begin(f[0], f[1], f[2], a.s));
    tg[r] = 0.f;
    if constexpr (MODE != 0) tg[r] = ld_stream_f1(a.tgt + e0 + (long long)r * a.oW);
  }
  if (__all_sync(0xffffffffu, tier == 2))
    exact_body<MODE, GV, VS, NR, 1>(a, vb, su, wy, tg, e0, valid, L);
  else if (__all_sync(0xffffffffu, tier >= 1))
    exact_body<MODE, GV, VS, NR, 2>(a, vb, su, wy, tg, e0, valid, L);
  else
    exact_body<MODE, GV, VS, NR, 0>(a, vb, su, wy, tg, e0, valid, L);
}

#ifndef BRICKX_MIN_CTAS
#define BRICKX_MIN_CTAS 2
#endif
#ifndef BRICKX_NR
#define BRICKX_NR 1
#endif

template <int MODE, bool GV, int VS>
__global__ void __launch_bounds__(kBrickThreads, BRICKX_MIN_CTAS) k_cage_brick_exact(const BrickArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BrickSmem m;
  brick_fill<true>(m, smem_raw, a);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int b = blockIdx.y;
  int t = blockIdx.x * kBrickWarps + wid;
  if (t >= a.ntask) return;
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
  if (za >= zb || ya >= yb) return;

  const int x = xs * 32 + lane;
  const bool valid = x < a.oW;
  const int xc = valid ? x : a.oW - 1;
  const int ix0 = m.ix[xc];
  const float2 wx = m.wx[xc];
  const long long plane = (long long)a.cD * a.cH * a.cW;
  const float *cb = a.cage + (long long)b * 3 * plane;
  float *gb = (MODE != 0) ? a.gcage + (long long)b * 3 * plane : nullptr;
  float *sA = m.lane + threadIdx.x;  // sA[(j, k, c)]: gradient of the cell's corner (y: j, z: k), still to be spread along x
  ExactLane L;
  L.lsum = 0.0, L.ni = 0, L.nu = 0;
  int ciz = -1, ciy = -1;  // the cell the corner registers hold

  // segmented reduction over the lanes of one cage column (consecutive lanes), heads add to the cage gradient
  bool same[5];
#pragma unroll
  for (int q = 0; q < 5; q++) {
    const int ok = __shfl_down_sync(0xffffffffu, ix0, 1 << q);
    same[q] = lane + (1 << q) < 32 && ok == ix0;
  }
  const int ix_prev = __shfl_up_sync(0xffffffffu, ix0, 1);
  const bool head = lane == 0 || ix_prev != ix0;
  auto flush = [&]() {
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int k = 0; k < 2; k++)
#pragma unroll
        for (int c = 0; c < 3; c++) {
          float *slot = sA + ((j * 2 + k) * 3 + c) * kBrickThreads;
          const float g = *slot;
          *slot = 0.f;
          float lo = g * wx.x, hi = g * wx.y;
#pragma unroll
          for (int q = 0; q < 5; q++) {
            const float l2 = __shfl_down_sync(0xffffffffu, lo, 1 << q), h2 = __shfl_down_sync(0xffffffffu, hi, 1 << q);
            if (same[q]) lo += l2, hi += h2;
          }
          if (head && ciz + k < a.cD && ciy + j < a.cH) {
            float *q = gb + c * plane + ((long long)(ciz + k) * a.cH + ciy + j) * a.cW;
            if (lo != 0.0f) red_add(q + ix0, lo);
            if (hi != 0.0f && ix0 + 1 < a.cW) red_add(q + ix0 + 1, hi);
          }
        }
  };
  if constexpr (MODE != 0) {
#pragma unroll
    for (int i = 0; i < 12; i++) sA[i * kBrickThreads] = 0.f;
  }

  const float *vb = opaque(a.vol + (long long)b * a.vB);
  const long long orow = (long long)a.oH * a.oW;
  const long long eb = (long long)b * a.oD * orow + xc;

  for (int z = za; z < zb; z++) {
    const float2 wz = m.wz[z];
    const int iz = m.iz[z];
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int c = 0; c < 3; c++) L.S[j][c] = 0.f;
    const long long ro = eb + (long long)z * orow;
    int y = ya;
    while (y < yb) {
      const int iy = m.iy[y];
      if (iz != ciz || iy != ciy) {  // warp-uniform: at most twice per axis and task
        if constexpr (MODE != 0) {
          if (ciz >= 0) flush();
        }
        ciz = iz, ciy = iy;
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const int cx = ix0 + (k & 1), cy = iy + ((k >> 1) & 1), cz = iz + (k >> 2);
          const bool in = cx < a.cW && cy < a.cH && cz < a.cD;
          const long long o = ((long long)cz * a.cH + cy) * a.cW + cx;
#pragma unroll
          for (int c = 0; c < 3; c++) L.V[k][c] = in ? __ldg(cb + c * plane + o) : 0.f;
        }
      }
      int n = 1;
      if (BRICKX_NR > 1 && y + BRICKX_NR <= yb && m.iy[y + BRICKX_NR - 1] == iy) n = BRICKX_NR;
      if (n == BRICKX_NR && BRICKX_NR > 1) {
        float2 wy[BRICKX_NR];
#pragma unroll
        for (int r = 0; r < BRICKX_NR; r++) wy[r] = m.wy[y + r];
        exact_rows<MODE, GV, VS, BRICKX_NR>(a, vb, wx, wz, wy, ro + (long long)y * a.oW, valid, L);
      } else {
        const float2 wy[1] = {m.wy[y]};
        exact_rows<MODE, GV, VS, 1>(a, vb, wx, wz, wy, ro + (long long)y * a.oW, valid, L);
      }
      y += n;
      if constexpr (MODE != 0) {
        if (y >= yb || m.iy[y] != ciy) {  // the rows of this z inside the cell are done: fold along z
#pragma unroll
          for (int j = 0; j < 2; j++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
              float *q = sA + ((j * 2) * 3 + c) * kBrickThreads;
              q[0] = fmaf(wz.x, L.S[j][c], q[0]);
              q[3 * kBrickThreads] = fmaf(wz.y, L.S[j][c], q[3 * kBrickThreads]);
              L.S[j][c] = 0.f;
            }
        }
      }
    }
  }
  if constexpr (MODE != 0) flush();

  if constexpr (MODE == 2) {
    double ls = L.lsum;
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
}

template <int MODE, bool GV, int VS>
static int exact_launch_v(const BrickArgs &a, cudaStream_t st) {
  const size_t smem = brick_smem_bytes(a);
  if (smem > 48 * 1024)
    VOLSTN_CUDA_TRY(cudaFuncSetAttribute(k_cage_brick_exact<MODE, GV, VS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nblk = (a.ntask + kBrickWarps - 1) / kBrickWarps;
  const long long cplane = 3LL * a.cD * a.cH * a.cW, oplane = (long long)a.oD * a.oH * a.oW;
  for (int b0 = 0; b0 < a.B; b0 += 65535) {
    BrickArgs c = a;
    const int nb = a.B - b0 < 65535 ? a.B - b0 : 65535;
    c.vol += (long long)b0 * a.vB;
    c.cage += (long long)b0 * cplane;
    if (c.gcage) c.gcage += (long long)b0 * cplane;
    if (c.tgt) c.tgt += (long long)b0 * oplane;
    if (c.out) c.out += (long long)b0 * oplane;
    if (c.loss) c.loss += b0;
    if (c.iou) c.iou += 2LL * b0;
    k_cage_brick_exact<MODE, GV, VS><<<dim3(nblk, nb), kBrickThreads, smem, st>>>(c);
    int rc = after_launch();
    if (rc) return rc;
  }
  return VOLSTN_OK;
}

template <int MODE, bool GV>
static int exact_launch_t(const BrickArgs &a, cudaStream_t st) {
  if (a.iW == a.iH && !getenv("VOLSTN_BRICK_NO_VS")) {
    if (a.iW == 32) return exact_launch_v<MODE, GV, 32>(a, st);
    if (a.iW == 64) return exact_launch_v<MODE, GV, 64>(a, st);
    if (a.iW == 128) return exact_launch_v<MODE, GV, 128>(a, st);
  }
  return exact_launch_v<MODE, GV, 0>(a, st);
}

int brick_exact_launch(const BrickArgs &a, int mode, bool grad_vol, cudaStream_t st) {
  switch (mode) {
    case 0: return exact_launch_t<0, false>(a, st);
    case 1: return grad_vol ? exact_launch_t<1, true>(a, st) : exact_launch_t<1, false>(a, st);
    default: return grad_vol ? exact_launch_t<2, true>(a, st) : exact_launch_t<2, false>(a, st);
  }
}

}  // namespace volstn
