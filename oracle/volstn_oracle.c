/*
 * oracle/volstn_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C CPU restatement of the arithmetic of the reference's volumetric sampler,
 *   stn3dtorch/generic/BilinearSamplerVolumetric.c     (paths relative to /root/reference)
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared object, and only as the checker or the timed CPU baseline.  The
 * product (alignet_b200/, include/) never links, imports or falls back to it.
 *
 * PARITY STATUS: pinned.  tests/test_oracle_golden.py checks every function below bit-for-bit
 * (float32 and float64) against the .npz files in tests/golden, which were produced by the reference's own
 * C file compiled unmodified (oracle/ref_shim.c -> oracle/_ref/, script
 * tests/golden/make_golden.py), and tests/test_oracle_vs_ref.py repeats the comparison live on
 * fresh random inputs wherever oracle/_ref/ is present.
 *
 * The file is written as one corner-indexed loop instead of the reference's eight spelled-out
 * corners; corner k has  x = x0 + (k&1),  y = y0 + ((k>>1)&1),  z = z0 + ((k>>2)&1), i.e.
 *   k:  0=FrontTopLeft 1=FrontTopRight 2=FrontBottomLeft 3=FrontBottomRight
 *       4=BackTopLeft  5=BackTopRight  6=BackBottomLeft  7=BackBottomRight      (...c:521-529)
 * What must be (and is) preserved exactly is the ORDER of floating point operations:
 *   coordinate      ((g + 1) * (size-1)) / 2 ; i0 = floor ; w = 1 - (c - i0)    ...c:507-517
 *   corner weight   ((wx' * wy') * wz')                                         ...c:566-573
 *   forward  G=4    sum over k = 0,1,2,3,4,5,6,7                                ...c:566-573
 *   forward  G=3    sum over k = 0,4,1,5,2,6,3,7   (TLF,TLB,TRF,TRB,...)        ...c:149-156
 *   backward        dot_k += v_k*go ; gradVol_k += (((wx'*wy')*wz')*go)         ...c:730-789
 *   gradGrid        sums over k = 0,4,1,5,2,6,3,7 of ((dot_k*a)*b)*(+-1)        ...c:792-817
 *                   then (d * (size-1)) / 2, 4th component = 0 (G=4)            ...c:819-822
 * Compile with -ffp-contract=off (oracle/Makefile).
 *
 * One deliberate, documented difference: the reference forms addresses in `int` (...c:519-529)
 * and converts floor() to `int` with x86 cvttsd2si semantics (out-of-range / NaN -> INT_MIN).
 * We form addresses in `long` (identical wherever the reference does not overflow) and spell the
 * x86 conversion out in floor_to_int() so that the behaviour is defined C.
 */
#ifndef VOLSTN_ORACLE_BODY

#include <limits.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

typedef struct oracle_tensor5 {
  void *data;
  long size[5];   /* B, D, H, W, C|G */
  long stride[5]; /* element strides; the last one must be 1 (reference assumes it: ...c:492-495, 556) */
} oracle_tensor5;

/* (int)floor(c) as the reference's x86-64 build performs it: cvttsd2si yields the "integer
 * indefinite" value INT_MIN for NaN and for anything outside [-2^31, 2^31). */
static int floor_to_int(double c) {
  double f = floor(c);
  if (!(f >= -2147483648.0 && f < 2147483648.0)) return INT_MIN;
  return (int)f;
}

/* corner visiting orders */
static const int ORDER_XYZ[8] = {0, 1, 2, 3, 4, 5, 6, 7}; /* G=4 forward, G=4 backward channel loop */
static const int ORDER_ZXY[8] = {0, 4, 1, 5, 2, 6, 3, 7}; /* G=3 forward, all gradGrid sums, G=3 channel loop */

#define VOLSTN_ORACLE_BODY

#define real float
#define FN(name) name##_f32
#include "volstn_oracle.c"
#undef real
#undef FN

#define real double
#define FN(name) name##_f64
#include "volstn_oracle.c"
#undef real
#undef FN

int oracle_abi_version(void) { return 1; }

#else /* ---------------------------- per-type body ---------------------------- */

typedef struct FN(voxel_setup) {
  int i0[3];     /* x0, y0, z0  (floor indices)                     */
  real w[3];     /* wx, wy, wz  (weight of the LOW corner per axis) */
  int in[8];     /* corner k inside the volume?                     */
  long off[8];   /* element offset of corner k inside a sample (strides of the tensor given) */
  real cw[8];    /* ((wx' * wy') * wz') for corner k                */
} FN(voxel_setup);

/* ...c:507-517 (G=4) and ...c:63-73 (G=3): identical expressions. `n` is the volume extent. */
static void FN(axis)(real g, long n, int *i0, real *w) {
  real c = (g + 1) * (real)(n - 1) / 2;
  *i0 = floor_to_int((double)c);
  *w = 1 - (c - (real)*i0);
}

static void FN(setup)(real zf, real yf, real xf, const long size[5], const long vstride[5],
                      FN(voxel_setup) * s) {
  int k;
  FN(axis)(xf, size[3], &s->i0[0], &s->w[0]);
  FN(axis)(yf, size[2], &s->i0[1], &s->w[1]);
  FN(axis)(zf, size[1], &s->i0[2], &s->w[2]);
  for (k = 0; k < 8; k++) {
    /* int arithmetic exactly as the reference's masks: `x0+1 >= 0 && x0+1 <= W-1` (...c:542-550).
     * x0 == INT_MIN makes x0+1 negative; x0 == INT_MAX cannot occur (floor_to_int). */
    long x = (long)s->i0[0] + (k & 1);
    long y = (long)s->i0[1] + ((k >> 1) & 1);
    long z = (long)s->i0[2] + ((k >> 2) & 1);
    real wx = (k & 1) ? (1 - s->w[0]) : s->w[0];
    real wy = (k & 2) ? (1 - s->w[1]) : s->w[1];
    real wz = (k & 4) ? (1 - s->w[2]) : s->w[2];
    s->in[k] = x >= 0 && x <= size[3] - 1 && y >= 0 && y <= size[2] - 1 && z >= 0 && z <= size[1] - 1;
    s->off[k] = z * vstride[1] + y * vstride[2] + x * vstride[3];
    s->cw[k] = wx * wy * wz;
  }
}

/* Forward: reference ...c:442-583 (G=4) and ...c:9-167 (G=3). */
int FN(oracle_sampler_forward)(int G, const oracle_tensor5 *vol, const oracle_tensor5 *grid,
                               oracle_tensor5 *out) {
  const int *order = (G == 4) ? ORDER_XYZ : ORDER_ZXY;
  const real *V = (const real *)vol->data;
  const real *Gd = (const real *)grid->data;
  real *O = (real *)out->data;
  long b, z, y, x, t;
  long C = vol->size[4];
  if (G != 3 && G != 4) return -1;
  for (b = 0; b < vol->size[0]; b++)
    for (z = 0; z < out->size[1]; z++)
      for (y = 0; y < out->size[2]; y++)
        for (x = 0; x < out->size[3]; x++) {
          const real *g = Gd + b * grid->stride[0] + z * grid->stride[1] + y * grid->stride[2] + x * grid->stride[3];
          real *o = O + b * out->stride[0] + z * out->stride[1] + y * out->stride[2] + x * out->stride[3];
          const real *vb = V + b * vol->stride[0];
          FN(voxel_setup) s;
          FN(setup)(g[0], g[1], g[2], vol->size, vol->stride, &s);
          for (t = 0; t < C; t++) {
            real acc = 0;
            int j;
            for (j = 0; j < 8; j++) {
              int k = order[j];
              real v = s.in[k] ? vb[s.off[k] + t] : 0; /* out-of-bounds corner reads as 0 (...c:531-564) */
              real term = s.cw[k] * v;
              acc = (j == 0) ? term : acc + term;
            }
            o[t] = acc;
          }
        }
  return 0;
}

/* Backward: reference ...c:585-829 (G=4) and ...c:169-439 (G=3).
 * gradVol is ACCUMULATED into (the Lua wrapper zero-fills it first, BilinearSamplerVolumetric.lua:101-104);
 * gradGrid is ASSIGNED.  only_grid reproduces the reference's (never enabled) onlyGrid switch (...c:594). */
int FN(oracle_sampler_backward)(int G, const oracle_tensor5 *vol, const oracle_tensor5 *grid,
                                oracle_tensor5 *gradVol, oracle_tensor5 *gradGrid,
                                const oracle_tensor5 *gradOut, int only_grid) {
  const int *chan_order = (G == 4) ? ORDER_XYZ : ORDER_ZXY;
  const real *V = (const real *)vol->data;
  const real *Gd = (const real *)grid->data;
  const real *GO = (const real *)gradOut->data;
  real *GV = (real *)gradVol->data;
  real *GG = (real *)gradGrid->data;
  long b, z, y, x, t;
  long C = vol->size[4];
  if (G != 3 && G != 4) return -1;
  for (b = 0; b < vol->size[0]; b++)
    for (z = 0; z < gradOut->size[1]; z++)
      for (y = 0; y < gradOut->size[2]; y++)
        for (x = 0; x < gradOut->size[3]; x++) {
          const real *g = Gd + b * grid->stride[0] + z * grid->stride[1] + y * grid->stride[2] + x * grid->stride[3];
          const real *go = GO + b * gradOut->stride[0] + z * gradOut->stride[1] + y * gradOut->stride[2] + x * gradOut->stride[3];
          real *gg = GG + b * gradGrid->stride[0] + z * gradGrid->stride[1] + y * gradGrid->stride[2] + x * gradGrid->stride[3];
          const real *vb = V + b * vol->stride[0];
          real *gvb = GV + b * gradVol->stride[0];
          FN(voxel_setup) s, sg;
          real dot[8];
          real dz, dy, dx;
          int j, k;
          FN(setup)(g[0], g[1], g[2], vol->size, vol->stride, &s);
          FN(setup)(g[0], g[1], g[2], vol->size, gradVol->stride, &sg); /* gradVol has its own strides (...c:686) */
          for (k = 0; k < 8; k++) dot[k] = 0;
          for (t = 0; t < C; t++) {
            real gov = go[t];
            for (j = 0; j < 8; j++) {
              k = chan_order[j];
              if (s.in[k]) {
                dot[k] += vb[s.off[k] + t] * gov;
                if (!only_grid) gvb[sg.off[k] + t] += s.cw[k] * gov;
              }
            }
          }
          dz = dy = dx = 0;
          for (j = 0; j < 8; j++) {
            real wxs, wys, wzs, ty, tx, tz;
            k = ORDER_ZXY[j];
            wxs = (k & 1) ? (1 - s.w[0]) : s.w[0];
            wys = (k & 2) ? (1 - s.w[1]) : s.w[1];
            wzs = (k & 4) ? (1 - s.w[2]) : s.w[2];
            ty = dot[k] * wxs * wzs * ((k & 2) ? 1 : -1); /* ...c:792-799 */
            tx = dot[k] * wys * wzs * ((k & 1) ? 1 : -1); /* ...c:801-808 */
            tz = dot[k] * wys * wxs * ((k & 4) ? 1 : -1); /* ...c:810-817 */
            dy = (j == 0) ? ty : dy + ty;
            dx = (j == 0) ? tx : dx + tx;
            dz = (j == 0) ? tz : dz + tz;
          }
          gg[0] = dz * (real)(vol->size[1] - 1) / 2; /* ...c:819 */
          gg[1] = dy * (real)(vol->size[2] - 1) / 2; /* ...c:820 */
          gg[2] = dx * (real)(vol->size[3] - 1) / 2; /* ...c:821 */
          if (G == 4) gg[3] = 0;                     /* ...c:822 */
        }
  return 0;
}

/* Debug dump used by the bit-exact index/mask parity test: idx[v*3+{0,1,2}] = (z0, y0, x0) and
 * mask[v] bit k = corner k in bounds, for v in raster order over (b, z, y, x). */
int FN(oracle_sampler_indices)(const long vol_size[5], const oracle_tensor5 *grid, int32_t *idx,
                               uint8_t *mask) {
  const real *Gd = (const real *)grid->data;
  long fake_stride[5] = {0, 0, 0, 0, 1};
  long b, z, y, x, v = 0;
  for (b = 0; b < grid->size[0]; b++)
    for (z = 0; z < grid->size[1]; z++)
      for (y = 0; y < grid->size[2]; y++)
        for (x = 0; x < grid->size[3]; x++, v++) {
          const real *g = Gd + b * grid->stride[0] + z * grid->stride[1] + y * grid->stride[2] + x * grid->stride[3];
          FN(voxel_setup) s;
          int k;
          uint8_t m = 0;
          FN(setup)(g[0], g[1], g[2], vol_size, fake_stride, &s);
          for (k = 0; k < 8; k++) m |= (uint8_t)(s.in[k] << k);
          idx[v * 3 + 0] = s.i0[2];
          idx[v * 3 + 1] = s.i0[1];
          idx[v * 3 + 2] = s.i0[0];
          mask[v] = m;
        }
  return 0;
}

#endif /* VOLSTN_ORACLE_BODY */
