/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY (never linked or imported by the product).
 *
 * Compiles the reference's own CPU implementation of the volumetric sampler,
 *   /root/reference/stn3dtorch/generic/BilinearSamplerVolumetric.c
 * UNMODIFIED, from where it lies (the path is given on the command line as REF_GENERIC_FILE;
 * no reference source is copied into this repository).  The reference file is a Torch7
 * "generic" file: it expects TH / luaT / Lua to be present (stn3dtorch/init.c:1-9).  None of
 * those exist in this image, so this shim supplies the ~10 names the file really uses and
 * nothing else:
 *
 *   real, THTensor{size,stride}, THTensor_(data), luaT_checkudata, lua_tonumber, lua_State,
 *   luaL_Reg, luaT_pushmetatable, luaT_registeratname, lua_pop, nn_(), torch_Tensor
 *
 * Build (see oracle/Makefile):   gcc -O2 -ffp-contract=off -DREF_REAL=float|double ...
 * -ffp-contract=off is mandatory: the parity contract is the reference's expression order
 * evaluated without FMA contraction (hard part 4).
 *
 * Exported entry point:  ref_call(fn, tensors...)  with fn =
 *   0  BilinearSamplerVolumetric_updateOutput      (reference ...c:442)   G=4
 *   1  BilinearSamplerVolumetric_updateGradInput   (reference ...c:585)   G=4
 *   2  BilinearSamplerBHWD_updateOutput            (reference ...c:9)     G=3
 *   3  BilinearSamplerBHWD_updateGradInput         (reference ...c:169)   G=3
 */
#include <math.h>
#include <stddef.h>

#ifndef REF_REAL
#error "define REF_REAL to float or double"
#endif
#ifndef REF_GENERIC_FILE
#error "define REF_GENERIC_FILE to the absolute path of the reference's generic .c file"
#endif

typedef REF_REAL real;

/* The two TH tensor fields the reference dereferences directly (...c:15-37, 449-473). */
typedef struct THTensor {
  long *size;
  long *stride;
  real *data;
} THTensor;

/* A "Lua stack" that is just the argument slots luaT_checkudata() is asked for (2..6). */
typedef struct lua_State {
  void *slot[8];
} lua_State;

typedef int (*lua_CFunction)(lua_State *);
typedef struct luaL_Reg {
  const char *name;
  lua_CFunction func;
} luaL_Reg;

#define torch_Tensor "shim.Tensor"
#define REF_CAT_(a, b) a##b
#define REF_CAT(a, b) REF_CAT_(a, b)
#define nn_(NAME) REF_CAT(nn_shim_, NAME)
#define THTensor_(NAME) REF_CAT(THShimTensor_, NAME)

static inline void *luaT_checkudata(lua_State *L, int idx, const char *tname) {
  (void)tname;
  return L->slot[idx];
}
static inline double lua_tonumber(lua_State *L, int idx) {
  (void)L;
  (void)idx;
  return 0.0; /* focal_length: read and ignored by the reference (...c:447,592) */
}
static inline real *THTensor_(data)(THTensor *t) { return t->data; }
static inline void luaT_pushmetatable(lua_State *L, const char *tname) {
  (void)L;
  (void)tname;
}
static inline void luaT_registeratname(lua_State *L, const luaL_Reg *regs, const char *name) {
  (void)L;
  (void)regs;
  (void)name;
}
static inline void lua_pop(lua_State *L, int n) {
  (void)L;
  (void)n;
}

/* Select the body of the reference file (its #ifndef TH_GENERIC_FILE guard, ...c:1-3). */
#define TH_GENERIC_FILE "generic/BilinearSamplerVolumetric.c"
#include REF_GENERIC_FILE

/* ------------------------------------------------------------------------------------------
 * ctypes-facing entry point.  t[i] describe up to five 5-D tensors in the order of the
 * reference's Lua stack slots 2..6 (fwd: vol, grid, out; bwd: vol, grid, gradVol, gradGrid,
 * gradOut -- ...c:444-446, 587-591).  Sizes/strides are in elements.
 * ---------------------------------------------------------------------------------------- */
typedef struct ref_tensor5 {
  void *data;
  long size[5];
  long stride[5];
} ref_tensor5;

int ref_call(int fn, int ntensors, ref_tensor5 *t) {
  THTensor th[5];
  lua_State L;
  int i;
  if (ntensors < 3 || ntensors > 5) return -1;
  for (i = 0; i < 8; i++) L.slot[i] = NULL;
  for (i = 0; i < ntensors; i++) {
    th[i].size = t[i].size;
    th[i].stride = t[i].stride;
    th[i].data = (real *)t[i].data;
    L.slot[2 + i] = &th[i];
  }
  switch (fn) {
    case 0: return nn_(BilinearSamplerVolumetric_updateOutput)(&L);
    case 1: return nn_(BilinearSamplerVolumetric_updateGradInput)(&L);
    case 2: return nn_(BilinearSamplerBHWD_updateOutput)(&L);
    case 3: return nn_(BilinearSamplerBHWD_updateGradInput)(&L);
    default: return -2;
  }
}

int ref_sizeof_real(void) { return (int)sizeof(real); }

/* keep the reference's registration table referenced so -Wunused does not hide a real problem */
void ref_touch_registration(void) { nn_(BilinearSamplerVolumetric_init)(NULL); }
