#ifndef TH_GENERIC_FILE
#define TH_GENERIC_FILE "generic/volstn_host.c"
#else
/*
 * generic/volstn_host.c -- the torch.FloatTensor / torch.DoubleTensor functions of libvolstn, instantiated for
 * both types by THGenerateFloatTypes.h exactly like the reference instantiates
 * stn3dtorch/generic/BilinearSamplerVolumetric.c (stn3dtorch/init.c:4-9).  Every function is a lua_CFunction
 * with the reference's stack layout and forwards to the HOST-buffer C ABI; no arithmetic happens here.
 */

#if defined(TH_REAL_IS_FLOAT)
#define VOLSTN_REAL_CODE VOLSTN_F32
#elif defined(TH_REAL_IS_DOUBLE)
#define VOLSTN_REAL_CODE VOLSTN_F64
#else
#error "libvolstn registers Float and Double only (stn3dtorch/init.c:8-9,18-19)"
#endif

/* a 5-D Torch tensor as an ABI descriptor (sizes and ELEMENT strides, like THTensor) */
static void nn_(volstn_desc5)(volstn_tensor5 *d, THTensor *t, const char *what) {
  int i;
  if (THTensor_(nDimension)(t) != 5) THError("%s: 5-D tensor expected, got %d-D", what, THTensor_(nDimension)(t));
  d->data = (void *)THTensor_(data)(t);
  for (i = 0; i < 5; i++) {
    d->size[i] = (int64_t)THTensor_(size)(t, i);
    d->stride[i] = (int64_t)THTensor_(stride)(t, i);
  }
}

/* ---- sampler: (self, inputImages, grids, output [, flags])  -- generic/BilinearSamplerVolumetric.c:444-447 ---- */
static int nn_(volstn_sampler_updateOutput)(lua_State *L, int G) {
  THTensor *inputImages = (THTensor *)luaT_checkudata(L, 2, torch_Tensor);
  THTensor *grids = (THTensor *)luaT_checkudata(L, 3, torch_Tensor);
  THTensor *output = (THTensor *)luaT_checkudata(L, 4, torch_Tensor);
  const unsigned flags = volstn_glue_flags(lua_tonumber(L, 5));
  volstn_tensor5 v, g, o;
  nn_(volstn_desc5)(&v, inputImages, "inputImages");
  nn_(volstn_desc5)(&g, grids, "grids");
  nn_(volstn_desc5)(&o, output, "output");
  VOLSTN_GLUE_CHECK("BilinearSampler.updateOutput",
                    volstn_b200_host_sampler_forward(volstn_glue_ctx(), VOLSTN_REAL_CODE, G, &v, &g, &o, flags));
  return 1;
}

/* ---- (self, inputImages, grids, gradInputImages, gradGrids, gradOutput [, flags])  -- ...c:587-592 ----
 * flags = 0 (what the reference's own Lua file passes): gradInputImages is ACCUMULATED into and gradGrids assigned,
 * like ...c:730-789, 819-822 -- the Lua side has zeroed both (BilinearSamplerVolumetric.lua:101-104). */
static int nn_(volstn_sampler_updateGradInput)(lua_State *L, int G) {
  THTensor *inputImages = (THTensor *)luaT_checkudata(L, 2, torch_Tensor);
  THTensor *grids = (THTensor *)luaT_checkudata(L, 3, torch_Tensor);
  THTensor *gradInputImages = (THTensor *)luaT_checkudata(L, 4, torch_Tensor);
  THTensor *gradGrids = (THTensor *)luaT_checkudata(L, 5, torch_Tensor);
  THTensor *gradOutput = (THTensor *)luaT_checkudata(L, 6, torch_Tensor);
  const unsigned flags = volstn_glue_flags(lua_tonumber(L, 7));
  volstn_tensor5 v, g, gv, gg, go;
  nn_(volstn_desc5)(&v, inputImages, "inputImages");
  nn_(volstn_desc5)(&g, grids, "grids");
  nn_(volstn_desc5)(&gv, gradInputImages, "gradInputImages");
  nn_(volstn_desc5)(&gg, gradGrids, "gradGrids");
  nn_(volstn_desc5)(&go, gradOutput, "gradOutput");
  VOLSTN_GLUE_CHECK("BilinearSampler.updateGradInput",
                    volstn_b200_host_sampler_backward(volstn_glue_ctx(), VOLSTN_REAL_CODE, G, &v, &g, &gv, &gg, &go, flags));
  return 1;
}

static int nn_(BilinearSamplerVolumetric_updateOutput)(lua_State *L) { return nn_(volstn_sampler_updateOutput)(L, 4); }
static int nn_(BilinearSamplerVolumetric_updateGradInput)(lua_State *L) { return nn_(volstn_sampler_updateGradInput)(L, 4); }
/* the three-component twins the reference registers under the BHWD names (...c:9-439, 834-835) */
static int nn_(BilinearSamplerBHWD_updateOutput)(lua_State *L) { return nn_(volstn_sampler_updateOutput)(L, 3); }
static int nn_(BilinearSamplerBHWD_updateGradInput)(lua_State *L) { return nn_(volstn_sampler_updateGradInput)(L, 3); }

/* ---- nn.Cumsum: (self, input, output) and (self, gradOutput, gradInput); contiguous B x N x d1 .. dN ----
 * New native boundary for Cumsum.lua:12-22 / 39-48 (pure Lua over tensor:cumsum / :index in the reference). */
static int nn_(volstn_cumsum)(lua_State *L, int reverse) {
  THTensor *src = (THTensor *)luaT_checkudata(L, 2, torch_Tensor);
  THTensor *dst = (THTensor *)luaT_checkudata(L, 3, torch_Tensor);
  int64_t dims[8];
  const int nd = THTensor_(nDimension)(src);
  int i, rc;
  if (nd < 3 || nd > 10 || THTensor_(size)(src, 1) != nd - 2) THError("nn.Cumsum: B x N x d1 x ... x dN expected");
  if (!THTensor_(isContiguous)(src) || !THTensor_(isContiguous)(dst) || THTensor_(nElement)(dst) != THTensor_(nElement)(src))
    THError("nn.Cumsum: contiguous tensors of equal size expected");
  for (i = 0; i < nd - 2; i++) dims[i] = (int64_t)THTensor_(size)(src, i + 2);
  rc = reverse ? volstn_b200_host_cumsum_backward(volstn_glue_ctx(), VOLSTN_REAL_CODE, THTensor_(data)(src), THTensor_(data)(dst),
                                                  (int64_t)THTensor_(size)(src, 0), nd - 2, dims)
               : volstn_b200_host_cumsum_forward(volstn_glue_ctx(), VOLSTN_REAL_CODE, THTensor_(data)(src), THTensor_(data)(dst),
                                                 (int64_t)THTensor_(size)(src, 0), nd - 2, dims);
  VOLSTN_GLUE_CHECK(reverse ? "Cumsum.updateGradInput" : "Cumsum.updateOutput", rc);
  return 1;
}
static int nn_(Cumsum_updateOutput)(lua_State *L) { return nn_(volstn_cumsum)(L, 0); }
static int nn_(Cumsum_updateGradInput)(lua_State *L) { return nn_(volstn_cumsum)(L, 1); }

/* ---- nn.VolumetricGridGenerator: (self, transformMatrix B x 4 x 4, output B x D x H x W x 4) and
 * (self, gradGrid B x D x H x W x 4, gradInput B x 4 x 4); contiguous.  New native boundary for
 * VolumetricGridGenerator.lua:58-60 / 79-82 (torch.bmm / baddbmm over a B-replicated base grid in the reference). */
static int nn_(VolumetricGridGenerator_updateOutput)(lua_State *L) {
  THTensor *T = (THTensor *)luaT_checkudata(L, 2, torch_Tensor);
  THTensor *out = (THTensor *)luaT_checkudata(L, 3, torch_Tensor);
  if (THTensor_(nDimension)(T) != 3 || THTensor_(size)(T, 1) != 4 || THTensor_(size)(T, 2) != 4 || THTensor_(nDimension)(out) != 5 ||
      THTensor_(size)(out, 4) != 4 || THTensor_(size)(out, 0) != THTensor_(size)(T, 0) || !THTensor_(isContiguous)(T) ||
      !THTensor_(isContiguous)(out))
    THError("please input affine transform matrices (bx4x4)"); /* VolumetricGridGenerator.lua:44-47 */
  VOLSTN_GLUE_CHECK("VolumetricGridGenerator.updateOutput",
                    volstn_b200_host_gridgen_forward(volstn_glue_ctx(), VOLSTN_REAL_CODE, THTensor_(data)(T), THTensor_(data)(out),
                                                     (int64_t)THTensor_(size)(out, 0), (int64_t)THTensor_(size)(out, 1),
                                                     (int64_t)THTensor_(size)(out, 2), (int64_t)THTensor_(size)(out, 3)));
  return 1;
}
static int nn_(VolumetricGridGenerator_updateGradInput)(lua_State *L) {
  THTensor *gg = (THTensor *)luaT_checkudata(L, 2, torch_Tensor);
  THTensor *gT = (THTensor *)luaT_checkudata(L, 3, torch_Tensor);
  if (THTensor_(nDimension)(gg) != 5 || THTensor_(size)(gg, 4) != 4 || THTensor_(nDimension)(gT) != 3 || THTensor_(size)(gT, 1) != 4 ||
      THTensor_(size)(gT, 2) != 4 || THTensor_(size)(gT, 0) != THTensor_(size)(gg, 0) || !THTensor_(isContiguous)(gg) ||
      !THTensor_(isContiguous)(gT))
    THError("VolumetricGridGenerator: gradGrid b x D x H x W x 4 and gradInput b x 4 x 4 (contiguous) expected");
  VOLSTN_GLUE_CHECK("VolumetricGridGenerator.updateGradInput",
                    volstn_b200_host_gridgen_backward(volstn_glue_ctx(), VOLSTN_REAL_CODE, THTensor_(data)(gg), THTensor_(data)(gT),
                                                      (int64_t)THTensor_(size)(gg, 0), (int64_t)THTensor_(size)(gg, 1),
                                                      (int64_t)THTensor_(size)(gg, 2), (int64_t)THTensor_(size)(gg, 3)));
  return 1;
}

/* the reference's four names first (generic/BilinearSamplerVolumetric.c:831-836), then the four new ones */
/* ---- nn.AffineSamplerVolumetric on host tensors (the loader-side augmentation of dataset_2d.lua:285-311 and the projector
 * of stn3dtorch/test.lua:15-25 run on CPU tensors): nn.VolumetricGridGenerator + nn.BilinearSamplerVolumetric in one call,
 * the B x D x H x W x 4 grid never exists.  Float only (a new module; the reference's Double registration covers its four
 * sampler functions).  (self, inputImages, T, output [, flags]) / (self, inputImages, T, gradInputImages, gradT, gradOutput
 * [, flags]); T and gradT are contiguous B x 4 x 4. ---- */
#if defined(TH_REAL_IS_FLOAT)
static void nn_(volstn_affine_T)(THTensor *T, THTensor *images, const char *what) {
  if (THTensor_(nDimension)(T) != 3 || THTensor_(size)(T, 1) != 4 || THTensor_(size)(T, 2) != 4 ||
      THTensor_(size)(T, 0) != THTensor_(size)(images, 0) || !THTensor_(isContiguous)(T))
    THError("%s: please input affine transform matrices (bx4x4)", what);
}
static int nn_(AffineSamplerVolumetric_updateOutput)(lua_State *L) {
  THTensor *inputImages = (THTensor *)luaT_checkudata(L, 2, torch_Tensor);
  THTensor *T = (THTensor *)luaT_checkudata(L, 3, torch_Tensor);
  THTensor *output = (THTensor *)luaT_checkudata(L, 4, torch_Tensor);
  const unsigned flags = volstn_glue_flags(lua_tonumber(L, 5));
  volstn_tensor5 v, o;
  nn_(volstn_desc5)(&v, inputImages, "inputImages");
  nn_(volstn_desc5)(&o, output, "output");
  nn_(volstn_affine_T)(T, inputImages, "AffineSamplerVolumetric");
  VOLSTN_GLUE_CHECK("AffineSamplerVolumetric.updateOutput",
                    volstn_b200_host_affine_sampler_forward(volstn_glue_ctx(), VOLSTN_REAL_CODE, THTensor_(data)(T), &v, &o, flags));
  return 1;
}
static int nn_(AffineSamplerVolumetric_updateGradInput)(lua_State *L) {
  THTensor *inputImages = (THTensor *)luaT_checkudata(L, 2, torch_Tensor);
  THTensor *T = (THTensor *)luaT_checkudata(L, 3, torch_Tensor);
  THTensor *gradInputImages = (THTensor *)luaT_checkudata(L, 4, torch_Tensor);
  THTensor *gradT = (THTensor *)luaT_checkudata(L, 5, torch_Tensor);
  THTensor *gradOutput = (THTensor *)luaT_checkudata(L, 6, torch_Tensor);
  const unsigned flags = volstn_glue_flags(lua_tonumber(L, 7));
  volstn_tensor5 v, gv, go;
  nn_(volstn_desc5)(&v, inputImages, "inputImages");
  nn_(volstn_desc5)(&go, gradOutput, "gradOutput");
  nn_(volstn_affine_T)(T, inputImages, "AffineSamplerVolumetric");
  nn_(volstn_affine_T)(gradT, inputImages, "AffineSamplerVolumetric (gradT)");
  if (!(flags & VOLSTN_BWD_ONLY_GRID)) nn_(volstn_desc5)(&gv, gradInputImages, "gradInputImages");
  VOLSTN_GLUE_CHECK("AffineSamplerVolumetric.updateGradInput",
                    volstn_b200_host_affine_sampler_backward(volstn_glue_ctx(), VOLSTN_REAL_CODE, THTensor_(data)(T), &v,
                                                             (flags & VOLSTN_BWD_ONLY_GRID) ? NULL : &gv, THTensor_(data)(gradT), &go,
                                                             flags));
  return 1;
}
#endif

static const struct luaL_Reg nn_(BilinearSamplerVolumetric__)[] = {
    {"BilinearSamplerVolumetric_updateOutput", nn_(BilinearSamplerVolumetric_updateOutput)},
    {"BilinearSamplerVolumetric_updateGradInput", nn_(BilinearSamplerVolumetric_updateGradInput)},
    {"BilinearSamplerBHWD_updateOutput", nn_(BilinearSamplerBHWD_updateOutput)},
    {"BilinearSamplerBHWD_updateGradInput", nn_(BilinearSamplerBHWD_updateGradInput)},
    {"Cumsum_updateOutput", nn_(Cumsum_updateOutput)},
    {"Cumsum_updateGradInput", nn_(Cumsum_updateGradInput)},
    {"VolumetricGridGenerator_updateOutput", nn_(VolumetricGridGenerator_updateOutput)},
    {"VolumetricGridGenerator_updateGradInput", nn_(VolumetricGridGenerator_updateGradInput)},
#if defined(TH_REAL_IS_FLOAT)
    {"AffineSamplerVolumetric_updateOutput", nn_(AffineSamplerVolumetric_updateOutput)},
    {"AffineSamplerVolumetric_updateGradInput", nn_(AffineSamplerVolumetric_updateGradInput)},
#endif
    {NULL, NULL}};

static void nn_(BilinearSamplerVolumetric_init)(lua_State *L) { /* ...c:838-843 */
  luaT_pushmetatable(L, torch_Tensor);
  luaT_registeratname(L, nn_(BilinearSamplerVolumetric__), "nn");
  lua_pop(L, 1);
}

#undef VOLSTN_REAL_CODE
#endif
