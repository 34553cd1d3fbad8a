/*
 * lua/init.cu -- drop-in replacement of stn3dtorch/init.cu + BilinearSamplerVolumetric.cu + utils.c: the
 * torch.CudaTensor library `libcuvolstn` (`require 'libcuvolstn'`, stn3dtorch/init.lua:6).  Same entry point
 * (luaopen_libcuvolstn, init.cu:10-15) and the same function names in the `nn` table of torch.CudaTensor
 * (BilinearSamplerVolumetric.cu:1074-1088).  The launchers of the reference (…cu:695-742, 1009-1072, 179-228,
 * 511-575) become one call each into the device-pointer C ABI on THC's current stream (…cu:706): asynchronous, no
 * allocation, no synchronisation (except the fused training step, which returns its loss), nothing retained.  This file
 * contains no device code; it is a .cu only because the
 * reference's CMakeLists.txt builds `init.cu` with CUDA_ADD_LIBRARY.
 */
#include <cuda_runtime.h>

#include "luaT.h"
#include "THC.h"

#include "volstn_glue.h"

/* cutorch keeps its THCState in Lua: cutorch.getState() (what stn3dtorch/utils.c:3-11 does) */
static THCState *volstn_cutorch_state(lua_State *L) {
  THCState *state;
  lua_getglobal(L, "cutorch");
  lua_getfield(L, -1, "getState");
  lua_call(L, 0, 1);
  state = (THCState *)lua_touserdata(L, -1);
  lua_pop(L, 2);
  return state;
}

static void cunn_volstn_desc5(THCState *state, volstn_tensor5 *d, THCudaTensor *t, const char *what) {
  int i;
  if (THCudaTensor_nDimension(state, t) != 5) THError("%s: 5-D tensor expected, got %d-D", what, THCudaTensor_nDimension(state, t));
  d->data = (void *)THCudaTensor_data(state, t);
  for (i = 0; i < 5; i++) {
    d->size[i] = (int64_t)THCudaTensor_size(state, t, i);
    d->stride[i] = (int64_t)THCudaTensor_stride(state, t, i);
  }
}

/* (self, inputImages, grids, output [, flags])  -- BilinearSamplerVolumetric.cu:695-742 */
static int cunn_volstn_sampler_updateOutput(lua_State *L, int G) {
  THCState *state = volstn_cutorch_state(L);
  THCudaTensor *inputImages = (THCudaTensor *)luaT_checkudata(L, 2, "torch.CudaTensor");
  THCudaTensor *grids = (THCudaTensor *)luaT_checkudata(L, 3, "torch.CudaTensor");
  THCudaTensor *output = (THCudaTensor *)luaT_checkudata(L, 4, "torch.CudaTensor");
  const unsigned flags = volstn_glue_flags(lua_tonumber(L, 5));
  volstn_tensor5 v, g, o;
  cunn_volstn_desc5(state, &v, inputImages, "inputImages");
  cunn_volstn_desc5(state, &g, grids, "grids");
  cunn_volstn_desc5(state, &o, output, "output");
  VOLSTN_GLUE_CHECK("BilinearSampler.updateOutput",
                    volstn_b200_sampler_forward(VOLSTN_F32, G, &v, &g, &o, flags, (void *)THCState_getCurrentStream(state)));
  return 1;
}

/* (self, inputImages, grids, gradInputImages, gradGrids, gradOutput [, flags])  -- …cu:1009-1072 */
static int cunn_volstn_sampler_updateGradInput(lua_State *L, int G) {
  THCState *state = volstn_cutorch_state(L);
  THCudaTensor *inputImages = (THCudaTensor *)luaT_checkudata(L, 2, "torch.CudaTensor");
  THCudaTensor *grids = (THCudaTensor *)luaT_checkudata(L, 3, "torch.CudaTensor");
  THCudaTensor *gradInputImages = (THCudaTensor *)luaT_checkudata(L, 4, "torch.CudaTensor");
  THCudaTensor *gradGrids = (THCudaTensor *)luaT_checkudata(L, 5, "torch.CudaTensor");
  THCudaTensor *gradOutput = (THCudaTensor *)luaT_checkudata(L, 6, "torch.CudaTensor");
  const unsigned flags = volstn_glue_flags(lua_tonumber(L, 7));
  volstn_tensor5 v, g, gv, gg, go;
  cunn_volstn_desc5(state, &v, inputImages, "inputImages");
  cunn_volstn_desc5(state, &g, grids, "grids");
  cunn_volstn_desc5(state, &gv, gradInputImages, "gradInputImages");
  cunn_volstn_desc5(state, &gg, gradGrids, "gradGrids");
  cunn_volstn_desc5(state, &go, gradOutput, "gradOutput");
  VOLSTN_GLUE_CHECK("BilinearSampler.updateGradInput",
                    volstn_b200_sampler_backward(VOLSTN_F32, G, &v, &g, &gv, &gg, &go, flags,
                                                 (void *)THCState_getCurrentStream(state)));
  return 1;
}

static int cunn_BilinearSamplerVolumetric_updateOutput(lua_State *L) { return cunn_volstn_sampler_updateOutput(L, 4); }
static int cunn_BilinearSamplerVolumetric_updateGradInput(lua_State *L) { return cunn_volstn_sampler_updateGradInput(L, 4); }
static int cunn_BilinearSamplerBHWD_updateOutput(lua_State *L) { return cunn_volstn_sampler_updateOutput(L, 3); }
static int cunn_BilinearSamplerBHWD_updateGradInput(lua_State *L) { return cunn_volstn_sampler_updateGradInput(L, 3); }

/* nn.Cumsum on CudaTensors: (self, input, output) / (self, gradOutput, gradInput) */
static int cunn_volstn_cumsum(lua_State *L, int reverse) {
  THCState *state = volstn_cutorch_state(L);
  THCudaTensor *src = (THCudaTensor *)luaT_checkudata(L, 2, "torch.CudaTensor");
  THCudaTensor *dst = (THCudaTensor *)luaT_checkudata(L, 3, "torch.CudaTensor");
  int64_t dims[8];
  const int nd = THCudaTensor_nDimension(state, src);
  int i, rc;
  if (nd < 3 || nd > 10 || THCudaTensor_size(state, src, 1) != nd - 2) THError("nn.Cumsum: B x N x d1 x ... x dN expected");
  if (!THCudaTensor_isContiguous(state, src) || !THCudaTensor_isContiguous(state, dst) ||
      THCudaTensor_nElement(state, dst) != THCudaTensor_nElement(state, src))
    THError("nn.Cumsum: contiguous tensors of equal size expected");
  for (i = 0; i < nd - 2; i++) dims[i] = (int64_t)THCudaTensor_size(state, src, i + 2);
  rc = reverse ? volstn_b200_cumsum_backward(VOLSTN_F32, THCudaTensor_data(state, src), THCudaTensor_data(state, dst),
                                             (int64_t)THCudaTensor_size(state, src, 0), nd - 2, dims,
                                             (void *)THCState_getCurrentStream(state))
               : volstn_b200_cumsum_forward(VOLSTN_F32, THCudaTensor_data(state, src), THCudaTensor_data(state, dst),
                                            (int64_t)THCudaTensor_size(state, src, 0), nd - 2, dims,
                                            (void *)THCState_getCurrentStream(state));
  VOLSTN_GLUE_CHECK(reverse ? "Cumsum.updateGradInput" : "Cumsum.updateOutput", rc);
  return 1;
}
static int cunn_Cumsum_updateOutput(lua_State *L) { return cunn_volstn_cumsum(L, 0); }
static int cunn_Cumsum_updateGradInput(lua_State *L) { return cunn_volstn_cumsum(L, 1); }

/* nn.VolumetricGridGenerator on CudaTensors: (self, T, output) / (self, gradGrid, gradInput, workspace).
 * `workspace` is a module-owned CudaTensor (the Torch7 idiom for scratch memory); it is resized here. */
static int cunn_VolumetricGridGenerator_updateOutput(lua_State *L) {
  THCState *state = volstn_cutorch_state(L);
  THCudaTensor *T = (THCudaTensor *)luaT_checkudata(L, 2, "torch.CudaTensor");
  THCudaTensor *out = (THCudaTensor *)luaT_checkudata(L, 3, "torch.CudaTensor");
  if (THCudaTensor_nDimension(state, T) != 3 || THCudaTensor_size(state, T, 1) != 4 || THCudaTensor_size(state, T, 2) != 4 ||
      THCudaTensor_nDimension(state, out) != 5 || THCudaTensor_size(state, out, 4) != 4 ||
      THCudaTensor_size(state, out, 0) != THCudaTensor_size(state, T, 0) || !THCudaTensor_isContiguous(state, T) ||
      !THCudaTensor_isContiguous(state, out))
    THError("please input affine transform matrices (bx4x4)");
  VOLSTN_GLUE_CHECK("VolumetricGridGenerator.updateOutput",
                    volstn_b200_gridgen_forward(VOLSTN_F32, THCudaTensor_data(state, T), THCudaTensor_data(state, out),
                                                (int64_t)THCudaTensor_size(state, out, 0), (int64_t)THCudaTensor_size(state, out, 1),
                                                (int64_t)THCudaTensor_size(state, out, 2), (int64_t)THCudaTensor_size(state, out, 3),
                                                (void *)THCState_getCurrentStream(state)));
  return 1;
}
static int cunn_VolumetricGridGenerator_updateGradInput(lua_State *L) {
  THCState *state = volstn_cutorch_state(L);
  THCudaTensor *gg = (THCudaTensor *)luaT_checkudata(L, 2, "torch.CudaTensor");
  THCudaTensor *gT = (THCudaTensor *)luaT_checkudata(L, 3, "torch.CudaTensor");
  THCudaTensor *ws = (THCudaTensor *)luaT_checkudata(L, 4, "torch.CudaTensor");
  int64_t B, D, H, W;
  size_t bytes;
  if (THCudaTensor_nDimension(state, gg) != 5 || THCudaTensor_size(state, gg, 4) != 4 || THCudaTensor_nDimension(state, gT) != 3 ||
      THCudaTensor_size(state, gT, 1) != 4 || THCudaTensor_size(state, gT, 2) != 4 ||
      THCudaTensor_size(state, gT, 0) != THCudaTensor_size(state, gg, 0) || !THCudaTensor_isContiguous(state, gg) ||
      !THCudaTensor_isContiguous(state, gT))
    THError("VolumetricGridGenerator: gradGrid b x D x H x W x 4 and gradInput b x 4 x 4 (contiguous) expected");
  B = (int64_t)THCudaTensor_size(state, gg, 0), D = (int64_t)THCudaTensor_size(state, gg, 1);
  H = (int64_t)THCudaTensor_size(state, gg, 2), W = (int64_t)THCudaTensor_size(state, gg, 3);
  bytes = volstn_b200_gridgen_backward_workspace(B, D, H, W);
  THCudaTensor_resize1d(state, ws, (long)((bytes + 3) / 4 + 2)); /* float elements; +2 for 8-byte alignment slack */
  {
    char *p = (char *)THCudaTensor_data(state, ws);
    char *aligned = (char *)(((size_t)p + 7) & ~(size_t)7);
    VOLSTN_GLUE_CHECK("VolumetricGridGenerator.updateGradInput",
                      volstn_b200_gridgen_backward(VOLSTN_F32, THCudaTensor_data(state, gg), THCudaTensor_data(state, gT), B, D, H, W,
                                                   aligned, bytes, (void *)THCState_getCurrentStream(state)));
  }
  return 1;
}

/* ---------------------------------------------------------------------------------------------------------
 * The fused modules of ALIGNet's warp head (no single reference function: the sub-graph models/alignet_2d.lua:136-147 +
 * models/alignet.lua:4-39 lifted to 3-D, and the step of train.lua:123-126 with model.lua:34 and util.lua:110-120).
 * ------------------------------------------------------------------------------------------------------- */
static void cunn_volstn_cage(lua_State *L, THCState *state, THCudaTensor *cage, THCudaTensor *source, int64_t dims[3]) {
  int i;
  (void)L;
  if (THCudaTensor_nDimension(state, cage) != 5 || THCudaTensor_size(state, cage, 1) != 3 ||
      THCudaTensor_size(state, cage, 0) != THCudaTensor_size(state, source, 0) || !THCudaTensor_isContiguous(state, cage))
    THError("nn.CageWarp: cage must be a contiguous B x 3 x cD x cH x cW tensor");
  for (i = 0; i < 3; i++) dims[i] = (int64_t)THCudaTensor_size(state, cage, i + 2);
}

/* nn.CageWarpVolumetric: (self, source, cage, output [, flags]) */
static int cunn_CageWarpVolumetric_updateOutput(lua_State *L) {
  THCState *state = volstn_cutorch_state(L);
  THCudaTensor *source = (THCudaTensor *)luaT_checkudata(L, 2, "torch.CudaTensor");
  THCudaTensor *cage = (THCudaTensor *)luaT_checkudata(L, 3, "torch.CudaTensor");
  THCudaTensor *output = (THCudaTensor *)luaT_checkudata(L, 4, "torch.CudaTensor");
  const unsigned flags = volstn_glue_flags(lua_tonumber(L, 5));
  volstn_tensor5 v, o;
  int64_t dims[3];
  cunn_volstn_desc5(state, &v, source, "source");
  cunn_volstn_desc5(state, &o, output, "output");
  cunn_volstn_cage(L, state, cage, source, dims);
  VOLSTN_GLUE_CHECK("CageWarpVolumetric.updateOutput",
                    volstn_b200_cage_warp_forward(VOLSTN_F32, THCudaTensor_data(state, cage), dims, &v, &o, flags,
                                                  (void *)THCState_getCurrentStream(state)));
  return 1;
}

/* (self, source, cage, gradSource, gradCage, gradOutput [, flags]): gradCage is assigned; gradSource is accumulated into
 * (VOLSTN_BWD_ZERO_GRADVOL zero-fills it first) and ignored with VOLSTN_BWD_ONLY_GRID */
static int cunn_CageWarpVolumetric_updateGradInput(lua_State *L) {
  THCState *state = volstn_cutorch_state(L);
  THCudaTensor *source = (THCudaTensor *)luaT_checkudata(L, 2, "torch.CudaTensor");
  THCudaTensor *cage = (THCudaTensor *)luaT_checkudata(L, 3, "torch.CudaTensor");
  THCudaTensor *gradSource = (THCudaTensor *)luaT_checkudata(L, 4, "torch.CudaTensor");
  THCudaTensor *gradCage = (THCudaTensor *)luaT_checkudata(L, 5, "torch.CudaTensor");
  THCudaTensor *gradOutput = (THCudaTensor *)luaT_checkudata(L, 6, "torch.CudaTensor");
  const unsigned flags = volstn_glue_flags(lua_tonumber(L, 7));
  volstn_tensor5 v, gv, go;
  int64_t dims[3];
  cunn_volstn_desc5(state, &v, source, "source");
  cunn_volstn_desc5(state, &go, gradOutput, "gradOutput");
  cunn_volstn_cage(L, state, cage, source, dims);
  if (!THCudaTensor_isContiguous(state, gradCage) || THCudaTensor_nElement(state, gradCage) != THCudaTensor_nElement(state, cage))
    THError("nn.CageWarp: gradCage must be contiguous and of the cage's size");
  if (!(flags & VOLSTN_BWD_ONLY_GRID)) cunn_volstn_desc5(state, &gv, gradSource, "gradSource");
  VOLSTN_GLUE_CHECK("CageWarpVolumetric.updateGradInput",
                    volstn_b200_cage_warp_backward(VOLSTN_F32, THCudaTensor_data(state, cage), dims, &v,
                                                   (flags & VOLSTN_BWD_ONLY_GRID) ? NULL : &gv, THCudaTensor_data(state, gradCage), &go,
                                                   flags, (void *)THCState_getCurrentStream(state)));
  return 1;
}

/* nn.CageWarpSmoothL1Volumetric: (self, source, cage, target, output, gradSource, gradCage, workspace, stats, gradScale [, flags]).
 * One kernel: forward, SmoothL1Criterion and its gradient (gradScale = 1 / nElement with sizeAverage, THNN), the
 * intersection / union counts at 0.5 and the backward.  `output` may be an EMPTY tensor (the forward value is then not
 * stored).  `workspace` is a module-owned CudaTensor of at least 6 B + 2 floats (device accumulators: B doubles + 2 B 64-bit
 * counts); `stats` is a module-owned HOST torch.DoubleTensor B x 3 that receives (SmoothL1 sum, intersection, union) per
 * sample.  Like criterion:forward in train.lua:123, which returns a Lua number, the call synchronises with its stream. */
static int cunn_CageWarpSmoothL1Volumetric_forwardBackward(lua_State *L) {
  THCState *state = volstn_cutorch_state(L);
  THCudaTensor *source = (THCudaTensor *)luaT_checkudata(L, 2, "torch.CudaTensor");
  THCudaTensor *cage = (THCudaTensor *)luaT_checkudata(L, 3, "torch.CudaTensor");
  THCudaTensor *target = (THCudaTensor *)luaT_checkudata(L, 4, "torch.CudaTensor");
  THCudaTensor *output = (THCudaTensor *)luaT_checkudata(L, 5, "torch.CudaTensor");
  THCudaTensor *gradSource = (THCudaTensor *)luaT_checkudata(L, 6, "torch.CudaTensor");
  THCudaTensor *gradCage = (THCudaTensor *)luaT_checkudata(L, 7, "torch.CudaTensor");
  THCudaTensor *workspace = (THCudaTensor *)luaT_checkudata(L, 8, "torch.CudaTensor");
  THDoubleTensor *stats = (THDoubleTensor *)luaT_checkudata(L, 9, "torch.DoubleTensor");
  const float grad_scale = (float)lua_tonumber(L, 10);
  const unsigned flags = volstn_glue_flags(lua_tonumber(L, 11));
  cudaStream_t st = (cudaStream_t)THCState_getCurrentStream(state);
  volstn_tensor5 v, t, o, gv;
  int64_t dims[3];
  long B, b;
  const int keep = THCudaTensor_nElement(state, output) > 0;
  char *ws;
  double *loss, *host;
  uint64_t *iou;
  cunn_volstn_desc5(state, &v, source, "source");
  cunn_volstn_desc5(state, &t, target, "target");
  cunn_volstn_cage(L, state, cage, source, dims);
  B = THCudaTensor_size(state, source, 0);
  if (keep) cunn_volstn_desc5(state, &o, output, "output");
  if (!(flags & VOLSTN_BWD_ONLY_GRID)) cunn_volstn_desc5(state, &gv, gradSource, "gradSource");
  if (!THCudaTensor_isContiguous(state, gradCage) || THCudaTensor_nElement(state, gradCage) != THCudaTensor_nElement(state, cage))
    THError("nn.CageWarp: gradCage must be contiguous and of the cage's size");
  if (THDoubleTensor_nDimension(stats) != 2 || THDoubleTensor_size(stats, 0) != B || THDoubleTensor_size(stats, 1) != 3 ||
      !THDoubleTensor_isContiguous(stats))
    THError("nn.CageWarpSmoothL1Volumetric: stats must be a contiguous B x 3 torch.DoubleTensor");
  THCudaTensor_resize1d(state, workspace, 6 * B + 2);
  ws = (char *)(((size_t)THCudaTensor_data(state, workspace) + 7) & ~(size_t)7);
  loss = (double *)ws;
  iou = (uint64_t *)(ws + 8 * (size_t)B);
  VOLSTN_GLUE_CHECK("CageWarpSmoothL1Volumetric.forwardBackward",
                    volstn_b200_cage_warp_step(VOLSTN_F32, THCudaTensor_data(state, cage), dims, &v, &t, keep ? &o : NULL,
                                               (flags & VOLSTN_BWD_ONLY_GRID) ? NULL : &gv, THCudaTensor_data(state, gradCage), loss,
                                               iou, grad_scale, flags, (void *)st));
  host = (double *)malloc(24 * (size_t)(B > 0 ? B : 1));
  if (!host) THError("nn.CageWarpSmoothL1Volumetric: out of memory");
  if (cudaMemcpyAsync(host, ws, 24 * (size_t)B, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
    free(host);
    THError("nn.CageWarpSmoothL1Volumetric: reading the loss back failed");
  }
  for (b = 0; b < B; b++) {
    double *row = THDoubleTensor_data(stats) + 3 * b;
    const uint64_t *cnt = (const uint64_t *)(host + B) + 2 * b;
    row[0] = host[b];
    row[1] = (double)cnt[0];
    row[2] = (double)cnt[1];
  }
  free(host);
  return 1;
}

/* nn.BilinearSamplerSmoothL1Volumetric: (self, inputImages, grids, target, output, gradInputImages, gradGrids, workspace, stats,
 * gradScale [, flags]) -- the training step around a plain sampler (a network that emits a dense field): forward, SmoothL1 and
 * its gradient, IoU counts and the backward in one kernel; gradGrids has the bits of the module chain.  Buffers and the
 * synchronisation as for nn.CageWarpSmoothL1Volumetric above. */
static int cunn_BilinearSamplerSmoothL1Volumetric_forwardBackward(lua_State *L) {
  THCState *state = volstn_cutorch_state(L);
  THCudaTensor *inputImages = (THCudaTensor *)luaT_checkudata(L, 2, "torch.CudaTensor");
  THCudaTensor *grids = (THCudaTensor *)luaT_checkudata(L, 3, "torch.CudaTensor");
  THCudaTensor *target = (THCudaTensor *)luaT_checkudata(L, 4, "torch.CudaTensor");
  THCudaTensor *output = (THCudaTensor *)luaT_checkudata(L, 5, "torch.CudaTensor");
  THCudaTensor *gradInputImages = (THCudaTensor *)luaT_checkudata(L, 6, "torch.CudaTensor");
  THCudaTensor *gradGrids = (THCudaTensor *)luaT_checkudata(L, 7, "torch.CudaTensor");
  THCudaTensor *workspace = (THCudaTensor *)luaT_checkudata(L, 8, "torch.CudaTensor");
  THDoubleTensor *stats = (THDoubleTensor *)luaT_checkudata(L, 9, "torch.DoubleTensor");
  const float grad_scale = (float)lua_tonumber(L, 10);
  const unsigned flags = volstn_glue_flags(lua_tonumber(L, 11));
  cudaStream_t st = (cudaStream_t)THCState_getCurrentStream(state);
  volstn_tensor5 v, g, t, o, gv, gg;
  long B, b;
  const int keep = THCudaTensor_nElement(state, output) > 0;
  char *ws;
  double *loss, *host;
  uint64_t *iou;
  cunn_volstn_desc5(state, &v, inputImages, "inputImages");
  cunn_volstn_desc5(state, &g, grids, "grids");
  cunn_volstn_desc5(state, &t, target, "target");
  cunn_volstn_desc5(state, &gg, gradGrids, "gradGrids");
  B = THCudaTensor_size(state, inputImages, 0);
  if (keep) cunn_volstn_desc5(state, &o, output, "output");
  if (!(flags & VOLSTN_BWD_ONLY_GRID)) cunn_volstn_desc5(state, &gv, gradInputImages, "gradInputImages");
  if (THDoubleTensor_nDimension(stats) != 2 || THDoubleTensor_size(stats, 0) != B || THDoubleTensor_size(stats, 1) != 3 ||
      !THDoubleTensor_isContiguous(stats))
    THError("nn.BilinearSamplerSmoothL1Volumetric: stats must be a contiguous B x 3 torch.DoubleTensor");
  THCudaTensor_resize1d(state, workspace, 6 * B + 2);
  ws = (char *)(((size_t)THCudaTensor_data(state, workspace) + 7) & ~(size_t)7);
  loss = (double *)ws;
  iou = (uint64_t *)(ws + 8 * (size_t)B);
  VOLSTN_GLUE_CHECK("BilinearSamplerSmoothL1Volumetric.forwardBackward",
                    volstn_b200_sampler_step(VOLSTN_F32, (int)g.size[4], &v, &g, &t, keep ? &o : NULL,
                                             (flags & VOLSTN_BWD_ONLY_GRID) ? NULL : &gv, &gg, loss, iou, grad_scale, flags, (void *)st));
  host = (double *)malloc(24 * (size_t)(B > 0 ? B : 1));
  if (!host) THError("nn.BilinearSamplerSmoothL1Volumetric: out of memory");
  if (cudaMemcpyAsync(host, ws, 24 * (size_t)B, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
    free(host);
    THError("nn.BilinearSamplerSmoothL1Volumetric: reading the loss back failed");
  }
  for (b = 0; b < B; b++) {
    double *row = THDoubleTensor_data(stats) + 3 * b;
    const uint64_t *cnt = (const uint64_t *)(host + B) + 2 * b;
    row[0] = host[b];
    row[1] = (double)cnt[0];
    row[2] = (double)cnt[1];
  }
  free(host);
  return 1;
}

/* nn.AffineSamplerVolumetric on CudaTensors: (self, inputImages, T, output [, flags]) /
 * (self, inputImages, T, gradInputImages, gradT, gradOutput, workspace [, flags]); `workspace` is a module-owned CudaTensor
 * that is resized here (partial sums of the deterministic gradT reduction). */
static void cunn_volstn_affine_T(THCState *state, THCudaTensor *T, THCudaTensor *images) {
  if (THCudaTensor_nDimension(state, T) != 3 || THCudaTensor_size(state, T, 1) != 4 || THCudaTensor_size(state, T, 2) != 4 ||
      THCudaTensor_size(state, T, 0) != THCudaTensor_size(state, images, 0) || !THCudaTensor_isContiguous(state, T))
    THError("please input affine transform matrices (bx4x4)");
}
static int cunn_AffineSamplerVolumetric_updateOutput(lua_State *L) {
  THCState *state = volstn_cutorch_state(L);
  THCudaTensor *inputImages = (THCudaTensor *)luaT_checkudata(L, 2, "torch.CudaTensor");
  THCudaTensor *T = (THCudaTensor *)luaT_checkudata(L, 3, "torch.CudaTensor");
  THCudaTensor *output = (THCudaTensor *)luaT_checkudata(L, 4, "torch.CudaTensor");
  const unsigned flags = volstn_glue_flags(lua_tonumber(L, 5));
  volstn_tensor5 v, o;
  cunn_volstn_desc5(state, &v, inputImages, "inputImages");
  cunn_volstn_desc5(state, &o, output, "output");
  cunn_volstn_affine_T(state, T, inputImages);
  VOLSTN_GLUE_CHECK("AffineSamplerVolumetric.updateOutput",
                    volstn_b200_affine_sampler_forward(VOLSTN_F32, THCudaTensor_data(state, T), &v, &o, flags,
                                                       (void *)THCState_getCurrentStream(state)));
  return 1;
}
static int cunn_AffineSamplerVolumetric_updateGradInput(lua_State *L) {
  THCState *state = volstn_cutorch_state(L);
  THCudaTensor *inputImages = (THCudaTensor *)luaT_checkudata(L, 2, "torch.CudaTensor");
  THCudaTensor *T = (THCudaTensor *)luaT_checkudata(L, 3, "torch.CudaTensor");
  THCudaTensor *gradInputImages = (THCudaTensor *)luaT_checkudata(L, 4, "torch.CudaTensor");
  THCudaTensor *gradT = (THCudaTensor *)luaT_checkudata(L, 5, "torch.CudaTensor");
  THCudaTensor *gradOutput = (THCudaTensor *)luaT_checkudata(L, 6, "torch.CudaTensor");
  THCudaTensor *ws = (THCudaTensor *)luaT_checkudata(L, 7, "torch.CudaTensor");
  const unsigned flags = volstn_glue_flags(lua_tonumber(L, 8));
  volstn_tensor5 v, gv, go;
  size_t bytes;
  char *aligned;
  cunn_volstn_desc5(state, &v, inputImages, "inputImages");
  cunn_volstn_desc5(state, &go, gradOutput, "gradOutput");
  cunn_volstn_affine_T(state, T, inputImages);
  cunn_volstn_affine_T(state, gradT, inputImages);
  if (!(flags & VOLSTN_BWD_ONLY_GRID)) cunn_volstn_desc5(state, &gv, gradInputImages, "gradInputImages");
  bytes = volstn_b200_affine_sampler_backward_workspace(go.size[0], go.size[1], go.size[2], go.size[3]);
  THCudaTensor_resize1d(state, ws, (long)((bytes + 3) / 4 + 2));
  aligned = (char *)(((size_t)THCudaTensor_data(state, ws) + 7) & ~(size_t)7);
  VOLSTN_GLUE_CHECK("AffineSamplerVolumetric.updateGradInput",
                    volstn_b200_affine_sampler_backward(VOLSTN_F32, THCudaTensor_data(state, T), &v,
                                                        (flags & VOLSTN_BWD_ONLY_GRID) ? NULL : &gv, THCudaTensor_data(state, gradT),
                                                        &go, aligned, bytes, flags, (void *)THCState_getCurrentStream(state)));
  return 1;
}

static const struct luaL_Reg cunn_BilinearSamplerVolumetric__[] = {
    {"BilinearSamplerVolumetric_updateOutput", cunn_BilinearSamplerVolumetric_updateOutput},
    {"BilinearSamplerVolumetric_updateGradInput", cunn_BilinearSamplerVolumetric_updateGradInput},
    {"BilinearSamplerBHWD_updateOutput", cunn_BilinearSamplerBHWD_updateOutput},
    {"BilinearSamplerBHWD_updateGradInput", cunn_BilinearSamplerBHWD_updateGradInput},
    {"Cumsum_updateOutput", cunn_Cumsum_updateOutput},
    {"Cumsum_updateGradInput", cunn_Cumsum_updateGradInput},
    {"VolumetricGridGenerator_updateOutput", cunn_VolumetricGridGenerator_updateOutput},
    {"VolumetricGridGenerator_updateGradInput", cunn_VolumetricGridGenerator_updateGradInput},
    {"CageWarpVolumetric_updateOutput", cunn_CageWarpVolumetric_updateOutput},
    {"CageWarpVolumetric_updateGradInput", cunn_CageWarpVolumetric_updateGradInput},
    {"CageWarpSmoothL1Volumetric_forwardBackward", cunn_CageWarpSmoothL1Volumetric_forwardBackward},
    {"BilinearSamplerSmoothL1Volumetric_forwardBackward", cunn_BilinearSamplerSmoothL1Volumetric_forwardBackward},
    {"AffineSamplerVolumetric_updateOutput", cunn_AffineSamplerVolumetric_updateOutput},
    {"AffineSamplerVolumetric_updateGradInput", cunn_AffineSamplerVolumetric_updateGradInput},
    {NULL, NULL}};

static void cunn_BilinearSamplerVolumetric_init(lua_State *L) { /* …cu:1083-1088 */
  luaT_pushmetatable(L, "torch.CudaTensor");
  luaT_registeratname(L, cunn_BilinearSamplerVolumetric__, "nn");
  lua_pop(L, 1);
}

LUA_EXTERNC DLL_EXPORT int luaopen_libcuvolstn(lua_State *L);

int luaopen_libcuvolstn(lua_State *L) {
  lua_newtable(L);
  cunn_BilinearSamplerVolumetric_init(L);
  return 1;
}
