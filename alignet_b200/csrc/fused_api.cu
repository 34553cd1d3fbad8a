// fused_api.cu -- host side of the ROW kernels (rowwarp.cu): plans from tensor descriptors and the C ABI of
// the fused modules.
//
//   strided / planar nn.BilinearSamplerVolumetric   row_sampler_forward / row_sampler_backward (called from the
//                                                   dispatchers of sampler.cu when the packed kernels do not apply)
//   volstn_b200_affine_sampler_*                    nn.VolumetricGridGenerator + nn.BilinearSamplerVolumetric in one
//                                                   pass (the `projector` of stn3dtorch/test.lua:15-25)
//   volstn_b200_cage_warp_*                         cage up-sample + source warp in one pass
//                                                   (models/alignet_2d.lua:136-147 lifted to 3-D)
#include <cstring>

#include "cagebrick.cuh"
#include "rowwarp.cuh"
#include "volstn_internal.h"

namespace volstn {
namespace {

bool fits31(int64_t v) { return v >= 0 && v < 0x7fffffffLL; }

// largest in-sample element offset of a descriptor (dims 1..4); -1 if a stride is negative
int64_t sample_extent(const volstn_tensor5 *t) {
  int64_t e = 0;
  for (int i = 1; i < 5; i++) {
    if (t->stride[i] < 0) return -1;
    if (t->size[i] > 0) e += (t->size[i] - 1) * t->stride[i];
  }
  return e;
}

bool row_tensor_ok(const volstn_tensor5 *t) {
  if (!t || !t->data) return false;
  for (int i = 0; i < 5; i++)
    if (!fits31(t->size[i])) return false;
  if (t->stride[0] < 0) return false;
  const int64_t e = sample_extent(t);
  return e >= 0 && e < 0x7fffff00LL;
}

void zero_args(RowArgs &a) { memset(&a, 0, sizeof(a)); }

// volume (gather source), out-like tensor (out or gradOut) and, for the backward, gradVol
bool plan_volume(RowArgs &a, RowPlan &p, const volstn_tensor5 *vol, const volstn_tensor5 *outlike,
                 const volstn_tensor5 *gradVol) {
  if (!row_tensor_ok(vol) || !row_tensor_ok(outlike)) return false;
  if (vol->size[1] < 1 || vol->size[2] < 1 || vol->size[3] < 1) return false;
  if (vol->size[1] > (1 << 22) || vol->size[2] > (1 << 22) || vol->size[3] > (1 << 22)) return false;
  if (gradVol) {
    if (!row_tensor_ok(gradVol)) return false;
    for (int i = 0; i < 5; i++)
      if (gradVol->size[i] != vol->size[i] || (vol->size[i] > 1 && gradVol->stride[i] != vol->stride[i])) return false;
  }
  SamplerArgs<float> &s = a.s;
  s.vol = static_cast<const float *>(vol->data);
  s.gvol = gradVol ? static_cast<float *>(gradVol->data) : nullptr;
  s.B = (int)vol->size[0];
  s.iD = (int)vol->size[1], s.iH = (int)vol->size[2], s.iW = (int)vol->size[3], s.C = (int)vol->size[4];
  s.oD = (int)outlike->size[1], s.oH = (int)outlike->size[2], s.oW = (int)outlike->size[3];
  s.n = (long long)s.oD * s.oH * s.oW;
  s.dm1 = (float)(s.iD - 1), s.hm1 = (float)(s.iH - 1), s.wm1 = (float)(s.iW - 1);
  a.vB = vol->stride[0];
  a.vD = (int)vol->stride[1], a.vH = (int)vol->stride[2], a.vW = (int)vol->stride[3], a.vC = (int)vol->stride[4];
  a.oB = outlike->stride[0];
  a.osD = (int)outlike->stride[1], a.osH = (int)outlike->stride[2], a.osW = (int)outlike->stride[3];
  a.osC = (int)outlike->stride[4];
  p.c1 = (s.C == 1 && a.vW == 1) ? 1 : 0;
  p.two_rows = s.oW <= 32 ? 1 : 0;
  // fast tiers (sampler.cuh::VoxelSetup::begin / finish_fast) for THIS volume's strides
  memcpy(&s.fbx, &s.wm1, 4);
  memcpy(&s.fby, &s.hm1, 4);
  memcpy(&s.fbz, &s.dm1, 4);
  const unsigned sW = p.c1 ? 1u : (unsigned)a.vW;
  s.fneg = (int)(0u - 0x4B000000u * ((unsigned)a.vD + (unsigned)a.vH + sW));
  s.border_ok = 1;
  return true;
}

void plan_debug_flags(RowArgs &a, unsigned flags) {
  if (flags & VOLSTN_DEBUG_NO_FAST) a.s.fbx = a.s.fby = a.s.fbz = 0, a.s.border_ok = 0;
  if (flags & VOLSTN_DEBUG_NO_BORDER) a.s.border_ok = 0;
}

void plan_geometry(RowArgs &a, const RowPlan &p) {
  row_geometry(a.s.B, (int64_t)a.s.oD * a.s.oH, p.src, p.two_rows, &a.nblk, &a.rows_per_warp);
}

bool aos4(const volstn_tensor5 *g) {
  return g->size[4] == 4 && g->stride[4] == 1 && g->stride[3] == 4 && g->stride[2] % 4 == 0 && g->stride[1] % 4 == 0 &&
         g->stride[0] % 4 == 0 && reinterpret_cast<uintptr_t>(g->data) % 16 == 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// nn.BilinearSamplerVolumetric with strided / planar tensors
// ---------------------------------------------------------------------------------------------
// returns -1 when the row kernels cannot take the call (the caller falls back to the generic kernels)
int row_sampler_forward(int G, const volstn_tensor5 *vol, const volstn_tensor5 *grid, const volstn_tensor5 *out,
                        unsigned flags, cudaStream_t st) {
  RowArgs a;
  RowPlan p;
  zero_args(a);
  memset(&p, 0, sizeof(p));
  p.src = ROW_SRC_GRID;
  if (!plan_volume(a, p, vol, out, nullptr) || !row_tensor_ok(grid)) return -1;
  a.s.grid = static_cast<const float *>(grid->data);
  a.s.out = static_cast<float *>(out->data);
  a.gB = grid->stride[0];
  a.gD = (int)grid->stride[1], a.gH = (int)grid->stride[2], a.gW = (int)grid->stride[3], a.gC = (int)grid->stride[4];
  a.g3 = G == 3;
  p.grid_aos4 = G == 4 && aos4(grid);
  plan_debug_flags(a, flags);
  plan_geometry(a, p);
  return row_forward(a, p, a.s.B, st);
}

int row_sampler_backward(int G, const volstn_tensor5 *vol, const volstn_tensor5 *grid, const volstn_tensor5 *gradVol,
                         const volstn_tensor5 *gradGrid, const volstn_tensor5 *gradOut, bool og, bool ov, unsigned flags,
                         cudaStream_t st) {
  RowArgs a;
  RowPlan p;
  zero_args(a);
  memset(&p, 0, sizeof(p));
  p.src = ROW_SRC_GRID;
  if (!plan_volume(a, p, vol, gradOut, og ? nullptr : gradVol) || !row_tensor_ok(grid)) return -1;
  if (!ov && !row_tensor_ok(gradGrid)) return -1;
  a.s.grid = static_cast<const float *>(grid->data);
  a.s.gout = static_cast<const float *>(gradOut->data);
  a.gB = grid->stride[0];
  a.gD = (int)grid->stride[1], a.gH = (int)grid->stride[2], a.gW = (int)grid->stride[3], a.gC = (int)grid->stride[4];
  a.g3 = G == 3;
  p.grid_aos4 = G == 4 && aos4(grid);
  if (!ov) {
    a.s.ggrid = static_cast<float *>(gradGrid->data);
    a.ggB = gradGrid->stride[0];
    a.ggD = (int)gradGrid->stride[1], a.ggH = (int)gradGrid->stride[2], a.ggW = (int)gradGrid->stride[3];
    a.ggC = (int)gradGrid->stride[4];
    p.grid_aos4 = p.grid_aos4 && aos4(gradGrid);
  }
  plan_debug_flags(a, flags);
  plan_geometry(a, p);
  return row_backward(a, p, a.s.B, og, ov, st);
}

namespace {

// shape rules shared by the fused modules: `out`-like tensor B x oD x oH x oW x C against the volume
int check_fused(int dtype, const volstn_tensor5 *vol, const volstn_tensor5 *outlike) {
  if (dtype == VOLSTN_F64) return VOLSTN_ERR_UNSUPPORTED;  // the fused modules are float only
  if (dtype != VOLSTN_F32) return VOLSTN_ERR_ARG;
  if (!vol || !outlike) return VOLSTN_ERR_ARG;
  for (int i = 0; i < 5; i++)
    if (vol->size[i] < 0 || outlike->size[i] < 0) return VOLSTN_ERR_ARG;
  if (outlike->size[0] != vol->size[0] || outlike->size[4] != vol->size[4]) return VOLSTN_ERR_SHAPE;
  return VOLSTN_OK;
}

int zero_grad_vol(const volstn_tensor5 *gradVol, cudaStream_t st) {
  if (!desc_dense(gradVol)) return VOLSTN_ERR_LAYOUT;
  VOLSTN_CUDA_TRY(cudaMemsetAsync(gradVol->data, 0, (size_t)desc_numel(gradVol) * 4, st));
  return VOLSTN_OK;
}


// the accumulators of a training step -- gradCage (floats), the per-sample loss sums (doubles) and IoU counts (64-bit) --
// zeroed by ONE launch instead of three memset nodes (about 10 us of a 160 us step when the calls are not graph-captured)
__global__ void k_zero_step(float *gcage, size_t ncage, double *loss, unsigned long long *iou, int B) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t k = i; k < ncage; k += stride) gcage[k] = 0.f;
  if (i < (size_t)B) loss[i] = 0.0;
  if (iou && i < 2 * (size_t)B) iou[i] = 0ull;
}

int zero_step(float *gcage, size_t ncage, double *loss, uint64_t *iou, int64_t B, cudaStream_t st) {
  size_t need = ncage > 2 * (size_t)B ? ncage : 2 * (size_t)B;
  unsigned blocks = (unsigned)((need + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  if (blocks < (unsigned)((2 * B + 255) / 256)) blocks = (unsigned)((2 * B + 255) / 256);
  k_zero_step<<<blocks, 256, 0, st>>>(gcage, ncage, loss, reinterpret_cast<unsigned long long *>(iou), (int)B);
  return after_launch();
}

// VOLSTN_FAST_ARITH on contiguous single-channel tensors: the brick kernels (cagebrick.cu).  Returns false when they do
// not apply (the row kernels take the call).  `outlike`: out (forward), gradOut (backward) or the target (step).
bool brick_args(BrickArgs &k, unsigned flags, const void *cage, const int64_t cage_dims[3], const volstn_tensor5 *vol,
                const volstn_tensor5 *outlike, const volstn_tensor5 *gradVol) {
  if (!(flags & VOLSTN_FAST_ARITH) || (flags & VOLSTN_ALGO_MASK) == VOLSTN_ALGO_ROW) return false;
  if (flags & (VOLSTN_DEBUG_NO_FAST | VOLSTN_DEBUG_NO_BORDER)) return false;
  if (vol->size[4] != 1 || !vol->data || !outlike->data) return false;
  volstn_tensor5 v1 = *vol, o1 = *outlike;  // contiguous inside a sample; the batch stride is free
  v1.size[0] = o1.size[0] = 1;
  if (!desc_contiguous(&v1) || !desc_contiguous(&o1)) return false;
  if (outlike->size[0] > 1 && outlike->stride[0] != outlike->size[1] * outlike->size[2] * outlike->size[3]) return false;
  if (vol->stride[0] < 0) return false;
  memset(&k, 0, sizeof(k));
  for (int i = 1; i < 4; i++)
    if (vol->size[i] > (1 << 22) || outlike->size[i] > (1 << 22)) return false;
  k.vol = static_cast<const float *>(vol->data);
  k.cage = static_cast<const float *>(cage);
  k.vB = vol->stride[0];
  k.B = (int)vol->size[0];
  k.iD = (int)vol->size[1], k.iH = (int)vol->size[2], k.iW = (int)vol->size[3];
  k.oD = (int)outlike->size[1], k.oH = (int)outlike->size[2], k.oW = (int)outlike->size[3];
  k.cD = (int)cage_dims[0], k.cH = (int)cage_dims[1], k.cW = (int)cage_dims[2];
  if (gradVol) {
    if (!gradVol->data) return false;
    for (int i = 0; i < 5; i++)
      if (gradVol->size[i] != vol->size[i] || (vol->size[i] > 1 && gradVol->stride[i] != vol->stride[i])) return false;
    k.gv_delta = static_cast<const float *>(gradVol->data) - k.vol;
  }
  if (!brick_supported(k)) return false;
  brick_plan(k);
  return true;
}

}  // namespace
}  // namespace volstn

using namespace volstn;

extern "C" {

// =============================================================================================
// f3: grid generator folded into the sampler
// =============================================================================================
int volstn_b200_affine_sampler_forward(int dtype, const void *T, const volstn_tensor5 *vol, const volstn_tensor5 *out,
                                       unsigned flags, void *cuda_stream) {
  int rc = check_fused(dtype, vol, out);
  if (rc) return rc;
  if (out->size[1] < 2 || out->size[2] < 2 || out->size[3] < 2) return VOLSTN_ERR_SHAPE;  // VolumetricGridGenerator.lua:6-8
  if (desc_numel(out) == 0) return VOLSTN_OK;
  if (!T) return VOLSTN_ERR_ARG;
  if ((flags & VOLSTN_ALGO_MASK) != VOLSTN_ALGO_ROW && vol->data && out->data && vol->size[1] >= 1 && vol->size[2] >= 1 &&
      vol->size[3] >= 1) {
    // contiguous tensors: the flat voxel mapping of the packed kernels
    if ((rc = affine_forward_packed(static_cast<const float *>(T), vol, out, flags, static_cast<cudaStream_t>(cuda_stream))) != -1)
      return rc;
  }
  RowArgs a;
  RowPlan p;
  zero_args(a);
  memset(&p, 0, sizeof(p));
  p.src = ROW_SRC_AFFINE;
  if (!plan_volume(a, p, vol, out, nullptr)) return VOLSTN_ERR_LAYOUT;
  a.s.out = static_cast<float *>(out->data);
  a.T = static_cast<const float *>(T);
  plan_debug_flags(a, flags);
  plan_geometry(a, p);
  return row_forward(a, p, a.s.B, static_cast<cudaStream_t>(cuda_stream));
}

size_t volstn_b200_affine_sampler_backward_workspace(int64_t B, int64_t oD, int64_t oH, int64_t oW) {
  if (B <= 0 || oD <= 0 || oH <= 0) return 0;
  int nblk, rpw;
  row_geometry(B, oD * oH, ROW_SRC_AFFINE, oW <= 32, &nblk, &rpw);
  return (size_t)B * nblk * 16 * sizeof(double);
}

int volstn_b200_affine_sampler_backward(int dtype, const void *T, const volstn_tensor5 *vol,
                                        const volstn_tensor5 *gradVol, void *gradT, const volstn_tensor5 *gradOut,
                                        void *workspace, size_t workspace_bytes, unsigned flags, void *cuda_stream) {
  int rc = check_fused(dtype, vol, gradOut);
  if (rc) return rc;
  const bool og = flags & VOLSTN_BWD_ONLY_GRID, ov = flags & VOLSTN_BWD_ONLY_VOLUME;
  if (og && ov) return VOLSTN_ERR_ARG;
  if (gradOut->size[1] < 2 || gradOut->size[2] < 2 || gradOut->size[3] < 2) return VOLSTN_ERR_SHAPE;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (!og) {
    if (!gradVol) return VOLSTN_ERR_ARG;
    for (int i = 0; i < 5; i++)
      if (gradVol->size[i] != vol->size[i]) return VOLSTN_ERR_SHAPE;
    if ((flags & VOLSTN_BWD_ZERO_GRADVOL) && desc_numel(gradVol) && (rc = zero_grad_vol(gradVol, st))) return rc;
  }
  const int64_t B = vol->size[0];
  if (!ov) {
    if (B && !gradT) return VOLSTN_ERR_ARG;
    if (B && workspace_bytes < volstn_b200_affine_sampler_backward_workspace(B, gradOut->size[1], gradOut->size[2],
                                                                              gradOut->size[3]))
      return VOLSTN_ERR_WORKSPACE;
  }
  if (desc_numel(gradOut) == 0) {
    if (!ov && B) VOLSTN_CUDA_TRY(cudaMemsetAsync(gradT, 0, (size_t)B * 64, st));
    return VOLSTN_OK;
  }
  if (!T || (!ov && !workspace)) return VOLSTN_ERR_ARG;
  RowArgs a;
  RowPlan p;
  zero_args(a);
  memset(&p, 0, sizeof(p));
  p.src = ROW_SRC_AFFINE;
  if (!plan_volume(a, p, vol, gradOut, og ? nullptr : gradVol)) return VOLSTN_ERR_LAYOUT;
  a.s.gout = static_cast<const float *>(gradOut->data);
  a.T = static_cast<const float *>(T);
  a.partial = static_cast<double *>(workspace);
  a.gradT = static_cast<float *>(gradT);
  plan_debug_flags(a, flags);
  plan_geometry(a, p);
  return row_backward(a, p, a.s.B, og, ov, st);
}

// =============================================================================================
// f1: cage up-sample + source warp
// =============================================================================================
static int check_cage(const int64_t cage_dims[3]) {
  if (!cage_dims) return VOLSTN_ERR_ARG;
  for (int i = 0; i < 3; i++)
    if (cage_dims[i] < 1 || cage_dims[i] > 64) return cage_dims[i] < 1 ? VOLSTN_ERR_SHAPE : VOLSTN_ERR_UNSUPPORTED;
  if (cage_dims[2] > 20) return VOLSTN_ERR_UNSUPPORTED;  // backward: 3 * (cW + 1) (cell, channel) items on <= 64 lanes-trips
  return VOLSTN_OK;
}

int volstn_b200_cage_warp_forward(int dtype, const void *cage, const int64_t cage_dims[3], const volstn_tensor5 *vol,
                                  const volstn_tensor5 *out, unsigned flags, void *cuda_stream) {
  int rc = check_fused(dtype, vol, out);
  if (rc) return rc;
  if ((rc = check_cage(cage_dims))) return rc;
  if (out->size[1] < 2 || out->size[2] < 2 || out->size[3] < 2) return VOLSTN_ERR_SHAPE;
  if (desc_numel(out) == 0) return VOLSTN_OK;
  if (!cage) return VOLSTN_ERR_ARG;
  {
    BrickArgs k;
    if (brick_args(k, flags, cage, cage_dims, vol, out, nullptr)) {
      k.out = static_cast<float *>(out->data);
      return brick_launch(k, 0, false, static_cast<cudaStream_t>(cuda_stream));
    }
  }
  RowArgs a;
  RowPlan p;
  zero_args(a);
  memset(&p, 0, sizeof(p));
  p.src = ROW_SRC_CAGE;
  if (!plan_volume(a, p, vol, out, nullptr)) return VOLSTN_ERR_LAYOUT;
  a.s.out = static_cast<float *>(out->data);
  a.cage = static_cast<const float *>(cage);
  a.cD = (int)cage_dims[0], a.cH = (int)cage_dims[1], a.cW = (int)cage_dims[2];
  p.fast = (flags & VOLSTN_FAST_ARITH) ? 1 : 0;
  plan_debug_flags(a, flags);
  plan_geometry(a, p);
  return row_forward(a, p, a.s.B, static_cast<cudaStream_t>(cuda_stream));
}

int volstn_b200_cage_warp_backward(int dtype, const void *cage, const int64_t cage_dims[3], const volstn_tensor5 *vol,
                                   const volstn_tensor5 *gradVol, void *gradCage, const volstn_tensor5 *gradOut,
                                   unsigned flags, void *cuda_stream) {
  int rc = check_fused(dtype, vol, gradOut);
  if (rc) return rc;
  if ((rc = check_cage(cage_dims))) return rc;
  const bool og = flags & VOLSTN_BWD_ONLY_GRID, ov = flags & VOLSTN_BWD_ONLY_VOLUME;
  if (og && ov) return VOLSTN_ERR_ARG;
  if (gradOut->size[1] < 2 || gradOut->size[2] < 2 || gradOut->size[3] < 2) return VOLSTN_ERR_SHAPE;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (!og) {
    if (!gradVol) return VOLSTN_ERR_ARG;
    for (int i = 0; i < 5; i++)
      if (gradVol->size[i] != vol->size[i]) return VOLSTN_ERR_SHAPE;
    if ((flags & VOLSTN_BWD_ZERO_GRADVOL) && desc_numel(gradVol) && (rc = zero_grad_vol(gradVol, st))) return rc;
  }
  const int64_t B = vol->size[0];
  const size_t cage_bytes = (size_t)B * 3 * cage_dims[0] * cage_dims[1] * cage_dims[2] * 4;
  if (!ov && cage_bytes) {
    if (!gradCage) return VOLSTN_ERR_ARG;
    VOLSTN_CUDA_TRY(cudaMemsetAsync(gradCage, 0, cage_bytes, st));  // gradCage is ASSIGNED (like gradGrid, ...c:819-822)
  }
  if (desc_numel(gradOut) == 0) return VOLSTN_OK;
  if (!cage) return VOLSTN_ERR_ARG;
  if (!ov) {
    BrickArgs k;
    if (brick_args(k, flags, cage, cage_dims, vol, gradOut, og ? nullptr : gradVol)) {
      k.tgt = static_cast<const float *>(gradOut->data);
      k.gcage = static_cast<float *>(gradCage);
      return brick_launch(k, 1, !og, st);
    }
  }
  RowArgs a;
  RowPlan p;
  zero_args(a);
  memset(&p, 0, sizeof(p));
  p.src = ROW_SRC_CAGE;
  if (!plan_volume(a, p, vol, gradOut, og ? nullptr : gradVol)) return VOLSTN_ERR_LAYOUT;
  a.s.gout = static_cast<const float *>(gradOut->data);
  a.cage = static_cast<const float *>(cage);
  a.gcage = static_cast<float *>(gradCage);
  a.cD = (int)cage_dims[0], a.cH = (int)cage_dims[1], a.cW = (int)cage_dims[2];
  p.fast = (flags & VOLSTN_FAST_ARITH) ? 1 : 0;
  plan_debug_flags(a, flags);
  plan_geometry(a, p);
  return row_backward(a, p, a.s.B, og, ov, st);
}

// =============================================================================================
// f4: the training step of the cage warp in one pass -- forward, SmoothL1Criterion (model.lua:34), its gradient,
// the IoU counts of util.lua:110-120 and the backward to the cage
// =============================================================================================
int volstn_b200_cage_warp_step(int dtype, const void *cage, const int64_t cage_dims[3], const volstn_tensor5 *vol,
                               const volstn_tensor5 *target, const volstn_tensor5 *out, const volstn_tensor5 *gradVol,
                               void *gradCage, double *loss_sum, uint64_t *iou_counts, float grad_scale, unsigned flags,
                               void *cuda_stream) {
  int rc = check_fused(dtype, vol, target);
  if (rc) return rc;
  if ((rc = check_cage(cage_dims))) return rc;
  if (flags & VOLSTN_BWD_ONLY_VOLUME) return VOLSTN_ERR_ARG;
  const bool og = flags & VOLSTN_BWD_ONLY_GRID;
  if (vol->size[4] != 1) return VOLSTN_ERR_UNSUPPORTED;  // single-channel volumes (ALIGNet's occupancy grids)
  if (target->size[1] < 2 || target->size[2] < 2 || target->size[3] < 2) return VOLSTN_ERR_SHAPE;
  if (out)
    for (int i = 0; i < 5; i++)
      if (out->size[i] != target->size[i] || (target->size[i] > 1 && out->stride[i] != target->stride[i]))
        return VOLSTN_ERR_LAYOUT;  // the optional forward output shares the target's layout
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const int64_t B = vol->size[0];
  if (!og) {
    if (!gradVol) return VOLSTN_ERR_ARG;
    for (int i = 0; i < 5; i++)
      if (gradVol->size[i] != vol->size[i]) return VOLSTN_ERR_SHAPE;
    if ((flags & VOLSTN_BWD_ZERO_GRADVOL) && desc_numel(gradVol) && (rc = zero_grad_vol(gradVol, st))) return rc;
  }
  if (B == 0) return VOLSTN_OK;
  if (!gradCage || !loss_sum) return VOLSTN_ERR_ARG;
  const size_t cage_bytes = (size_t)B * 3 * cage_dims[0] * cage_dims[1] * cage_dims[2] * 4;
  if ((rc = zero_step(static_cast<float *>(gradCage), cage_bytes / 4, loss_sum, iou_counts, B, st))) return rc;
  if (desc_numel(target) == 0) return VOLSTN_OK;
  if (!cage || (out && !out->data)) return VOLSTN_ERR_ARG;
  {
    BrickArgs k;
    if (brick_args(k, flags, cage, cage_dims, vol, target, og ? nullptr : gradVol)) {
      k.tgt = static_cast<const float *>(target->data);
      k.out = out ? static_cast<float *>(out->data) : nullptr;
      k.gcage = static_cast<float *>(gradCage);
      k.loss_norm = grad_scale;
      k.loss = loss_sum;
      k.iou = reinterpret_cast<unsigned long long *>(iou_counts);
      return brick_launch(k, 2, !og, st);
    }
  }
  RowArgs a;
  RowPlan p;
  zero_args(a);
  memset(&p, 0, sizeof(p));
  p.src = ROW_SRC_CAGE;
  if (!plan_volume(a, p, vol, target, og ? nullptr : gradVol)) return VOLSTN_ERR_LAYOUT;
  if (!p.c1) return VOLSTN_ERR_LAYOUT;
  a.s.gout = static_cast<const float *>(target->data);
  a.s.out = out ? static_cast<float *>(out->data) : nullptr;
  a.cage = static_cast<const float *>(cage);
  a.gcage = static_cast<float *>(gradCage);
  a.cD = (int)cage_dims[0], a.cH = (int)cage_dims[1], a.cW = (int)cage_dims[2];
  a.loss_norm = grad_scale;
  a.loss = loss_sum;
  a.iou = reinterpret_cast<unsigned long long *>(iou_counts);
  p.fast = (flags & VOLSTN_FAST_ARITH) ? 1 : 0;
  plan_debug_flags(a, flags);
  plan_geometry(a, p);
  return row_cage_step(a, p, a.s.B, og, st);
}

// the same step for the plain sampler: the sampling grid is a tensor (any strides; G = 4 or 3) and its gradient is stored
int volstn_b200_sampler_step(int dtype, int G, const volstn_tensor5 *vol, const volstn_tensor5 *grid,
                             const volstn_tensor5 *target, const volstn_tensor5 *out, const volstn_tensor5 *gradVol,
                             const volstn_tensor5 *gradGrid, double *loss_sum, uint64_t *iou_counts, float grad_scale,
                             unsigned flags, void *cuda_stream) {
  int rc = check_fused(dtype, vol, target);
  if (rc) return rc;
  if (G != 3 && G != 4) return VOLSTN_ERR_ARG;
  if (!grid || !gradGrid) return VOLSTN_ERR_ARG;
  if (flags & VOLSTN_BWD_ONLY_VOLUME) return VOLSTN_ERR_ARG;
  const bool og = flags & VOLSTN_BWD_ONLY_GRID;
  if (vol->size[4] != 1) return VOLSTN_ERR_UNSUPPORTED;
  if (grid->size[0] != vol->size[0] || grid->size[4] != G) return VOLSTN_ERR_SHAPE;  // lua:33-34
  for (int i = 0; i < 5; i++)
    if (gradGrid->size[i] != grid->size[i]) return VOLSTN_ERR_SHAPE;
  for (int i = 0; i < 4; i++)
    if (target->size[i] != grid->size[i]) return VOLSTN_ERR_SHAPE;  // lua:36-41
  if (out)
    for (int i = 0; i < 5; i++)
      if (out->size[i] != target->size[i] || (target->size[i] > 1 && out->stride[i] != target->stride[i]))
        return VOLSTN_ERR_LAYOUT;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const int64_t B = vol->size[0];
  if (!og) {
    if (!gradVol) return VOLSTN_ERR_ARG;
    for (int i = 0; i < 5; i++)
      if (gradVol->size[i] != vol->size[i]) return VOLSTN_ERR_SHAPE;
    if ((flags & VOLSTN_BWD_ZERO_GRADVOL) && desc_numel(gradVol) && (rc = zero_grad_vol(gradVol, st))) return rc;
  }
  if (B == 0) return VOLSTN_OK;
  if (!loss_sum) return VOLSTN_ERR_ARG;
  if ((rc = zero_step(nullptr, 0, loss_sum, iou_counts, B, st))) return rc;
  if (desc_numel(target) == 0) return VOLSTN_OK;
  if (out && !out->data) return VOLSTN_ERR_ARG;
  RowArgs a;
  RowPlan p;
  zero_args(a);
  memset(&p, 0, sizeof(p));
  p.src = ROW_SRC_GRID;
  if (!plan_volume(a, p, vol, target, og ? nullptr : gradVol) || !row_tensor_ok(grid) || !row_tensor_ok(gradGrid))
    return VOLSTN_ERR_LAYOUT;
  if (!p.c1) return VOLSTN_ERR_LAYOUT;
  a.s.grid = static_cast<const float *>(grid->data);
  a.s.gout = static_cast<const float *>(target->data);
  a.s.out = out ? static_cast<float *>(out->data) : nullptr;
  a.s.ggrid = static_cast<float *>(gradGrid->data);
  a.gB = grid->stride[0];
  a.gD = (int)grid->stride[1], a.gH = (int)grid->stride[2], a.gW = (int)grid->stride[3], a.gC = (int)grid->stride[4];
  a.ggB = gradGrid->stride[0];
  a.ggD = (int)gradGrid->stride[1], a.ggH = (int)gradGrid->stride[2], a.ggW = (int)gradGrid->stride[3];
  a.ggC = (int)gradGrid->stride[4];
  a.g3 = G == 3;
  p.grid_aos4 = G == 4 && aos4(grid) && aos4(gradGrid);
  a.loss_norm = grad_scale;
  a.loss = loss_sum;
  a.iou = reinterpret_cast<unsigned long long *>(iou_counts);
  plan_debug_flags(a, flags);
  plan_geometry(a, p);
  return row_cage_step(a, p, a.s.B, og, st);
}

}  // extern "C"
