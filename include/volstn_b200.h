/*
 * volstn_b200.h -- C ABI of the B200-native volumetric warping hot path.
 *
 * This header is the drop-in boundary: it sits where stn3dtorch/init.c (CPU library `libvolstn`)
 * and stn3dtorch/init.cu (CUDA library `libcuvolstn`) sit in the reference.  Every entry point
 * names the reference interface it replaces (paths relative to the reference root; "...c" is
 * stn3dtorch/generic/BilinearSamplerVolumetric.c, "...cu" is stn3dtorch/BilinearSamplerVolumetric.cu).
 * INTEGRATION.md shows the luaT glue a maintainer adds on the reference side, and
 * alignet_b200/nn.py is the Python mirror of the Lua nn modules that the tests drive.
 *
 * Conventions
 *  - plain C: pointers, sizes, element strides; no Torch / PyTorch types.
 *  - tensors are 5-D descriptors  B x D x H x W x {C | G}  exactly like the reference's
 *    THTensor (size[], stride[], data).  Strides are in ELEMENTS.  The reference assumes a unit stride
 *    on the last dimension (`+ t`, `+ 1`, `+ 2`: ...c:492-495, 556); this library honours ANY stride
 *    there as well, so the `B x D x H x W x C` view of a contiguous `B x C x D x H x W` tensor can be
 *    passed directly instead of through the materialised nn.Transpose copies of stn3dtorch/test.lua:17,21
 *    and models/alignet.lua:11,14,21.
 *  - volume layout channels-last `B x D x H x W x C`; grid `B x D x H x W x G`, components
 *    (z, y, x[, dis]) in [-1, 1]; -1 -> index 0, +1 -> index size-1; zero padding outside.
 *  - volstn_b200_* take DEVICE pointers and a CUDA stream (pass the caller's current stream, as
 *    the reference does with THCState_getCurrentStream, ...cu:706); they launch asynchronously,
 *    never allocate, never synchronise, never retain pointers.
 *  - volstn_b200_host_* take HOST pointers (the reference's torch.FloatTensor path, init.c:13-21):
 *    inputs are staged to the GPU, computed there and copied back inside the call, which returns
 *    after the results are in the host buffers.  There is NO CPU compute path in this library.
 *  - every function returns a volstn_status (0 = ok); nothing is thrown, nothing printed.  The
 *    reference's error convention is `printf + THError("aborting")` (...cu:736-740); the glue
 *    turns a non-zero status into exactly that.
 *  - re-entrant: no global mutable state other than an atomic launch counter.
 */
#ifndef VOLSTN_B200_H
#define VOLSTN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VOLSTN_B200_ABI_VERSION 1

typedef enum volstn_status {
  VOLSTN_OK = 0,
  VOLSTN_ERR_ARG = 1,         /* NULL pointer, bad G, bad dtype, negative size                     */
  VOLSTN_ERR_SHAPE = 2,       /* the asserts of BilinearSamplerVolumetric.lua:26-42 do not hold     */
  VOLSTN_ERR_LAYOUT = 3,      /* a tensor this entry point needs contiguous / dense / non-aliasing is not */
  VOLSTN_ERR_CUDA = 4,        /* a CUDA call failed; see volstn_b200_last_cuda_error()              */
  VOLSTN_ERR_UNSUPPORTED = 5, /* e.g. a sample with >= 2^31 elements                                */
  VOLSTN_ERR_NO_DEVICE = 6,   /* no sm_100 device: the library refuses to run (no fallback)         */
  VOLSTN_ERR_WORKSPACE = 7    /* workspace too small                                                */
} volstn_status;

typedef enum volstn_dtype {
  VOLSTN_F32 = 0, /* torch.FloatTensor / torch.CudaTensor (init.c:8-9; ...cu is float only) */
  VOLSTN_F64 = 1  /* torch.DoubleTensor (init.c:18-19)                                      */
} volstn_dtype;

typedef struct volstn_tensor5 {
  void *data;        /* device pointer (volstn_b200_*) or host pointer (volstn_b200_host_*) */
  int64_t size[5];   /* B, D, H, W, C (volume-like) or G (grid-like)                        */
  int64_t stride[5]; /* element strides, ANY value on all five dimensions (the reference needs
                        stride[4] == 1, ...c:492-495; this library does not)                 */
} volstn_tensor5;

/* flags for volstn_b200_sampler_backward ---------------------------------------------------- */
#define VOLSTN_BWD_ONLY_GRID 0x1u     /* skip gradVol: the reference's onlyGrid switch (...c:594, ...cu:230,1079) */
#define VOLSTN_BWD_ONLY_VOLUME 0x2u   /* skip gradGrid (its grid is a constant, e.g. the cage up-sample)        */
#define VOLSTN_BWD_ZERO_GRADVOL 0x4u  /* zero-fill gradVol inside the call (replaces the Lua :zero() of
                                         BilinearSamplerVolumetric.lua:103); without it gradVol is
                                         ACCUMULATED into, exactly like the reference                       */
/* HOST entry points only, OPT-IN (the reference's CPU code re-reads its inputs in updateGradInput, ...c:587-640, and
 * that is what flags = 0 does): keep device copies of the forward call's volume and grid in the context (KEEP,
 * forward) and let the next backward call on the SAME host tensors skip their upload (REUSE, backward) -- 320 of
 * the 704 MB a 64^3 x 64 step uploads.  Matched by host pointer, shape, strides and dtype only: the caller promises
 * not to have modified the tensors between the two calls. */
#define VOLSTN_HOST_KEEP_INPUTS 0x8u
#define VOLSTN_HOST_REUSE_INPUTS 0x10u
/* HOST sampler entry points: return as soon as the copies and kernels are enqueued; the outputs are valid (and the
 * inputs may be modified again) after volstn_b200_host_wait().  With pinned host memory the uploads of the next call
 * then overlap the downloads of this one. */
#define VOLSTN_HOST_ASYNC 0x20u
/* algorithm selection (bits 8..11), 0 = automatic.  Results are identical up to the summation
 * order of gradVol; exposed so that bench.py / profiles can time each variant. */
#define VOLSTN_ALGO_SHIFT 8
#define VOLSTN_ALGO_MASK 0xF00u
#define VOLSTN_ALGO_AUTO 0x000u
#define VOLSTN_ALGO_DIRECT 0x100u /* one red.global per in-bounds corner (vector red for C = 2, 4)   */
#define VOLSTN_ALGO_TILE 0x200u   /* shared-memory privatised gradVol tile per CTA, falls back per CTA.  ACCURACY: the tile
                                     accumulates in fixed point scaled to max|gradOut| of each 256-voxel brick, so a contribution
                                     is rounded by at most 2^-23 of that maximum (gradVol's contract, 1e-4 * max|ref|, holds; a
                                     contribution 2^23 times smaller than the brick's largest gradOut is lost).  AUTO takes this
                                     path only when the output has >= 128 voxels per volume cell (ALIGNet's cage up-sample);
                                     VOLSTN_ALGO_DIRECT never does */
#define VOLSTN_ALGO_QUAD 0x400u   /* backward, C = 1: x0/x0+1 contributions as one 16-byte reduction on the aligned quad (AUTO
                                     chooses this per warp when the warp's cells are scattered; DIRECT never does) */
#define VOLSTN_ALGO_TMA 0x600u    /* single-channel packed float tensors: a CTA per 8^3 output brick, the bounding box of its volume cells
                                     staged in shared memory by ONE TMA box load (hardware zero fill = the reference's padding rule),
                                     gradVol privatised in fixed point and flushed by ONE TMA box reduction; out / gradGrid keep the
                                     reference's bits, gradVol as VOLSTN_ALGO_TILE.  For sheared warp fields; opt-in */
#define VOLSTN_ALGO_ROW 0x500u    /* the row kernels (a warp per output row; the path of strided / planar tensors) even for packed tensors */
/* 0x300 was VOLSTN_ALGO_PAIR (packed fp32x2 arithmetic): measured 12-25 % slower in round 1 and removed from the
 * library in round 2 (record: profiles/r01_kbench_history.txt); the value now selects the automatic path */
/* debugging / A-B timing: force the slower but always-valid code paths (results are identical) */
#define VOLSTN_DEBUG_NO_FAST 0x10000u   /* every warp takes the exact predicated path            */
#define VOLSTN_DEBUG_NO_BORDER 0x20000u /* warps touching a high face take the exact path (no border tier) */
/* Cage entry points (volstn_b200_cage_warp_forward / _backward / _step), single-channel volumes: FAST ARITHMETIC.  The
 * cage up-sample is evaluated in its separable form, the trilinear value and its three coordinate derivatives come from
 * the seven nested linear interpolations, every sum is an FMA chain.  Same mathematics as the default path, different
 * rounding: the up-sampled field moves by a few ulp, so `out` is within 1e-5 * max|out| of the reference chain instead
 * of bit-identical, and where a sampling coordinate sits on a cell boundary (the identity cage: every voxel) the source
 * warp may take the neighbouring cell -- with an interpolation weight within a few ulp of 0 or 1, i.e. the same value,
 * and the OTHER one-sided derivative for gradCage (the function is only C0 there; the reference's own CUDA path differs
 * from its CPU path in the same way, its kernels being compiled with FMA contraction).  That includes the LOW faces of
 * the source volume: a coordinate less than size * 2^-23 (a few ulp) below 0 is read as 0.
 * Kernels: contiguous tensors (the training case) run the brick kernels of csrc/cagebrick.cu -- a warp per cage cell and
 * 32-wide x slot, the cage folded along x per task and along z per z, the gradient folded back the same way -- other
 * layouts the row kernels of csrc/rowwarp.cu with the same contract (VOLSTN_ALGO_ROW forces them, for A/B timing).
 * Off by default. */
#define VOLSTN_FAST_ARITH 0x40000u
/* voxels per thread (bits 12..15), 0 = automatic */
#define VOLSTN_VPT_SHIFT 12
#define VOLSTN_VPT_MASK 0xF000u

/* ---------------------------------------------------------------------------------------------
 * Sampler (nn.BilinearSamplerVolumetric), DEVICE pointers
 * ------------------------------------------------------------------------------------------- */

/* Trilinear gather.  Replaces
 *   G = 4: nn_(BilinearSamplerVolumetric_updateOutput)  ...c:442-583 ; cunn_... ...cu:695-742
 *   G = 3: nn_(BilinearSamplerBHWD_updateOutput)        ...c:9-167   ; cunn_... ...cu:179-228
 * Lua stack order (self, vol, grid, out): ...c:444-446.  `out` must be B x gD x gH x gW x C
 * (BilinearSamplerVolumetric.lua:73); every element of it is overwritten.
 * out and the floor/mask computation are bit-identical to the reference's CPU float (double)
 * build compiled without FMA contraction. */
int volstn_b200_sampler_forward(int dtype, int G, const volstn_tensor5 *vol, const volstn_tensor5 *grid,
                                const volstn_tensor5 *out, unsigned flags, void *cuda_stream);

/* gradInput via scatter-add + gradGrid via the 8-corner derivative.  Replaces
 *   G = 4: nn_(BilinearSamplerVolumetric_updateGradInput) ...c:585-829 ; cunn_... ...cu:1009-1072
 *   G = 3: nn_(BilinearSamplerBHWD_updateGradInput)       ...c:169-439 ; cunn_... ...cu:511-575
 * Lua stack order (self, vol, grid, gradVol, gradGrid, gradOut): ...c:587-591.
 * gradGrid is assigned (4th component 0 for G = 4, ...c:822); gradVol is accumulated into unless
 * VOLSTN_BWD_ZERO_GRADVOL.  gradGrid is bit-identical to the reference CPU build; gradVol differs
 * only by floating-point summation order (atomics), <= 1e-4 * max|ref|. */
int volstn_b200_sampler_backward(int dtype, int G, const volstn_tensor5 *vol, const volstn_tensor5 *grid,
                                 const volstn_tensor5 *gradVol, const volstn_tensor5 *gradGrid,
                                 const volstn_tensor5 *gradOut, unsigned flags, void *cuda_stream);

/* Debug dump for the bit-exact parity test: idx[v*3 + {0,1,2}] = (z0, y0, x0) as int32 and
 * mask[v] bit k = corner k in bounds (k: bit0 = +x, bit1 = +y, bit2 = +z; ...c:542-550), for v in
 * raster order over (b, z, y, x) of the grid.  vol_size = {B, D, H, W, C}. */
int volstn_b200_sampler_indices(int dtype, int G, const int64_t vol_size[5], const volstn_tensor5 *grid,
                                int32_t *idx, uint8_t *mask, void *cuda_stream);

/* ---------------------------------------------------------------------------------------------
 * Fused modules, DEVICE pointers, float only.  They are NEW
 * entry points -- the reference has no counterpart to bind -- whose results are, bit for bit, those of
 * the chain of un-fused calls they replace (gradients of reductions: within the stated tolerance).
 * ------------------------------------------------------------------------------------------- */

/* out = BilinearSamplerVolumetric(vol, VolumetricGridGenerator(oD, oH, oW)(T)) without the 16 B/voxel grid
 * ever touching memory: the `projector` of stn3dtorch/test.lua:15-25 and the loader-side affine
 * augmentation (dataset_2d.lua:285-311).  T: B x 4 x 4 (row 3 unused); out: B x oD x oH x oW x C.
 * Replaces volstn_b200_gridgen_forward + volstn_b200_sampler_forward (G = 4). */
int volstn_b200_affine_sampler_forward(int dtype, const void *T, const volstn_tensor5 *vol,
                                       const volstn_tensor5 *out, unsigned flags, void *cuda_stream);
/* gradVol (scatter-add; VOLSTN_BWD_ZERO_GRADVOL / VOLSTN_BWD_ONLY_GRID as for the sampler) and
 * gradT = VolumetricGridGenerator:updateGradInput(BilinearSamplerVolumetric gradGrid) (assigned; row 3 is zero
 * because gradGrid[..., 3] is, ...c:822).  Deterministic two-stage reduction; workspace as reported below. */
int volstn_b200_affine_sampler_backward(int dtype, const void *T, const volstn_tensor5 *vol,
                                        const volstn_tensor5 *gradVol, void *gradT, const volstn_tensor5 *gradOut,
                                        void *workspace, size_t workspace_bytes, unsigned flags, void *cuda_stream);
size_t volstn_b200_affine_sampler_backward_workspace(int64_t B, int64_t oD, int64_t oH, int64_t oW);

/* ALIGNet's warp: the B x 3 x cD x cH x cW control cage (models/alignet_2d.lua:136, after nn.Cumsum and
 * the +beta) is up-sampled to the resolution of `out` by sampling it at the identity field
 * (models/alignet_2d.lua:138-147, alignet.lua:27-39 lifted to 3-D) and the source volume is sampled at the
 * resulting field -- in one pass, the full-resolution field never exists.  Replaces
 *   gridgen_forward(I) ; sampler_forward(cage volume B x cD x cH x cW x 4, identity field) ; sampler_forward(vol, field).
 * cage_dims = {cD, cH, cW}, each <= 64, cW <= 20.  out: B x oD x oH x oW x C, extents >= 2. */
int volstn_b200_cage_warp_forward(int dtype, const void *cage, const int64_t cage_dims[3], const volstn_tensor5 *vol,
                                  const volstn_tensor5 *out, unsigned flags, void *cuda_stream);
/* gradVol as for the sampler; gradCage (B x 3 x cD x cH x cW, ASSIGNED) = the cage up-sample's gradInput[1] of the
 * source warp's gradGrid.  VOLSTN_BWD_ONLY_GRID skips gradVol (the source is data, models/alignet_2d.lua:23-24): the
 * descriptor is then not read and may be NULL (also in volstn_b200_cage_warp_step); VOLSTN_BWD_ONLY_VOLUME skips gradCage. */
int volstn_b200_cage_warp_backward(int dtype, const void *cage, const int64_t cage_dims[3], const volstn_tensor5 *vol,
                                   const volstn_tensor5 *gradVol, void *gradCage, const volstn_tensor5 *gradOut,
                                   unsigned flags, void *cuda_stream);

/* The training step of the cage warp in ONE pass (single-channel volumes):
 *     out      = cage_warp_forward(cage, vol)                                   (optionally stored: `out` may be NULL)
 *     loss_sum[b] = sum over the voxels of sample b of SmoothL1(out - target)     (nn.SmoothL1Criterion, model.lua:34:
 *                   z = |x|, z < 1 ? 0.5 z^2 : z - 0.5, accumulated in double like THNN's accreal; NOT yet divided)
 *     gradOut  = grad_scale * clamp(out - target, -1, 1)                        (SmoothL1Criterion_updateGradInput;
 *                   grad_scale = 1 / nElement for sizeAverage = true, the Torch default)
 *     gradVol, gradCage = cage_warp_backward(gradOut)                           (gradCage assigned, gradVol as usual)
 *     iou_counts[b] = {#(out > 0.5 and target > 0.5), #(out > 0.5 or target > 0.5)}   (computeIOU, util.lua:110-120;
 *                   exact integer counts, may be NULL)
 * Neither `out` nor gradOut ever has to exist in memory: the step reads the source gathers and the target (8 B/voxel).
 * `out`, when given, must have the target's shape and strides. */
int volstn_b200_cage_warp_step(int dtype, const void *cage, const int64_t cage_dims[3], const volstn_tensor5 *vol,
                               const volstn_tensor5 *target, const volstn_tensor5 *out, const volstn_tensor5 *gradVol,
                               void *gradCage, double *loss_sum, uint64_t *iou_counts, float grad_scale, unsigned flags,
                               void *cuda_stream);

/* The same step for the plain sampler (a network that emits a dense field): out = sampler_forward(vol, grid), loss,
 * gradOut, gradVol / gradGrid = sampler_backward -- one pass, single-channel volumes, any strides, G = 4 or 3.
 * gradGrid is assigned; flags, loss_sum, iou_counts, grad_scale and `out` as for volstn_b200_cage_warp_step. */
int volstn_b200_sampler_step(int dtype, int G, const volstn_tensor5 *vol, const volstn_tensor5 *grid,
                             const volstn_tensor5 *target, const volstn_tensor5 *out, const volstn_tensor5 *gradVol,
                             const volstn_tensor5 *gradGrid, double *loss_sum, uint64_t *iou_counts, float grad_scale,
                             unsigned flags, void *cuda_stream);

/* ---------------------------------------------------------------------------------------------
 * Grid generator (nn.VolumetricGridGenerator), DEVICE pointers, contiguous tensors
 * ------------------------------------------------------------------------------------------- */

/* out[b,z,y,x,:] = T[b] (4x4) . (zn, yn, xn, 1),  zn = -1 + z/(D-1)*2 evaluated in double and
 * rounded to `dtype` like the Lua constructor (VolumetricGridGenerator.lua:12-23).  Replaces
 * PGG:updateOutput's batchGrid replication + torch.bmm (VolumetricGridGenerator.lua:50-60); no
 * B-replicated base grid is materialised.  T: B x 4 x 4, out: B x D x H x W x 4. */
int volstn_b200_gridgen_forward(int dtype, const void *T, void *out, int64_t B, int64_t D, int64_t H,
                                int64_t W, void *cuda_stream);

/* gradT[b] (4x4) = sum_v gradGrid[b,v,:]^T (x) base[v,:].  Replaces PGG:updateGradInput's
 * zero() + baddbmm (VolumetricGridGenerator.lua:79-82).  Deterministic (fixed reduction tree).
 * workspace: device scratch of at least volstn_b200_gridgen_backward_workspace() bytes. */
int volstn_b200_gridgen_backward(int dtype, const void *gradGrid, void *gradT, int64_t B, int64_t D,
                                 int64_t H, int64_t W, void *workspace, size_t workspace_bytes,
                                 void *cuda_stream);
size_t volstn_b200_gridgen_backward_workspace(int64_t B, int64_t D, int64_t H, int64_t W);

/* ---------------------------------------------------------------------------------------------
 * Cumsum (nn.Cumsum), DEVICE pointers, contiguous  B x N x d[0] x ... x d[N-1]  tensors
 * ------------------------------------------------------------------------------------------- */

/* Channel i is prefix-summed along spatial axis i.  Replaces Cumsum:updateOutput's per-channel
 * split/squeeze/cumsum (Cumsum.lua:12-22) with one launch.  Prefixes accumulate in double and are
 * rounded to `dtype` -- TH's CPU cumsum (accreal = double), the parity reference of this repository
 * (results must match the reference's CPU path).  Known deviation: for torch.CudaTensor the reference
 * runs cutorch's scan, which accumulates in float; a float prefix can differ from the double-accumulated
 * one by one ulp.  Both entry-point families (device and host) use the CPU semantics. */
int volstn_b200_cumsum_forward(int dtype, const void *in, void *out, int64_t B, int N, const int64_t *dims,
                               void *cuda_stream);

/* gradIn = suffix sum of gradOut along the same axes.  Replaces Cumsum:updateGradInput's
 * flip / cumsum / flip / copy per channel (Cumsum.lua:39-48) with one launch. */
int volstn_b200_cumsum_backward(int dtype, const void *gradOut, void *gradIn, int64_t B, int N,
                                const int64_t *dims, void *cuda_stream);

/* ---------------------------------------------------------------------------------------------
 * HOST-buffer entry points: what `require 'libvolstn'` binds for torch.FloatTensor /
 * torch.DoubleTensor (init.c:13-21; lua/init.c in this repository is that binding).  Pointers in the
 * descriptors are HOST pointers (pinned memory gives full PCIe bandwidth; pageable memory works).  The
 * batch is cut into slices that are pipelined H2D -> kernel -> D2H over several streams; the call
 * returns when the outputs are in host memory (unless VOLSTN_HOST_ASYNC).  A context owns the device
 * staging buffers and streams and may be reused across calls from one thread at a time; create one per
 * thread for concurrent use.
 * Sampler tensors may be strided on every dimension, as the reference's CPU code honours them
 * (...c:460-473, 607-630): dense slices move with one copy, others with pitched copies over their
 * contiguous runs (a layout that would need more than 32768 copies per slice is VOLSTN_ERR_LAYOUT, and so
 * is an OUTPUT tensor whose elements alias each other).  Inputs may be broadcast (stride 0).
 * ------------------------------------------------------------------------------------------- */
typedef struct volstn_host_ctx volstn_host_ctx;

int volstn_b200_host_ctx_create(int device, volstn_host_ctx **ctx);
int volstn_b200_host_ctx_destroy(volstn_host_ctx *ctx);
/* pin / unpin a caller-owned host buffer (cudaHostRegister); optional, for bandwidth only */
int volstn_b200_host_pin(void *ptr, size_t bytes);
int volstn_b200_host_unpin(void *ptr);
/* asynchronous calls (VOLSTN_HOST_ASYNC): ticket of the most recently enqueued call, and wait until the call with
 * that ticket (and everything before it) is complete; ticket 0 waits for everything */
uint64_t volstn_b200_host_ticket(volstn_host_ctx *ctx);
int volstn_b200_host_wait(volstn_host_ctx *ctx, uint64_t ticket);
/* NUMA placement of host buffers: memory node next to `device` (-1 if the platform does not say), and pin the
 * CALLING THREAD to the CPUs next to it so that memory it allocates and first touches afterwards is node-local
 * (a no-op without topology information) */
int volstn_b200_host_numa_node(int device);
int volstn_b200_host_bind_thread(int device);

int volstn_b200_host_sampler_forward(volstn_host_ctx *ctx, int dtype, int G, const volstn_tensor5 *vol,
                                     const volstn_tensor5 *grid, const volstn_tensor5 *out, unsigned flags);
int volstn_b200_host_sampler_backward(volstn_host_ctx *ctx, int dtype, int G, const volstn_tensor5 *vol,
                                      const volstn_tensor5 *grid, const volstn_tensor5 *gradVol,
                                      const volstn_tensor5 *gradGrid, const volstn_tensor5 *gradOut,
                                      unsigned flags);
int volstn_b200_host_gridgen_forward(volstn_host_ctx *ctx, int dtype, const void *T, void *out, int64_t B,
                                     int64_t D, int64_t H, int64_t W);
int volstn_b200_host_gridgen_backward(volstn_host_ctx *ctx, int dtype, const void *gradGrid, void *gradT,
                                      int64_t B, int64_t D, int64_t H, int64_t W);
/* the fused modules on HOST tensors (contiguous, float): see volstn_b200_affine_sampler_* / volstn_b200_cage_warp_*.
 * gradVol is always zero-filled first (the module's :zero()); workspace is owned by the context.  VOLSTN_F64 is
 * refused with VOLSTN_ERR_UNSUPPORTED: these are new modules, the reference's Double registration (init.c:18-19)
 * covers the four sampler functions only, and those accept VOLSTN_F64 above. */
int volstn_b200_host_affine_sampler_forward(volstn_host_ctx *ctx, int dtype, const void *T, const volstn_tensor5 *vol,
                                            const volstn_tensor5 *out, unsigned flags);
int volstn_b200_host_affine_sampler_backward(volstn_host_ctx *ctx, int dtype, const void *T, const volstn_tensor5 *vol,
                                             const volstn_tensor5 *gradVol, void *gradT,
                                             const volstn_tensor5 *gradOut, unsigned flags);
int volstn_b200_host_cage_warp_forward(volstn_host_ctx *ctx, int dtype, const void *cage, const int64_t cage_dims[3],
                                       const volstn_tensor5 *vol, const volstn_tensor5 *out, unsigned flags);
int volstn_b200_host_cage_warp_backward(volstn_host_ctx *ctx, int dtype, const void *cage, const int64_t cage_dims[3],
                                        const volstn_tensor5 *vol, const volstn_tensor5 *gradVol, void *gradCage,
                                        const volstn_tensor5 *gradOut, unsigned flags);
int volstn_b200_host_cumsum_forward(volstn_host_ctx *ctx, int dtype, const void *in, void *out, int64_t B,
                                    int N, const int64_t *dims);
int volstn_b200_host_cumsum_backward(volstn_host_ctx *ctx, int dtype, const void *gradOut, void *gradIn,
                                     int64_t B, int N, const int64_t *dims);

/* ---------------------------------------------------------------------------------------------
 * Utilities
 * ------------------------------------------------------------------------------------------- */
int volstn_b200_abi_version(void);
const char *volstn_b200_strerror(int status);
/* cudaError_t of the most recent failing CUDA call on the calling thread (0 if none) and its text */
int volstn_b200_last_cuda_error(void);
const char *volstn_b200_last_cuda_error_string(void);
/* VOLSTN_OK iff `device` exists and is compute capability 10.x (the only code this library carries) */
int volstn_b200_device_check(int device);
/* number of kernels this library has launched in this process (monotonic, atomic) */
uint64_t volstn_b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VOLSTN_B200_H */
