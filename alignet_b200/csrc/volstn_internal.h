// volstn_internal.h -- declarations shared between translation units (not part of the ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/volstn_b200.h"

namespace volstn {

int sampler_forward_impl(int dtype, int G, const volstn_tensor5 *vol, const volstn_tensor5 *grid,
                         const volstn_tensor5 *out, unsigned flags, cudaStream_t st);
int sampler_backward_impl(int dtype, int G, const volstn_tensor5 *vol, const volstn_tensor5 *grid,
                          const volstn_tensor5 *gradVol, const volstn_tensor5 *gradGrid,
                          const volstn_tensor5 *gradOut, unsigned flags, cudaStream_t st);

inline bool desc_contiguous(const volstn_tensor5 *t) {
  int64_t expect = 1;
  for (int i = 4; i >= 0; i--) {
    if (t->size[i] != 1 && t->stride[i] != expect) return false;
    expect *= t->size[i];
  }
  return true;
}

// dense = some permutation of a contiguous tensor (every element of the memory block belongs to it exactly once),
// e.g. the B x D x H x W x C view of a contiguous B x C x D x H x W tensor: a flat memset covers it
inline bool desc_dense(const volstn_tensor5 *t) {
  int order[5] = {0, 1, 2, 3, 4};
  for (int i = 0; i < 5; i++)
    for (int j = i + 1; j < 5; j++)
      if (t->stride[order[j]] < t->stride[order[i]] ||
          (t->stride[order[j]] == t->stride[order[i]] && t->size[order[j]] < t->size[order[i]])) {
        const int tmp = order[i];
        order[i] = order[j];
        order[j] = tmp;
      }
  int64_t expect = 1;
  for (int i = 0; i < 5; i++) {
    const int d = order[i];
    if (t->size[d] == 1) continue;
    if (t->stride[d] != expect) return false;
    expect *= t->size[d];
  }
  return true;
}

// row kernels (rowwarp.cu / fused_api.cu) for float tensors with arbitrary B/D/H strides, channels-last or planar;
// return -1 when they cannot take the call
int row_sampler_forward(int G, const volstn_tensor5 *vol, const volstn_tensor5 *grid, const volstn_tensor5 *out,
                        unsigned flags, cudaStream_t st);
int row_sampler_backward(int G, const volstn_tensor5 *vol, const volstn_tensor5 *grid, const volstn_tensor5 *gradVol,
                         const volstn_tensor5 *gradGrid, const volstn_tensor5 *gradOut, bool only_grid, bool only_vol,
                         unsigned flags, cudaStream_t st);

// grid generator folded into the PACKED forward (sampler.cu); -1 when it does not apply
int affine_forward_packed(const float *T, const volstn_tensor5 *vol, const volstn_tensor5 *out, unsigned flags,
                          cudaStream_t st);

inline int64_t desc_numel(const volstn_tensor5 *t) {
  return t->size[0] * t->size[1] * t->size[2] * t->size[3] * t->size[4];
}

}  // namespace volstn
