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

inline int64_t desc_numel(const volstn_tensor5 *t) {
  return t->size[0] * t->size[1] * t->size[2] * t->size[3] * t->size[4];
}

}  // namespace volstn
