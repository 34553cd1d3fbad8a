// oracle/ref_cuda_shim.cu -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the reference's own CUDA kernels, stn3dtorch/BilinearSamplerVolumetric.cu, UNMODIFIED from
// /root/reference (REF_CU_FILE, given by oracle/Makefile) for sm_100a behind the two fake headers in
// oracle/fake_thc/, and exports the four launchers through a plain C entry point.  The result,
// oracle/_ref/libcuvolstn_ref.so, is the GPU INCUMBENT: the sm_20-era kernels recompiled for the B200,
// timed beside ours in tests/test_gpu_incumbent.py and checked there against the CPU reference first
// (its warp-synchronous shared-memory reductions carry no __syncwarp).  Nothing in
// alignet_b200/ links or loads it.
#include <lua.h>
#include "THCGeneral.h"

int volstn_ref_cuda_error = 0;
static THCState g_state = {0};
THCState *getCutorchState(lua_State *L) { (void)L; return &g_state; }

#include REF_CU_FILE

// which: 0 Volumetric fwd (vol, grid, out), 1 Volumetric bwd (vol, grid, gradVol, gradGrid, gradOut),
//        2 BHWD (G = 3) fwd, 3 BHWD bwd.  tensors[i] fills Lua stack slot 2 + i (slot 1 is `self`).
extern "C" int ref_cuda_call(int which, int ntensors, THCudaTensor *tensors, void *stream) {
  lua_State L;
  for (int i = 0; i < 8; i++) L.arg[i] = 0;
  for (int i = 0; i < ntensors && i < 6; i++) L.arg[2 + i] = &tensors[i];
  g_state.stream = (cudaStream_t)stream;
  volstn_ref_cuda_error = 0;
  switch (which) {
    case 0: cunn_BilinearSamplerVolumetric_updateOutput(&L); break;
    case 1: cunn_BilinearSamplerVolumetric_updateGradInput(&L); break;
    case 2: cunn_BilinearSamplerBHWD_updateOutput(&L); break;
    case 3: cunn_BilinearSamplerBHWD_updateGradInput(&L); break;
    default: return -1;
  }
  return volstn_ref_cuda_error;
}
