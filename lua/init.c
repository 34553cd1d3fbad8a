/*
 * lua/init.c -- drop-in replacement of stn3dtorch/init.c: the CPU-tensor library `libvolstn`
 * (`require 'libvolstn'`, stn3dtorch/init.lua:4).  Same entry point (luaopen_libvolstn, init.c:13-21), same global
 * table `volstn`, same four function names registered into the `nn` table of torch.FloatTensor and
 * torch.DoubleTensor (generic/BilinearSamplerVolumetric.c:831-844) -- but instead of compiling the scalar loops of
 * generic/BilinearSamplerVolumetric.c it links libvolstn_b200.so: host tensors are staged through the B200
 * (include/volstn_b200.h, volstn_b200_host_*).  There is no CPU compute path.
 *
 * Build: lua/CMakeLists.txt (Torch7 tree) -- or, without Torch7, against the stand-in headers of
 * tests/luat_harness/include (what __graft_entry__.build() and tests/test_lua_glue*.py do).
 */
#include "TH.h"
#include "luaT.h"

#include "volstn_glue.h"

#define torch_(NAME) TH_CONCAT_3(torch_, Real, NAME)
#define torch_Tensor TH_CONCAT_STRING_3(torch., Real, Tensor)
#define nn_(NAME) TH_CONCAT_3(nn_, Real, NAME)

#include "generic/volstn_host.c"
#include "THGenerateFloatTypes.h"

LUA_EXTERNC DLL_EXPORT int luaopen_libvolstn(lua_State *L);

int luaopen_libvolstn(lua_State *L) {
  lua_newtable(L);
  lua_pushvalue(L, -1);
  lua_setglobal(L, "volstn");
  nn_FloatBilinearSamplerVolumetric_init(L);
  nn_DoubleBilinearSamplerVolumetric_init(L);
  return 1;
}
