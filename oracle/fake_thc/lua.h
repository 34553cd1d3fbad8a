/* oracle/fake_thc/lua.h -- TEST INFRASTRUCTURE ONLY.  The ~10 names of Lua / luaT that the reference's
 * stn3dtorch/BilinearSamplerVolumetric.cu touches, so that the file compiles UNMODIFIED from
 * /root/reference without Torch7 (which does not exist in this image).  See oracle/ref_cuda_shim.cu. */
#ifndef VOLSTN_FAKE_LUA_H
#define VOLSTN_FAKE_LUA_H
typedef struct lua_State { void *arg[8]; } lua_State; /* arg[i] = Lua stack slot i */
struct luaL_Reg { const char *name; int (*func)(lua_State *); };
static inline void *luaT_checkudata(lua_State *L, int idx, const char *tname) { (void)tname; return L->arg[idx]; }
static inline void luaT_pushmetatable(lua_State *L, const char *tname) { (void)L; (void)tname; }
static inline void luaT_registeratname(lua_State *L, const struct luaL_Reg *r, const char *name) { (void)L; (void)r; (void)name; }
#define lua_pop(L, n) ((void)0)
#endif
