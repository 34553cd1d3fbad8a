/* tests/luat_harness/include/lua.h -- TEST INFRASTRUCTURE: the handful of Lua 5.1 C-API names lua/init.c and
 * lua/init.cu use, backed by tests/luat_harness/harness.c (no Lua interpreter exists in this image).  Signatures follow
 * lua.h of Lua 5.1 / LuaJIT 2 (the macros lua_newtable / lua_setglobal / lua_getglobal / lua_pop are functions here). */
#ifndef HARNESS_LUA_H
#define HARNESS_LUA_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct lua_State lua_State;
typedef int (*lua_CFunction)(lua_State *L);
typedef double lua_Number;
lua_Number lua_tonumber(lua_State *L, int idx);
void *lua_touserdata(lua_State *L, int idx);
int lua_gettop(lua_State *L);
void lua_newtable(lua_State *L);
void lua_pushvalue(lua_State *L, int idx);
void lua_setglobal(lua_State *L, const char *name);
void lua_getglobal(lua_State *L, const char *name);
void lua_getfield(lua_State *L, int idx, const char *k);
void lua_call(lua_State *L, int nargs, int nresults);
void lua_settop(lua_State *L, int idx);
#define lua_pop(L, n) lua_settop(L, -(n)-1)
#ifdef __cplusplus
}
#endif
#endif
