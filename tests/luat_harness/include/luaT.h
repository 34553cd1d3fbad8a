/* TEST INFRASTRUCTURE: the luaT names the glue uses (torch7/lib/luaT/luaT.h): luaT_checkudata raises a Lua error on a
 * type mismatch, luaT_pushmetatable pushes the metatable of a registered type, luaT_registeratname fills the table
 * `name` of the metatable on top of the stack with the functions of `methods`. */
#ifndef HARNESS_LUAT_H
#define HARNESS_LUAT_H
#ifdef __cplusplus
extern "C" {
#endif
#include "lua.h"
#include "lauxlib.h"
#ifdef __cplusplus
}
#define LUA_EXTERNC extern "C"
#else
#define LUA_EXTERNC extern
#endif
#define DLL_EXPORT __attribute__((visibility("default")))
#ifdef __cplusplus
extern "C" {
#endif
void *luaT_checkudata(lua_State *L, int ud, const char *tname);
int luaT_pushmetatable(lua_State *L, const char *tname);
void luaT_registeratname(lua_State *L, const struct luaL_Reg *methods, const char *name);
#ifdef __cplusplus
}
#endif
#endif
