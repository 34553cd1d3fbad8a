/* TEST INFRASTRUCTURE: luaL_Reg as in lauxlib.h of Lua 5.1 */
#ifndef HARNESS_LAUXLIB_H
#define HARNESS_LAUXLIB_H
#include "lua.h"
typedef struct luaL_Reg {
  const char *name;
  lua_CFunction func;
} luaL_Reg;
#endif
