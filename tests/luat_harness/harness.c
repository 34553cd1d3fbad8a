/*
 * tests/luat_harness/harness.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A stand-in for the Lua interpreter + luaT + TH that lets the REAL glue (lua/init.c, lua/init.cu) be compiled and
 * driven in an image without Torch7 -- the same trick oracle/ref_shim.c uses to drive the reference's own
 * lua_CFunctions.  A fake lua_State holds
 *   - the argument slots 1..9 of the call being made (tensors with their luaT type name, numbers, nil),
 *   - a small value stack for the few pushes the glue performs (tables, the cutorch.getState() call),
 *   - the registration table that luaT_registeratname() fills (type name, table name, function name, function).
 * THError() / luaT type errors longjmp back into harness_call(), which reports them like a Lua pcall would.
 */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "THC.h"
#include "luaT.h"

enum { K_NIL = 0, K_TENSOR, K_NUMBER, K_TABLE, K_FUNC, K_USERDATA };

typedef struct value {
  int kind;
  void *p;
  const char *tname;
  double num;
} value;

struct lua_State {
  value arg[16];
  int nargs;
  value stk[32];
  int top; /* entries in stk; stack index nargs + 1 + i */
  const char *meta_on_top;
  THCState cutorch;
  struct {
    char type[40], table[16];
    const char *name;
    lua_CFunction fn;
  } reg[96];
  int nreg;
  char global_set[64];
  char error[512];
  jmp_buf jb;
  int armed;
};

static __thread lua_State *g_cur;

static void raise_error(lua_State *L, const char *msg) {
  snprintf(L->error, sizeof(L->error), "%s", msg);
  if (L->armed) longjmp(L->jb, 1);
  fprintf(stderr, "harness: error outside a protected call: %s\n", msg);
  abort();
}

void THError(const char *fmt, ...) {
  char buf[400];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (!g_cur) {
    fprintf(stderr, "THError without a running call: %s\n", buf);
    abort();
  }
  raise_error(g_cur, buf);
}

/* ---- Lua C API subset ---- */
static value *at(lua_State *L, int idx) {
  static value nil = {K_NIL, NULL, NULL, 0};
  if (idx > 0) {
    if (idx <= L->nargs) return &L->arg[idx];
    idx -= L->nargs + 1;
    return (idx < L->top) ? &L->stk[idx] : &nil;
  }
  if (idx < 0 && -idx <= L->top) return &L->stk[L->top + idx];
  return &nil;
}
static void push(lua_State *L, value v) {
  if (L->top >= 32) raise_error(L, "harness: value stack overflow");
  L->stk[L->top++] = v;
}
lua_Number lua_tonumber(lua_State *L, int idx) {
  const value *v = at(L, idx);
  return v->kind == K_NUMBER ? v->num : 0.0; /* nil / none -> 0, like Lua */
}
void *lua_touserdata(lua_State *L, int idx) {
  const value *v = at(L, idx);
  return (v->kind == K_USERDATA || v->kind == K_TENSOR) ? v->p : NULL;
}
int lua_gettop(lua_State *L) { return L->nargs + L->top; }
void lua_newtable(lua_State *L) {
  value v = {K_TABLE, NULL, NULL, 0};
  push(L, v);
}
void lua_pushvalue(lua_State *L, int idx) { push(L, *at(L, idx)); }
void lua_setglobal(lua_State *L, const char *name) {
  if (L->top < 1) raise_error(L, "lua_setglobal: empty stack");
  snprintf(L->global_set, sizeof(L->global_set), "%s", name);
  L->top--;
}
void lua_getglobal(lua_State *L, const char *name) {
  value v = {K_TABLE, NULL, name, 0};
  if (strcmp(name, "cutorch") != 0) raise_error(L, "lua_getglobal: only `cutorch` exists in the harness");
  push(L, v);
}
void lua_getfield(lua_State *L, int idx, const char *k) {
  const value *t = at(L, idx);
  value v = {K_FUNC, NULL, k, 0};
  if (t->kind != K_TABLE || !t->tname || strcmp(t->tname, "cutorch") != 0 || strcmp(k, "getState") != 0)
    raise_error(L, "lua_getfield: only cutorch.getState exists in the harness");
  push(L, v);
}
void lua_call(lua_State *L, int nargs, int nresults) {
  value r = {K_USERDATA, NULL, NULL, 0};
  if (nargs != 0 || nresults != 1 || L->top < 1 || L->stk[L->top - 1].kind != K_FUNC) raise_error(L, "lua_call: unexpected call");
  L->top--;
  r.p = &L->cutorch;
  push(L, r);
}
void lua_settop(lua_State *L, int idx) {
  if (idx >= 0) raise_error(L, "lua_settop: only lua_pop is supported");
  {
    const int n = -idx - 1;
    if (n > L->top) raise_error(L, "lua_pop: stack underflow");
    L->top -= n;
  }
}

/* ---- luaT subset ---- */
void *luaT_checkudata(lua_State *L, int ud, const char *tname) {
  const value *v = at(L, ud);
  char buf[160];
  if (v->kind != K_TENSOR || !v->tname || strcmp(v->tname, tname) != 0) {
    snprintf(buf, sizeof(buf), "bad argument #%d (%s expected, got %s)", ud, tname,
             v->kind == K_TENSOR ? v->tname : (v->kind == K_NUMBER ? "number" : "nil"));
    raise_error(L, buf);
  }
  return v->p;
}
int luaT_pushmetatable(lua_State *L, const char *tname) {
  value v = {K_TABLE, NULL, tname, 0};
  push(L, v);
  L->meta_on_top = tname;
  return 1;
}
void luaT_registeratname(lua_State *L, const struct luaL_Reg *methods, const char *name) {
  if (L->top < 1 || L->stk[L->top - 1].kind != K_TABLE || !L->stk[L->top - 1].tname)
    raise_error(L, "luaT_registeratname: no metatable on top of the stack");
  for (; methods->name; methods++) {
    if (L->nreg >= 96) raise_error(L, "harness: registration table full");
    snprintf(L->reg[L->nreg].type, sizeof(L->reg[0].type), "%s", L->stk[L->top - 1].tname);
    snprintf(L->reg[L->nreg].table, sizeof(L->reg[0].table), "%s", name);
    L->reg[L->nreg].name = methods->name;
    L->reg[L->nreg].fn = methods->func;
    L->nreg++;
  }
}

void THCudaTensor_resize1d(THCState *s, THCudaTensor *t, long size0) {
  (void)s;
  if (size0 > t->capacity) THError("harness: workspace tensor too small (%ld > %ld elements)", size0, t->capacity);
  t->nDimension = 1;
  t->size[0] = size0;
  t->stride[0] = 1;
}

/* ---- what the Python side drives ---- */
#if defined(HARNESS_OPEN_CPU)
LUA_EXTERNC int luaopen_libvolstn(lua_State *L);
#define HARNESS_OPEN luaopen_libvolstn
#elif defined(HARNESS_OPEN_CUDA)
LUA_EXTERNC int luaopen_libcuvolstn(lua_State *L);
#define HARNESS_OPEN luaopen_libcuvolstn
#else
#error "define HARNESS_OPEN_CPU or HARNESS_OPEN_CUDA"
#endif

lua_State *harness_new(void) { return (lua_State *)calloc(1, sizeof(lua_State)); }
void harness_free(lua_State *L) { free(L); }

/* require 'libvolstn' / 'libcuvolstn': returns what the luaopen function returns, -1 on a raised error */
int harness_open(lua_State *L) {
  int r;
  L->nargs = 0;
  L->top = 0;
  L->error[0] = 0;
  g_cur = L;
  L->armed = 1;
  if (setjmp(L->jb)) {
    L->armed = 0;
    return -1;
  }
  r = HARNESS_OPEN(L);
  L->armed = 0;
  return r;
}
int harness_stack_height(lua_State *L) { return L->top; } /* values left on the stack by the last call */
const char *harness_global_set(lua_State *L) { return L->global_set; }
int harness_nreg(lua_State *L) { return L->nreg; }
const char *harness_reg_type(lua_State *L, int i) { return L->reg[i].type; }
const char *harness_reg_table(lua_State *L, int i) { return L->reg[i].table; }
const char *harness_reg_name(lua_State *L, int i) { return L->reg[i].name; }
const char *harness_error(lua_State *L) { return L->error; }

void harness_args_clear(lua_State *L, int nargs) {
  int i;
  for (i = 0; i < 16; i++) {
    L->arg[i].kind = K_NIL;
    L->arg[i].p = NULL;
    L->arg[i].tname = NULL;
  }
  L->nargs = nargs;
}
void harness_arg_tensor(lua_State *L, int slot, const char *tname, void *tensor) {
  L->arg[slot].kind = K_TENSOR;
  L->arg[slot].p = tensor;
  L->arg[slot].tname = tname;
}
void harness_arg_number(lua_State *L, int slot, double v) {
  L->arg[slot].kind = K_NUMBER;
  L->arg[slot].num = v;
}
void harness_set_stream(lua_State *L, void *stream) { L->cutorch.current_stream = stream; }

/* tensor.nn.<name>(self, ...) for a tensor of luaT type `type`: -2 = no such function registered, -1 = error raised */
int harness_call(lua_State *L, const char *type, const char *name) {
  volatile int i;
  int r;
  for (i = 0; i < L->nreg; i++)
    if (!strcmp(L->reg[i].type, type) && !strcmp(L->reg[i].table, "nn") && !strcmp(L->reg[i].name, name)) break;
  if (i == L->nreg) return -2;
  L->top = 0;
  L->error[0] = 0;
  g_cur = L;
  L->armed = 1;
  if (setjmp(L->jb)) {
    L->armed = 0;
    return -1;
  }
  r = L->reg[i].fn(L);
  L->armed = 0;
  return r;
}
