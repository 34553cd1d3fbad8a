"""The drop-in boundary itself (SURVEY 8 b): lua/init.c and lua/init.cu -- the replacements of stn3dtorch/init.c and
stn3dtorch/init.cu -- compiled against stand-in Torch7 headers and opened through a fake Lua state
(tests/luat_harness).  No GPU needed here: registration, names, stack discipline, argument checking.  The calls
themselves are exercised on the GPU in tests/test_gpu_lua_glue.py."""
import os
import re

import numpy as np
import pytest

from tests.luat_harness import luat

REF_GENERIC = "/root/reference/stn3dtorch/generic/BilinearSamplerVolumetric.c"
REF_CU = "/root/reference/stn3dtorch/BilinearSamplerVolumetric.cu"


@pytest.fixture(scope="module")
def built():
    import importlib.util
    spec = importlib.util.spec_from_file_location("build_glue", os.path.join(luat.ROOT, "lua", "build_glue.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from alignet_b200 import build as b
    b.build()
    return mod.build()


def test_luaopen_libvolstn_registers_the_reference_names(built):
    h = luat.Harness(luat.DROPIN_CPU)
    assert h.opened == 1                      # `return 1` (the table), init.c:20
    assert h.global_table() == "volstn"       # init.c:15-17
    assert h.stack_after_open == 1            # the new table stays, every metatable was popped (...c:840-842)
    regs = h.registrations()
    assert all(table == "nn" for _, table, _ in regs)
    for tname in ("torch.FloatTensor", "torch.DoubleTensor"):   # init.c:18-19
        names = [n for t, _, n in regs if t == tname]
        assert names[:4] == luat.REFERENCE_NAMES                 # generic/...c:831-836, same order
        assert names[4:8] == luat.NEW_NAMES
        assert names[8:] == (luat.AFFINE_NAMES if tname == "torch.FloatTensor" else [])   # the fused host module is float only
    assert {t for t, _, _ in regs} == {"torch.FloatTensor", "torch.DoubleTensor"}
    h.close()


def test_luaopen_libcuvolstn_registers_the_reference_names(built):
    h = luat.Harness(luat.DROPIN_CUDA)
    assert h.opened == 1 and h.stack_after_open == 1             # init.cu:12-14
    regs = h.registrations()
    assert [t for t, _, _ in regs] == ["torch.CudaTensor"] * 14  # ...cu:1083-1088
    assert [n for _, _, n in regs][:4] == luat.REFERENCE_NAMES   # ...cu:1074-1080 (the OnlyGrid line is commented out there)
    assert [n for _, _, n in regs][4:8] == luat.NEW_NAMES
    assert [n for _, _, n in regs][8:] == luat.FUSED_NAMES + luat.AFFINE_NAMES
    h.close()


@pytest.mark.skipif(not os.path.exists(REF_GENERIC), reason="reference sources not present on this box")
def test_names_match_the_reference_source_text(built):
    """The literal list in the harness against the reference's registration tables as they stand in its sources."""
    def table(path):
        src = open(path).read()
        body = src[src.rindex("luaL_Reg"):]
        body = body[:body.index("NULL")]
        body = "\n".join(l for l in body.splitlines() if not l.strip().startswith("//"))
        return re.findall(r'\{"(\w+)"', body)
    assert table(REF_GENERIC) == luat.REFERENCE_NAMES
    assert sorted(table(REF_CU)) == sorted(luat.REFERENCE_NAMES)   # the .cu lists the BHWD pair first; a table has no order


@pytest.mark.skipif(not os.path.exists(luat.REFERENCE_CPU), reason="oracle/_ref/libvolstn_refinit.so not built")
def test_reference_init_c_registers_the_same_through_the_same_harness(built):
    """stn3dtorch/init.c itself, compiled unmodified behind the same stand-in headers: same global, same types, same
    four names -- the drop-in registers a superset."""
    ref, ours = luat.Harness(luat.REFERENCE_CPU), luat.Harness(luat.DROPIN_CPU)
    assert ref.opened == ours.opened == 1 and ref.global_table() == ours.global_table() == "volstn"
    assert ref.stack_after_open == ours.stack_after_open
    assert set(ref.registrations()) <= set(ours.registrations())
    assert [n for t, _, n in ref.registrations() if t == "torch.FloatTensor"] == luat.REFERENCE_NAMES
    ref.close(), ours.close()


def test_type_dispatch_and_argument_errors_are_lua_errors(built):
    h = luat.Harness(luat.DROPIN_CPU)
    v = np.zeros((1, 2, 2, 2, 1), np.float32)
    g = np.zeros((1, 2, 2, 2, 4), np.float32)
    o = np.zeros((1, 2, 2, 2, 1), np.float32)
    # a Double tensor handed to the Float function: luaT_checkudata raises (generic/...c:444-446)
    with pytest.raises(luat.LuaError, match="torch.FloatTensor expected"):
        h.call("BilinearSamplerVolumetric_updateOutput", v, g.astype(np.float64), o)
    with pytest.raises(luat.LuaError, match="expected, got nil"):
        h.call("BilinearSamplerVolumetric_updateOutput", v, g)
    with pytest.raises(luat.LuaError, match="5-D tensor expected"):
        h.call("BilinearSamplerVolumetric_updateOutput", v[0], g, o)
    with pytest.raises(luat.LuaError, match="nil value"):
        h.call("BilinearSamplerVolumetric_updateGradInputOnlyGrid", v, g, o)   # commented out in the reference too
    # module-contract violations come back as the reference's `printf + THError("aborting")` (...cu:736-740)
    with pytest.raises(luat.LuaError, match="aborting"):
        h.call("BilinearSamplerVolumetric_updateOutput", v, np.zeros((2, 2, 2, 2, 4), np.float32), o)   # batch mismatch, lua:33
    with pytest.raises(luat.LuaError, match="bx4x4"):
        h.call("VolumetricGridGenerator_updateOutput", np.zeros((2, 3, 4), np.float32), np.zeros((2, 2, 2, 2, 4), np.float32))
    with pytest.raises(luat.LuaError, match="B x N x d1"):
        h.call("Cumsum_updateOutput", np.zeros((2, 2, 4, 4, 4), np.float32), np.zeros((2, 2, 4, 4, 4), np.float32))
    h.close()


def test_lua_files_call_only_registered_functions(built):
    """No Lua interpreter exists in this image, so the Lua classes cannot be run: check statically that every native function a
    file of lua/ calls (`<tensor>.nn.<Name>(self, ...)`, the dispatch idiom of BilinearSamplerVolumetric.lua:74,109) is
    registered by the glue library of the tensor types the class supports, with the argument count the C function reads."""
    import glob
    import re
    cpu, cuda = luat.Harness(luat.DROPIN_CPU), luat.Harness(luat.DROPIN_CUDA)
    reg_float = {n for t, _, n in cpu.registrations() if t == "torch.FloatTensor"}
    reg_cuda = {n for _, _, n in cuda.registrations()}
    cpu.close(); cuda.close()
    called = {}
    for path in glob.glob(os.path.join(luat.ROOT, "lua", "*.lua")):
        src = open(path).read()
        src = re.sub(r"--\[\[.*?\]\]", "", src, flags=re.S)              # block comments
        src = "\n".join(l.split("--")[0] for l in src.splitlines())      # line comments
        for m in re.finditer(r"\.nn\.([A-Za-z0-9_]+)\(", src):
            called.setdefault(m.group(1), set()).add(os.path.basename(path))
    assert called, "no native calls found: the pattern is stale"
    cuda_only = {"CageWarpVolumetric.lua", "CageWarpSmoothL1Volumetric.lua", "BilinearSamplerSmoothL1Volumetric.lua"}
    for name, files in called.items():
        assert name in reg_cuda, (name, files)                          # every class works on torch.CudaTensor
        if not files <= cuda_only:
            assert name in reg_float, (name, files)                     # ... and the others on torch.FloatTensor too
    # and nothing registered is orphaned except the G = 3 twins, which the reference registers and never wraps
    orphans = (reg_cuda | reg_float) - set(called)
    assert orphans == {"BilinearSamplerBHWD_updateOutput", "BilinearSamplerBHWD_updateGradInput"}, orphans


def test_lua_call_sites_match_the_stack_slots_the_c_functions_read():
    """Static: the number of arguments at every `x.nn.<Name>(self, ...)` call site of lua/*.lua against the highest stack slot
    the registered C function reads (luaT_checkudata / lua_tonumber).  The last slot of every function is the optional flags
    number (nil reads as 0), so a call may pass one argument fewer."""
    import glob
    import re

    def strip_lua(src):
        src = re.sub(r"--\[\[.*?\]\]", "", src, flags=re.S)
        return "\n".join(l.split("--")[0] for l in src.splitlines())

    def call_args(src, start):              # arguments of the call whose '(' is at `start`
        depth, n, i = 0, 1, start
        while True:
            c = src[i]
            if c in "([{":
                depth += 1
            elif c in ")]}":
                depth -= 1
                if depth == 0:
                    return n
            elif c == "," and depth == 1:
                n += 1
            i += 1

    csrc = open(os.path.join(luat.ROOT, "lua", "init.cu")).read() + open(os.path.join(luat.ROOT, "lua", "generic", "volstn_host.c")).read()

    def bodies(fname):                      # bodies of `static int cunn_<fname>(lua_State *L...)` / `nn_(<fname>)(...)`, brace-matched
        out = []
        for m in re.finditer(r"static int (?:cunn_" + fname + r"|nn_\(" + fname + r"\))\(lua_State \*L[^)]*\) \{", csrc):
            depth, i = 1, m.end()
            while depth:
                depth += {"{": 1, "}": -1}.get(csrc[i], 0)
                i += 1
            out.append(csrc[m.end():i])
        return out

    def max_slot(name):
        slots = []
        for body in bodies(name):
            for helper in re.findall(r"return (?:cunn_|nn_\()(volstn_[A-Za-z_]+)\)?\(L", body):      # thin wrappers (sampler, cumsum)
                body += "".join(bodies(helper))
            slots.append(max(int(x) for x in re.findall(r"(?:luaT_checkudata|lua_tonumber)\(L, (\d+)", body)))
        assert slots, name
        return slots

    checked = 0
    for path in glob.glob(os.path.join(luat.ROOT, "lua", "*.lua")):
        src = strip_lua(open(path).read())
        for m in re.finditer(r"\.nn\.([A-Za-z0-9_]+)\(", src):
            nargs = call_args(src, m.end() - 1)
            for top in max_slot(m.group(1)):
                if nargs in (top, top - 1):
                    break
            else:
                raise AssertionError((os.path.basename(path), m.group(1), nargs, max_slot(m.group(1))))
            checked += 1
    assert checked >= 14, checked
