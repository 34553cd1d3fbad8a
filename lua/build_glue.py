"""Compile the luaT glue (lua/init.c -> libvolstn.so, lua/init.cu -> libcuvolstn.so) against the stand-in Torch7
headers of tests/luat_harness/ and link it with alignet_b200/csrc/libvolstn_b200.so.

    python lua/build_glue.py

No Lua, luaT, TH or THC exists in this image, so this is how the glue is proven to compile and how the tests drive
its lua_CFunctions (tests/test_lua_glue.py, tests/test_gpu_lua_glue.py).  In a real Torch7 tree the same two files are
built by lua/CMakeLists.txt against the real headers; nothing in them refers to the harness.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "build")
HARNESS = os.path.join(ROOT, "tests", "luat_harness")
CSRC = os.path.join(ROOT, "alignet_b200", "csrc")
CPU_SO = os.path.join(OUT, "libvolstn.so")
CUDA_SO = os.path.join(OUT, "libcuvolstn.so")

INC = ["-I", os.path.join(HARNESS, "include"), "-I", os.path.join(ROOT, "include"), "-I", HERE]
INPUTS = [os.path.join(HERE, f) for f in ("init.c", "init.cu", "volstn_glue.h", os.path.join("generic", "volstn_host.c"))] + \
         [os.path.join(HARNESS, "harness.c")] + \
         [os.path.join(HARNESS, "include", f) for f in ("TH.h", "THC.h", "THGenerateFloatTypes.h", "lua.h", "lauxlib.h", "luaT.h")] + \
         [os.path.join(ROOT, "include", "volstn_b200.h"), os.path.abspath(__file__)]


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("glue build failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)


def up_to_date() -> bool:
    if not (os.path.exists(CPU_SO) and os.path.exists(CUDA_SO)):
        return False
    newest = max(os.path.getmtime(p) for p in INPUTS)
    return min(os.path.getmtime(CPU_SO), os.path.getmtime(CUDA_SO)) >= newest


def build(force: bool = False):
    if not force and up_to_date():
        return CPU_SO, CUDA_SO
    os.makedirs(OUT, exist_ok=True)
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    link = ["-L", CSRC, "-lvolstn_b200"]
    # libvolstn: plain C, as luarocks / cmake would build stn3dtorch/init.c
    _run(["gcc", "-O2", "-Wall", "-Wextra", "-Werror", "-fPIC", "-shared", "-DHARNESS_OPEN_CPU", *INC, os.path.join(HERE, "init.c"),
          os.path.join(HARNESS, "harness.c"), "-o", CPU_SO, *link, "-Wl,-rpath,$ORIGIN/../../alignet_b200/csrc", "-lpthread"])
    # libcuvolstn: init.cu through nvcc (C++), as CUDA_ADD_LIBRARY would
    obj, hobj = os.path.join(OUT, "init_cu.o"), os.path.join(OUT, "harness_cuda.o")
    _run([nvcc, "-O2", "-x", "cu", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC,-Wall,-Werror", *INC, "-c",
          os.path.join(HERE, "init.cu"), "-o", obj])
    _run(["gcc", "-O2", "-Wall", "-fPIC", "-DHARNESS_OPEN_CUDA", *INC, "-c", os.path.join(HARNESS, "harness.c"), "-o", hobj])
    _run([nvcc, "-shared", "-cudart", "static", "-o", CUDA_SO, obj, hobj, *link, "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../alignet_b200/csrc"])
    return CPU_SO, CUDA_SO


if __name__ == "__main__":
    print(*build(force=True), sep="\n")
