"""Build alignet_b200/csrc/libvolstn_b200.so -- the C-ABI library of hand-written sm_100a kernels.

    python -m alignet_b200.build [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU.  The library is built IN-TREE (it is git-ignored
but travels to the GPU box with the repo snapshot) and links the CUDA runtime statically, so it
depends on nothing but the driver.  There is exactly one code path: sm_100a SASS (plus compute_100a
PTX is NOT embedded) -- loading it on any other device fails loudly.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(CSRC, "build")
LIB = os.path.join(CSRC, "libvolstn_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

SOURCES = ["abi.cu", "sampler.cu", "sampler_tile.cu", "sampler_tma.cu", "rowwarp.cu", "cagebrick.cu", "fused_api.cu", "gridgen.cu", "cumsum.cu", "host_api.cu"]
HEADERS = ["volstn_common.cuh", "sampler.cuh", "rowwarp.cuh", "cagebrick.cuh", "volstn_internal.h", os.path.join(INCLUDE, "volstn_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # never contract a multiply and an add: parity is defined on the reference's unfused operation order.  The
    # scalar __f*_rn intrinsics already guarantee it; the packed __fmul2_rn/__fadd2_rn ones do NOT (ptxas
    # fused them into FFMA2), so contraction is switched off for the whole library.
    "-fmad=false",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newest_input() -> float:
    paths = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    paths += [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    paths.append(os.path.abspath(__file__))
    return max(os.path.getmtime(p) for p in paths)


def up_to_date() -> bool:
    return os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_input()


def build(force: bool = False, verbose: bool = False, defines=(), out: str = None) -> str:
    """``defines`` / ``out``: a tuning VARIANT of the library (e.g. ("VOLSTN_BWD_CTAS=5",) ->
    csrc/variants/libvolstn_b200_ctas5.so) for the sweep scripts under tools/ (they load it through VOLSTN_B200_LIB)."""
    if out is not None:
        return _build(list(defines), out, os.path.join(BUILD, "variant_" + os.path.basename(out).replace(".so", "")), verbose)
    if not force and up_to_date():
        return LIB
    return _build([], LIB, BUILD, verbose)


def _build(defines, LIB, BUILD, verbose):
    nvcc = _nvcc()
    os.makedirs(BUILD, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]

    def compile_one(src: str):
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *["-D" + d for d in defines], "-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    objs = []
    log = []
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        for src, obj, r in ex.map(compile_one, srcs):
            log.append(f"==== {src}\n{r.stdout}{r.stderr}")
            if r.returncode != 0:
                sys.stderr.write(log[-1])
                raise RuntimeError(f"nvcc failed on {src}")
            objs.append(obj)
    with open(os.path.join(BUILD, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
            "-Xcompiler", "-fPIC", "-o", LIB, *objs]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("-D", dest="defines", action="append", default=[], help="build a tuning variant with this macro")
    ap.add_argument("--out", default=None, help="path of the variant library")
    a = ap.parse_args()
    print(build(a.force, a.verbose, a.defines, a.out))
