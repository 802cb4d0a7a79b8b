"""In-tree build of the native parts of pysparse_b200.

    python -m pysparse_b200.build [--force] [--verbose]

* ``pysparse_b200/libpsp_b200.so``  -- CUDA kernels + C-ABI (``include/psp_b200.h``),
  nvcc, sm_100a only (cross-compiles without a GPU).
* ``pysparse_b200/{sparse/spmatrix,itsolvers/krylov,precon/precon}.so`` -- the
  CPython extension modules (C, ``pysparse_b200/host/*.c``) that mirror the
  reference's ``spmatrix`` / ``krylov`` / ``precon`` modules and call the C-ABI.

Artefacts are git-ignored but travel to the GPU box with ``gpurun``.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
OBJ = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(PKG, "libpsp_b200.so")

CU_SOURCES = ["psp_runtime.cu", "psp_spmv.cu", "psp_runs.cu", "psp_solvers.cu", "psp_precon.cu", "psp_correq.cu", "psp_dist.cu", "psp_multi.cu", "psp_persist.cu", "psp_persist_bicg.cu",
              "psp_persist_gmres.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-I" + os.path.join(ROOT, "include")]
NVCC_FLAGS += os.environ.get("PSP_NVCC_EXTRA", "").split()      # e.g. -DPSP_PERSIST_PROF (diagnostic builds)

EXT_MODULES = {
    # output (relative to PKG)         sources (relative to HOST)
    "sparse/spmatrix": ["spmatrixmodule.c"],
    "itsolvers/krylov": ["krylovmodule.c"],
    "precon/precon": ["preconmodule.c"],
    "eigen/correq": ["correqmodule.c"],
}


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA library cannot be built")


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps if os.path.exists(d))


def _run(cmd: list[str], log: str | None = None, verbose: bool = False) -> None:
    if verbose:
        print(" ".join(cmd), flush=True)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        sys.stderr.write(p.stdout)
        raise RuntimeError(f"command failed ({p.returncode}): {' '.join(cmd)}")


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "psp_b200.h"))
    sources = [os.path.join(CSRC, s) for s in CU_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if not force and _newer(LIB, sources + headers):
        return LIB
    nvcc = _nvcc()
    objs = []

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        if force or not _newer(obj, [src] + headers):
            _run([nvcc, *NVCC_FLAGS, "-c", src, "-o", obj], log=obj + ".log", verbose=verbose)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-ldl"],
         verbose=verbose)
    return LIB


def build_extensions(force: bool = False, verbose: bool = False) -> list[str]:
    import numpy

    suffix = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    inc = [sysconfig.get_paths()["include"], numpy.get_include(), os.path.join(ROOT, "include"), HOST]
    outs = []
    hdrs = [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".h")] if os.path.isdir(HOST) else []
    for rel, srcs in EXT_MODULES.items():
        srcs_abs = [os.path.join(HOST, s) for s in srcs]
        if not all(os.path.exists(s) for s in srcs_abs):
            continue
        out = os.path.join(PKG, rel + suffix)
        outs.append(out)
        if not force and _newer(out, srcs_abs + hdrs + [os.path.join(ROOT, "include", "psp_b200.h")]):
            continue
        depth = rel.count("/")
        rpath = "$ORIGIN/" + "/".join([".."] * depth) if depth else "$ORIGIN"
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-Wall", "-Wno-unused-function",
               "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
               *["-I" + i for i in inc], *srcs_abs, "-o", out,
               "-L" + PKG, "-lpsp_b200", "-Wl,-rpath," + rpath, "-lm"]
        _run(cmd, verbose=verbose)
    return outs


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_cuda(force, verbose)
    build_extensions(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print("built:", LIB)
