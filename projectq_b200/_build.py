"""Builds the native pieces in-tree: libpqb200.so (CUDA, sm_100a) and the pybind11 shim module.

Called by ``__graft_entry__.build()`` and usable as ``python -m projectq_b200._build``.  nvcc cross-compiles for
sm_100a without a GPU; the resulting .so files are git-ignored but travel to the GPU box with the snapshot.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libpqb200.so")
SHIM = os.path.join(HERE, "_pqb_shim" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
LIB_SOURCES = ["kernels.cu", "engine.cpp", "fuser.cpp", "devmem.cpp", "dist.cpp", "fdpass.cpp", "capi.cpp"]
LIB_HEADERS = ["kernels.cuh", "engine.h", "fuser.h", "devmem.h", "dist.h", "fdpass.h", "bits.h", "../../include/pqb200.h"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build_lib(force=False):
    srcs = [os.path.join(CSRC, s) for s in LIB_SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in LIB_HEADERS]
    if not force and not _stale(LIB, deps):
        return LIB
    objs = []
    odir = os.path.join(HERE, "build")
    os.makedirs(odir, exist_ok=True)
    for s in srcs:
        o = os.path.join(odir, os.path.basename(s) + ".o")
        if force or _stale(o, [s] + deps[len(srcs):]):
            _run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                  "-Xcompiler", "-fPIC,-fvisibility=hidden", "-I", os.path.join(ROOT, "include"),
                  "-x", "cu" if s.endswith(".cu") else "c++", "-c", s, "-o", o])
        objs.append(o)
    _run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-cudart", "shared", "-ldl"])
    return LIB


def build_shim(force=False):
    import pybind11

    src = os.path.join(CSRC, "shim.cpp")
    if not force and not _stale(SHIM, [src, os.path.join(ROOT, "include", "pqb200.h"), LIB]):
        return SHIM
    _run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden",
          "-I", sysconfig.get_paths()["include"], "-I", pybind11.get_include(), "-I", os.path.join(ROOT, "include"),
          src, "-o", SHIM, "-L", HERE, "-lpqb200", "-Wl,-rpath,$ORIGIN"])
    return SHIM


def build_all(force=False):
    build_lib(force)
    build_shim(force)
    return LIB, SHIM


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
