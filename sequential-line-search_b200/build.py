"""Compile libslsgp.so (the C-ABI CUDA library) in-tree for sm_100a with nvcc.

The shared object lands in sequential-line-search_b200/lib/ so that it travels with the repository snapshot to the
GPU box; there is no JIT cache and no torch extension machinery involved (the library has no torch dependency).
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libslsgp.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    hdr = os.path.join(os.path.dirname(PKG_DIR), "include", "slsgp.h")
    return any(os.path.getmtime(s) > t for s in _sources() + [hdr])


def build(force: bool = False, verbose: bool = False) -> str:
    """Build libslsgp.so if it is missing or older than its sources. Returns the library path."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libslsgp.so cannot be built (and there is no CPU fallback)")
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, os.path.join(CSRC, "slsgp.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


HOST_DIR = os.path.join(PKG_DIR, "host")
HOST_LIB_PATH = os.path.join(LIB_DIR, "libsls_b200_host.so")


def build_host(force: bool = False) -> str:
    """Build libsls_b200_host.so: the C++ mirror of the reference's Regressor / acquisition_func interface above the
    C ABI (g++, links libslsgp.so with an $ORIGIN rpath). Uses the system Eigen when EIGEN_INC names it, else the
    repository's eigen-lite subset."""
    build()
    cmd = ["make", "-s", "-C", HOST_DIR] + (["-B"] if force else [])
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("host layer build failed:\n" + res.stdout + res.stderr)
    return HOST_LIB_PATH


PY_SRC = os.path.join(PKG_DIR, "python", "pySequentialLineSearch.cpp")


def python_module_path() -> str:
    import sysconfig
    return os.path.join(LIB_DIR, "pySequentialLineSearch" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_python_module(force: bool = False) -> str:
    """Build pySequentialLineSearch (pybind11): the reference's Python module over the host layer. Import it with
    sequential-line-search_b200/lib on sys.path."""
    build_host()
    out = python_module_path()
    deps = [PY_SRC, HOST_LIB_PATH] + [os.path.join(HOST_DIR, "include", "sequential-line-search", f)
                                      for f in os.listdir(os.path.join(HOST_DIR, "include", "sequential-line-search"))]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    import pybind11
    import sysconfig
    eigen = os.environ.get("EIGEN_INC", os.path.join(os.path.dirname(PKG_DIR), "include", "eigen-lite"))
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"],
           "-I" + os.path.join(HOST_DIR, "include"), "-I" + eigen, "-o", out, PY_SRC, "-L" + LIB_DIR, "-lsls_b200_host", "-lslsgp",
           "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("pySequentialLineSearch build failed:\n" + res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
    print(build_host(force=True))
    print(build_python_module(force=True))
