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
    hdr = os.path.join(os.path.dirname(PKG_DIR), "include", "slsgp.h")
    return not _is_current(LIB_PATH, _sources() + [hdr])


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
    _stamp(LIB_PATH, _sources() + [os.path.join(os.path.dirname(PKG_DIR), "include", "slsgp.h")])
    return LIB_PATH


HOST_DIR = os.path.join(PKG_DIR, "host")
HOST_LIB_PATH = os.path.join(LIB_DIR, "libsls_b200_host.so")


def _tree_hash(paths) -> str:
    """Content hash of source files. Staleness is decided by content, not by mtime: the snapshot that travels to the GPU
    box does not preserve modification times, and a spurious rebuild there costs GPU-minutes."""
    import hashlib
    h = hashlib.sha256()
    root = os.path.dirname(PKG_DIR)
    for p in sorted(paths):
        h.update(os.path.relpath(p, root).encode())  # relative: the repository lives under a different path on the GPU box
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _is_current(out: str, sources) -> bool:
    stamp = out + ".srchash"
    return os.path.exists(out) and os.path.exists(stamp) and open(stamp).read().strip() == _tree_hash(sources)


def _stamp(out: str, sources) -> None:
    with open(out + ".srchash", "w") as f:
        f.write(_tree_hash(sources))


def _host_sources():
    inc = os.path.join(HOST_DIR, "include", "sequential-line-search")
    src = os.path.join(HOST_DIR, "src")
    return ([os.path.join(inc, f) for f in os.listdir(inc)] + [os.path.join(src, f) for f in os.listdir(src)] +
            [os.path.join(HOST_DIR, "Makefile"), os.path.join(os.path.dirname(PKG_DIR), "include", "slsgp.h")])


def build_nlopt() -> None:
    """libnlopt.a + nlopt.hpp from the NLopt sources the reference vendors (third_party/nlopt/Makefile). A no-op where those
    sources are absent (the GPU box): the prebuilt files travel with the snapshot. With them the host layer offers the
    Reference / Hybrid search drivers; without them it is built Native-only."""
    res = subprocess.run(["make", "-s", "-C", os.path.join(os.path.dirname(PKG_DIR), "third_party", "nlopt")], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("NLopt build failed:\n" + res.stdout + res.stderr)


def build_host(force: bool = False) -> str:
    """Build libsls_b200_host.so: the C++ mirror of the reference's Regressor / acquisition_func interface above the
    C ABI (g++, links libslsgp.so with an $ORIGIN rpath). Uses the system Eigen when EIGEN_INC names it, else the
    repository's eigen-lite subset."""
    build()
    build_nlopt()
    if not force and _is_current(HOST_LIB_PATH, _host_sources()):
        return HOST_LIB_PATH
    res = subprocess.run(["make", "-s", "-B", "-C", HOST_DIR], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("host layer build failed:\n" + res.stdout + res.stderr)
    _stamp(HOST_LIB_PATH, _host_sources())
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
    deps = [PY_SRC] + _host_sources()
    if not force and _is_current(out, deps):
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
    _stamp(out, deps)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
    print(build_host(force=True))
    print(build_python_module(force=True))
