"""sequential-line-search_b200: the B200 (sm_100a) Gaussian-process hot path of sequential-line-search.

Contents: csrc/ (CUDA kernels + the C ABI of include/slsgp.h), build.py (nvcc driver), binding.py (ctypes stub),
sharding.py (host-side split / winner selection of the multi-GPU candidate sweep), host/ (the C++ mirror of the
reference's Regressor / acquisition_func interface above the C ABI) and hostlib.py (its ctypes view, for tests).
The directory name is not a Python identifier; import it with
    importlib.import_module("sequential-line-search_b200")
"""
from .build import build, build_host, build_python_module, python_module_path, LIB_PATH, LIB_DIR, HOST_LIB_PATH  # noqa: F401
from .binding import *  # noqa: F401,F403
from .binding import Context, SlsgpError, load_library, API_SYMBOLS  # noqa: F401
from . import sharding  # noqa: F401
from . import hostlib  # noqa: F401
