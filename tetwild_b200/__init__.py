"""tetwild_b200 -- TetWild's data-parallel hot path on NVIDIA B200 (sm_100a).

The product is `libtetwild_gpu.so` (hand-written CUDA behind the C ABI of include/tetwild_gpu.h). This package is the
thin Python host mirror used by the tests and bench.py: ctypes in, numpy / torch buffers out. There is no CPU
fallback: importing works anywhere, but every compute entry point needs the CUDA library and a B200-class GPU and
raises loudly otherwise.
"""
from .api import (  # noqa: F401
    Context,
    Surface,
    TetMesh,
    Winding,
    TetWildGPUError,
    lib_path,
    load_library,
    MAX_ENERGY,
    NO_FACET,
)
