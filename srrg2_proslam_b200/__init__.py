"""srrg2_proslam_b200 -- B200-native (sm_100a) CUDA frontend for srrg2_proslam.

csrc/   hand-written CUDA kernels + the C ABI (include/pslam_cuda.h) -> libpslam_cuda.so
host/   C++ mirror of the srrg2 Configurable plugin classes above the C ABI -> libpslam_plugin.so
capi.py ctypes harness over the C ABI (tests, bench)
"""
from . import capi  # noqa: F401
