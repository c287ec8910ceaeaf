"""eventql_b200 - B200-native columnar scan / filter / GROUP BY behind EventQL's operator surface.

  eventql_b200.plan   query plans in the wire shape of include/evqgpu.h (pure Python)
  eventql_b200.capi   ctypes binding of the C ABI (libevqgpu.so: CUDA kernels for sm_100a)
  eventql_b200.build  in-tree build of the native libraries
"""
from . import plan  # noqa: F401

__all__ = ["plan"]
