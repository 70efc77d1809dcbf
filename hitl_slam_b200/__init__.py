"""hitl_slam_b200 — B200-native (sm_100a) back-end for HitL-SLAM's data-parallel hot path.

The product is the C-ABI library lib/libhitl_gpu.so (include/hitl_gpu.h) plus the C++ host
mirror lib/libhitl_host.so; this Python package is only plumbing (ctypes bindings, the
synthetic-trajectory generator, the in-tree build).  No CPU fallback exists.
"""
from .capi import HitlGpu, HostLib, HostSession, HitlError, default_min_cos, ABI_SYMBOLS, KDNODE  # noqa: F401
