"""evplp_b200 -- B200-native (sm_100a) energy-compensated VPL hot path.

The product is libevplp_b200.so (CUDA kernels behind the C ABI of include/evplp.h) plus the
C++ host classes in evplp_b200/host/.  This Python package is plumbing for tests and
bench.py: ctypes bindings (`_capi`), a handle wrapper (`device.Device`) and numpy scene
containers (`scene`).  Nothing here computes on the CPU.
"""
from . import _capi as capi  # noqa: F401
from .device import Device  # noqa: F401
from .scene import Camera, Material, Mesh, Scene, cornell_scene, make_params  # noqa: F401
